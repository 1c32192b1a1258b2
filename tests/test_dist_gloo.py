"""N > 1 host logic on CPU: world_size-2 gloo run of the shard / all-gather plumbing of
pygenray_b200.dist.  The per-rank compute is stood in by the CPU oracle (there is no CPU product
path); what is tested is that block-cyclic sharding + one all-gather reproduces the single-process
arrival table exactly, including ragged shard sizes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_rays, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import make_env, oracle_env_from
    from oracle import cases
    from oracle import oracle as orc
    from pygenray_b200 import dist as pdist

    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    angles = np.linspace(-15, 15, n_rays)
    mine = pdist.shard_indices(n_rays, rank, world)
    o = orc.shoot_fan(oenv, 1000.0, 0.0, 60e3, angles_deg=angles[mine], n_threads=2)
    rec = torch.zeros((len(mine), pdist.RECORD_WIDTH), dtype=torch.float64)
    rec[:, :3] = torch.from_numpy(o["final"])
    rec[:, 3] = torch.from_numpy(o["n_bottom"].astype(np.float64))
    rec[:, 4] = torch.from_numpy(o["n_surface"].astype(np.float64))
    rec[:, 5] = torch.from_numpy(o["status"].astype(np.float64))
    table = pdist.all_gather_arrivals(rec, n_rays, rank, world)
    np.save(out_path.format(rank=rank), table.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [200, 97])
def test_sharded_fan_equals_single_process(tmp_path, n_rays):
    sys.path.insert(0, ROOT)
    from conftest import make_env, oracle_env_from
    from oracle import cases
    from oracle import oracle as orc
    from pygenray_b200 import dist as pdist

    world, port = 2, 29500 + (os.getpid() + n_rays) % 2000
    out = str(tmp_path / "table_{rank}.npy")
    mp.spawn(_worker, args=(world, port, n_rays, out), nprocs=world, join=True)
    t0, t1 = np.load(out.format(rank=0)), np.load(out.format(rank=1))
    assert np.array_equal(t0, t1, equal_nan=True)
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    o = orc.shoot_fan(oenv, 1000.0, 0.0, 60e3, angles_deg=np.linspace(-15, 15, n_rays))
    assert np.array_equal(t0[:, :3], o["final"], equal_nan=True)
    assert np.array_equal(t0[:, 3], o["n_bottom"]) and np.array_equal(t0[:, 5], o["status"])
    # shards partition the fan in cyclic blocks of 32
    a, b = pdist.shard_indices(n_rays, 0, 2), pdist.shard_indices(n_rays, 1, 2)
    assert np.array_equal(np.sort(np.concatenate([a, b])), np.arange(n_rays))
    assert a[0] == 0 and b[0] == 32 and (np.diff(a)[:31] == 1).all()
