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
    import workloads as cases
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
    import workloads as cases
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


# ---- eigenray brackets and ensemble realisations over ranks ----------------------------------------------------
def _np_bracket(z_end, thetas, receiver_depths):
    """numpy restatement of eigenrays.py:62-79 in the layout of pygenray_b200.find_brackets (CPU tensors)."""
    z, th = z_end.numpy(), thetas.numpy()
    ks = [np.where(np.diff(np.sign(z + rd)))[0] for rd in receiver_depths]
    counts = np.array([len(k) for k in ks], dtype=np.int64)
    k = np.concatenate(ks).astype(np.int64)
    br = dict(z1=z[k], z2=z[k + 1], th1=th[k], th2=th[k + 1], rd=np.repeat(np.asarray(receiver_depths, float), counts),
              rcv=np.repeat(np.arange(len(ks), dtype=np.int32), counts), k=k)
    return {kk: torch.from_numpy(np.ascontiguousarray(v)) for kk, v in br.items()}, counts


def _oracle_refine(oenv, br, source_depth, source_range, receiver_range, rtol=1e-9, terminate_backwards=True,
                   ztol=1.0, max_iter=20):
    from oracle import oracle as orc

    h = {k: v.numpy() for k, v in br.items()}
    o = orc.refine_brackets(oenv, h["z1"], h["z2"], h["th1"], h["th2"], h["rd"], source_depth, source_range,
                            receiver_range, rtol=rtol, terminate_backwards=terminate_backwards, ztol=ztol,
                            max_iter=max_iter, n_threads=2)
    return dict(theta=torch.from_numpy(o["theta"]), final=torch.from_numpy(o["final"].T.copy()),
                n_bottom=torch.from_numpy(o["n_bottom"]), n_surface=torch.from_numpy(o["n_surface"]),
                status=torch.from_numpy(o["status"]), shots=torch.from_numpy(o["shots"]))


def _fan_table(orc, oenv, angles, R):
    o = orc.shoot_fan(oenv, 1000.0, 0.0, R, angles_deg=angles, n_threads=2)
    t = np.zeros((len(angles), 6))
    t[:, :3] = o["final"]; t[:, 3] = o["n_bottom"]; t[:, 4] = o["n_surface"]; t[:, 5] = o["status"]
    t[::17, :3] = np.nan  # stand-in for dropped rays (the table is compacted to survivors before bracketing)
    t[::17, 5] = 4
    return torch.from_numpy(t)


def _eigen_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import make_env, oracle_env_from
    import workloads as cases
    from oracle import oracle as orc
    from pygenray_b200 import dist as pdist

    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    internal = np.linspace(-16, 16, 300)
    table = _fan_table(orc, oenv, internal, 80e3)
    rds = np.linspace(300.0, 4000.0, 25)
    rec, counts = pdist.find_eigenrays_sharded(oenv, table, -internal, rds, 1000.0, 0.0, 80e3, rank, world, ztol=1e-2,
                                               max_iter=20, _bracket=_np_bracket, _refine=_oracle_refine)
    np.save(out_path.format(rank=rank), rec.numpy())
    # ensemble sharding: 5 "realisations" (ragged over 2 ranks) stood in by scaled launch-angle sets
    def shoot(dens, zs, r0, R, angles, **kw):
        first, count = dens
        fin = np.stack([orc.shoot_fan(oenv, zs, r0, R, angles_deg=np.asarray(angles) * (1 + 0.01 * (first + i)), n_threads=2)
                        for i in range(count)])
        return dict(final=torch.from_numpy(np.stack([f["final"].T for f in fin], 1).copy()),
                    n_bottom=torch.from_numpy(np.stack([f["n_bottom"] for f in fin])),
                    n_surface=torch.from_numpy(np.stack([f["n_surface"] for f in fin])),
                    status=torch.from_numpy(np.stack([f["status"] for f in fin])))

    class Dens(tuple):
        n_env = property(lambda self: self[1])

    first, count = pdist.shard_realisations(5, rank, world)
    table5, _ = pdist.shoot_ensemble_sharded(Dens((first, count)), 5, 1000.0, 0.0, 40e3, np.linspace(-10, 10, 21), rank,
                                             world, _shoot=shoot)
    np.save(out_path.format(rank=rank).replace("rec_", "ens_"), table5.numpy())
    dist.destroy_process_group()


def test_sharded_eigenrays_and_ensembles_equal_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from conftest import make_env, oracle_env_from
    import workloads as cases
    from oracle import oracle as orc
    from pygenray_b200 import dist as pdist

    world, port = 2, 29500 + (os.getpid() + 777) % 2000
    out = str(tmp_path / "rec_{rank}.npy")
    mp.spawn(_eigen_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = np.load(out.format(rank=0)), np.load(out.format(rank=1))
    assert np.array_equal(r0, r1, equal_nan=True) and r0.shape[1] == pdist.EIGEN_WIDTH
    # single process: same fan, same brackets, every bracket refined here
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    internal = np.linspace(-16, 16, 300)
    table = _fan_table(orc, oenv, internal, 80e3)
    assert (table[:, 5] != 0).sum() > 0  # some rays are dropped: k indexes the compacted fan
    keep = np.flatnonzero(table[:, 5].numpy() == 0)
    rds = np.linspace(300.0, 4000.0, 25)
    br, counts = _np_bracket(-table[keep, 1], torch.from_numpy(-internal[keep]), rds)
    ref = _oracle_refine(oenv, br, 1000.0, 0.0, 80e3, ztol=1e-2, max_iter=20)
    n_br = len(br["k"])
    assert r0.shape[0] == n_br > 64  # more than one cyclic block per rank
    assert np.array_equal(r0[:, 0], br["rcv"].numpy()) and np.array_equal(r0[:, 1], br["k"].numpy())
    assert np.array_equal(r0[:, 2], ref["theta"].numpy(), equal_nan=True)
    assert np.array_equal(r0[:, 3:6], ref["final"].numpy().T, equal_nan=True)
    assert np.array_equal(r0[:, 8], ref["status"].numpy()) and np.array_equal(r0[:, 9], ref["shots"].numpy())
    assert (r0[:, 8] == 0).sum() > 10
    e0, e1 = np.load(out.format(rank=0).replace("rec_", "ens_")), np.load(out.format(rank=1).replace("rec_", "ens_"))
    assert np.array_equal(e0, e1, equal_nan=True) and e0.shape == (5, 21, 6)
    for i in range(5):
        o = orc.shoot_fan(oenv, 1000.0, 0.0, 40e3, angles_deg=np.linspace(-10, 10, 21) * (1 + 0.01 * i))
        assert np.array_equal(e0[i, :, :3], o["final"], equal_nan=True) and np.array_equal(e0[i, :, 5], o["status"])
    assert pdist.shard_realisations(5, 0, 2) == (0, 3) and pdist.shard_realisations(5, 1, 2) == (3, 2)


def test_sharded_calls_reject_bad_arguments():
    sys.path.insert(0, ROOT)
    from pygenray_b200 import dist as pdist

    with pytest.raises(ValueError, match="compact"):
        pdist.shoot_fan_sharded(None, 0.0, 0.0, 1.0, np.zeros(4), 0, 1, compact=True)
    with pytest.raises(ValueError, match="rows"):
        pdist.all_gather_rows(torch.zeros((3, 6), dtype=torch.float64), 4, 0, 1)
    with pytest.raises(ValueError):
        pdist.all_gather_arrivals(torch.zeros((4, 5), dtype=torch.float64), 4, 0, 1)
    assert pdist.all_gather_rows(torch.zeros((0, 10), dtype=torch.float64), 0, 1, 4).shape == (0, 10)  # no collective
