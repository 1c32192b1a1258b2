"""Test infrastructure (lives under tests/ because it calls the oracle): extended GPU-vs-oracle fuzz (same generator as tests/fuzz_cases.py, more seeds).  Usage: fuzz_gpu.py LO HI"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pygenray_b200 as pr
from conftest import make_env, oracle_env_from, assert_same_fan
from fuzz_cases import random_case
from oracle import oracle as orc
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    k = random_case(seed)
    env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
    oenv = oracle_env_from(env, k["flat"])
    rs = np.linspace(k["r0"], k["r1"], k["n_save"]) if k["n_save"] else None
    o = orc.shoot_fan(oenv, k["zs"], k["r0"], k["r1"], angles_deg=k["angles"], range_save=rs, rtol=k["rtol"],
                      terminate_backwards=k["terminate_backwards"])
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=k["flat"]))
    res = pr.shoot_fan_arrays(denv, k["zs"], k["r0"], k["r1"], angles=k["angles"], num_range_save=k["n_save"], counters=True,
                              rtol=k["rtol"], terminate_backwards=k["terminate_backwards"])
    g = dict(final=res["final"].T.copy(), n_bottom=res["n_bottom"], n_surface=res["n_surface"], status=res["status"],
             counters=res["counters"].T.astype(np.int64), traj=np.transpose(res["traj"], (1, 0, 2)) if k["n_save"] else None)
    denv.close()
    try:
        assert_same_fan(o, g, f"seed {seed}")
    except AssertionError as e:
        bad += 1; print("MISMATCH", seed, str(e)[:200], flush=True)
print(f"seeds {lo}..{hi - 1}: mismatches {bad}")
