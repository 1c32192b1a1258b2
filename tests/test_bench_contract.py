"""bench.py on a machine without a GPU: the reference arm (`--impl reference`) runs here, and its JSON line must carry
the keys of the driver contract with the same `config` object the GPU arm prints."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                        "--steps", "1", "--warmup", "1", "--ref-rays", "64", "--no-python-ref"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1 and lines[0].startswith("{"), p.stdout[:400]  # stdout is the ONE JSON line, nothing else
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "ray_km_per_s" and d["unit"] == "ray-km/s" and d["dtype"] == "f64"
    assert d["value"] > 1e4 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "64 rays" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "ray-km/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_both_arms_share_one_config_object_and_workload_definition():
    sys.path.insert(0, ROOT)
    import bench

    for name, n_total, scaling in (("cfg3", 1 << 20, "strong"), ("cfg2", 10000, "weak")):
        w = bench.workload(name, 0, 1)
        c = bench.config_block(w, 1)
        assert c["rays_total"] == n_total and c["scaling"] == scaling and c["workload"].startswith(name)
        assert len(w["angles"]) == n_total and np.array_equal(w["angles"], w["all_angles"])
        sub = bench.sample_angles(w, 64)
        assert len(sub) == 64 and sub[0] == w["all_angles"][0] and sub[-1] == w["all_angles"][-1]
    # strong scaling: the ranks' shares partition ONE fan; weak scaling: the fan grows with the ranks
    shares = [bench.workload("cfg3", r, 4)["angles"] for r in range(4)]
    assert sum(len(s) for s in shares) == 1 << 20 and np.array_equal(np.sort(np.concatenate(shares)), bench.workload("cfg3")["all_angles"])
    weak = [bench.workload("cfg2", r, 4) for r in range(4)]
    assert weak[1]["n_total"] == 40000 and sum(len(w["angles"]) for w in weak) == 40000
    assert all(abs(len(w["angles"]) - 10000) <= 32 for w in weak)  # cyclic blocks of 32 angles
    flops = bench.algorithmic_flops(np.array([[10], [2], [1], [1], [4], [1]], dtype=np.uint32))
    assert flops == 38 * 10 + 200 * 2 + 40 * 1 + 168 * 1 + 60 * 4


def test_stdout_is_reserved_for_the_json_line():
    """Libraries write to the C stdout (NCCL prints its version banner there under torchrun): bench.py keeps a private
    handle for the JSON line and points file descriptor 1 at stderr for everything else."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._reserve_stdout(); "
            "os.write(1, b'NCCL version x.y\\n'); print('stray print'); bench._emit({'a': 1})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == '{"a": 1}\n'
    assert "NCCL version x.y" in p.stderr and "stray print" in p.stderr


def test_reference_arm_under_torchrun_only_rank_zero_works():
    """N > 1: the driver launches the reference arm under torchrun as well; rank 0 alone runs and prints, the other ranks
    exit 0 at once with nothing on stdout."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout == "", (p.stdout[:200], p.stderr[-500:])
    env["RANK"] = env["LOCAL_RANK"] = "0"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                        "cfg2", "--steps", "1", "--warmup", "1", "--ref-rays", "64", "--no-python-ref"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["n_gpus"] == 2
