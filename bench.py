#!/usr/bin/env python
"""bench.py -- ray-km/s of the ray-fan hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg1|cfg3]

One "step" = one pass of the hot path over one fan: every launch angle of the workload integrated
from the source to the receiver range, trajectories resampled on the save grid, arrival table out.
Default workload = BASELINE.json configs[1]: range-dependent 200x500 gridded SSP with a seamount,
10 000-ray fan (+-15 deg) to 500 km, 1001 save points per ray  (5e6 ray-km per step).

  value      ray-km/s with inputs resident in HBM (CUDA events around the kernels, max over ranks)
  e2e        same metric through the public API pygenray_b200.shoot_rays(): host launch angles in,
             host RayFan out, H2D/D2H inside the timed region
  roofline   FP64 FMA pipe: algorithmic flops from per-ray device counters / kernel time, against a
             DFMA microbenchmark measured in the same run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the CPU oracle port (oracle/prt_oracle.c) on the box's host cores, same workload
  --impl reference  times that CPU port alone (the reference itself is Python and cannot travel)
N > 1 (torchrun): weak scaling -- every rank shoots its own 10k-ray fan (interleaved launch angles of
an N x 10k fan) and the arrival tables are all-gathered over NCCL inside the step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from oracle import cases  # noqa: E402  (synthetic workload definitions shared with the tests)


# ------------------------------------------------------------------------------------------------
def workload(name, rank=0, world=1):
    """-> dict(arrays, flatearth, source_depth, source_range, receiver_range, angles, n_save, label)."""
    if name == "cfg1":
        a = cases.config1_arrays()
        n, lim, R, ns, flat = 101, 12.0, 100e3, 1001, True
        label = "cfg1: Munk range-independent 100x6000 grid (flat-earth), 101 rays +-12deg, 100 km, 1001 save pts"
    elif name == "cfg3":
        a = cases.config3_arrays()
        n, lim, R, ns, flat = 1 << 20, 15.0, 1000e3, 0, False
        label = "cfg3: Munk + internal waves 2001x1201 grid, 1,048,576 rays +-15deg, 1000 km, arrival table only"
    elif name == "cfg5":
        n_env = 128  # per GPU (1024 realisations over 8 GPUs)
        a = cases.config2_arrays()
        a["c_stack"] = np.stack([cases.config5_realisation(rank * n_env + i)["c"] for i in range(n_env)])
        n, lim, R, ns, flat = 4096, 15.0, 500e3, 0, False
        label = ("cfg5: tomography ensemble, 128 perturbed realisations (per GPU) of the 200x500 grid x 4096 rays "
                 "+-15deg, 500 km, arrival tables only")
        return dict(arrays=a, flatearth=flat, source_depth=1000.0, source_range=0.0, receiver_range=R,
                    angles=np.linspace(-lim, lim, n), n_total=n * n_env * world, n_save=ns, label=label, name=name,
                    rays_per_gpu=n * n_env, n_env=n_env)
    else:
        a = cases.config2_arrays()
        n, lim, R, ns, flat = 10000, 15.0, 500e3, 1001, False
        label = "cfg2: range-dependent 200x500 gridded SSP + seamount bathymetry, 10k rays +-15deg, 500 km, 1001 save pts"
    # weak scaling: an N x n-ray fan dealt to the ranks in cyclic blocks of 32 angles (pygenray_b200.dist)
    from pygenray_b200.dist import shard_indices

    all_angles = np.linspace(-lim, lim, n * world)
    mine = shard_indices(n * world, rank, world)
    return dict(arrays=a, flatearth=flat, source_depth=1000.0, source_range=0.0, receiver_range=R,
                angles=np.ascontiguousarray(all_angles[mine]), n_total=n * world, n_save=ns, label=label,
                name=name, rays_per_gpu=n)


def make_env(pr, a, flat):
    ssp = pr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
    bathy = pr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
    return pr.OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, flat_earth_transform=flat)


def algorithmic_flops(cnt):
    """W = 38 n_rhs + 200 n_attempts + 40 n_accepted + 168 n_dense + 60 n_bisect."""
    c = cnt.astype(np.float64).sum(axis=1)
    return 38 * c[0] + 200 * c[1] + 40 * c[2] + 168 * c[5] + 60 * c[4]


class ClockSampler(threading.Thread):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        smax = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def oracle_env(w):
    """Oracle-side environment from the same arrays through the same (host) unpacking as the product."""
    import pygenray_b200 as pr  # host-only pieces: environment object, _unpack_envi, spline coefficients
    from oracle import oracle as orc

    env = make_env(pr, w["arrays"], w["flatearth"])
    arrs = pr._unpack_envi(env, flatearth=w["flatearth"])
    ba_x, ba_c = pr.bottom_angle_ppoly(arrs[5], arrs[6])
    return orc, orc.OracleEnv(arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], arrs[5], ba_x, ba_c)


def time_oracle(w, angles, threads=0):
    orc, oenv = oracle_env(w)
    rs = np.linspace(w["source_range"], w["receiver_range"], w["n_save"]) if w["n_save"] else None
    t0 = time.perf_counter()
    out = orc.shoot_fan(oenv, w["source_depth"], w["source_range"], w["receiver_range"], angles_deg=angles,
                        range_save=rs, n_threads=threads, counters=False)
    dt = time.perf_counter() - t0
    km = float((out["status"] == 0).sum()) * abs(w["receiver_range"] - w["source_range"]) / 1e3
    return km / dt, dt, orc.num_threads() if threads <= 0 else threads


def run_reference(args):
    """--impl reference: the CPU port of the reference's path, all host threads, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.workload)
    angles = w["angles"] if args.ref_rays <= 0 else w["angles"][:: max(1, len(w["angles"]) // args.ref_rays)]
    for _ in range(args.warmup):
        time_oracle(w, angles[:: max(1, len(angles) // 64)])
    vals, dts = [], []
    cores = 0
    for _ in range(args.steps):
        v, dt, cores = time_oracle(w, angles)
        vals.append(v); dts.append(dt)
    value = float(np.sum([v * d for v, d in zip(vals, dts)]) / np.sum(dts))
    sample = f"{len(angles)} of {len(w['angles'])} rays of the workload per step, {cores} pthreads"
    line = {
        "impl": "reference", "metric": "ray_km_per_s", "value": value, "unit": "ray-km/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "rays_per_step": int(len(angles))},
        "cpu_baseline": {"value": value, "unit": "ray-km/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ray-km/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python (scipy solve_ivp + numba); it cannot travel to the GPU box, so this arm "
                "times the plain-C restatement oracle/prt_oracle.c (about 40x faster per core than the Python "
                "reference measured in the build container: 5.1-5.5 k ray-km/s/core, BASELINE.md)",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import pygenray_b200 as pr
    from pygenray_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        from pygenray_b200.dist import bind_to_gpu_numa
        if os.environ.get("PRT_NUMA_BIND", "1") != "0":
            numa = bind_to_gpu_numa(local_rank)  # before the first pinned allocation (first touch places it)
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.require_gpu()

    w = workload(args.workload, rank, world)
    env = make_env(pr, w["arrays"], w["flatearth"])
    arrs = pr._unpack_envi(env, flatearth=w["flatearth"])
    n_env = w.get("n_env", 0)
    staging = None
    if n_env:
        # the realisations are staged twice, once per route, and timed (untimed setup of the benchmark proper):
        # host route = np.gradient + interleave + one H2D copy of {c, dc/dz}; device route = the sound speeds are
        # already in HBM (as perturbations generated on the GPU would be), dc/dz formed by prt_stage_grid_kernel
        stack = w["arrays"]["c_stack"]
        t0 = time.perf_counter()
        cp_stack = np.gradient(stack, arrs[3], axis=2)
        host_env = pr.DeviceEnsemble(stack, cp_stack, *arrs[2:], device=local_rank)
        t_host = time.perf_counter() - t0
        host_env.close()
        stack_d = torch.from_numpy(stack).to(dev)
        for _ in range(2):  # first calls pay the lazy load of the staging kernel and a fresh cudaMalloc
            pr.DeviceEnsemble.from_device(stack_d, *arrs[2:]).close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        denv = pr.DeviceEnsemble.from_device(stack_d, *arrs[2:])
        torch.cuda.synchronize()
        staging = {"host_route_ms": 1e3 * t_host, "device_route_ms": 1e3 * (time.perf_counter() - t0),
                   "bytes": int(stack.nbytes) * 2}
        del stack_d
    else:
        denv = pr.DeviceEnvironment(*arrs, device=local_rank)
    angles = w["angles"]
    n = len(angles) * max(n_env, 1)
    R_km = abs(w["receiver_range"] - w["source_range"]) / 1e3
    n_save = w["n_save"]

    # ---- resident inputs / outputs for the device-timed leg
    import ctypes as C
    ang_d = torch.from_numpy(angles).to(dev)
    rs_d = torch.from_numpy(np.linspace(w["source_range"], w["receiver_range"], n_save)).to(dev) if n_save else None
    traj_d = torch.empty((3, n, n_save), dtype=torch.float64, device=dev) if n_save else None
    final_d = torch.empty((3, n), dtype=torch.float64, device=dev)
    nb_d = torch.empty(n, dtype=torch.int32, device=dev)
    ns_d = torch.empty(n, dtype=torch.int32, device=dev)
    st_d = torch.empty(n, dtype=torch.int32, device=dev)
    cnt_d = torch.empty((6, n), dtype=torch.int32, device=dev)
    flips_d = torch.empty(n, dtype=torch.int32, device=dev) if n_save else None
    from pygenray_b200 import dist as pdist
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def launch(flips):
        if n_env:
            return L.prt_shoot_ensemble(denv.handle, n_env, len(angles), p(ang_d), w["source_depth"], w["source_range"],
                                        w["receiver_range"], 1e-9, 1e-6, 1, n_save, p(rs_d), p(traj_d), p(final_d),
                                        p(nb_d), p(ns_d), p(st_d), p(cnt_d), 1, flips, 1 << 20,
                                        C.c_void_p(stream.cuda_stream))
        return L.prt_shoot_fan(denv.handle, n, p(ang_d), None, w["source_depth"], w["source_range"],
                               w["receiver_range"], 1e-9, 1e-6, 1, n_save, p(rs_d), p(traj_d), p(final_d), p(nb_d),
                               p(ns_d), p(st_d), p(cnt_d), 1, flips, 1 << 20, C.c_void_p(stream.cuda_stream))

    def device_step():
        _lib.check(launch(p(flips_d)), "prt_shoot_fan")
        if world > 1:  # the path's only exchange: all-gather the arrival records (6 doubles per ray)
            rec = pdist.records_from_result(dict(final=final_d, n_bottom=nb_d, n_surface=ns_d, status=st_d))
            if n_env:  # ensembles shard by realisation: plain concatenation
                dist.all_gather_into_tensor(torch.empty((world * n, 6), dtype=torch.float64, device=dev), rec)
            else:
                pdist.all_gather_arrivals(rec, w["n_total"], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    fp64_peak = L.prt_measure_fp64_peak(local_rank, 2000)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    step_ms, kern_ms = [], []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        device_step()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        step_ms.append(float(ms.item()))
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    ok = int((st_d == 0).sum().item())
    ok_t = torch.tensor([ok], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ok_t)
    total_km = float(ok_t.item()) * R_km
    ms_per_step = float(np.mean(step_ms))
    value = total_km / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (no gather), for the roofline
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        rc = launch(None)
        e1.record(stream)
        torch.cuda.synchronize()
        _lib.check(rc, "prt_shoot_fan")
        kern_ms.append(e0.elapsed_time(e1))
    cnt = cnt_d.cpu().numpy().astype(np.uint32)
    flops = algorithmic_flops(cnt)
    k_ms = float(np.mean(kern_ms))
    achieved = flops / (k_ms * 1e-3) / 1e12
    out_bytes = (3 * n * n_save + 3 * n) * 8 + 3 * n * 4 + 6 * n * 4

    # ---- end to end through the public API (host in, host out)
    e2e_dts = []
    fan = None
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        if n_save:
            fan = pr.shoot_rays(w["source_depth"], w["source_range"], -angles, w["receiver_range"], n_save, env,
                                rtol=1e-9, debug=False, flatearth=w["flatearth"])
            n_ok = len(fan)
        elif n_env:
            res = pr.shoot_ensemble_arrays(denv, w["source_depth"], w["source_range"], w["receiver_range"], angles)
            n_ok = int((res["status"] == 0).sum())
        else:
            res = pr.shoot_fan_arrays(denv, w["source_depth"], w["source_range"], w["receiver_range"], angles=angles)
            n_ok = int((res["status"] == 0).sum())
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        cnt_ok = torch.tensor([float(n_ok)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt_ok)
        if i >= args.warmup:
            e2e_dts.append((float(dt.item()), float(cnt_ok.item())))
    e2e_value = float(np.mean([c * R_km / d for d, c in e2e_dts]))
    h2d = n * 8 + n_save * 8
    d2h = out_bytes - 6 * n * 4 + (n * 4 if n_save else 0)

    # ---- eigenrays/s (BASELINE config 4): 64 receiver depths at the receiver range of this fan, ztol 1e-3,
    # brackets from this rank's fan; bracketing + regula-falsi refinement timed end to end on the device
    eig = None
    if n_save and not args.no_eigen:
        from pygenray_b200.eigenrays import find_brackets, refine_brackets
        rds = np.linspace(200.0, 4000.0, 64)
        zs_end = fan.zs[:, -1].copy()
        thetas = fan.thetas.copy()
        def eig_leg(ztol):
            dts = []
            for i in range(1 + 2):
                barrier()
                t0 = time.perf_counter()
                br_, counts_ = find_brackets(zs_end, thetas, rds, device=local_rank)
                ref_ = refine_brackets(denv, br_, w["source_depth"], w["source_range"], w["receiver_range"],
                                       ztol=ztol, max_iter=20)
                st_ = ref_["status"].cpu().numpy()
                barrier()
                if i >= 1:
                    dts.append(time.perf_counter() - t0)
            found_ = int((st_ == 0).sum())
            tot = torch.tensor([float(found_)], dtype=torch.float64, device=dev)
            dtm = torch.tensor([float(np.mean(dts))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tot)
                dist.all_reduce(dtm, op=dist.ReduceOp.MAX)
            return br_, counts_, ref_, found_, float(tot.item() / dtm.item()), 1e3 * float(dtm.item())

        br, counts, ref, found, rate, ms = eig_leg(1e-3)
        _, _, ref1, found1, rate1, ms1 = eig_leg(1.0)
        eig = {"value": rate, "unit": "eigenrays/s", "receivers": 64, "ztol_m": 1e-3,
               "brackets_rank0": int(counts.sum()), "found_rank0": found, "ms": ms,
               "shots_per_bracket": float(ref["shots"].float().mean().item()),
               "note": "on this 11 m grid the arrival depth carries ~1e-2 m of step-sequence noise (DESIGN.md "
                       "section 5), so most brackets cannot reach 1e-3 m and run to max_iter, as in the reference",
               "ztol_1m": {"value": rate1, "found_rank0": found1, "ms": ms1,
                           "shots_per_bracket": float(ref1["shots"].float().mean().item())}}
        if rank == 0:
            host = {k: v.cpu().numpy() for k, v in br.items()}
            sub = slice(0, min(len(host["z1"]), 2048))
            orc, oenv = oracle_env(w)
            t0 = time.perf_counter()
            o = orc.refine_brackets(oenv, host["z1"][sub], host["z2"][sub], host["th1"][sub], host["th2"][sub],
                                    host["rd"][sub], w["source_depth"], w["source_range"], w["receiver_range"],
                                    ztol=1e-3, max_iter=20)
            dtc = time.perf_counter() - t0
            eig["cpu_baseline"] = {"value": float((o["status"] == 0).sum() / dtc), "unit": "eigenrays/s",
                                   "cores": orc.num_threads(), "kind": "port",
                                   "sample": f"first {sub.stop} brackets, {dtc:.1f} s"}
            eig["parity_vs_cpu"] = bool(np.array_equal(o["theta"], ref["theta"].cpu().numpy()[sub], equal_nan=True))

    clocks = sampler.stop() if rank == 0 else None  # sampled over the device-timed, kernel-only and e2e legs
    if rank == 0:
        # CPU oracle on the same workload (bounded: the whole 10k fan takes seconds in C)
        sub = angles if len(angles) <= 20000 else angles[:: len(angles) // 4096]
        cpu_v, cpu_dt, cores = time_oracle(w, sub)
        line = {
            "metric": "ray_km_per_s", "value": value, "unit": "ray-km/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["label"], "rays_per_gpu": n, "range_km": R_km, "n_save": n_save, "rtol": 1e-9,
                       "atol": 1e-6, "l2": "256 MiB buffer zeroed between timed steps (untimed)",
                       **({"ensemble_staging": staging} if staging else {}),
                       **({"numa_node_rank0": numa} if world > 1 else {}),
                       "parallelism": f"launch angles interleaved over {world} rank(s); arrival tables all-gathered"
                       if world > 1 else "single GPU, one thread per ray"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / (fp64_peak / 1e12) if fp64_peak > 0 else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from the
                         # ncu --set full capture summarised in profiles/r01_ncu_fan.md (11.8 MB + 200.3 MB)
                         "traffic": 212.0e6 if (w["name"] == "cfg2" and world == 1) else None,
                         "traffic_algorithmic_bytes": out_bytes,
                         "kernel": "prt_fan_kernel", "kernel_ms": k_ms, "algorithmic_flops_per_launch": flops,
                         "peak_source": "DFMA microbenchmark (prt_measure_fp64_peak) in this run; "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "hbm_output_GBps": out_bytes / (k_ms * 1e-3) / 1e9},
            "cpu_baseline": {"value": cpu_v, "unit": "ray-km/s", "cores": cores, "kind": "port",
                             "sample": f"{len(sub)} rays of the same fan, {cpu_dt:.1f} s, oracle/prt_oracle.c"},
            "e2e": {"value": e2e_value, "unit": "ray-km/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "api": "pygenray_b200.shoot_rays" if n_save else
                           ("pygenray_b200.shoot_ensemble_arrays" if n_env else "pygenray_b200.shoot_fan_arrays")},
            "gpu_launches": int(launches), "clocks": clocks,
            "work": {"rays_ok": int(ok_t.item()), "nfev": int(cnt[0].sum()), "attempts": int(cnt[1].sum()),
                     "accepted": int(cnt[2].sum()), "segments": int(cnt[3].sum()), "bisect": int(cnt[4].sum()),
                     "dense": int(cnt[5].sum())},
            "wall_s_timed_region": t_wall,
            "eigenrays": eig,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg5"])
    ap.add_argument("--no-eigen", action="store_true", help="skip the eigenrays/s leg")
    ap.add_argument("--ref-rays", type=int, default=0, help="reference arm: sub-sample the fan to about this many rays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
