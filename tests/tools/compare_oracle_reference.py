"""Regenerate the numbers of DESIGN.md section 5 (oracle vs reference vs the reference's own jit / no-jit spread) from the
committed golden vectors: three-way distances of the arrival depth / time on configs 2 and 3, bounce-count detail of the
chaotic rays, eigenray comparison on config 4.  Run on the CPU: python tests/tools/compare_oracle_reference.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import golden, make_env, oracle_env_from
import workloads as cases
from oracle import oracle as orc

np.set_printoptions(precision=3, linewidth=160)


def three_way(name, env, zs, R):
    g, gn = golden(name + "_jit.npz"), golden(name + "_nojit.npz")
    oenv = oracle_env_from(env, False)
    o = orc.shoot_fan(oenv, zs, 0.0, R, angles_deg=g["al_internal_angles"], range_save=g["al_range_save"])
    nb_g, nb_n = g["al_n_bottom"] + g["al_n_surface"], gn["al_n_bottom"] + gn["al_n_surface"]
    ok = g["al_ok"] & gn["al_ok"] & (o["status"] == 0)
    same = ok & (g["al_n_bottom"] == gn["al_n_bottom"]) & (g["al_n_surface"] == gn["al_n_surface"])
    free = same & (nb_g == 0)
    print(f"== {name}: {len(ok)} rays, kept by all three {ok.sum()}, variants agree on bounces {same.sum()}, bounce-free {free.sum()}")
    for rows, label in ((same, "rays on which the variants agree"), (free, "bounce-free rays")):
        for comp, nm in ((1, "depth [m]"), (0, "time [s]")):
            for a, b, lab in ((o["final"], g["al_final"], "oracle-jit  "), (o["final"], gn["al_final"], "oracle-nojit"),
                              (g["al_final"], gn["al_final"], "jit-nojit   ")):
                d = np.abs(a - b)[rows, comp]
                print(f"   {label:34s} {nm:9s} {lab} median {np.median(d):9.3g}  p90 {np.percentile(d, 90):9.3g}  max {d.max():9.3g}")
    print("   RHS evaluations oracle / reference:", o["counters"][:, 0].sum() / g["al_nfev"].sum())
    for k in range(len(ok)):
        b = (int(g["al_n_bottom"][k]), int(g["al_n_surface"][k])); bn = (int(gn["al_n_bottom"][k]), int(gn["al_n_surface"][k]))
        bo = (int(o["n_bottom"][k]), int(o["n_surface"][k]))
        if b != bo or bn != bo or not g["al_ok"][k] or not gn["al_ok"][k]:
            print(f"   ray {k:2d} {g['al_internal_angles'][k]:8.3f} deg: jit {b} kept {bool(g['al_ok'][k])} | nojit {bn} kept {bool(gn['al_ok'][k])} | oracle {bo} status {o['status'][k]}")


def config4():
    g, gn = golden("ref_cfg4_sub_jit.npz"), golden("ref_cfg4_sub_nojit.npz")
    oenv = oracle_env_from(make_env(cases.config3_arrays()), False)
    user = g["user_angles"]
    o = orc.shoot_fan(oenv, 1000.0, 0.0, 1e6, angles_deg=-user)
    z_o, z_j, z_n = -o["final"][:, 1], g["fan_zs"][:, -1], gn["fan_zs"][:, -1]
    print("== config 4 fan: arrival depth max |oracle-jit|, |oracle-nojit|, |jit-nojit| [m]:", np.abs(z_o - z_j).max(), np.abs(z_o - z_n).max(), np.abs(z_j - z_n).max())
    print("   reference at ztol 1e-3 raised:", str(g["z3_raised"]), str(gn["z3_raised"]))
    for tag, ztol in (("z1", 1.0), ("z05", 0.05), ("z3", 1e-3)):
        for i, rd in enumerate(g["receiver_depths"]):
            k = np.where(np.diff(np.sign(z_o + rd)))[0]
            ref = orc.refine_brackets(oenv, z_o[k], z_o[k + 1], user[k], user[k + 1], np.full(len(k), rd), 1000.0, 0.0, 1e6, ztol=ztol, max_iter=20)
            conv = ref["status"] == 0
            line = f"   ztol {ztol:g} rd {rd:6.0f}: brackets {k}, oracle found {conv.sum()} (shots {ref['shots']})"
            if tag != "z3":
                pre = f"{tag}_e{i}_"
                line += f" | reference found jit {int(g[pre + 'found'])} nojit {int(gn[pre + 'found'])}"
                if conv.sum() == int(g[pre + "found"]) == int(gn[pre + "found"]):
                    line += (f" | dtheta oracle-jit {np.abs(np.radians(ref['theta'][conv] - g[pre + 'launch'])).max():.2e} rad,"
                             f" jit-nojit {np.abs(np.radians(gn[pre + 'launch'] - g[pre + 'launch'])).max():.2e} rad")
            print(line)


def _three_lines(label, O, G, N):
    for comp, nm in ((1, "depth [m]"), (0, "time [s]")):
        for a, b, lab in ((O, G, "oracle-jit  "), (O, N, "oracle-nojit"), (G, N, "jit-nojit   ")):
            d = np.abs(a - b)[:, comp]
            print(f"   {label:34s} {nm:9s} {lab} median {np.median(d):9.3g}  p90 {np.percentile(d, 90):9.3g}  "
                  f"p99 {np.percentile(d, 99):9.3g}  max {d.max():9.3g}")


def config5():
    g, gn = golden("ref_cfg5_sub_jit.npz"), golden("ref_cfg5_sub_nojit.npz")
    O, G, N = [], [], []
    for i in (0, 7):
        oenv = oracle_env_from(make_env(cases.config5_realisation(i)), False)
        P = f"m{i}_al_"
        o = orc.shoot_fan(oenv, 1000.0, 0.0, 500e3, angles_deg=g[P + "internal_angles"], range_save=g[P + "range_save"])
        m = g[P + "ok"] & gn[P + "ok"] & (o["status"] == 0)
        m &= (g[P + "n_bottom"] == gn[P + "n_bottom"]) & (g[P + "n_surface"] == gn[P + "n_surface"])
        print(f"== config 5 member {i}: {len(m)} rays, kept by all three with the variants agreeing on bounces {m.sum()}, oracle counts equal "
              f"{np.array_equal(o['n_bottom'][m], g[P + 'n_bottom'][m]) and np.array_equal(o['n_surface'][m], g[P + 'n_surface'][m])}, "
              f"RHS evaluations oracle / reference {o['counters'][:, 0].sum() / g[P + 'nfev'].sum():.4f}")
        O.append(o["final"][m]); G.append(g[P + "final"][m]); N.append(gn[P + "final"][m])
    _three_lines("config 5, members 0 and 7", np.concatenate(O), np.concatenate(G), np.concatenate(N))


def fuzz_short():
    from fuzz_cases import short_case

    g, gn = golden("ref_fuzz_short_jit.npz"), golden("ref_fuzz_short_nojit.npz")
    O, G, N, n, kept_equal, counts_equal = [], [], [], 0, 0, 0
    for seed in range(40):
        k = short_case(seed)
        oenv = oracle_env_from(make_env(k["arrays"], flat=k["flat"], lat=k["lat"]), k["flat"])
        P = f"s{seed}_"
        o = orc.shoot_fan(oenv, k["zs"], k["r0"], k["r1"], angles_deg=k["angles"], range_save=g[P + "range_save"], rtol=k["rtol"],
                          terminate_backwards=k["terminate_backwards"])
        m = g[P + "ok"]
        n += len(m); kept_equal += int(((o["status"] == 0) == m).sum())
        counts_equal += int(((o["n_bottom"] == g[P + "n_bottom"]) & (o["n_surface"] == g[P + "n_surface"]))[m].sum())
        O.append(o["final"][m]); G.append(g[P + "final"][m]); N.append(gn[P + "final"][m])
    O, G, N = np.concatenate(O), np.concatenate(G), np.concatenate(N)
    print(f"== 40 random short-range environments: {n} rays, kept / dropped equal on {kept_equal}, {len(O)} kept, bounce counts equal on {counts_equal}")
    _three_lines("random short-range environments", O, G, N)


def fuzz_smooth():
    from fuzz_cases import smooth_case

    g, gn = golden("ref_fuzz_smooth_jit.npz"), golden("ref_fuzz_smooth_nojit.npz")
    rows = {}
    n = counts_equal = nfev_equal = 0
    for seed in range(24):
        k = smooth_case(seed)
        oenv = oracle_env_from(make_env(k["arrays"], flat=k["flat"], lat=k["lat"]), k["flat"])
        P = f"s{seed}_"
        o = orc.shoot_fan(oenv, k["zs"], k["r0"], k["r1"], angles_deg=k["angles"], range_save=g[P + "range_save"], rtol=k["rtol"],
                          terminate_backwards=True)
        nb = g[P + "n_bottom"] + g[P + "n_surface"]
        n += len(nb)
        counts_equal += int(((o["n_bottom"] == g[P + "n_bottom"]) & (o["n_surface"] == g[P + "n_surface"])).sum())
        nfev_equal += int((o["counters"][:, 0] == g[P + "nfev"])[nb == 0].sum())
        for i, b in enumerate(nb):
            rows.setdefault(min(int(b), 6), []).append((o["final"][i], g[P + "final"][i], gn[P + "final"][i]))
    print(f"== 24 random smooth environments: {n} rays, all kept, bounce counts equal on {counts_equal}, RHS evaluations equal on "
          f"{nfev_equal} of {len(rows[0])} bounce-free rays")
    for b in sorted(rows):
        O, G, N = (np.array([r[i] for r in rows[b]]) for i in range(3))
        _three_lines(f"{b}{'+' if b == 6 else ''} reflections ({len(O)} rays)", O, G, N)


if __name__ == "__main__":
    three_way("ref_cfg2_sub", make_env(cases.config2_arrays()), 1000.0, 500e3)
    three_way("ref_cfg3_sub", make_env(cases.config3_arrays()), 1000.0, 1000e3)
    config4()
    config5()
    fuzz_short()
    fuzz_smooth()
