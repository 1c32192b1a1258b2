// core_host.cpp -- TEST HARNESS: compiles the product's per-ray device routine (prt_core.h) for
// the host with g++ so that `pytest -m "not gpu"` can compare it bit-for-bit with the CPU oracle
// on a machine without a GPU.  Not part of the product: libprt.so never contains or calls this.
#include <stdint.h>

#include "../../pygenray_b200/csrc/prt_core.h"
#include "../../pygenray_b200/csrc/prt_env_layout.h"

extern "C" {

struct hh_env {
    prt_env_blob blob;
    prt_env_view view;
};

__attribute__((visibility("default"))) void* hh_env_create(int n_r, int n_z, const double* rin, const double* zin,
                                                            const double* cin, const double* cpin, int n_b,
                                                            const double* depth_ranges, const double* depths,
                                                            int n_ba, const double* ba_x, const double* ba_c)
{
    hh_env* e = new hh_env;
    if (prt_env_pack(n_r, n_z, rin, zin, cin, cpin, n_b, depth_ranges, depths, n_ba, ba_x, ba_c, e->blob)) {
        delete e;
        return nullptr;
    }
    e->view = prt_env_rebase(e->blob.view, e->blob.bytes.data());
    return e;
}

__attribute__((visibility("default"))) void hh_env_destroy(void* e) { delete static_cast<hh_env*>(e); }

// np.gradient(f, z, edge_order=1) through the coefficient tables the device staging kernel uses (same operations)
__attribute__((visibility("default"))) void hh_gradient(const double* z, int n, const double* f, double* out)
{
    const prt_gradient_tables g = prt_gradient_coefficients(z, n);
    for (int j = 0; j < n; ++j) {
        if (j == 0) out[j] = (f[1] - f[0]) / g.dz_first;
        else if (j == n - 1) out[j] = (f[j] - f[j - 1]) / g.dz_last;
        else if (g.uniform) out[j] = (f[j + 1] - f[j - 1]) / g.two_dz;
        else out[j] = (g.a[j] * f[j - 1] + g.b[j] * f[j]) + g.c[j] * f[j + 1];
    }
}

// staging-time decisions about the cell guess: out = {z_cubic, spec_off, z_uniform}
__attribute__((visibility("default"))) void hh_env_flags(void* e, int* out)
{
    const prt_env_view& V = static_cast<hh_env*>(e)->view;
    out[0] = V.z_cubic;
    out[1] = V.spec_off;
    out[2] = V.z_uniform + 2 * V.r_uniform;
}

// same output layout as prt_shoot_fan: traj [3][n_rays][n_save], final [3][n_rays], counters [6][n_rays]
__attribute__((visibility("default"))) int hh_shoot_fan(void* env, int64_t n_rays, const double* angles_deg,
                                                         const double* y0s, double source_depth, double source_range,
                                                         double receiver_range, double rtol, double atol,
                                                         int terminate_backwards, int n_save, const double* range_save,
                                                         double* out_traj, double* out_final, int32_t* out_nb,
                                                         int32_t* out_ns, int32_t* out_status, uint32_t* out_counters,
                                                         int max_bounces)
{
    const prt_env_view& V = static_cast<hh_env*>(env)->view;
    const prt_envp E{V, V.grid, prt_z_mode(V)};
    for (int64_t i = 0; i < n_rays; ++i) {
        double T, z, p;
        if (y0s) {
            T = y0s[3 * i]; z = y0s[3 * i + 1]; p = y0s[3 * i + 2];
        } else {
            prt_initial_state(E, source_depth, source_range, angles_deg[i], T, z, p);
        }
        prt_traj_sink sink{nullptr, nullptr, nullptr, 1, 1.0};
        prt_ray_result res;
        uint32_t cnt[PRT_CNT_N];
        prt_shot S;
        S.x_end = receiver_range; S.rtol = rtol; S.atol = atol; S.terminate_backwards = terminate_backwards;
        S.n_save = 0; S.max_bounces = max_bounces; S.range_save = nullptr;
        if (out_traj && n_save > 0) {
            sink.T = out_traj + ((size_t)0 * n_rays + i) * n_save;
            sink.z = out_traj + ((size_t)1 * n_rays + i) * n_save;
            sink.p = out_traj + ((size_t)2 * n_rays + i) * n_save;
            for (int k = 0; k < n_save; ++k) sink.T[k] = sink.z[k] = sink.p[k] = NAN;
            S.n_save = n_save; S.range_save = range_save;
            prt_integrate_ray<true>(E, T, z, p, source_range, S, sink, res, cnt);
        } else {
            prt_integrate_ray<false>(E, T, z, p, source_range, S, sink, res, cnt);
        }
        if (out_traj && n_save > 0 && res.status != PRT_RAY_OK)  // as prt_fan_kernel does for dropped rays
            for (int k = 0; k < n_save; ++k) sink.T[k] = sink.z[k] = sink.p[k] = NAN;
        out_final[0 * n_rays + i] = res.T;
        out_final[1 * n_rays + i] = res.z;
        out_final[2 * n_rays + i] = res.p;
        out_nb[i] = res.n_bottom; out_ns[i] = res.n_surface; out_status[i] = res.status;
        if (out_counters)
            for (int k = 0; k < PRT_CNT_N; ++k) out_counters[(size_t)k * n_rays + i] = cnt[k];
    }
    return 0;
}
}
