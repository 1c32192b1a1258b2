// prt_env_layout.h -- host-side packing of one environment into the single buffer the kernels read.
//
// Input = the seven arrays of the reference's _unpack_envi (/root/reference/src/pygenray/
// launch_rays.py:717-742) + the bottom-angle cubic spline as PPoly coefficients.  Output = one
// contiguous, 256-byte-aligned blob laid out for 16-byte loads:
//   grid   [n_r][n_z] {c, dc/dz}   -- both fields of a node in one LDG.128
//   rpair  [n_r-1] {r[i], r[i+1]}  -- both edges of a cell in one LDG.128 (same for z, bathymetry)
//   rrecip [n_r-1] 1/(r[i+1]-r[i]) -- interpolation weights without a division (same for z, bathymetry)
// and a prt_env_view whose pointers are offsets into the blob until prt_env_rebase() is applied.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "prt_core.h"

struct prt_env_blob {
    std::vector<unsigned char> bytes;  // the layout from byte `grid_bytes` on (everything when grid_bytes == 0)
    prt_env_view view;                 // pointers hold byte OFFSETS into the full layout
    size_t total_bytes = 0, grid_bytes = 0;
};

// np.gradient(f, z, axis=depth, edge_order=1) as coefficient tables (numpy/lib/_function_base_impl.py gradient):
// uniform spacing (all diff(z) equal): interior (f[j+1] - f[j-1]) / (2 dz); otherwise a f[j-1] + b f[j] + c f[j+1];
// one-sided first differences at the two ends.  Same operations in the same order as numpy, hence the same bits.
struct prt_gradient_tables {
    int uniform;
    double two_dz, dz_first, dz_last;
    std::vector<double> a, b, c;  // [n_z], interior entries only
};

static inline prt_gradient_tables prt_gradient_coefficients(const double* z, int n)
{
    prt_gradient_tables g;
    std::vector<double> d((size_t)n - 1);
    for (int j = 0; j < n - 1; ++j) d[j] = z[j + 1] - z[j];
    g.uniform = 1;
    for (int j = 1; j < n - 1; ++j)
        if (d[j] != d[0]) g.uniform = 0;
    g.two_dz = 2.0 * d[0];
    g.dz_first = d[0];
    g.dz_last = g.uniform ? d[0] : d[n - 2];
    g.a.assign(n, 0.0); g.b.assign(n, 0.0); g.c.assign(n, 0.0);
    for (int j = 1; j < n - 1; ++j) {
        const double dx1 = d[j - 1], dx2 = d[j];
        g.a[j] = -(dx2) / (dx1 * (dx1 + dx2));
        g.b[j] = (dx2 - dx1) / (dx1 * dx2);
        g.c[j] = dx1 / (dx2 * (dx1 + dx2));
    }
    return g;
}

static inline size_t prt_align256(size_t n) { return (n + 255) & ~(size_t)255; }

static inline void prt_pack_pairs(prt_d2* dst, const double* a, int n)
{
    for (int i = 0; i < n - 1; ++i) {
        dst[i].x = a[i];
        dst[i].y = a[i + 1];
    }
}

static inline void prt_pack_recip(double* dst, const double* a, int n)
{
    for (int i = 0; i < n - 1; ++i) dst[i] = 1.0 / (a[i + 1] - a[i]);
}

// {lo, hi}, {a[i], 1/(a[i+1]-a[i])} per cell; lo = -inf for the first cell, hi = +inf for the last
static inline void prt_pack_cells(prt_d2* dst, const double* a, int n)
{
    for (int i = 0; i < n - 1; ++i) {
        dst[2 * i].x = (i == 0) ? -INFINITY : a[i];
        dst[2 * i].y = (i == n - 2) ? INFINITY : a[i + 1];
        dst[2 * i + 1].x = a[i];
        dst[2 * i + 1].y = 1.0 / (a[i + 1] - a[i]);
    }
}

// {a[i], 1/(a[i+1]-a[i])} per cell
static inline void prt_pack_xr(prt_d2* dst, const double* a, int n)
{
    for (int i = 0; i < n - 1; ++i) {
        dst[i].x = a[i];
        dst[i].y = 1.0 / (a[i + 1] - a[i]);
    }
}

// bathymetry cells: {lo, hi} padded like prt_pack_cells, {x[i], 1/(x[i+1]-x[i])}, {depth[i], depth[i+1]}
static inline void prt_pack_bcells(prt_d2* dst, const double* x, const double* d, int n)
{
    for (int i = 0; i < n - 1; ++i) {
        dst[3 * i].x = (i == 0) ? -INFINITY : x[i];
        dst[3 * i].y = (i == n - 2) ? INFINITY : x[i + 1];
        dst[3 * i + 1].x = x[i];
        dst[3 * i + 1].y = 1.0 / (x[i + 1] - x[i]);
        dst[3 * i + 2].x = d[i];
        dst[3 * i + 2].y = d[i + 1];
    }
}

static inline double prt_inv_mean_step(const double* a, int n)
{
    double span = a[n - 1] - a[0];
    if (!(span > 0.0)) return 0.0;
    return (double)(n - 1) / span;
}

// Quality of the cell guess trunc(g(v)) over one axis: the largest distance, in cells, between g and the true
// real-valued index at the nodes and cell centres.  A lookup misses when v lies within that distance of a node.
template <class G>
static inline double prt_guess_error(const double* a, int n, G g)
{
    double worst = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        const double e0 = fabs(g(a[i]) - (double)i);
        const double e1 = fabs(g(0.5 * (a[i] + a[i + 1])) - ((double)i + 0.5));
        const double e2 = fabs(g(a[i + 1]) - (double)(i + 1));
        worst = fmax(worst, fmax(e0, fmax(e1, e2)));
    }
    return worst;
}

// least-squares q(u) = b1 + b2 u + b3 u^2 such that s (1 + q(s / n)) ~ index, s = (a[j] - a[0]) * inv
static inline void prt_fit_guess(const double* a, int n, double inv, double un, double b[3])
{
    long double M[3][4] = {{0}};
    for (int j = 1; j < n; ++j) {
        const long double s = (long double)(a[j] - a[0]) * inv, u = s * un;
        const long double phi[3] = {s, s * u, s * u * u};
        const long double y = (long double)j - s;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) M[r][c] += phi[r] * phi[c];
            M[r][3] += phi[r] * y;
        }
    }
    for (int c = 0; c < 3; ++c) {  // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < 3; ++r)
            if (fabsl(M[r][c]) > fabsl(M[piv][c])) piv = r;
        for (int k = 0; k < 4; ++k) { long double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
        if (M[c][c] == 0.0L) { b[0] = b[1] = b[2] = 0.0; return; }
        for (int r = 0; r < 3; ++r) {
            if (r == c) continue;
            const long double f = M[r][c] / M[c][c];
            for (int k = c; k < 4; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int r = 0; r < 3; ++r) b[r] = (double)(M[r][3] / M[r][r]);
}

// returns 0 on success, nonzero on invalid sizes
// n_env >= 1 realisations of the sound-speed field (cin/cpin: [n_env][n_r][n_z]) share one set of
// coordinate, bathymetry and spline tables (tomography ensembles, BASELINE config 5)
static inline int prt_env_pack(int n_r, int n_z, const double* rin, const double* zin, const double* cin,
                               const double* cpin, int n_b, const double* depth_ranges, const double* depths,
                               int n_ba, const double* ba_x, const double* ba_c, prt_env_blob& out, int n_env = 1)
{
    // cin == cpin == nullptr: tables only -- the grid (first region of the layout) is filled on the device by
    // prt_stage_grid_kernel; out.bytes then holds the layout from the end of the grid region on (out.grid_bytes)
    const bool host_grid = (cin != nullptr);
    if (n_r < 2 || n_z < 2 || n_b < 2 || n_ba < 2 || n_env < 1) return 1;
    size_t o_grid = 0;
    size_t o_rpair = o_grid + prt_align256(sizeof(prt_d2) * (size_t)n_r * n_z * (size_t)n_env);
    size_t o_zpair = o_rpair + prt_align256(sizeof(prt_d2) * (size_t)(n_r - 1));
    size_t o_bx = o_zpair + prt_align256(sizeof(prt_d2) * (size_t)(n_z - 1));
    size_t o_bv = o_bx + prt_align256(sizeof(prt_d2) * (size_t)(n_b - 1));
    size_t o_rrec = o_bv + prt_align256(sizeof(prt_d2) * (size_t)(n_b - 1));
    size_t o_zrec = o_rrec + prt_align256(sizeof(double) * (size_t)(n_r - 1));
    size_t o_brec = o_zrec + prt_align256(sizeof(double) * (size_t)(n_z - 1));
    size_t o_rcell = o_brec + prt_align256(sizeof(double) * (size_t)(n_b - 1));
    size_t o_zcell = o_rcell + prt_align256(sizeof(prt_d2) * 2 * (size_t)(n_r - 1));
    size_t o_bcell = o_zcell + prt_align256(sizeof(prt_d2) * 2 * (size_t)(n_z - 1));
    size_t o_rxr = o_bcell + prt_align256(sizeof(prt_d2) * 3 * (size_t)(n_b - 1));
    size_t o_zxr = o_rxr + prt_align256(sizeof(prt_d2) * (size_t)(n_r - 1));
    size_t o_bax = o_zxr + prt_align256(sizeof(prt_d2) * (size_t)(n_z - 1));
    size_t o_bac = o_bax + prt_align256(sizeof(double) * (size_t)n_ba);
    size_t total = o_bac + prt_align256(sizeof(double) * 4 * (size_t)(n_ba - 1));
    out.total_bytes = total;
    out.grid_bytes = host_grid ? 0 : o_rpair;
    out.bytes.assign(total - out.grid_bytes, 0);
    unsigned char* base = out.bytes.data() - out.grid_bytes;  // offsets below stay those of the full layout
    if (host_grid) {
        prt_d2* grid = reinterpret_cast<prt_d2*>(base + o_grid);
        const size_t n_nodes = (size_t)n_r * n_z * (size_t)n_env;
        for (size_t k = 0; k < n_nodes; ++k) {
            grid[k].x = cin[k];
            grid[k].y = cpin[k];
        }
    }
    prt_pack_pairs(reinterpret_cast<prt_d2*>(base + o_rpair), rin, n_r);
    prt_pack_pairs(reinterpret_cast<prt_d2*>(base + o_zpair), zin, n_z);
    prt_pack_pairs(reinterpret_cast<prt_d2*>(base + o_bx), depth_ranges, n_b);
    prt_pack_pairs(reinterpret_cast<prt_d2*>(base + o_bv), depths, n_b);
    prt_pack_cells(reinterpret_cast<prt_d2*>(base + o_rcell), rin, n_r);
    prt_pack_cells(reinterpret_cast<prt_d2*>(base + o_zcell), zin, n_z);
    prt_pack_bcells(reinterpret_cast<prt_d2*>(base + o_bcell), depth_ranges, depths, n_b);
    prt_pack_xr(reinterpret_cast<prt_d2*>(base + o_rxr), rin, n_r);
    prt_pack_xr(reinterpret_cast<prt_d2*>(base + o_zxr), zin, n_z);
    prt_pack_recip(reinterpret_cast<double*>(base + o_rrec), rin, n_r);
    prt_pack_recip(reinterpret_cast<double*>(base + o_zrec), zin, n_z);
    prt_pack_recip(reinterpret_cast<double*>(base + o_brec), depth_ranges, n_b);
    memcpy(base + o_bax, ba_x, sizeof(double) * n_ba);
    memcpy(base + o_bac, ba_c, sizeof(double) * 4 * (size_t)(n_ba - 1));

    prt_env_view& v = out.view;
    v.n_r = n_r; v.n_z = n_z; v.n_b = n_b; v.n_ba = n_ba;
    v.n_env = n_env;
    v.grid_stride = (n_env > 1) ? (int64_t)n_r * n_z : 0;
    v.grid = reinterpret_cast<const prt_d2*>(o_grid);
    v.rpair = reinterpret_cast<const prt_d2*>(o_rpair);
    v.zpair = reinterpret_cast<const prt_d2*>(o_zpair);
    v.bxpair = reinterpret_cast<const prt_d2*>(o_bx);
    v.bvpair = reinterpret_cast<const prt_d2*>(o_bv);
    v.rcell = reinterpret_cast<const prt_d2*>(o_rcell);
    v.zcell = reinterpret_cast<const prt_d2*>(o_zcell);
    v.bcell = reinterpret_cast<const prt_d2*>(o_bcell);
    v.rxr = reinterpret_cast<const prt_d2*>(o_rxr);
    v.zxr = reinterpret_cast<const prt_d2*>(o_zxr);
    v.rrecip = reinterpret_cast<const double*>(o_rrec);
    v.zrecip = reinterpret_cast<const double*>(o_zrec);
    v.brecip = reinterpret_cast<const double*>(o_brec);
    v.ba_x = reinterpret_cast<const double*>(o_bax);
    v.ba_c = reinterpret_cast<const double*>(o_bac);
    v.r0 = rin[0]; v.r_inv = prt_inv_mean_step(rin, n_r);
    v.z0 = zin[0]; v.z_inv = prt_inv_mean_step(zin, n_z);
    v.b0 = depth_ranges[0]; v.b_inv = prt_inv_mean_step(depth_ranges, n_b);
    v.r_off = -v.r0 * v.r_inv; v.z_off = -v.z0 * v.z_inv; v.b_off = -v.b0 * v.b_inv;
    // How good are the guesses?  (They only decide speed: a trial step whose guess missed is redone exactly.)
    // Linear guess within 1e-6 cells: nothing to do.  Smooth depth stretching (flat-earth): refine with the
    // fitted quadratic.  Anything worse: the speculative trial step would mostly be wasted, switch it off.
    const double tol_ok = 1e-4;
    v.z_cubic = 0; v.z_b1 = v.z_b2 = v.z_b3 = 0.0; v.z_un = 1.0 / (double)n_z;
    v.spec_off = 0;
    const double r_inv = v.r_inv, r_off = v.r_off, z_inv = v.z_inv, z_off = v.z_off, b_inv = v.b_inv, b_off = v.b_off;
    const double err_r = prt_guess_error(rin, n_r, [=](double x) { return fma(x, r_inv, r_off); });
    const double err_b = prt_guess_error(depth_ranges, n_b, [=](double x) { return fma(x, b_inv, b_off); });
    double err_z = prt_guess_error(zin, n_z, [=](double x) { return fma(x, z_inv, z_off); });
    if (err_z > 1e-6) {
        double b[3];
        const double un = v.z_un;
        prt_fit_guess(zin, n_z, z_inv, un, b);
        const double err_c = prt_guess_error(zin, n_z, [=](double x) {
            const double s = fma(x, z_inv, z_off), u = s * un;
            return fma(s, fma(fma(b[2], u, b[1]), u, b[0]), s);
        });
        if (err_c < err_z) {
            v.z_cubic = 1; v.z_b1 = b[0]; v.z_b2 = b[1]; v.z_b3 = b[2];
            err_z = err_c;
        }
    }
    if (err_r > tol_ok || err_b > tol_ok || err_z > tol_ok) v.spec_off = 1;
    // exactly uniform depth grid?  (every lower edge reproduced by one fma, every reciprocal width identical)
    v.z_uniform = 0;
    v.z_dz = zin[1] - zin[0];
    v.z_r0 = 1.0 / (zin[1] - zin[0]);
    if (!v.z_cubic && !v.spec_off) {
        bool uniform = true;
        for (int j = 0; j < n_z - 1 && uniform; ++j)
            uniform = (fma((double)j, v.z_dz, v.z0) == zin[j]) && (1.0 / (zin[j + 1] - zin[j]) == v.z_r0);
        v.z_uniform = uniform ? 1 : 0;
    }
    v.r_uniform = 0;
    v.r_dx = rin[1] - rin[0];
    v.r_r0 = 1.0 / (rin[1] - rin[0]);
    if (v.z_uniform) {
        bool uniform = true;
        for (int i = 0; i < n_r - 1 && uniform; ++i)
            uniform = (fma((double)i, v.r_dx, v.r0) == rin[i]) && (1.0 / (rin[i + 1] - rin[i]) == v.r_r0);
        v.r_uniform = uniform ? 1 : 0;
    }
    // ray_bounding_box_event tolerances (integration_processes.py:280-303)
    const double tol = 1e-6;
    v.bb_zmax = zin[n_z - 1] + tol;
    v.bb_zmin = zin[0] - tol;
    v.bb_rmin = rin[0] - tol;
    v.bb_rmax = rin[n_r - 1] + tol;
    return 0;
}

// turn the offsets of `view` into pointers relative to `base` (host or device address)
static inline prt_env_view prt_env_rebase(const prt_env_view& view, const void* base)
{
    prt_env_view v = view;
    const unsigned char* b = static_cast<const unsigned char*>(base);
    v.grid = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.grid));
    v.rpair = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.rpair));
    v.zpair = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.zpair));
    v.bxpair = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.bxpair));
    v.bvpair = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.bvpair));
    v.rcell = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.rcell));
    v.zcell = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.zcell));
    v.bcell = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.bcell));
    v.rxr = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.rxr));
    v.zxr = reinterpret_cast<const prt_d2*>(b + reinterpret_cast<size_t>(view.zxr));
    v.rrecip = reinterpret_cast<const double*>(b + reinterpret_cast<size_t>(view.rrecip));
    v.zrecip = reinterpret_cast<const double*>(b + reinterpret_cast<size_t>(view.zrecip));
    v.brecip = reinterpret_cast<const double*>(b + reinterpret_cast<size_t>(view.brecip));
    v.ba_x = reinterpret_cast<const double*>(b + reinterpret_cast<size_t>(view.ba_x));
    v.ba_c = reinterpret_cast<const double*>(b + reinterpret_cast<size_t>(view.ba_c));
    return v;
}
