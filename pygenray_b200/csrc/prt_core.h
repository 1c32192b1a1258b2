// prt_core.h -- per-ray integrator of the B200 ray-fan path (device code; header-only).
//
// One call of prt_integrate_ray() == the reference's _shoot_ray_array + _interpolate_ray for one
// launch angle (/root/reference/src/pygenray/launch_rays.py:325-484, 745-784), i.e. a chain of
// scipy.integrate.solve_ivp(RK45, rtol, atol=1e-6, 4 terminal events, dense_output=True) segments
// (launch_rays.py:593-681; scipy 1.18.1 integrate/_ivp/{rk,ivp,common,base}.py) joined by specular
// reflections.  The structure here is NOT the reference's: there is no solver object, no list of
// interpolants and no post-hoc resampling -- the ray is a register-resident state machine
// (prt_ray) whose phases a warp executes in lock step, the sound-speed lookup is shared between c
// and dc/dz, events reuse the sound speed of the FSAL stage, asin is replaced by exact threshold
// tests, and save points are emitted while stepping.
//
// Rounding contract ("PRT arithmetic"): IEEE binary64, every written
// operation is one correctly rounded +,-,*,/,sqrt or fma; fma only where PRT_FMA is written
// (nvcc -fmad=false, g++ -ffp-contract=off).  The same contract is restated independently in
// oracle/prt_oracle.c, which is how bit-for-bit parity against the CPU oracle is testable.
//
// The functions are __host__ __device__ only so that tests/host_harness can compile this header
// with g++ and compare it bit-for-bit with the oracle on a machine without a GPU; the shipped
// library (libprt.so) only ever calls them from kernels and has no CPU execution path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PRT_HD __host__ __device__ __forceinline__
#define PRT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PRT_HD inline
#define PRT_HD_NOINLINE
#endif

// explicit re-convergence of the lanes in `mask` (all of them must arrive); nothing to do on the host
#if defined(__CUDA_ARCH__)
#define PRT_CONVERGE(mask) __syncwarp(mask)
#else
#define PRT_CONVERGE(mask) ((void)(mask))
#endif

struct alignas(16) prt_d2 {
    double x, y;
};

// -DPRT_BOUNDS_CHECK: every table index of the lookup functions is compared with the table size and violations are
// counted (prt_debug_oob(); compute-sanitizer is closed on the measurement pool, so the memcheck pass over the parity
// cases is done with this build: profiles/r02_bounds_check.md).  Compiled out otherwise.
#if defined(PRT_BOUNDS_CHECK) && defined(__CUDACC__)
__device__ unsigned long long prt_dbg_oob_count;
#endif
#if defined(PRT_BOUNDS_CHECK) && defined(__CUDA_ARCH__)
#define PRT_CHECK_IDX(idx, n) do { if (!((long long)(idx) >= 0 && (long long)(idx) < (long long)(n))) atomicAdd(&prt_dbg_oob_count, 1ULL); } while (0)
#else
#define PRT_CHECK_IDX(idx, n) do { } while (0)
#endif

#if defined(__CUDA_ARCH__)
PRT_HD prt_d2 prt_ld2(const prt_d2* p)
{
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    prt_d2 r;
    r.x = v.x;
    r.y = v.y;
    return r;
}
PRT_HD double prt_ld(const double* p) { return __ldg(p); }
#define PRT_FMA(a, b, c) __fma_rn((a), (b), (c))
PRT_HD double prt_bits_to_double(uint64_t b) { return __longlong_as_double((long long)b); }
PRT_HD uint64_t prt_double_to_bits(double d) { return (uint64_t)__double_as_longlong(d); }
PRT_HD uint32_t prt_hi32(double d) { return (uint32_t)__double2hiint(d); }
PRT_HD uint32_t prt_lo32(double d) { return (uint32_t)__double2loint(d); }
PRT_HD double prt_from_hilo(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
#else
#include <string.h>
PRT_HD prt_d2 prt_ld2(const prt_d2* p) { return *p; }
PRT_HD double prt_ld(const double* p) { return *p; }
#define PRT_FMA(a, b, c) fma((a), (b), (c))
PRT_HD double prt_bits_to_double(uint64_t b)
{
    double d;
    memcpy(&d, &b, 8);
    return d;
}
PRT_HD uint64_t prt_double_to_bits(double d)
{
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}
PRT_HD uint32_t prt_hi32(double d) { return (uint32_t)(prt_double_to_bits(d) >> 32); }
PRT_HD uint32_t prt_lo32(double d) { return (uint32_t)prt_double_to_bits(d); }
PRT_HD double prt_from_hilo(uint32_t hi, uint32_t lo) { return prt_bits_to_double(((uint64_t)hi << 32) | lo); }
#endif

// ---------------------------------------------------------------------------------------------
// Per-ray status: why the reference would have returned None for this launch angle
// (launch_rays.py:427-430, 443-456, 474-477; exceptions raised inside solve_ivp / interp1d).
// (values are those of include/prt.h; when prt.h was included first its macros are used)
#ifndef PRT_RAY_OK
enum {
    PRT_RAY_OK = 0,
    PRT_RAY_STEP_TOO_SMALL = 1,
    PRT_RAY_VERTICAL = 2,
    PRT_RAY_BBOX = 3,
    PRT_RAY_BACKWARDS = 4,
    PRT_RAY_BOTTOM_OOB = 5,
    PRT_RAY_BRENTQ_SIGN = 6,
    PRT_RAY_BRENTQ_ITER = 7,
    PRT_RAY_MAX_BOUNCES = 9
};
enum { PRT_CNT_NFEV = 0, PRT_CNT_ATTEMPTS, PRT_CNT_ACCEPTED, PRT_CNT_SEGMENTS, PRT_CNT_BISECT, PRT_CNT_DENSE, PRT_CNT_N };
#endif

// Device view of one environment (all pointers device memory, one allocation; see prt_env.cu).
struct prt_env_view {
    int32_t n_r, n_z, n_b, n_ba;
    const prt_d2* grid;     // [n_r][n_z] {c, dc/dz} interleaved: one 16-byte load per node
    const prt_d2* rpair;    // [n_r-1] {rin[i], rin[i+1]}   : cell edges in one 16-byte load
    const prt_d2* zpair;    // [n_z-1] {zin[j], zin[j+1]}
    const prt_d2* bxpair;   // [n_b-1] {depth_ranges[i], depth_ranges[i+1]}
    const prt_d2* bvpair;   // [n_b-1] {depths[i], depths[i+1]}
    const prt_d2* rcell;    // [n_r-1] x 2 prt_d2: {lo, hi}, {rin[i], 1/(rin[i+1]-rin[i])}; lo/hi = cell edges with
    const prt_d2* zcell;    // [n_z-1]   -inf / +inf at the two ends of the grid: "cell i holds v" <=> lo < v <= hi
    const double* rrecip;   // [n_r-1] 1 / (rin[i+1] - rin[i])
    const double* zrecip;   // [n_z-1] 1 / (zin[j+1] - zin[j])
    const double* brecip;   // [n_b-1] 1 / (depth_ranges[i+1] - depth_ranges[i])
    const double* ba_x;     // [n_ba] bottom-angle spline breakpoints
    const double* ba_c;     // [4][n_ba-1] PPoly coefficients, highest power first
    int64_t grid_stride;    // ensembles: nodes between consecutive realisations' grids (0 for one environment)
    int32_t n_env;          // number of realisations sharing the coordinate tables (1 for one environment)
    double r0, r_inv, z0, z_inv, b0, b_inv;  // cell-index guesses: (v - x0) * inv (any value is correct)
    double r_off, z_off, b_off;              // -x0 * inv: the same guesses as one fma, v * inv + off
    // smooth non-uniform depth grids (flat-earth transform): the linear guess s is refined to s + s q(s / n_z) with a
    // least-squares quadratic q fitted at staging time (prt_env_layout.h); z_cubic == 0 for uniform grids
    double z_b1, z_b2, z_b3, z_un;
    int32_t z_cubic;
    // exactly uniform depth grids (np.arange / np.linspace with a representable step): zin[j] == fma(j, z_dz, z0) and
    // 1/(zin[j+1]-zin[j]) == z_r0 bit for bit for every cell (checked at staging time) -- the stage lookups then
    // COMPUTE the lower edge and take the reciprocal as a constant instead of loading {zin[j], 1/dz_j} from a table:
    // one scattered load less per lookup, and the kernel is bound by the L1 data pipe (profiles/r02_ncu_fan.md)
    int32_t z_uniform;
    double z_dz, z_r0;
    int32_t r_uniform;  // the same for the range axis of the sound-speed grid: rin[i] == fma(i, r_dx, r0), 1/dr == r_r0
    double r_dx, r_r0;
    int32_t spec_off;       // 1: guesses too poor on this environment (irregular grids) -> every trial step uses the exact lookup
    const prt_d2* rxr;      // [n_r-1] {rin[i], 1/(rin[i+1]-rin[i])}: all the speculative stage lookup needs per axis
    const prt_d2* zxr;      // [n_z-1] {zin[j], 1/(zin[j+1]-zin[j])}
    const prt_d2* bcell;    // [n_b-1] x 3 prt_d2: {lo, hi} (padded like rcell), {x[i], 1/(x[i+1]-x[i])}, {depth[i], depth[i+1]}
    double bb_zmax, bb_zmin, bb_rmin, bb_rmax;  // zin[-1]+1e-6, zin[0]-1e-6, rin[0]-1e-6, rin[-1]+1e-6
};

// grid mode of an environment: 0 linear cell guesses + edge tables, 1 depth guess refined by the fitted quadratic,
// 2 exactly uniform depth grid (depth-cell edges computed), 3 exactly uniform depth AND range grids (both computed)
PRT_HD int prt_z_mode(const prt_env_view& v) { return v.z_uniform ? (v.r_uniform ? 3 : 2) : (v.z_cubic ? 1 : 0); }

// One ray's environment: the shared tables plus this ray's sound-speed grid (ensembles: one grid per
// realisation behind one set of tables).  The tables are referenced, never copied, so that on the device
// every field is read straight from the kernel's __grid_constant__ parameter.
struct prt_envp {
    const prt_env_view& v;
    const prt_d2* grid;
    int zm;   // grid mode (prt_z_mode), a compile-time constant in each kernel instantiation
};

// np.nextafter(t, +inf) (scipy rk.py:119) by integer increment of the bit pattern (select-only: it opens
// every step attempt)
PRT_HD double prt_next_up(double t)
{
    const uint64_t b = prt_double_to_bits(t);
    const uint64_t nb = (t > 0.0) ? b + 1 : ((t < 0.0) ? b - 1 : (uint64_t)1);  // 0 -> smallest subnormal
    const double nu = prt_bits_to_double(nb);
    return (fabs(t) < (double)INFINITY) ? nu : t;  // NaN, +-inf: unchanged (never reached: t is a finite range)
}

// ---------------------------------------------------------------------------------------------
// Python builtin min/max (scipy rk.py:155-163, common.py:117-134)
PRT_HD double prt_py_min(double a, double b) { return (b < a) ? b : a; }
PRT_HD double prt_py_max(double a, double b) { return (b > a) ? b : a; }
// np.maximum(np.abs(a), np.abs(b)) (rk.py:148-149, common.py:117) on the bit patterns: non-negative doubles order like
// unsigned integers and every NaN pattern lies above infinity, so the integer maximum IS the NaN-propagating maximum
// of the magnitudes -- integer instructions instead of three FP64 compares per component on the FP64 pipe
PRT_HD double prt_max_abs(double a, double b)
{
    const uint64_t x = prt_double_to_bits(a) & 0x7fffffffffffffffULL, y = prt_double_to_bits(b) & 0x7fffffffffffffffULL;
    return prt_bits_to_double(x > y ? x : y);
}

// ---------------------------------------------------------------------------------------------
// Elementary functions of the rounding contract

#define PRT_POW10_INIT {1.0, 0x1.ddb680117ab12p-1, 0x1.bdb8cdadbe120p-1, 0x1.9fdf8bcce533ep-1, 0x1.8406003b2ae5cp-1, \
                        0x1.6a09e667f3bcdp-1, 0x1.51cb453b9536cp-1, 0x1.3b2c47bff8329p-1, 0x1.2611186bae674p-1, \
                        0x1.125fbee250664p-1}
#if defined(__CUDACC__)
__constant__ double prt_pow10_dev[10] = PRT_POW10_INIT;
#endif
static const double prt_pow10_host[10] = PRT_POW10_INIT;
#if defined(__CUDA_ARCH__)
#define PRT_POW10_TAB(i) prt_pow10_dev[i]
#else
#define PRT_POW10_TAB(i) prt_pow10_host[i]
#endif

// x^(-1/5): stands for error_norm ** (-1/5), scipy rk.py:155,163.  Hot path = normal finite x; NaN,
// infinity and subnormals take the cold prologue.
PRT_HD_NOINLINE double prt_pow_m02_special(double x, double& post)
{
    post = 1.0;
    if (x != x) return x;
    if (x == (double)INFINITY) return x;
    if (x < 0x1p-1022) {
        post = 0x1p20;
        return x * 0x1p100;
    }
    return x;
}

PRT_HD double prt_pow_m02(double x)
{
    double post = 1.0;
    if (!(x >= 0x1p-1022 && x < (double)INFINITY)) {
        x = prt_pow_m02_special(x, post);
        if (x != x) return x;
        if (x == (double)INFINITY) return 0.0;
    }
    const uint64_t b = prt_double_to_bits(x);
    const int e = (int)((b >> 52) & 0x7ff) - 1023;
    const double m = prt_bits_to_double((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    // e = 5 q + r with r in 0..4 (floor division; e + 1025 > 0 keeps the integer division unsigned)
    const unsigned eu = (unsigned)(e + 1025);
    const int q = (int)(eu / 5u) - 205;
    const int r = (int)(eu - 5u * (eu / 5u));
    double y = PRT_FMA(PRT_FMA(0.05264997, m, -0.28570902), m, 1.23209767);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const double y2 = y * y;
        const double y4 = y2 * y2;
        const double y5 = y4 * y;
        const double res = PRT_FMA(-m, y5, 1.0);
        y = PRT_FMA(y * 0.2, res, y);
    }
    // 2^(-r/5)
    const double tr = (r == 0) ? 1.0
                    : (r == 1) ? 0x1.bdb8cdadbe120p-1
                    : (r == 2) ? 0x1.8406003b2ae5cp-1
                    : (r == 3) ? 0x1.51cb453b9536cp-1
                               : 0x1.2611186bae674p-1;
    y = y * tr;
    const double p2 = prt_bits_to_double((uint64_t)(1023 - q) << 52);
    return (y * p2) * post;
}

// x^(-1/10): the controller's error_norm ** (-1/5) on the mean square x = error_norm^2 (no sqrt).
// Branch-free: the biased exponent is clamped to [953, 1092] (2^-70 <= x < 2^70), outside of which the
// step-size factor 0.9 x^(-1/10) is pinned at its limits anyway (0 and subnormals -> >= 119 -> MAX_FACTOR;
// huge, inf and NaN bit patterns -> <= 2^-7 -> MIN_FACTOR), so no value of x needs a special case.
PRT_HD double prt_pow_m01(double x)
{
    const uint32_t hi = prt_hi32(x);
    int ef = (int)((hi >> 20) & 0x7ffu);
    ef = ef < 953 ? 953 : ef;
    ef = ef > 1092 ? 1092 : ef;
    const double m = prt_from_hilo((hi & 0x000fffffu) | 0x3ff00000u, prt_lo32(x));
    const unsigned eu = (unsigned)(ef + 7);  // e + 1030 with e = ef - 1023 = 10 q + r, r in 0..9
    const unsigned q10 = eu / 10u;           // q + 103
    const int r = (int)(eu - 10u * q10);
    double y = PRT_FMA(PRT_FMA(0.024999779689669537, m, -0.14120119176598264), m, 1.1157678629591425);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const double y2 = y * y;
        const double y4 = y2 * y2;
        const double y8 = y4 * y4;
        const double y10 = y8 * y2;
        const double res = PRT_FMA(-m, y10, 1.0);
        y = PRT_FMA(y * 0.1, res, y);
    }
    y = y * PRT_POW10_TAB(r);  // 2^(-r/10)
    return y * prt_from_hilo((uint32_t)(1126u - q10) << 20, 0u);  // 2^(-q)
}

// 1/sqrt(w), normal w > 0: integer-shift seed on the high word + 4 Newton steps (<= 2.3 ulp); see oracle orc_rsqrt.
// Takes hw = w / 2 (what the Newton step multiplies by): halving is exact and commutes with the rounding of the
// product that formed w, so callers form hw directly and w itself is never needed; hi32(w) = hi32(hw) + 2^20.
PRT_HD double prt_rsqrt_half(double hw)
{
    double y = prt_from_hilo(0x5FE6EB50u - ((prt_hi32(hw) + 0x00100000u) >> 1), 0u);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const double t = y * y;
        const double r = PRT_FMA(-hw, t, 1.5);
        y = y * r;
    }
    return y;
}

// 1/x, normal x > 0: integer seed on the high word (5 %) + 4 Newton steps (<= 1 ulp); see oracle orc_rcp.
// Used where the reference divides by a quantity that is known before the dividend is (error scales, step
// length): the reciprocal is formed off the critical path and the division becomes a multiplication.
PRT_HD double prt_rcp(double x)
{
    double y = prt_from_hilo(0x7FDE6238u - prt_hi32(x), 0u);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const double e = PRT_FMA(-x, y, 1.0);
        y = PRT_FMA(y, e, y);
    }
    return y;
}

PRT_HD double prt_k_sin(double r)
{
    double z = r * r;
    double s = 1.0 / 355687428096000.0;
    s = PRT_FMA(s, z, -1.0 / 1307674368000.0);
    s = PRT_FMA(s, z, 1.0 / 6227020800.0);
    s = PRT_FMA(s, z, -1.0 / 39916800.0);
    s = PRT_FMA(s, z, 1.0 / 362880.0);
    s = PRT_FMA(s, z, -1.0 / 5040.0);
    s = PRT_FMA(s, z, 1.0 / 120.0);
    s = PRT_FMA(s, z, -1.0 / 6.0);
    return PRT_FMA(r * z, s, r);
}

PRT_HD double prt_k_cos(double r)
{
    double z = r * r;
    double s = 1.0 / 20922789888000.0;
    s = PRT_FMA(s, z, -1.0 / 87178291200.0);
    s = PRT_FMA(s, z, 1.0 / 479001600.0);
    s = PRT_FMA(s, z, -1.0 / 3628800.0);
    s = PRT_FMA(s, z, 1.0 / 40320.0);
    s = PRT_FMA(s, z, -1.0 / 720.0);
    s = PRT_FMA(s, z, 1.0 / 24.0);
    s = PRT_FMA(s, z, -0.5);
    return PRT_FMA(z, s, 1.0);
}

#define PRT_PIO2_HI 0x1.921fb54442d18p+0
#define PRT_PIO2_LO 0x1.1a62633145c07p-54
#define PRT_TWO_OVER_PI 0.6366197723675814
#define PRT_RAD2DEG 57.29577951308232
#define PRT_DEG2RAD 0.017453292519943295

// sin(x): np.sin of launch_rays.py:143, 285, 480
PRT_HD_NOINLINE double prt_sin(double x)
{
    if (x != x || fabs(x) == (double)INFINITY) return (double)NAN;
    if (fabs(x) <= 0.7853981633974483) return prt_k_sin(x);
    double n = floor(PRT_FMA(x, PRT_TWO_OVER_PI, 0.5));
    double r = PRT_FMA(-n, PRT_PIO2_HI, x);
    r = PRT_FMA(-n, PRT_PIO2_LO, r);
    int k = (int)((long long)n & 3);
    if (k == 0) return prt_k_sin(r);
    if (k == 1) return prt_k_cos(r);
    if (k == 2) return -prt_k_sin(r);
    return -prt_k_cos(r);
}

PRT_HD double prt_asin_core(double a)
{
    double th = PRT_FMA(a * a, a * (1.0 / 6.0), a);
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
        double d = (prt_k_sin(th) - a) / prt_k_cos(th);
        th = th - d;
    }
    return th;
}

// asin(x): np.arcsin of integration_processes.py:333; |x| > 1 -> NaN
PRT_HD_NOINLINE double prt_asin(double x)
{
    if (x != x) return x;
    double a = fabs(x);
    if (a > 1.0) return NAN;
    double t;
    if (a <= 0.5) {
        t = prt_asin_core(a);
    } else {
        double s = sqrt((1.0 - a) * 0.5);
        t = PRT_FMA(-2.0, prt_asin_core(s), PRT_PIO2_HI);
    }
    return (x < 0.0) ? -t : t;
}

// ---------------------------------------------------------------------------------------------
// Cell lookup: k = clamp(np.searchsorted(a, v, side="left") - 1, 0, n-2)
// (integration_processes.py:152-157, 221-224), returning the cell edges {a[k], a[k+1]} with it.
// Guess from the mean spacing, hunt a few cells, bisect as a fallback: result is exact for any
// sorted grid; the guess only affects speed.
// cold path of prt_locate (NaN, grids far from uniform): lower_bound over the n nodes
// (node i = pair[i].x, node n-1 = pair[n-2].y)
PRT_HD_NOINLINE int prt_locate_slow(const prt_d2* pair, int n, double v)
{
    const int m = n - 2;
    if (v != v) return m;  // NaN sorts last in np.searchsorted
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        double a = (mid <= m) ? prt_ld(&pair[mid].x) : prt_ld(&pair[m].y);
        if (a < v) lo = mid + 1;
        else hi = mid;
    }
    int g = lo - 1;
    g = g > m ? m : g;
    g = g < 0 ? 0 : g;
    return g;
}

#if defined(__CUDA_ARCH__)
#define PRT_D2I(x) __double2int_rz(x)  // saturating; NaN -> 0
#else
#define PRT_D2I(x) (((x) != (x)) ? 0 : ((x) >= 2147483647.0 ? 2147483647 : ((x) <= -2147483648.0 ? (-2147483647 - 1) : (int)(x))))
#endif

// Fast path: guess the cell from the mean spacing, load its edges, verify; one neighbour step for
// gently non-uniform grids (flat-earth depths); anything else goes to the cold path.  The result is
// exactly clamp(searchsorted_left - 1) whatever the guess; cell g is right iff
//   (g == 0 || a[g] < v) && (g == n-2 || v <= a[g+1])      (false for NaN unless the grid has one cell)
PRT_HD int prt_locate(const prt_d2* pair, int n, double x0, double inv, double v, prt_d2& edges)
{
    const int m = n - 2;
    int g = PRT_D2I((v - x0) * inv);
    g = g < 0 ? 0 : g;
    g = g > m ? m : g;
    PRT_CHECK_IDX(g, n - 1);
    edges = prt_ld2(pair + g);
    if ((g == 0 || edges.x < v) && (g == m || v <= edges.y)) return g;
    g += (edges.x < v) ? 1 : -1;  // failed on the upper edge -> one cell up, else one cell down
    g = g < 0 ? 0 : g;
    g = g > m ? m : g;
    edges = prt_ld2(pair + g);
    if ((g == 0 || edges.x < v) && (g == m || v <= edges.y)) return g;
    g = prt_locate_slow(pair, n, v);
    edges = prt_ld2(pair + g);
    return g;
}

// c(x, z) and dc/dz(x, z): bilinear_interp (integration_processes.py:101-174) evaluated ONCE for
// both fields (the reference evaluates it twice with identical weights, :86-87).  The weight
// divisions use the per-cell reciprocals 1/(x[i+1]-x[i]) tabulated at staging time; the four-weight
// sum is regrouped as lerps (contract, mirrored in the oracle):
//   l0 = v00 + wy (v01 - v00),  l1 = v10 + wy (v11 - v10),  v = l0 + wx (l1 - l0)
PRT_HD void prt_bilerp(const prt_d2& v00, const prt_d2& v01, const prt_d2& v10, const prt_d2& v11, double wx,
                       double wy, double& c, double& cz)
{
    const double c0 = PRT_FMA(wy, v01.x - v00.x, v00.x), c1 = PRT_FMA(wy, v11.x - v10.x, v10.x);
    const double d0 = PRT_FMA(wy, v01.y - v00.y, v00.y), d1 = PRT_FMA(wy, v11.y - v10.y, v10.y);
    c = PRT_FMA(wx, c1 - c0, c0);
    cz = PRT_FMA(wx, d1 - d0, d0);
}

// guessed cell of v: trunc(v * inv + off) folded into [0, m] (negative and NaN guesses land on a valid cell;
// any cell is a correct GUESS)
PRT_HD unsigned prt_guess(double v, double inv, double off, int m)
{
    const unsigned g = (unsigned)PRT_D2I(PRT_FMA(v, inv, off));
    return g > (unsigned)m ? (unsigned)m : g;
}

// depth axis: optionally refined guess (see prt_env_view::z_cubic); s = the linear guess z * z_inv + z_off
PRT_HD unsigned prt_guess_z_from(const prt_envp& E, double s)
{
    const prt_env_view& V = E.v;
    if (E.zm == 1) {
        const double u = s * V.z_un;
        s = PRT_FMA(s, PRT_FMA(PRT_FMA(V.z_b3, u, V.z_b2), u, V.z_b1), s);
    }
    const unsigned g = (unsigned)PRT_D2I(s);
    const unsigned m = (unsigned)(V.n_z - 2);
    return g > m ? m : g;
}

PRT_HD unsigned prt_guess_z(const prt_envp& E, double z) { return prt_guess_z_from(E, PRT_FMA(z, E.v.z_inv, E.v.z_off)); }

// Exact lookup, one cache round trip in the common case: every load (cell records and the four corner
// nodes) is issued from the guessed indices at once; the guesses are verified afterwards (two compares per
// axis against edges padded with -inf / +inf at the grid ends, false for NaN) and the rare miss
// (cell-boundary rounding, non-uniform grids) takes the general lookup.
PRT_HD void prt_field(const prt_envp& E, double x, double z, double& c, double& cz)
{
    const prt_env_view& V = E.v;
    unsigned i = prt_guess(x, V.r_inv, V.r_off, V.n_r - 2);
    unsigned j = prt_guess_z(E, z);
    PRT_CHECK_IDX(i, V.n_r - 1);
    PRT_CHECK_IDX(j, V.n_z - 1);
    const prt_d2* rc = V.rcell + 2 * i;
    const prt_d2* zc = V.zcell + 2 * j;
    const prt_d2 rv = prt_ld2(rc), ra = prt_ld2(rc + 1);
    const prt_d2 zv = prt_ld2(zc), za = prt_ld2(zc + 1);
    const prt_d2* g = E.grid + (i * (unsigned)V.n_z + j);
    prt_d2 v00 = prt_ld2(g);
    prt_d2 v01 = prt_ld2(g + 1);
    prt_d2 v10 = prt_ld2(g + V.n_z);
    prt_d2 v11 = prt_ld2(g + V.n_z + 1);
    double wx, wy;
    if ((rv.x < x) & (x <= rv.y) & (zv.x < z) & (z <= zv.y)) {
        wx = (x - ra.x) * ra.y;
        wy = (z - za.x) * za.y;
    } else {
        prt_d2 re, ze;
        i = (unsigned)prt_locate(V.rpair, V.n_r, V.r0, V.r_inv, x, re);
        j = (unsigned)prt_locate(V.zpair, V.n_z, V.z0, V.z_inv, z, ze);
        PRT_CHECK_IDX(i, V.n_r - 1);
        PRT_CHECK_IDX(j, V.n_z - 1);
        g = E.grid + (i * (unsigned)V.n_z + j);
        v00 = prt_ld2(g);
        v01 = prt_ld2(g + 1);
        v10 = prt_ld2(g + V.n_z);
        v11 = prt_ld2(g + V.n_z + 1);
        wx = (x - re.x) * prt_ld(V.rrecip + i);
        wy = (z - ze.x) * prt_ld(V.zrecip + j);
    }
    prt_bilerp(v00, v01, v10, v11, wx, wy, c, cz);
}

// Speculative lookup for the Runge-Kutta stages: the same arithmetic from the guessed cell, but NO branch
// and no extra load -- the guess is proven on the interpolation weight itself.  w = (v - x[i]) * (1/(x[i+1]-x[i]))
// is computed from the true lower edge and width of cell i, so
//   w > 0          <=>  v > x[i]     (an IEEE difference has the exact sign and the reciprocal is positive),
//   w < 1 - 2^-40   =>  v < x[i+1]   (v >= x[i+1] gives w >= 1 - 2^-50 after the three roundings),
// and both together say that i IS clamp(searchsorted(v) - 1).  Only the lower test of the first DEPTH cell is
// waived (a ray overshoots the surface by part of a step before every surface reflection; the reference clamps
// there and extrapolates with the same formula) -- the other ends of the grids are reached once per ray at most
// (the receiver range is usually the last range node) and are left to the exact re-run.  `ok` collects
// the proofs and the caller re-runs the whole step attempt through the exact lookup when any failed
// (prt_ray::attempt): an abscissa on a node or within 1e-12 cell widths below one, and wrong guesses on
// non-uniform grids.  The stage code of an attempt is thus one branch-free block that the scheduler is free
// to interleave across stages.
#define PRT_W_HI (1.0 - 0x1p-40)
// The proof on the high word only (one integer compare, off the FP64 pipe): 1 <= hi(w) <= 0x3FEFFFFE  =>  w is a
// positive number below 1 - 2^-20 < PRT_W_HI.  Conservative on purpose: +0, and weights within 2^-20 of 1 (an
// abscissa within a millionth of a cell below a node) are sent to the exact re-run, which returns the same bits.
PRT_HD bool prt_w_proven(double w) { return (prt_hi32(w) - 1u) < 0x3FEFFFFEu; }
// first depth cell: any finite w <= -0 is also accepted (above the surface the reference clamps the cell and extrapolates)
PRT_HD bool prt_w_negative_finite(double w) { return (prt_hi32(w) - 0x80000000u) < 0x7FF00000u; }
struct prt_xcell {
    unsigned row;  // i * n_z: element index of grid[i][0] (32-bit on purpose: one IMAD.WIDE per node address)
    double wx;
};

PRT_HD prt_xcell prt_xlook(const prt_envp& E, double x, bool& ok)
{
    const prt_env_view& V = E.v;
    const unsigned i = prt_guess(x, V.r_inv, V.r_off, V.n_r - 2);
    PRT_CHECK_IDX(i, V.n_r - 1);
    prt_d2 ra;
    if (E.zm == 3) {
        ra.x = PRT_FMA((double)i, V.r_dx, V.r0);
        ra.y = V.r_r0;
    } else {
        ra = prt_ld2(V.rxr + i);
    }
    prt_xcell xc;
    xc.row = i * (unsigned)V.n_z;
    xc.wx = (x - ra.x) * ra.y;
    ok = ok & prt_w_proven(xc.wx);
    return xc;
}

// sg: the real-valued linear cell guess of z (any value is a correct guess; see PRT_FOLD_GUESS in prt_rk_try)
PRT_HD void prt_field_spec(const prt_envp& E, const prt_xcell& xc, double z, double sg, double& c, double& cz, bool& ok)
{
    const prt_env_view& V = E.v;
    const unsigned j = prt_guess_z_from(E, sg);
    const unsigned e0 = xc.row + j, e1 = e0 + (unsigned)V.n_z;
    PRT_CHECK_IDX(j, V.n_z - 1);
    PRT_CHECK_IDX(e1 + 1u, (long long)V.n_r * V.n_z);
    const prt_d2 *g0 = E.grid + e0, *g1 = E.grid + e1;
    prt_d2 za;
    if (E.zm >= 2) {
        za.x = PRT_FMA((double)j, V.z_dz, V.z0);
        za.y = V.z_r0;
    } else {
        za = prt_ld2(V.zxr + j);
    }
    const prt_d2 v00 = prt_ld2(g0), v01 = prt_ld2(g0 + 1), v10 = prt_ld2(g1), v11 = prt_ld2(g1 + 1);
    const double wy = (z - za.x) * za.y;
    ok = ok & (prt_w_proven(wy) | ((j == 0u) & prt_w_negative_finite(wy)));  // waiver: above the surface (overshoot before a reflection)
    prt_bilerp(v00, v01, v10, v11, xc.wx, wy, c, cz);
}

// bottom depth at x: linear_interp (integration_processes.py:177-235); same guess / verify scheme, one
// record {padded edges, x[i] and reciprocal, depth[i] and depth[i+1]} per bathymetry cell
PRT_HD double prt_bathy(const prt_envp& E, double x)
{
    const prt_env_view& V = E.v;
    const unsigned g = prt_guess(x, V.b_inv, V.b_off, V.n_b - 2);
    PRT_CHECK_IDX(g, V.n_b - 1);
    const prt_d2* bc = V.bcell + 3 * g;
    const prt_d2 e = prt_ld2(bc), a = prt_ld2(bc + 1);
    prt_d2 v = prt_ld2(bc + 2);
    double w = (x - a.x) * a.y;
    if (!((e.x < x) & (x <= e.y))) {
        prt_d2 xe;
        const int i = prt_locate(V.bxpair, V.n_b, V.b0, V.b_inv, x, xe);
        v = prt_ld2(V.bvpair + i);
        w = (x - xe.x) * prt_ld(V.brecip + i);
    }
    return PRT_FMA(w, v.y, (1.0 - w) * v.x);
}

// speculative bottom depth for the trial step (same proof on the weight as prt_field_spec)
PRT_HD double prt_bathy_spec(const prt_envp& E, double x, bool& ok)
{
    const prt_env_view& V = E.v;
    const unsigned g = prt_guess(x, V.b_inv, V.b_off, V.n_b - 2);
    PRT_CHECK_IDX(g, V.n_b - 1);
    const prt_d2* bc = V.bcell + 3 * g;
    const prt_d2 a = prt_ld2(bc + 1), v = prt_ld2(bc + 2);
    const double w = (x - a.x) * a.y;
    ok = ok & prt_w_proven(w);
    return PRT_FMA(w, v.y, (1.0 - w) * v.x);
}

// derivsrd (integration_processes.py:26-98) from c and dc/dz, with a single reciprocal square root:
//   u = 1 / (c^2 sqrt(arg)) = rsqrt(c^4 arg);  [dT, dz, dp] = [c u, (c^3 p) u, -cz u]
// (the Newton iteration works on c^4 arg / 2, formed as (c^2 (c^2 / 2)) arg: same bits as halving c^4 arg)
// (every output is ONE multiplication after u; c^3 p is formed while the Newton iteration runs)
PRT_HD void prt_rhs_tail(double c, double cz, double p, double& dT, double& dz, double& dp)
{
    const double c2 = c * c;
    const double arg = PRT_FMA(-c2, p * p, 1.0);
    const double c4h = c2 * (0.5 * c2);  // c^4 / 2, exactly half of c2 * c2
    const double hw = (arg <= 0.0) ? c4h * 1e-30 : c4h * arg;  // "if arg <= 0: arg = 1e-30"
    const double c3p = (c2 * c) * p;
    const double u = prt_rsqrt_half(hw);
    dT = c * u;
    dz = c3p * u;
    dp = (-cz) * u;
}

// the same for the speculative trial step: "if arg <= 0: arg = 1e-30" (a ray beyond vertical) is not selected on the
// dependent chain but proven absent on the high word of arg (0 < arg <= 1 here), else the attempt is re-run exactly
PRT_HD void prt_rhs_tail_spec(double c, double cz, double p, double& dT, double& dz, double& dp, bool& ok)
{
    const double c2 = c * c;
    const double arg = PRT_FMA(-c2, p * p, 1.0);
    const double c4h = c2 * (0.5 * c2);
    ok = ok & ((prt_hi32(arg) - 1u) < 0x7FEFFFFFu);  // positive and finite
    const double hw = c4h * arg;
    const double c3p = (c2 * c) * p;
    const double u = prt_rsqrt_half(hw);
    dT = c * u;
    dz = c3p * u;
    dp = (-cz) * u;
}

PRT_HD void prt_rhs(const prt_envp& E, double x, double z, double p, double& dT, double& dz, double& dp,
                    double& c_out)
{
    double c, cz;
    prt_field(E, x, z, c, cz);
    prt_rhs_tail(c, cz, p, dT, dz, dp);
    c_out = c;
}

// The same as ONE out-of-line function for the seven stages of the exact trial step (the cold path of an attempt and
// the rebuild of a parked step): 14 inlined copies of the general lookup were a third of the kernels' 132 KB of
// code, and the kernels stall on instruction fetch when their warps are in different phases (35 % of the stall
// samples of the eigenray refinement, profiles/r02_ncu_final.md): 92 KB now, refinement 14 % faster, fan unchanged.
// (Going further -- segment start-up and root search through the call as well -- gains another 5 % on the refinement
// and loses 1-9 % on the fans; measured, not shipped.)  Same arithmetic, same bits.
PRT_HD_NOINLINE void prt_rhs_cold(const prt_env_view* V, const prt_d2* grid, int zm, double x, double z, double p,
                                  double* out4)
{
    const prt_envp E{*V, grid, zm};
    prt_rhs(E, x, z, p, out4[0], out4[1], out4[2], out4[3]);
}

PRT_HD void prt_rhs_spec(const prt_envp& E, const prt_xcell& xc, double z, double p, double sg, double& dT, double& dz,
                         double& dp, double& c_out, bool& ok)
{
    double c, cz;
    prt_field_spec(E, xc, z, sg, c, cz, ok);
    prt_rhs_tail_spec(c, cz, p, dT, dz, dp, ok);
    c_out = c;
}

// The four event functions (integration_processes.py:238-303) as a 4-bit mask of "+1" values.
// theta = degrees(asin(p*c)) is only ever compared with 0 and +-(90 - 1e-3), so
//   theta < 0  <=>  -1 <= p*c < 0,   theta > 0  <=>  0 < p*c <= 1          (asin is odd, NaN outside)
//   |theta| > 90-1e-3  <=>  literal evaluation, needed only when |p*c| is within 1e-8 of 1.
#define PRT_EV_SURFACE 1
#define PRT_EV_BOTTOM 2
#define PRT_EV_VERTICAL 4
#define PRT_EV_BBOX 8

PRT_HD bool prt_ev_vertical(double pc)
{
    double a = fabs(pc);
    if (!(a > 0.99999999)) return false;  // sin(89.999 deg) = 1 - 1.5e-10
    if (a > 1.0) return false;            // asin -> NaN, comparison false
    return fabs(prt_asin(pc) * PRT_RAD2DEG) > (90.0 - 1e-3);
}

PRT_HD int prt_events(const prt_envp& E, double x, double z, double p, double c)
{
    double pc = p * c;
    int m = 0;
    if ((z < 0.0) && (pc < 0.0) && (pc >= -1.0)) m |= PRT_EV_SURFACE;
    if ((pc > 0.0) & (pc <= 1.0) & (z > prt_bathy(E, x))) m |= PRT_EV_BOTTOM;  // bathymetry looked up unconditionally
    if (prt_ev_vertical(pc)) m |= PRT_EV_VERTICAL;
    if ((z > E.v.bb_zmax) | (z < E.v.bb_zmin) | (x < E.v.bb_rmin) | (x > E.v.bb_rmax)) m |= PRT_EV_BBOX;
    return m;
}

// single event evaluated on demand (inside the root search)
PRT_HD bool prt_event_one(const prt_envp& E, int which, double x, double z, double p)
{
    if (which == 3) return (z > E.v.bb_zmax) | (z < E.v.bb_zmin) | (x < E.v.bb_rmin) | (x > E.v.bb_rmax);
    if (which == 0 && !(z < 0.0)) return false;
    double c, cz;
    prt_field(E, x, z, c, cz);
    double pc = p * c;
    if (which == 0) return (pc < 0.0) && (pc >= -1.0);
    if (which == 1) return (pc > 0.0) && (pc <= 1.0) && (z > prt_bathy(E, x));
    return prt_ev_vertical(pc);
}

// ---------------------------------------------------------------------------------------------
// Dormand-Prince tableau (scipy rk.py:538-565), zeros omitted
// Dormand-Prince tableau (scipy rk.py:538-565, zeros omitted) and dense-output matrix P (rk.py:554-565).
// On the device the coefficients live in __constant__ memory so that DFMA reads them as constant-bank
// operands (a 64-bit immediate costs two UMOVs per use: 7 % of all executed instructions before).
#define PRT_TABLEAU_INIT { \
    /*  0 A21 */ 1.0 / 5, \
    /*  1 A31 */ 3.0 / 40, \
    /*  2 A32 */ 9.0 / 40, \
    /*  3 A41 */ 44.0 / 45, \
    /*  4 A42 */ -56.0 / 15, \
    /*  5 A43 */ 32.0 / 9, \
    /*  6 A51 */ 19372.0 / 6561, \
    /*  7 A52 */ -25360.0 / 2187, \
    /*  8 A53 */ 64448.0 / 6561, \
    /*  9 A54 */ -212.0 / 729, \
    /* 10 A61 */ 9017.0 / 3168, \
    /* 11 A62 */ -355.0 / 33, \
    /* 12 A63 */ 46732.0 / 5247, \
    /* 13 A64 */ 49.0 / 176, \
    /* 14 A65 */ -5103.0 / 18656, \
    /* 15 B1  */ 35.0 / 384, \
    /* 16 B3  */ 500.0 / 1113, \
    /* 17 B4  */ 125.0 / 192, \
    /* 18 B5  */ -2187.0 / 6784, \
    /* 19 B6  */ 11.0 / 84, \
    /* 20 E1  */ -71.0 / 57600, \
    /* 21 E3  */ 71.0 / 16695, \
    /* 22 E4  */ -71.0 / 1920, \
    /* 23 E5  */ 17253.0 / 339200, \
    /* 24 E6  */ -22.0 / 525, \
    /* 25 E7  */ 1.0 / 40, \
    /* 26 C2  */ 1.0 / 5, \
    /* 27 C3  */ 3.0 / 10, \
    /* 28 C4  */ 4.0 / 5, \
    /* 29 C5  */ 8.0 / 9, \
    /* 30 P11 */ -8048581381.0 / 2820520608.0, \
    /* 31 P31 */ 131558114200.0 / 32700410799.0, \
    /* 32 P41 */ -1754552775.0 / 470086768.0, \
    /* 33 P51 */ 127303824393.0 / 49829197408.0, \
    /* 34 P61 */ -282668133.0 / 205662961.0, \
    /* 35 P71 */ 40617522.0 / 29380423.0, \
    /* 36 P12 */ 8663915743.0 / 2820520608.0, \
    /* 37 P32 */ -68118460800.0 / 10900136933.0, \
    /* 38 P42 */ 14199869525.0 / 1410260304.0, \
    /* 39 P52 */ -318862633887.0 / 49829197408.0, \
    /* 40 P62 */ 2019193451.0 / 616988883.0, \
    /* 41 P72 */ -110615467.0 / 29380423.0, \
    /* 42 P13 */ -12715105075.0 / 11282082432.0, \
    /* 43 P33 */ 87487479700.0 / 32700410799.0, \
    /* 44 P43 */ -10690763975.0 / 1880347072.0, \
    /* 45 P53 */ 701980252875.0 / 199316789632.0, \
    /* 46 P63 */ -1453857185.0 / 822651844.0, \
    /* 47 P73 */ 69997945.0 / 29380423.0 }
#if defined(__CUDACC__)
__constant__ double prt_tab_dev[48] = PRT_TABLEAU_INIT;
#endif
static const double prt_tab_host[48] = PRT_TABLEAU_INIT;
#if defined(__CUDA_ARCH__)
#define PRT_TAB(i) prt_tab_dev[i]
#else
#define PRT_TAB(i) prt_tab_host[i]
#endif
#define A21 PRT_TAB(0)
#define A31 PRT_TAB(1)
#define A32 PRT_TAB(2)
#define A41 PRT_TAB(3)
#define A42 PRT_TAB(4)
#define A43 PRT_TAB(5)
#define A51 PRT_TAB(6)
#define A52 PRT_TAB(7)
#define A53 PRT_TAB(8)
#define A54 PRT_TAB(9)
#define A61 PRT_TAB(10)
#define A62 PRT_TAB(11)
#define A63 PRT_TAB(12)
#define A64 PRT_TAB(13)
#define A65 PRT_TAB(14)
#define B1 PRT_TAB(15)
#define B3 PRT_TAB(16)
#define B4 PRT_TAB(17)
#define B5 PRT_TAB(18)
#define B6 PRT_TAB(19)
#define E1 PRT_TAB(20)
#define E3 PRT_TAB(21)
#define E4 PRT_TAB(22)
#define E5 PRT_TAB(23)
#define E6 PRT_TAB(24)
#define E7 PRT_TAB(25)
#define C2 PRT_TAB(26)
#define C3 PRT_TAB(27)
#define C4 PRT_TAB(28)
#define C5 PRT_TAB(29)
#define P11 PRT_TAB(30)
#define P31 PRT_TAB(31)
#define P41 PRT_TAB(32)
#define P51 PRT_TAB(33)
#define P61 PRT_TAB(34)
#define P71 PRT_TAB(35)
#define P12 PRT_TAB(36)
#define P32 PRT_TAB(37)
#define P42 PRT_TAB(38)
#define P52 PRT_TAB(39)
#define P62 PRT_TAB(40)
#define P72 PRT_TAB(41)
#define P13 PRT_TAB(42)
#define P33 PRT_TAB(43)
#define P43 PRT_TAB(44)
#define P53 PRT_TAB(45)
#define P63 PRT_TAB(46)
#define P73 PRT_TAB(47)

#define PRT_SQRT3 1.7320508075688772
#define PRT_EPS 2.220446049250313e-16

PRT_HD double prt_rms3(double a, double b, double c)
{
    double s = a * a;
    s = PRT_FMA(b, b, s);
    s = PRT_FMA(c, c, s);
    return sqrt(s) / PRT_SQRT3;
}

// Q = K^T P for one component (scipy rk.py:178-180), accumulated by prt_rk_try; Q0 = K1 because P[:,0] = e_1
struct prt_q3 {
    double q1, q2, q3;
};
// RkDenseOutput._call_impl (scipy rk.py:715-737) for one component; xx = (t - t_old) * (1/h)
PRT_HD double prt_dense_eval(double q0, const prt_q3& q, double h, double y_old, double p1, double p2, double p3,
                             double p4)
{
    double a = q0 * p1;
    a = PRT_FMA(q.q1, p2, a);
    a = PRT_FMA(q.q2, p3, a);
    a = PRT_FMA(q.q3, p4, a);
    return PRT_FMA(h, a, y_old);
}

// cubic spline of the bottom angle (launch_rays.py:397-399, 469) as host-prepared PPoly
PRT_HD double prt_bottom_angle(const prt_env_view& E, double x, bool& oob)
{
    const int n = E.n_ba;
    oob = false;
    double x_first = prt_ld(E.ba_x), x_last = prt_ld(E.ba_x + n - 1);
    if (!(x >= x_first) || !(x <= x_last)) {
        oob = true;
        return 0.0;
    }
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (prt_ld(E.ba_x + mid) < x) lo = mid + 1;
        else hi = mid;
    }
    int i = lo - 1;
    i = i < 0 ? 0 : i;
    i = i > n - 2 ? n - 2 : i;
    PRT_CHECK_IDX(i, n - 1);
    double s = x - prt_ld(E.ba_x + i);
    const int m = n - 1;
    double v = prt_ld(E.ba_c + i);
    v = PRT_FMA(v, s, prt_ld(E.ba_c + m + i));
    v = PRT_FMA(v, s, prt_ld(E.ba_c + 2 * m + i));
    v = PRT_FMA(v, s, prt_ld(E.ba_c + 3 * m + i));
    return v;
}

// y0 = [0, z_src, sin(radians(angle)) / c(src)]  (launch_rays.py:140-144, 284-285)
PRT_HD void prt_initial_state(const prt_envp& E, double source_depth, double source_range, double angle_deg,
                              double& T, double& z, double& p)
{
    double c, cz;
    prt_field(E, source_range, source_depth, c, cz);
    T = 0.0;
    z = source_depth;
    p = prt_sin(angle_deg * PRT_DEG2RAD) / c;
}

// np.argmin(np.abs(range_save - v)) for ascending range_save, scanning forward from `start`
// (the minimum is never below start)
PRT_HD int prt_argmin_from(const double* rs, int n, int start, double v)
{
    int best = start;
    double bd = fabs(prt_ld(rs + start) - v);
    for (int i = start + 1; i < n; ++i) {
        double d = fabs(prt_ld(rs + i) - v);
        if (d < bd) {
            bd = d;
            best = i;
        } else
            break;
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// Trajectory sink: stores save point i of this ray.
struct prt_traj_sink {
    double* T;       // &out_T[ray][0]
    double* z;
    double* p;
    int64_t stride;  // element stride between consecutive save points of one ray
#if defined(PRT_BOUNDS_CHECK)
    int n_save;      // bounds-check builds only
#endif
    double sgn;      // +1: ODE convention; -1: stored convention (-z, -p) of ray_objects.py:49-52
    PRT_HD void put(int i, double vT, double vz, double vp) const
    {
        PRT_CHECK_IDX(i, n_save);
        T[(int64_t)i * stride] = vT;
        z[(int64_t)i * stride] = vz * sgn;
        p[(int64_t)i * stride] = vp * sgn;
    }
};

struct prt_ray_result {
    double T, z, p;
    int32_t n_bottom, n_surface, status;
};

// problem constants shared by every ray of a launch
struct prt_shot {
    double x_end, rtol, atol;
    int32_t terminate_backwards, n_save, max_bounces;
    const double* range_save;
};

// ---------------------------------------------------------------------------------------------
// One trial step of RK45: rk_step (scipy rk.py:14-71) + the error estimate and the controller power of
// _step_impl (rk.py:148-165).  Everything the accept / reject decision and an accepted step need.
struct prt_try {
    double k7T, k7z, k7p;  // f_new (FSAL stage)
    double nT, nz, np_;    // y_new
    double c_new;          // sound speed at (t + h, y_new), by-product of the FSAL stage
    double ms;             // mean square of the scaled error = error_norm^2
    double pw;             // ms^(-1/10) = error_norm^(-1/5)
    prt_q3 qT, qz, qp;     // dense-output rows Q = K^T P and rcp(h): only filled by prt_rk_try<.., DENSE = true>
    double inv_h;
    int g_new;             // event functions at (t + h, y_new) as a mask, the vertical-ray test still open when ...
    bool vert_open;        // ... |p c| is within 1e-8 of 1 (then asin decides, prt_ev_vertical)
};

// EXACT = false: speculative cell lookups, no branch anywhere; `ok` tells whether every guess was right.
// EXACT = true : the general lookup (prt_field) in every stage; same arithmetic, hence the same bits
//                whenever the speculative version reports ok.
// DENSE = true : also forms the dense-output rows and rcp(h) (trajectory-saving kernels, where nearly every
//                accepted step needs them: inside this block they fill the stalls of the stage chain)
//
// Every linear combination of stage values the step needs -- the inputs of the later stages (rows of A), y_new (B),
// the error estimate (E) and, with DENSE, the rows of Q = K^T P -- is a sum written in stage order, k1 first
// (rk.py:59-71, 176-180).  They are accumulated here AS THE STAGE VALUES ARRIVE: each accumulator still receives
// its terms in stage order, one fma per term, so every sum has the bits of the sum written out after the last
// stage, but no stage value outlives the stage that produced it.  Without DENSE an attempt keeps 6 accumulators
// alive across the last stage instead of 17 stage values, which is what lets a fifth block of 128 threads be
// resident per SM in the arrival-table kernels (profiles/r02_variants.md).  The travel-time row never feeds a
// stage input (the right-hand side does not depend on T): it only has the B and E sums.
template <bool EXACT, bool DENSE>
PRT_HD void prt_rk_try(const prt_envp& E, double rtol, double atol, double t, double hh, double T, double z,
                       double p, double f0, double f1, double f2, prt_try& v, bool& ok)
{
    double kT, kz, kp, cc;
    // Depth-cell guess of a stage, real-valued: z_stage * z_inv + z_off.  With PRT_FOLD_GUESS it is formed as ONE fma
    // after the last stage value the stage input waits for -- (hh z_inv A_last) k_last + [z z_inv + z_off + hh z_inv
    // (rest of the stage sum)], the bracket being ready long before -- instead of three dependent operations (stage
    // sum, stage input, guess): the guess differs from the literal one by rounding only, and any guess is a correct
    // guess (the weight proves it), so results do not change; the FP64 chain of every stage gets 2 operations shorter.
#ifndef PRT_FOLD_GUESS
#define PRT_FOLD_GUESS 0
#endif
    const double hz = hh * E.v.z_inv, zz = PRT_FMA(z, E.v.z_inv, E.v.z_off);
#define PRT_GUESS(zs, a_last, k_last, a_part) \
    ((PRT_FOLD_GUESS && !EXACT) ? PRT_FMA(hz * (a_last), (k_last), PRT_FMA(hz, (a_part), zz)) : PRT_FMA((zs), E.v.z_inv, E.v.z_off))
#define PRT_STAGE(xs, zs, ps, sg, cout)                                          \
    do {                                                                         \
        if (EXACT) {                                                             \
            double o4_[4];                                                       \
            prt_rhs_cold(&E.v, E.grid, E.zm, (xs), (zs), (ps), o4_);             \
            kT = o4_[0]; kz = o4_[1]; kp = o4_[2]; cout = o4_[3];                \
        } else {                                                                 \
            const prt_xcell xc_ = prt_xlook(E, (xs), ok);                        \
            prt_rhs_spec(E, xc_, (zs), (ps), (sg), kT, kz, kp, cout, ok);        \
        }                                                                        \
    } while (0)
    // stage 2 (its travel-time component enters no sum: B2 = E2 = P2 = 0)
    double k2z, k2p, k3T, k3z, k3p, k4T, k4z, k4p, a, b, ap, zs_;
    zs_ = PRT_FMA(hh, A21 * f1, z);
    PRT_STAGE(PRT_FMA(C2, hh, t), zs_, PRT_FMA(hh, A21 * f2, p), PRT_FMA(zs_, E.v.z_inv, E.v.z_off), cc);
    k2z = kz; k2p = kp;
    ap = A31 * f1; a = PRT_FMA(A32, k2z, ap);
    b = A31 * f2; b = PRT_FMA(A32, k2p, b);
    zs_ = PRT_FMA(hh, a, z);
    PRT_STAGE(PRT_FMA(C3, hh, t), zs_, PRT_FMA(hh, b, p), PRT_GUESS(zs_, A32, k2z, ap), cc);
    k3T = kT; k3z = kz; k3p = kp;
    ap = A41 * f1; ap = PRT_FMA(A42, k2z, ap); a = PRT_FMA(A43, k3z, ap);
    b = A41 * f2; b = PRT_FMA(A42, k2p, b); b = PRT_FMA(A43, k3p, b);
    zs_ = PRT_FMA(hh, a, z);
    PRT_STAGE(PRT_FMA(C4, hh, t), zs_, PRT_FMA(hh, b, p), PRT_GUESS(zs_, A43, k3z, ap), cc);
    k4T = kT; k4z = kz; k4p = kp;
    ap = A51 * f1; ap = PRT_FMA(A52, k2z, ap); ap = PRT_FMA(A53, k3z, ap); a = PRT_FMA(A54, k4z, ap);
    b = A51 * f2; b = PRT_FMA(A52, k2p, b); b = PRT_FMA(A53, k3p, b); b = PRT_FMA(A54, k4p, b);
    zs_ = PRT_FMA(hh, a, z);
    PRT_STAGE(PRT_FMA(C5, hh, t), zs_, PRT_FMA(hh, b, p), PRT_GUESS(zs_, A54, k4z, ap), cc);
    // from here on the sums are carried, not the stage values
    double a6z = A61 * f1; a6z = PRT_FMA(A62, k2z, a6z); a6z = PRT_FMA(A63, k3z, a6z); a6z = PRT_FMA(A64, k4z, a6z);
    const double sg6 = PRT_GUESS(PRT_FMA(hh, PRT_FMA(A65, kz, a6z), z), A65, kz, a6z);
    a6z = PRT_FMA(A65, kz, a6z);
    double a6p = A61 * f2; a6p = PRT_FMA(A62, k2p, a6p); a6p = PRT_FMA(A63, k3p, a6p); a6p = PRT_FMA(A64, k4p, a6p); a6p = PRT_FMA(A65, kp, a6p);
    double sT = B1 * f0; sT = PRT_FMA(B3, k3T, sT); sT = PRT_FMA(B4, k4T, sT); sT = PRT_FMA(B5, kT, sT);
    double sz = B1 * f1; sz = PRT_FMA(B3, k3z, sz); sz = PRT_FMA(B4, k4z, sz); sz = PRT_FMA(B5, kz, sz);
    double sp = B1 * f2; sp = PRT_FMA(B3, k3p, sp); sp = PRT_FMA(B4, k4p, sp); sp = PRT_FMA(B5, kp, sp);
    double eT = E1 * f0; eT = PRT_FMA(E3, k3T, eT); eT = PRT_FMA(E4, k4T, eT); eT = PRT_FMA(E5, kT, eT);
    double ez = E1 * f1; ez = PRT_FMA(E3, k3z, ez); ez = PRT_FMA(E4, k4z, ez); ez = PRT_FMA(E5, kz, ez);
    double ep = E1 * f2; ep = PRT_FMA(E3, k3p, ep); ep = PRT_FMA(E4, k4p, ep); ep = PRT_FMA(E5, kp, ep);
    prt_q3 qT, qz, qp;
    if (DENSE) {
#define PRT_Q5(q, k1, k3, k4, k5)                                                                       \
    q.q1 = P11 * k1; q.q1 = PRT_FMA(P31, k3, q.q1); q.q1 = PRT_FMA(P41, k4, q.q1); q.q1 = PRT_FMA(P51, k5, q.q1); \
    q.q2 = P12 * k1; q.q2 = PRT_FMA(P32, k3, q.q2); q.q2 = PRT_FMA(P42, k4, q.q2); q.q2 = PRT_FMA(P52, k5, q.q2); \
    q.q3 = P13 * k1; q.q3 = PRT_FMA(P33, k3, q.q3); q.q3 = PRT_FMA(P43, k4, q.q3); q.q3 = PRT_FMA(P53, k5, q.q3)
        PRT_Q5(qT, f0, k3T, k4T, kT);
        PRT_Q5(qz, f1, k3z, k4z, kz);
        PRT_Q5(qp, f2, k3p, k4p, kp);
#undef PRT_Q5
    }
    // stages 6 and 7 (FSAL) share the abscissa t + h, hence the range cell and its weight
    const double x6 = t + hh;
    prt_xcell xc6;
    if (!EXACT) xc6 = prt_xlook(E, x6, ok);
    if (EXACT) {
        double o4[4];
        prt_rhs_cold(&E.v, E.grid, E.zm, x6, PRT_FMA(hh, a6z, z), PRT_FMA(hh, a6p, p), o4);
        kT = o4[0]; kz = o4[1]; kp = o4[2]; cc = o4[3];
    } else prt_rhs_spec(E, xc6, PRT_FMA(hh, a6z, z), PRT_FMA(hh, a6p, p), sg6, kT, kz, kp, cc, ok);
    const double sg7 = PRT_GUESS(PRT_FMA(hh, PRT_FMA(B6, kz, sz), z), B6, kz, sz);
    sT = PRT_FMA(B6, kT, sT); sz = PRT_FMA(B6, kz, sz); sp = PRT_FMA(B6, kp, sp);
    eT = PRT_FMA(E6, kT, eT); ez = PRT_FMA(E6, kz, ez); ep = PRT_FMA(E6, kp, ep);
    if (DENSE) {
        qT.q1 = PRT_FMA(P61, kT, qT.q1); qT.q2 = PRT_FMA(P62, kT, qT.q2); qT.q3 = PRT_FMA(P63, kT, qT.q3);
        qz.q1 = PRT_FMA(P61, kz, qz.q1); qz.q2 = PRT_FMA(P62, kz, qz.q2); qz.q3 = PRT_FMA(P63, kz, qz.q3);
        qp.q1 = PRT_FMA(P61, kp, qp.q1); qp.q2 = PRT_FMA(P62, kp, qp.q2); qp.q3 = PRT_FMA(P63, kp, qp.q3);
    }
    v.nT = PRT_FMA(hh, sT, T); v.nz = PRT_FMA(hh, sz, z); v.np_ = PRT_FMA(hh, sp, p);
    // error scales and their reciprocals (rk.py:148-152): known before the last stage, so the three
    // divisions of the reference become one reciprocal (of the product) formed in the shadow of that stage
    const double scT = PRT_FMA(prt_max_abs(T, v.nT), rtol, atol);
    const double scz = PRT_FMA(prt_max_abs(z, v.nz), rtol, atol);
    const double scp = PRT_FMA(prt_max_abs(p, v.np_), rtol, atol);
    const double sTz = scT * scz, szp = scz * scp, sTp = scT * scp;
    const double inv3 = prt_rcp(sTz * scp);
    const double iT = inv3 * szp, iz = inv3 * sTp, ip = inv3 * sTz;
    if (EXACT) {  // f_new (FSAL)
        double o4[4];
        prt_rhs_cold(&E.v, E.grid, E.zm, x6, v.nz, v.np_, o4);
        v.k7T = o4[0]; v.k7z = o4[1]; v.k7p = o4[2]; v.c_new = o4[3];
    } else prt_rhs_spec(E, xc6, v.nz, v.np_, sg7, v.k7T, v.k7z, v.k7p, v.c_new, ok);
#undef PRT_STAGE
#undef PRT_GUESS
    // the event functions at the step end (ivp.py:678-699) only need y_new and c_new: formed here, they overlap
    // the tail of the last stage and the controller instead of following the accept decision
    {
        const double pc = v.np_ * v.c_new;
        const double bd = EXACT ? prt_bathy(E, x6) : prt_bathy_spec(E, x6, ok);
        int m = 0;
        if ((v.nz < 0.0) & (pc < 0.0) & (pc >= -1.0)) m |= PRT_EV_SURFACE;
        if ((pc > 0.0) & (pc <= 1.0) & (v.nz > bd)) m |= PRT_EV_BOTTOM;
        if ((v.nz > E.v.bb_zmax) | (v.nz < E.v.bb_zmin) | (x6 < E.v.bb_rmin) | (x6 > E.v.bb_rmax)) m |= PRT_EV_BBOX;
        v.g_new = m;
        v.vert_open = (fabs(pc) > 0.99999999) & (fabs(pc) <= 1.0);
    }
    eT = PRT_FMA(E7, v.k7T, eT); ez = PRT_FMA(E7, v.k7z, ez); ep = PRT_FMA(E7, v.k7p, ep);
    // error_norm carried as its square (mean square of the scaled error): accept iff ms < 1 and
    // error_norm ** (-1/5) = ms ** (-1/10)  -- no square root, no division by sqrt(3) (rk.py:148-165)
    const double r0 = (eT * hh) * iT, r1 = (ez * hh) * iz, r2 = (ep * hh) * ip;
    double ms = r0 * r0;
    ms = PRT_FMA(r1, r1, ms);
    ms = PRT_FMA(r2, r2, ms);
    ms = ms * 0x1.5555555555555p-2;
    v.ms = ms;
    v.pw = prt_pow_m01(ms);
    if (DENSE) {
        qT.q1 = PRT_FMA(P71, v.k7T, qT.q1); qT.q2 = PRT_FMA(P72, v.k7T, qT.q2); qT.q3 = PRT_FMA(P73, v.k7T, qT.q3);
        qz.q1 = PRT_FMA(P71, v.k7z, qz.q1); qz.q2 = PRT_FMA(P72, v.k7z, qz.q2); qz.q3 = PRT_FMA(P73, v.k7z, qz.q3);
        qp.q1 = PRT_FMA(P71, v.k7p, qp.q1); qp.q2 = PRT_FMA(P72, v.k7p, qp.q2); qp.q3 = PRT_FMA(P73, v.k7p, qp.q3);
        v.qT = qT; v.qz = qz; v.qp = qp;
        v.inv_h = prt_rcp(hh);  // contract: dense-output abscissa (t - t_old) * rcp(h), rk.py:722
    }
}

#if defined(PRT_COUNT_COLD) && defined(__CUDACC__)
__device__ unsigned long long prt_dbg_cold;
#endif
// cold path of prt_ray::attempt: the same trial step through the general cell lookup
template <bool DENSE>
PRT_HD_NOINLINE void prt_rk_try_exact(const prt_env_view* V, const prt_d2* grid, double rtol, double atol, double t,
                                      double hh, double T, double z, double p, double f0, double f1, double f2,
                                      prt_try* out)
{
    const prt_envp E{*V, grid, prt_z_mode(*V)};
    bool ok = true;
#if defined(PRT_COUNT_COLD) && !defined(__CUDACC__)
    extern long long prt_cold_count;
    ++prt_cold_count;
#elif defined(PRT_COUNT_COLD) && defined(__CUDA_ARCH__)
    atomicAdd(&prt_dbg_cold, 1ULL);
#endif
    prt_rk_try<true, DENSE>(E, rtol, atol, t, hh, T, z, p, f0, f1, f2, *out, ok);
}

// ---------------------------------------------------------------------------------------------
// One ray as a small state machine.  The same per-ray arithmetic is driven either ray by ray
// (prt_integrate_ray, used for host-side parity tests) or warp-synchronously (prt_integrate_warp
// in prt_kernels.cu), where lanes always execute step attempts together and lanes that hit a
// boundary wait for one another so that the event root searches also run together.
enum { PRT_PH_INIT = 0, PRT_PH_STEP = 1, PRT_PH_ROOT = 2, PRT_PH_DONE = 3 };

// The last accepted step of a ray: produced by attempt(), consumed in the same loop iteration by
// advance() -- or, for a lane that fired an event, parked (in shared memory on the device) until the
// warp's root searches run.  Keeping it out of the persistent ray state is what lets 4+ blocks of 128
// threads be resident per SM: it is not live across the stage evaluations of the next attempts.
struct prt_step {
    double t_old, t_new, h, inv_h, yT, yz, yp, nf0, nf1, nf2;
    prt_q3 qT, qz, qp;
    double root;       // set by solve_roots
    int32_t active;    // mask of events that fired on this step
    int32_t event;     // earliest event (set by solve_roots)
    bool finished;     // the step reached x_end
};

// COUNT = false drops the work counters (5 registers the arrival-table kernels cannot spare at 5 blocks per SM)
template <bool SAVE, bool COUNT = true>
struct prt_ray {
    // persistent ODE state (solver.t, solver.y, solver.f, solver.h_abs)
    double t, T, z, p, f0, f1, f2, h_abs;
    double rs_cur;  // SAVE: range_save[cursor], loaded when the cursor moves (never on the path of a step decision)
    int32_t g, n_ts, cursor, n_surface, n_bottom, status, phase;
    bool rejected;
    uint32_t c_att, c_acc, c_seg, c_bis, c_dense;  // RHS evaluations = 2 c_seg + 6 c_att

    PRT_HD void start(double T0, double z0, double p0, double x0, const prt_shot& S)
    {
        t = x0; T = T0; z = z0; p = p0;
        n_surface = n_bottom = 0;
        status = PRT_RAY_OK;
        cursor = 0;
        rs_cur = (SAVE && S.n_save > 0) ? prt_ld(S.range_save) : 0.0;
        c_att = c_acc = c_seg = c_bis = c_dense = 0;
        rejected = false;
        phase = (x0 < S.x_end) ? PRT_PH_INIT : PRT_PH_DONE;  // launch_rays.py:404
    }

    // RK45.__init__ (rk.py:85-103) + select_initial_step (common.py:68-134) + g = [st.event(t0, y0)]
    PRT_HD void seg_init(const prt_envp& E, const prt_shot& S)
    {
        if (COUNT) ++c_seg;
        double c_here;
        prt_rhs(E, t, z, p, f0, f1, f2, c_here);
        g = prt_events(E, t, z, p, c_here);
        double interval = fabs(S.x_end - t);
        double s0 = PRT_FMA(fabs(T), S.rtol, S.atol), s1 = PRT_FMA(fabs(z), S.rtol, S.atol),
               s2 = PRT_FMA(fabs(p), S.rtol, S.atol);
        double d0 = prt_rms3(T / s0, z / s1, p / s2);
        double d1 = prt_rms3(f0 / s0, f1 / s1, f2 / s2);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (0.01 * d0) / d1;
        h0 = prt_py_min(h0, interval);
        double g0, g1, g2, cc;
        prt_rhs(E, t + h0, PRT_FMA(h0, f1, z), PRT_FMA(h0, f2, p), g0, g1, g2, cc);
        double d2 = prt_rms3((g0 - f0) / s0, (g1 - f1) / s1, (g2 - f2) / s2) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = prt_py_max(1e-6, h0 * 1e-3);
        else h1 = 1.0 / prt_pow_m02(0.01 / prt_py_max(d1, d2));
        h_abs = prt_py_min(prt_py_min(100.0 * h0, h1), interval);
        n_ts = 1;
        rejected = false;
        phase = PRT_PH_STEP;
    }

    // One attempt of _step_impl (rk.py:111-176).  Returns true when the step was accepted; then the
    // step record (st.t_old .. st.nf2, events, dense output if needed) is filled in.  `lanes` = the lanes
    // of the warp that make this attempt together (device only): the prologue is select-only and the lanes
    // are re-converged explicitly before the trial step, so that its ~700 instructions are never executed
    // once per divergent sub-group.
    // ROWS = false (warp driver of the arrival-table kernels): the dense-output rows of a step that fired an
    // event are left to the root phase, which rebuilds the whole step record anyway (prt_park_slot::rebuild).
    template <bool ROWS = true>
    PRT_HD bool attempt(const prt_envp& E, const prt_shot& S, prt_step& st, unsigned lanes = 0xffffffffu)
    {
        const double min_step = 10.0 * fabs(prt_next_up(t) - t);
        // TOO_SMALL_STEP -> solve_ivp status -1 (only after a rejection; before the first trial of a step
        // h_abs is raised to min_step instead, rk.py:118-124)
        const bool too_small = rejected & (h_abs < min_step);
        const double h_try = (h_abs < min_step) ? min_step : h_abs;
        double tn = t + h_try;
        tn = (tn - S.x_end > 0.0) ? S.x_end : tn;
        const double hh = tn - t;
        PRT_CONVERGE(lanes);
        if (too_small) {
            status = PRT_RAY_STEP_TOO_SMALL;
            phase = PRT_PH_DONE;
            return false;
        }
        if (COUNT) ++c_att;
        h_abs = fabs(hh);
        // the trial step: speculative cell lookups, one branch-free block; re-run through the general lookup
        // if any of its seven guesses missed (same arithmetic, so the same bits either way)
        prt_try v;
        bool ok = (E.v.spec_off == 0);  // irregular grids: straight to the exact lookup
        if (ok) prt_rk_try<false, SAVE>(E, S.rtol, S.atol, t, hh, T, z, p, f0, f1, f2, v, ok);
        if (!ok) {
            prt_try cold;
            prt_rk_try_exact<SAVE>(&E.v, E.grid, S.rtol, S.atol, t, hh, T, z, p, f0, f1, f2, &cold);
            v = cold;
        }
        // controller (rk.py:153-165); one power serves accepting and rejecting lanes.  ms == 0 needs no special
        // case: prt_pow_m01(0) >= 119, so the factor is MAX_FACTOR as in the reference
        if (!(v.ms < 1.0)) {
            h_abs *= prt_py_max(0.2, 0.9 * v.pw);
            rejected = true;
            return false;
        }
        double factor = prt_py_min(10.0, 0.9 * v.pw);
        if (rejected) factor = prt_py_min(1.0, factor);
        h_abs *= factor;
        rejected = false;
        if (COUNT) ++c_acc;
        const double nT = v.nT, nz = v.nz, np_ = v.np_;
        const double k7T = v.k7T, k7z = v.k7z, k7p = v.k7p;
        double c_new = v.c_new;
        st.t_old = t; st.t_new = tn; st.h = hh;
        st.yT = nT; st.yz = nz; st.yp = np_;
        st.nf0 = k7T; st.nf1 = k7z; st.nf2 = k7p;
        st.finished = (tn - S.x_end >= 0.0);  // base.py:205-206
        // events at the step end (ivp.py:678-699) were formed by the trial step from c(t + h, y_new) of the FSAL
        // stage: identical to the event functions at st.t_new unless the clipped step made t + h != st.t_new
        int g_new = v.g_new;
        if (v.vert_open && prt_ev_vertical(np_ * c_new)) g_new |= PRT_EV_VERTICAL;
        if (t + hh != tn) {
            double cz_unused;
            prt_field(E, tn, nz, c_new, cz_unused);
            g_new = prt_events(E, tn, nz, np_, c_new);
        }
        // find_active_events (ivp.py:134-157): up for direction +1, up|down for direction 0
        const int up = (~g) & g_new, down = g & (~g_new);
        st.active = (up & (PRT_EV_SURFACE | PRT_EV_BOTTOM)) | ((up | down) & (PRT_EV_VERTICAL | PRT_EV_BBOX));
        g = g_new;
        bool need_dense = (st.active != 0);
        if (SAVE && !need_dense) need_dense = (cursor < S.n_save - 1) && (rs_cur <= tn);
        if (COUNT && need_dense) ++c_dense;
        if (SAVE) {  // formed by the trial step
            st.qT = v.qT; st.qz = v.qz; st.qp = v.qp;
            st.inv_h = v.inv_h;
        } else if (need_dense && ROWS) {
            // arrival-table mode keeps no stage values: the rows of the rare step that needs them (an event fired)
            // come from repeating the trial step with DENSE -- same inputs, same arithmetic, same bits
            prt_try d;
            prt_rk_try_exact<true>(&E.v, E.grid, S.rtol, S.atol, t, hh, T, z, p, f0, f1, f2, &d);
            st.qT = d.qT; st.qz = d.qz; st.qp = d.qp;
            st.inv_h = d.inv_h;
        }
        return true;
    }

    PRT_HD void emit(const prt_traj_sink& sink, const prt_step& st, int i, double rsv) const
    {
        const double xx = (rsv - st.t_old) * st.inv_h;
        const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
        sink.put(i, prt_dense_eval(f0, st.qT, st.h, T, p1, p2, p3, p4), prt_dense_eval(f1, st.qz, st.h, z, p1, p2, p3, p4),
                 prt_dense_eval(f2, st.qp, st.h, p, p1, p2, p3, p4));
    }

    // accepted step without an st.event: stream the save points it owns, then move on
    PRT_HD void advance(const prt_shot& S, const prt_traj_sink& sink, const prt_step& st)
    {
        if (SAVE) {
            const int ub = S.n_save - 1;
            while ((cursor < ub) && (rs_cur <= st.t_new)) {
                emit(sink, st, cursor, rs_cur);
                ++cursor;
                rs_cur = prt_ld(S.range_save + cursor);  // cursor <= n_save - 1
            }
        }
        T = st.yT; z = st.yz; p = st.yp;
        f0 = st.nf0; f1 = st.nf1; f2 = st.nf2;
        t = st.t_new;
        ++n_ts;
        if (st.finished) phase = PRT_PH_DONE;  // sol.status == 0, launch_rays.py:425-426
    }

    PRT_HD bool event_at(const prt_envp& E, const prt_step& st, int e, double x) const
    {
        const double xx = (x - st.t_old) * st.inv_h;
        const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
        return prt_event_one(E, e, x, prt_dense_eval(f1, st.qz, st.h, z, p1, p2, p3, p4),
                             prt_dense_eval(f2, st.qp, st.h, p, p1, p2, p3, p4));
    }

    // handle_events (ivp.py:80-131): every st.event is terminal -> earliest st.root, ties to the lower
    // st.event index.  solve_event_equation = brentq(xtol = rtol = 4 eps, maxiter 100) on a +-1
    // function, which degenerates to bisection (|f| is always 1).
    PRT_HD void solve_roots(const prt_envp& E, prt_step& st)
    {
        double best_root = 0.0;
        int best = -1;
        for (int e = 0; e < 4; ++e) {
            if (!(st.active & (1 << e))) continue;
            double xpre = st.t_old, xcur = st.t_new, xblk = 0.0;
            bool fpre = event_at(E, st, e, st.t_old), fcur = event_at(E, st, e, st.t_new);
            if (COUNT) c_bis += 2;
            if (fpre == fcur) {
                status = PRT_RAY_BRENTQ_SIGN;
                return;
            }
            bool converged = false;
            for (int i = 0; i < 100; ++i) {
                if (fpre != fcur) xblk = xpre;
                const double delta = (4 * PRT_EPS + 4 * PRT_EPS * fabs(xcur)) / 2;
                const double sbis = (xblk - xcur) / 2;
                if (fabs(sbis) < delta) {
                    converged = true;
                    break;
                }
                xpre = xcur;
                fpre = fcur;
                if (fabs(sbis) > delta) xcur += sbis;
                else xcur += (sbis > 0 ? delta : -delta);
                fcur = event_at(E, st, e, xcur);
                if (COUNT) ++c_bis;
            }
            if (!converged) {
                status = PRT_RAY_BRENTQ_ITER;
                return;
            }
            if (best < 0 || xcur < best_root) {
                best = e;
                best_root = xcur;
            }
        }
        st.event = best;
        st.root = best_root;
    }

    // terminal step: _interpolate_ray bookkeeping for the st.finished segment (launch_rays.py:745-784),
    // state at the st.event, then the reflection (launch_rays.py:432-480)
    PRT_HD void finish_terminal(const prt_envp& E, const prt_shot& S, const prt_traj_sink& sink, const prt_step& st)
    {
        if (status != PRT_RAY_OK) {
            phase = PRT_PH_DONE;
            return;
        }
        const bool popped = (n_ts > 1) && (st.root == st.t_old);  // donot_append, ivp.py:701-710
        if (SAVE) {
            const int start = cursor > 0 ? cursor - 1 : 0;
            const int ub = prt_argmin_from(S.range_save, S.n_save, start, st.root);  // idx2
            if (!popped) {
                while (cursor < ub) {
                    emit(sink, st, cursor, prt_ld(S.range_save + cursor));
                    ++cursor;
                }
            }
            cursor = ub;  // the next segment owns indices >= idx2
            rs_cur = prt_ld(S.range_save + cursor);
        }
        if (!popped) {
            const double xx = (st.root - st.t_old) * st.inv_h;
            const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
            const double nT = prt_dense_eval(f0, st.qT, st.h, T, p1, p2, p3, p4);
            const double nz = prt_dense_eval(f1, st.qz, st.h, z, p1, p2, p3, p4);
            const double np_ = prt_dense_eval(f2, st.qp, st.h, p, p1, p2, p3, p4);
            T = nT; z = nz; p = np_;
        }  // else: sol.y[:, -1] stays the state at st.t_old
        t = st.root;
        phase = PRT_PH_DONE;
        if (st.event == 2) { status = PRT_RAY_VERTICAL; return; }
        if (st.event == 3) { status = PRT_RAY_BBOX; return; }
        double c, cz;
        prt_field(E, t, z, c, cz);
        const double theta = prt_asin(p * c) * PRT_RAD2DEG;  // ray_angle
        double theta_bounce;
        if (st.event == 0) {
            theta_bounce = -theta;
            ++n_surface;
        } else {
            bool oob;
            const double beta = prt_bottom_angle(E.v, t, oob);
            if (oob) { status = PRT_RAY_BOTTOM_OOB; return; }
            theta_bounce = 2.0 * beta - theta;
            ++n_bottom;
        }
        if (S.terminate_backwards && (fabs(theta_bounce) > 90.0)) { status = PRT_RAY_BACKWARDS; return; }
        p = prt_sin(theta_bounce * PRT_DEG2RAD) / c;
        if (n_surface + n_bottom >= S.max_bounces) { status = PRT_RAY_MAX_BOUNCES; return; }
        if (t < S.x_end) phase = PRT_PH_INIT;  // launch_rays.py:404
    }

    PRT_HD void result(const prt_shot& S, const prt_traj_sink& sink, prt_ray_result& res, uint32_t* cnt) const
    {
        if (status == PRT_RAY_OK) {
            if (SAVE && S.n_save > 0) sink.put(S.n_save - 1, T, z, p);  // exact final state, launch_rays.py:772-775
            res.T = T; res.z = z; res.p = p;
        } else {
            res.T = NAN; res.z = NAN; res.p = NAN;
        }
        res.n_bottom = n_bottom;
        res.n_surface = n_surface;
        res.status = status;
        if (COUNT && cnt) {
            cnt[PRT_CNT_NFEV] = 2u * c_seg + 6u * c_att; cnt[PRT_CNT_ATTEMPTS] = c_att; cnt[PRT_CNT_ACCEPTED] = c_acc;
            cnt[PRT_CNT_SEGMENTS] = c_seg; cnt[PRT_CNT_BISECT] = c_bis; cnt[PRT_CNT_DENSE] = c_dense;
        }
    }
};

// ray-by-ray driver (host parity harness; also the semantic definition of the warp driver)
template <bool SAVE>
PRT_HD void prt_integrate_ray(const prt_envp& E, double T, double z, double p, double x0, prt_shot S,
                              const prt_traj_sink& sink, prt_ray_result& res, uint32_t* cnt)
{
    if (S.rtol < 100.0 * PRT_EPS) S.rtol = 100.0 * PRT_EPS;  // validate_tol, common.py:44-60
    prt_ray<SAVE> r;
    r.start(T, z, p, x0, S);
    while (r.phase != PRT_PH_DONE) {
        if (r.phase == PRT_PH_INIT) r.seg_init(E, S);
        prt_step st;
        if (!r.attempt(E, S, st)) continue;
        if (st.active) {
            r.solve_roots(E, st);
            r.finish_terminal(E, S, sink, st);
        } else {
            r.advance(S, sink, st);
        }
    }
    r.result(S, sink, res, cnt);
}
