/*
 * prt_oracle.c -- CPU restatement of pygenray's ray-fan shooting + eigenray search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pygenray_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / CPU baseline.
 *
 * Plain C99, scalar, written to mirror the STRUCTURE of the reference (a solver object that is
 * stepped, per-step dense-output objects collected per segment, brentq called with a callback,
 * literal event functions, _interpolate_ray run after the fact over the stored segments), not the
 * structure of the CUDA kernel (which streams, fuses and shares sub-expressions).  Each function
 * cites the reference lines it restates.  Paths: "ref:" = /root/reference/src/pygenray/,
 * "scipy:" = scipy 1.18.1 scipy/integrate/_ivp/ (un-pinned third-party dependency of the
 * reference, pyproject.toml:27-34; the arithmetic core of the path lives there).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file against
 *   (a) the reference's own golden fixture tests/fixtures/munk_regression.npz (copied numbers
 *       live in tests/golden/), and
 *   (b) outputs of the unmodified reference run in the build container
 *       (oracle/make_golden.py -> tests/golden/ref_*.npz).
 *
 * Rounding contract ("PRT arithmetic"): IEEE-754 binary64, round-to-nearest-even, every
 * operation written below is one correctly rounded operation; fma() is used exactly where
 * written and nowhere else (compile with -ffp-contract=off); +,-,*,/,sqrt,fma only, plus the
 * elementary functions orc_pow_m02 / orc_pow_m01 / orc_rsqrt / orc_sin / orc_asin defined in this file from those
 * operations; three divisions of the reference are written as multiplications by a correctly
 * rounded reciprocal (interpolation weights, dense-output abscissa, fact/c^2 in derivsrd: marked
 * "contract:" below), as LLVM fast-math is free to do for the reference's own Numba code.
 * The reference evaluates the same expressions through NumPy/BLAS dot products and
 * LLVM fast-math code whose association is not specified, so agreement
 * with the reference is to tolerance, while agreement of the CUDA path with this file is
 * bit-for-bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* status codes (shared vocabulary with include/prt.h; values restated, not included)          */
enum {
    ORC_OK = 0,
    ORC_STEP_TOO_SMALL = 1,  /* solve_ivp status -1            ref:launch_rays.py:427-430 */
    ORC_VERTICAL = 2,        /* vertical_ray event             ref:launch_rays.py:443-448 */
    ORC_BBOX = 3,            /* ray_bounding_box_event         ref:launch_rays.py:451-456 */
    ORC_BACKWARDS = 4,       /* |theta_bounce| > 90            ref:launch_rays.py:474-477 */
    ORC_BOTTOM_OOB = 5,      /* interp1d ValueError out of bathymetry range (launch_rays.py:469) */
    ORC_BRENTQ_SIGN = 6,     /* brentq ValueError: f(a), f(b) same sign */
    ORC_BRENTQ_ITER = 7,     /* brentq RuntimeError: no convergence in 100 iterations */
    ORC_TS_NOT_INCREASING = 8, /* OdeSolution ValueError (scipy:common.py:178-181) */
    ORC_MAX_BOUNCES = 9      /* safety cap of this restatement (not in the reference) */
};

typedef struct {
    int32_t n_r, n_z, n_b, n_ba;
    const double *rin;          /* [n_r]                       */
    const double *zin;          /* [n_z]                       */
    const double *cin;          /* [n_r][n_z]                  */
    const double *cpin;         /* [n_r][n_z]                  */
    const double *depth_ranges; /* [n_b]                       */
    const double *depths;       /* [n_b]                       */
    const double *ba_x;         /* [n_ba] spline breakpoints   */
    const double *ba_c;         /* [4][n_ba-1] PPoly coeffs, highest power first */
} orc_env;

enum { CNT_NFEV = 0, CNT_ATTEMPTS, CNT_ACCEPTED, CNT_SEGMENTS, CNT_BISECT, CNT_DENSE, CNT_N };

/* ------------------------------------------------------------------------------------------ */
/* Python builtin min/max semantics: return the first unless the later one compares better.    */
static double py_min(double a, double b) { return (b < a) ? b : a; }
static double py_max(double a, double b) { return (b > a) ? b : a; }
/* numpy.maximum: propagates NaN */
static double np_maximum(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

/* ------------------------------------------------------------------------------------------ */
/* Deterministic elementary functions                                                           */

/* x^(-1/5) for x > 0 (normal or subnormal); replaces libm pow(error_norm, -0.2) of
 * scipy:rk.py:155,163.  <= 2 ulp.  inf -> 0, NaN -> NaN. */
ORC_API double orc_pow_m02(double x)
{
    static const double T[5] = {1.0, 0x1.bdb8cdadbe120p-1, 0x1.8406003b2ae5cp-1,
                                0x1.51cb453b9536cp-1, 0x1.2611186bae674p-1}; /* 2^(-r/5) */
    if (x != x) return x;
    if (x == INFINITY) return 0.0;
    double post = 1.0;
    if (x < 0x1p-1022) { x = x * 0x1p100; post = 0x1p20; }
    uint64_t b;
    memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    int q = (e >= 0) ? (e / 5) : -((4 - e) / 5);
    int r = e - 5 * q;
    double y = fma(fma(0.05264997, m, -0.28570902), m, 1.23209767);
    for (int it = 0; it < 3; ++it) {
        double y2 = y * y;
        double y4 = y2 * y2;
        double y5 = y4 * y;
        double res = fma(-m, y5, 1.0);
        y = fma(y * 0.2, res, y);
    }
    y = y * T[r];
    uint64_t pb = (uint64_t)(1023 - q) << 52;
    double p2;
    memcpy(&p2, &pb, 8);
    return (y * p2) * post;
}

/* 1/sqrt(w) for normal w > 0: integer-shift seed on the high word (3.4e-2) + 4 Newton steps
 * y <- y (1.5 - 0.5 w y^2); <= 2.3 ulp.  contract: stands for the 1/sqrt of
 * ref:integration_processes.py:95 (see derivsrd). */
ORC_API double orc_rsqrt(double w)
{
    uint64_t b;
    memcpy(&b, &w, 8);
    b = (uint64_t)(uint32_t)(0x5FE6EB50u - ((uint32_t)(b >> 32) >> 1)) << 32;
    double y;
    memcpy(&y, &b, 8);
    const double hw = 0.5 * w;
    for (int it = 0; it < 4; ++it) {
        double t = y * y;
        double r = fma(-hw, t, 1.5);
        y = y * r;
    }
    return y;
}

/* 1/x for normal x > 0: integer seed on the high word (5e-2) + 4 Newton steps y <- y + y (1 - x y);
 * <= 1 ulp.  contract: stands for the divisions by the error scale (scipy:rk.py:151) and by the step
 * length (scipy:rk.py:722), whose divisors are known before their dividends. */
ORC_API double orc_rcp(double x)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    b = (uint64_t)(uint32_t)(0x7FDE6238u - (uint32_t)(b >> 32)) << 32;
    double y;
    memcpy(&y, &b, 8);
    for (int it = 0; it < 4; ++it) {
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}

/* x^(-1/10): the step controller's error_norm ** (-1/5) expressed on the MEAN SQUARE
 * x = error_norm^2 (no square root); exponent split by 10 + 3 Newton steps; <= 2 ulp for
 * 2^-70 <= x < 2^70.  contract: the biased exponent is clamped to [953, 1092]; outside that range the
 * controller's factor 0.9 x^(-1/10) is pinned at MAX_FACTOR (x -> 0) or MIN_FACTOR (x -> inf, NaN)
 * whatever the exact power, so 0, subnormals, inf and NaN need no special case. */
ORC_API double orc_pow_m01(double x)
{
    static const double T[10] = {1.0, 0x1.ddb680117ab12p-1, 0x1.bdb8cdadbe120p-1, 0x1.9fdf8bcce533ep-1,
                                 0x1.8406003b2ae5cp-1, 0x1.6a09e667f3bcdp-1, 0x1.51cb453b9536cp-1,
                                 0x1.3b2c47bff8329p-1, 0x1.2611186bae674p-1, 0x1.125fbee250664p-1}; /* 2^(-r/10) */
    uint64_t b;
    memcpy(&b, &x, 8);
    int ef = (int)((b >> 52) & 0x7ff);
    if (ef < 953) ef = 953;
    if (ef > 1092) ef = 1092;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    unsigned eu = (unsigned)(ef + 7); /* e + 1030, e = ef - 1023 = 10 q + r */
    unsigned q10 = eu / 10u;          /* q + 103 */
    int r = (int)(eu - 10u * q10);
    double y = fma(fma(0.024999779689669537, m, -0.14120119176598264), m, 1.1157678629591425);
    for (int it = 0; it < 3; ++it) {
        double y2 = y * y;
        double y4 = y2 * y2;
        double y8 = y4 * y4;
        double y10 = y8 * y2;
        double res = fma(-m, y10, 1.0);
        y = fma(y * 0.1, res, y);
    }
    y = y * T[r];
    uint64_t pb = (uint64_t)(1126u - q10) << 52;
    double p2;
    memcpy(&p2, &pb, 8);
    return y * p2;
}

static double k_sin(double r)
{ /* |r| <= pi/4 (+slack); Taylor through r^17 */
    double z = r * r;
    double s = 1.0 / 355687428096000.0;
    s = fma(s, z, -1.0 / 1307674368000.0);
    s = fma(s, z, 1.0 / 6227020800.0);
    s = fma(s, z, -1.0 / 39916800.0);
    s = fma(s, z, 1.0 / 362880.0);
    s = fma(s, z, -1.0 / 5040.0);
    s = fma(s, z, 1.0 / 120.0);
    s = fma(s, z, -1.0 / 6.0);
    return fma(r * z, s, r);
}

static double k_cos(double r)
{ /* Taylor through r^16 */
    double z = r * r;
    double s = 1.0 / 20922789888000.0;
    s = fma(s, z, -1.0 / 87178291200.0);
    s = fma(s, z, 1.0 / 479001600.0);
    s = fma(s, z, -1.0 / 3628800.0);
    s = fma(s, z, 1.0 / 40320.0);
    s = fma(s, z, -1.0 / 720.0);
    s = fma(s, z, 1.0 / 24.0);
    s = fma(s, z, -0.5);
    return fma(z, s, 1.0);
}

#define PIO2_HI 0x1.921fb54442d18p+0
#define PIO2_LO 0x1.1a62633145c07p-54
#define TWO_OVER_PI 0.6366197723675814
#define RAD2DEG 57.29577951308232
#define DEG2RAD 0.017453292519943295

/* sin(x), x in radians, |x| < ~1e6; replaces np.sin of ref:launch_rays.py:143,285,480 */
ORC_API double orc_sin(double x)
{
    if (x != x || fabs(x) == INFINITY) return NAN;
    if (fabs(x) <= 0.7853981633974483) return k_sin(x);
    double n = floor(fma(x, TWO_OVER_PI, 0.5));
    double r = fma(-n, PIO2_HI, x);
    r = fma(-n, PIO2_LO, r);
    int k = (int)((long long)n & 3);
    switch (k) {
    case 0: return k_sin(r);
    case 1: return k_cos(r);
    case 2: return -k_sin(r);
    default: return -k_cos(r);
    }
}

static double asin_core(double a)
{ /* 0 <= a <= 0.5: Newton on sin(th) = a from the cubic Taylor guess */
    double th = fma(a * a, a * (1.0 / 6.0), a);
    for (int it = 0; it < 3; ++it) {
        double d = (k_sin(th) - a) / k_cos(th);
        th = th - d;
    }
    return th;
}

/* asin(x); |x| > 1 -> NaN (as np.arcsin); replaces np.arcsin of ref:integration_processes.py:333 */
ORC_API double orc_asin(double x)
{
    if (x != x) return x;
    double a = fabs(x);
    if (a > 1.0) return NAN;
    double t;
    if (a <= 0.5) {
        t = asin_core(a);
    } else {
        double s = sqrt((1.0 - a) * 0.5);
        t = fma(-2.0, asin_core(s), PIO2_HI);
    }
    return (x < 0.0) ? -t : t;
}

/* ------------------------------------------------------------------------------------------ */
/* np.searchsorted(a, v, side="left"): first i with a[i] >= v; NaN sorts last.                  */
static int searchsorted_left(const double *a, int n, double v)
{
    if (v != v) return n;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* ref:integration_processes.py:101-174 bilinear_interp */
static double bilinear_interp(double x, double y, const double *xg, int nx, const double *yg, int ny,
                              const double *values)
{
    int i = searchsorted_left(xg, nx, x) - 1;
    int j = searchsorted_left(yg, ny, y) - 1;
    if (i > nx - 2) i = nx - 2;
    if (i < 0) i = 0;
    if (j > ny - 2) j = ny - 2;
    if (j < 0) j = 0;
    /* contract: the weight division is a multiplication by the correctly rounded cell reciprocal */
    double wx = (x - xg[i]) * (1.0 / (xg[i + 1] - xg[i]));
    double wy = (y - yg[j]) * (1.0 / (yg[j + 1] - yg[j]));
    double v00 = values[(size_t)i * ny + j];
    double v10 = values[(size_t)(i + 1) * ny + j];
    double v01 = values[(size_t)i * ny + j + 1];
    double v11 = values[(size_t)(i + 1) * ny + j + 1];
    /* contract: the four-weight sum (1-wx)(1-wy) v00 + wx (1-wy) v10 + (1-wx) wy v01 + wx wy v11
     * regrouped as two depth lerps and one range lerp */
    double l0 = fma(wy, v01 - v00, v00);
    double l1 = fma(wy, v11 - v10, v10);
    return fma(wx, l1 - l0, l0);
}

/* ref:integration_processes.py:177-235 linear_interp */
static double linear_interp(double x, const double *xin, const double *yin, int n)
{
    int i = searchsorted_left(xin, n, x) - 1;
    if (i > n - 2) i = n - 2;
    if (i < 0) i = 0;
    double w = (x - xin[i]) * (1.0 / (xin[i + 1] - xin[i]));
    return fma(w, yin[i + 1], (1.0 - w) * yin[i]);
}

/* ref:integration_processes.py:306-334 ray_angle -> theta [deg], c */
static double ray_angle(const orc_env *E, double x, const double *y, double *c_out)
{
    double c = bilinear_interp(x, y[1], E->rin, E->n_r, E->zin, E->n_z, E->cin);
    *c_out = c;
    return orc_asin(y[2] * c) * RAD2DEG;
}

/* ref:integration_processes.py:26-98 derivsrd */
static void derivsrd(const orc_env *E, double x, const double *y, double *dydx, int64_t *cnt)
{
    double z = y[1], pz = y[2];
    double c = bilinear_interp(x, z, E->rin, E->n_r, E->zin, E->n_z, E->cin);
    double cp = bilinear_interp(x, z, E->rin, E->n_r, E->zin, E->n_z, E->cpin);
    double c2 = c * c;
    double arg = fma(-c2, pz * pz, 1.0);
    if (arg <= 0.0) arg = 1e-30;
    /* contract: fact / c^2 is formed once as u = 1 / (c^2 sqrt(arg)) = rsqrt(c^4 arg); then
     * [fact/c, c*pz*fact, -fact*cp/c^2] = [c u, (c^3 pz) u, -cp u] */
    double u = orc_rsqrt((c2 * c2) * arg);
    dydx[0] = c * u;
    dydx[1] = ((c2 * c) * pz) * u;
    dydx[2] = (-cp) * u;
    cnt[CNT_NFEV]++;
}

/* the four events, ref:integration_processes.py:238-303; each returns +-1.0 */
static double ev_surface(const orc_env *E, double x, const double *y)
{
    double c, th = ray_angle(E, x, y, &c);
    return ((y[1] < 0.0) && (th < 0.0)) ? 1.0 : -1.0;
}
static double ev_bottom(const orc_env *E, double x, const double *y)
{
    double bd = linear_interp(x, E->depth_ranges, E->depths, E->n_b);
    double c, th = ray_angle(E, x, y, &c);
    return ((y[1] > bd) && (th > 0.0)) ? 1.0 : -1.0;
}
static double ev_vertical(const orc_env *E, double x, const double *y)
{
    double c, th = ray_angle(E, x, y, &c);
    return (fabs(th) > (90.0 - 1e-3)) ? 1.0 : -1.0;
}
static double ev_bbox(const orc_env *E, double x, const double *y)
{
    double z = y[1], tol = 1e-6;
    int out = (z > E->zin[E->n_z - 1] + tol) | (z < E->zin[0] - tol) | (x < E->rin[0] - tol) |
              (x > E->rin[E->n_r - 1] + tol);
    return out ? 1.0 : -1.0;
}
typedef double (*event_fn)(const orc_env *, double, const double *);
static const event_fn EVENTS[4] = {ev_surface, ev_bottom, ev_vertical, ev_bbox};
static const int EVENT_DIR[4] = {1, 1, 0, 0}; /* ref:launch_rays.py:649-661 */

/* ------------------------------------------------------------------------------------------ */
/* RK45 tableau, scipy:rk.py:538-565                                                            */
static const double RK_C[6] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200,
                               -22.0 / 525, 1.0 / 40};
static const double RK_P[7][4] = {
    {1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0,
     701980252875.0 / 199316789632.0},
    {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

#define SAFETY 0.9
#define MIN_FACTOR 0.2
#define MAX_FACTOR 10.0
#define SQRT3 1.7320508075688772 /* 3 ** 0.5, scipy:common.py:63-65 */

/* dot of the first n stage rows with a coefficient vector, skipping exact zeros */
static double kdot(double K[7][3], int comp, const double *coef, int n)
{
    double acc = 0.0;
    int first = 1;
    for (int j = 0; j < n; ++j) {
        if (coef[j] == 0.0) continue;
        if (first) { acc = coef[j] * K[j][comp]; first = 0; }
        else acc = fma(coef[j], K[j][comp], acc);
    }
    return acc;
}

/* scipy:common.py:63-65 norm() for the 3-vector */
static double rms3(const double *v)
{
    double s = v[0] * v[0];
    s = fma(v[1], v[1], s);
    s = fma(v[2], v[2], s);
    return sqrt(s) / SQRT3;
}

/* One RkDenseOutput (scipy:rk.py:178-180, 715-737) */
typedef struct {
    double t_old, h, y_old[3], Q[3][4];
} dense_t;

static void dense_eval(const dense_t *d, double t, double *y)
{
    double x = (t - d->t_old) * orc_rcp(d->h); /* contract: division by h as a multiplication by rcp(h) */
    double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x; /* np.cumprod */
    for (int k = 0; k < 3; ++k) {
        double acc = d->Q[k][0] * p1;
        acc = fma(d->Q[k][1], p2, acc);
        acc = fma(d->Q[k][2], p3, acc);
        acc = fma(d->Q[k][3], p4, acc);
        y[k] = fma(d->h, acc, d->y_old[k]);
    }
}

/* The solver object (scipy:rk.py:85-176 RungeKutta, scipy:base.py:134-210 OdeSolver) */
typedef struct {
    const orc_env *E;
    double t, t_old, t_bound, y[3], y_old[3], f[3], h_abs, rtol, atol, K[7][3];
    int finished, failed;
    int64_t *cnt;
} solver_t;

/* scipy:common.py:68-134 select_initial_step (direction = +1, order = 4, max_step = inf) */
static double select_initial_step(solver_t *S)
{
    double interval_length = fabs(S->t_bound - S->t);
    if (interval_length == 0.0) return 0.0;
    double scale[3], a[3], b[3];
    for (int k = 0; k < 3; ++k) {
        scale[k] = fma(fabs(S->y[k]), S->rtol, S->atol);
        a[k] = S->y[k] / scale[k];
        b[k] = S->f[k] / scale[k];
    }
    double d0 = rms3(a), d1 = rms3(b), h0;
    if (d0 < 1e-5 || d1 < 1e-5) h0 = 1e-6;
    else h0 = (0.01 * d0) / d1;
    h0 = py_min(h0, interval_length);
    double y1[3], f1[3];
    for (int k = 0; k < 3; ++k) y1[k] = fma(h0, S->f[k], S->y[k]);
    derivsrd(S->E, S->t + h0, y1, f1, S->cnt);
    for (int k = 0; k < 3; ++k) a[k] = (f1[k] - S->f[k]) / scale[k];
    double d2 = rms3(a) / h0, h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, h0 * 1e-3);
    else h1 = 1.0 / orc_pow_m02(0.01 / py_max(d1, d2)); /* (.)**(1/5) */
    return py_min(py_min(100.0 * h0, h1), interval_length);
}

/* scipy:rk.py:85-103 RungeKutta.__init__ */
static void solver_init(solver_t *S, const orc_env *E, double t0, const double *y0, double t_bound,
                        double rtol, double atol, int64_t *cnt)
{
    S->E = E; S->t = t0; S->t_old = NAN; S->t_bound = t_bound; S->cnt = cnt;
    memcpy(S->y, y0, sizeof S->y);
    if (rtol < 100.0 * 2.220446049250313e-16) rtol = 100.0 * 2.220446049250313e-16; /* validate_tol */
    S->rtol = rtol; S->atol = atol;
    S->finished = 0; S->failed = 0;
    derivsrd(E, S->t, S->y, S->f, cnt);
    S->h_abs = select_initial_step(S);
}

/* scipy:rk.py:14-71 rk_step + scipy:rk.py:111-176 _step_impl; returns 1 on success */
static int solver_step_impl(solver_t *S)
{
    double t = S->t, *y = S->y;
    double min_step = 10.0 * fabs(nextafter(t, INFINITY) - t);
    double h_abs;
    if (S->h_abs < min_step) h_abs = min_step; else h_abs = S->h_abs; /* max_step = inf */
    int step_rejected = 0;
    double y_new[3], f_new[3], t_new, h;
    for (;;) {
        if (h_abs < min_step) return 0; /* TOO_SMALL_STEP */
        S->cnt[CNT_ATTEMPTS]++;
        h = h_abs;
        t_new = t + h;
        if (t_new - S->t_bound > 0.0) t_new = S->t_bound;
        h = t_new - t;
        h_abs = fabs(h);
        /* rk_step */
        memcpy(S->K[0], S->f, sizeof S->f);
        for (int s = 1; s < 6; ++s) {
            double ys[3];
            for (int k = 0; k < 3; ++k) ys[k] = fma(h, kdot(S->K, k, RK_A[s], s), y[k]);
            derivsrd(S->E, fma(RK_C[s], h, t), ys, S->K[s], S->cnt);
        }
        for (int k = 0; k < 3; ++k) y_new[k] = fma(h, kdot(S->K, k, RK_B, 6), y[k]);
        derivsrd(S->E, t + h, y_new, f_new, S->cnt);
        memcpy(S->K[6], f_new, sizeof f_new);
        /* error estimate */
        /* contract: error / scale as a multiplication by 1/scale, the three reciprocals formed from ONE
         * reciprocal of the product of the scales (they are known before the last stage value is) */
        double r[3], scale[3];
        for (int k = 0; k < 3; ++k) scale[k] = fma(np_maximum(fabs(y[k]), fabs(y_new[k])), S->rtol, S->atol);
        double s01 = scale[0] * scale[1], s12 = scale[1] * scale[2], s02 = scale[0] * scale[2];
        double inv3 = orc_rcp(s01 * scale[2]);
        double inv[3] = {inv3 * s12, inv3 * s02, inv3 * s01};
        for (int k = 0; k < 3; ++k) r[k] = (kdot(S->K, k, RK_E, 7) * h) * inv[k];
        /* contract: error_norm = sqrt(mean(r^2)) is carried as its square ms (no square root, no division):
         * accept iff ms < 1, and error_norm ** (-1/5) = ms ** (-1/10)   (scipy:rk.py:148-165) */
        double ms = r[0] * r[0];
        ms = fma(r[1], r[1], ms);
        ms = fma(r[2], r[2], ms);
        ms = ms * 0x1.5555555555555p-2;
        if (ms < 1.0) {
            double factor;
            if (ms == 0.0) factor = MAX_FACTOR;
            else factor = py_min(MAX_FACTOR, SAFETY * orc_pow_m01(ms));
            if (step_rejected) factor = py_min(1.0, factor);
            h_abs *= factor;
            break;
        } else {
            h_abs *= py_max(MIN_FACTOR, SAFETY * orc_pow_m01(ms));
            step_rejected = 1;
        }
    }
    memcpy(S->y_old, y, sizeof S->y_old);
    S->t = t_new;
    memcpy(S->y, y_new, sizeof y_new);
    S->h_abs = h_abs;
    memcpy(S->f, f_new, sizeof f_new);
    S->cnt[CNT_ACCEPTED]++;
    return 1;
}

/* scipy:base.py:179-210 OdeSolver.step */
static void solver_step(solver_t *S)
{
    if (S->t == S->t_bound) {
        S->t_old = S->t; S->t = S->t_bound; S->finished = 1;
        memcpy(S->y_old, S->y, sizeof S->y);
        return;
    }
    double t = S->t;
    if (!solver_step_impl(S)) { S->failed = 1; return; }
    S->t_old = t;
    if (S->t - S->t_bound >= 0.0) S->finished = 1;
}

/* scipy:rk.py:178-180 _dense_output_impl */
static void solver_dense_output(const solver_t *S, dense_t *d)
{
    d->t_old = S->t_old;
    d->h = S->t - S->t_old;
    memcpy(d->y_old, S->y_old, sizeof d->y_old);
    for (int k = 0; k < 3; ++k)
        for (int m = 0; m < 4; ++m) {
            double col[7];
            for (int j = 0; j < 7; ++j) col[j] = RK_P[j][m];
            d->Q[k][m] = kdot((double(*)[3])S->K, k, col, 7);
        }
    S->cnt[CNT_DENSE]++;
}

/* scipy/optimize/Zeros/brentq.c, literal (xtol = rtol = 4 eps, maxiter = 100; scipy:ivp.py:52-77) */
typedef struct { const orc_env *E; const dense_t *d; event_fn ev; int64_t *cnt; } evclo_t;
static double ev_closure(double t, void *p)
{
    evclo_t *c = (evclo_t *)p;
    double y[3];
    dense_eval(c->d, t, y);
    c->cnt[CNT_BISECT]++;
    return c->ev(c->E, t, y);
}
static double brentq(double (*f)(double, void *), double xa, double xb, double xtol, double rtol,
                     int iter, void *fd, int *err)
{
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    *err = 0;
    fpre = f(xpre, fd);
    fcur = f(xcur, fd);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { *err = ORC_BRENTQ_SIGN; return 0.; }
    for (int i = 0; i < iter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim = 3 * fabs(sbis) - delta;
            if (fabs(spre) < lim) lim = fabs(spre);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = scur; scur = sbis; }
        } else {
            spre = scur; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur, fd);
    }
    *err = ORC_BRENTQ_ITER;
    return xcur;
}

/* One solve_ivp result restricted to what the path reads (scipy:ivp.py:641-756) */
typedef struct {
    int n_ts;          /* len(sol.t) */
    double *ts;        /* sol.t; also OdeSolution.ts */
    int n_interp;
    dense_t *interp;   /* OdeSolution.interpolants */
    double y_last[3];  /* sol.y[:, -1] */
    int status;        /* 0 finished, 1 event, -1 failed */
    int event;         /* which t_events[i] is non-empty, -1 if none */
    double t_event;
    int err;           /* exception raised inside solve_ivp, as an ORC_ code (0 = none) */
} segment_t;

static void segment_free(segment_t *s) { free(s->ts); free(s->interp); }

/* ref:launch_rays.py:593-681 _shoot_ray_segment == scipy solve_ivp(RK45, events, dense_output) */
static void shoot_ray_segment(const orc_env *E, double x0, const double *y0, double receiver_range,
                              double rtol, double atol, int64_t *cnt, segment_t *out)
{
    solver_t S;
    solver_init(&S, E, x0, y0, receiver_range, rtol, atol, cnt);
    int cap = 64;
    out->ts = (double *)malloc(sizeof(double) * cap);
    out->interp = (dense_t *)malloc(sizeof(dense_t) * cap);
    out->n_ts = 1; out->ts[0] = x0; out->n_interp = 0;
    out->event = -1; out->err = 0; out->t_event = NAN;
    double y_last[3];
    memcpy(y_last, y0, sizeof y_last);
    double g[4], g_new[4];
    for (int e = 0; e < 4; ++e) g[e] = EVENTS[e](E, x0, y0);
    int status = -2; /* None */
    cnt[CNT_SEGMENTS]++;
    while (status == -2) {
        solver_step(&S);
        if (S.finished) status = 0;
        else if (S.failed) { status = -1; break; }
        double t_old = S.t_old, t = S.t, y[3];
        memcpy(y, S.y, sizeof y);
        if (out->n_interp + 1 >= cap) {
            cap *= 2;
            out->ts = (double *)realloc(out->ts, sizeof(double) * cap);
            out->interp = (dense_t *)realloc(out->interp, sizeof(dense_t) * cap);
        }
        dense_t *sol = &out->interp[out->n_interp];
        solver_dense_output(&S, sol);
        out->n_interp++;
        for (int e = 0; e < 4; ++e) g_new[e] = EVENTS[e](E, t, y);
        /* find_active_events scipy:ivp.py:134-157 */
        int active[4], n_active = 0;
        for (int e = 0; e < 4; ++e) {
            int up = (g[e] <= 0) && (g_new[e] >= 0);
            int down = (g[e] >= 0) && (g_new[e] <= 0);
            int m = (up && EVENT_DIR[e] > 0) || (down && EVENT_DIR[e] < 0) ||
                    ((up || down) && EVENT_DIR[e] == 0);
            if (m) active[n_active++] = e;
        }
        if (n_active > 0) {
            /* handle_events scipy:ivp.py:80-131: all events terminal -> keep the earliest root
             * (stable order: lower event index wins ties) */
            double best_root = 0;
            int best = -1;
            for (int a = 0; a < n_active; ++a) {
                evclo_t clo = {E, sol, EVENTS[active[a]], cnt};
                int err;
                double root = brentq(ev_closure, t_old, t, 4 * 2.220446049250313e-16,
                                     4 * 2.220446049250313e-16, 100, &clo, &err);
                if (err) { out->err = err; status = -1; goto done; }
                if (best < 0 || root < best_root) { best = active[a]; best_root = root; }
            }
            out->event = best;
            out->t_event = best_root;
            status = 1;
            t = best_root;
            dense_eval(sol, t, y);
        }
        memcpy(g, g_new, sizeof g);
        /* scipy:ivp.py:701-710 */
        int donot_append = (out->n_ts > 1) && (out->ts[out->n_ts - 1] == t);
        if (!donot_append) {
            out->ts[out->n_ts++] = t;
            memcpy(y_last, y, sizeof y_last);
        } else {
            out->n_interp--;
        }
    }
done:
    out->status = status;
    memcpy(out->y_last, y_last, sizeof y_last);
    /* OdeSolution.__init__ validity check scipy:common.py:174-181 */
    if (status >= 0 && !out->err) {
        int ok = (out->n_ts == 2 && out->ts[0] == out->ts[1]);
        if (!ok) {
            ok = 1;
            for (int i = 1; i < out->n_ts; ++i) if (!(out->ts[i] - out->ts[i - 1] > 0)) ok = 0;
        }
        if (!ok || out->n_ts != out->n_interp + 1) out->err = ORC_TS_NOT_INCREASING;
    }
}

/* scipy:common.py:202-212 OdeSolution._call_single */
static void odesolution_call(const segment_t *s, double t, double *y)
{
    int ind = searchsorted_left(s->ts, s->n_ts, t);
    int seg = ind - 1;
    if (seg < 0) seg = 0;
    if (seg > s->n_interp - 1) seg = s->n_interp - 1;
    dense_eval(&s->interp[seg], t, y);
}

/* np.argmin(np.abs(a - v)) : first minimum; NaN propagates as numpy does (first NaN wins) */
static int argmin_abs_diff(const double *a, int n, double v)
{
    int best = 0;
    double bv = fabs(a[0] - v);
    if (bv != bv) return 0;
    for (int i = 1; i < n; ++i) {
        double d = fabs(a[i] - v);
        if (d != d) return i;
        if (d < bv) { bv = d; best = i; }
    }
    return best;
}

/* scipy.interpolate.interp1d(kind="cubic") of the bottom angle, as host-prepared PPoly
 * (ref:launch_rays.py:397-399, 469).  Returns 0 and sets *oob when x is outside the knots. */
static double bottom_angle_eval(const orc_env *E, double x, int *oob)
{
    int n = E->n_ba;
    *oob = 0;
    if (!(x >= E->ba_x[0]) || !(x <= E->ba_x[n - 1])) { *oob = 1; return 0.0; }
    int i = searchsorted_left(E->ba_x, n, x) - 1; /* ba_x[i] < x <= ba_x[i+1] */
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    double s = x - E->ba_x[i];
    const double *c = E->ba_c;
    int m = n - 1;
    double v = c[0 * m + i];
    v = fma(v, s, c[1 * m + i]);
    v = fma(v, s, c[2 * m + i]);
    v = fma(v, s, c[3 * m + i]);
    return v;
}

/* optional per-thread bounce log for debugging/tests: rows of [x_event, T, z, p_before, p_after, kind] */
static __thread double *g_evlog = NULL;
static __thread int g_evlog_cap = 0, g_evlog_n = 0;
ORC_API void orc_set_event_log(double *buf, int cap) { g_evlog = buf; g_evlog_cap = cap; g_evlog_n = 0; }
ORC_API int orc_event_log_count(void) { return g_evlog_n; }

/*
 * ref:launch_rays.py:325-484 _shoot_ray_array followed by ref:launch_rays.py:745-784
 * _interpolate_ray (when out_traj != NULL).
 *   out_traj : [3][n_save] rows T, z, p in ODE sign convention (positive-down z)
 *   final    : sols[-1].y[:, -1]
 * returns an ORC_ status; on status != OK outputs are NaN-filled / undefined like the
 * reference's (None, None, None, None).
 */
ORC_API int orc_shoot_ray(const orc_env *E, const double *y0, double source_range, double receiver_range,
                          double rtol, double atol, int terminate_backwards, int n_save,
                          const double *range_save, double *out_traj, double *final, int32_t *n_bottom_out,
                          int32_t *n_surface_out, int64_t *cnt, int max_bounces)
{
    double x = source_range, y[3];
    memcpy(y, y0, sizeof y);
    int n_surface = 0, n_bottom = 0, status = ORC_OK;
    int n_sols = 0, cap = 16;
    segment_t *sols = (segment_t *)malloc(sizeof(segment_t) * cap);
    while (x < receiver_range) {
        if (n_sols == cap) { cap *= 2; sols = (segment_t *)realloc(sols, sizeof(segment_t) * cap); }
        segment_t *sol = &sols[n_sols++];
        shoot_ray_segment(E, x, y, receiver_range, rtol, atol, cnt, sol);
        if (sol->err) { status = sol->err; break; }
        if (sol->status == 0) break;
        if (sol->status == -1) { status = ORC_STEP_TOO_SMALL; break; }
        memcpy(y, sol->y_last, sizeof y);
        if (sol->event == 0 || sol->event == 1) x = sol->t_event;
        else if (sol->event == 2) { status = ORC_VERTICAL; break; }
        else { status = ORC_BBOX; break; }
        double c, theta = ray_angle(E, x, y, &c), theta_bounce;
        if (sol->event == 0) { theta_bounce = -theta; n_surface++; }
        else {
            int oob;
            double beta = bottom_angle_eval(E, x, &oob);
            if (oob) { status = ORC_BOTTOM_OOB; break; }
            theta_bounce = 2.0 * beta - theta;
            n_bottom++;
        }
        if (terminate_backwards && (fabs(theta_bounce) > 90.0)) { status = ORC_BACKWARDS; break; }
        double p_before = y[2];
        y[2] = orc_sin(theta_bounce * DEG2RAD) / c;
        if (g_evlog && g_evlog_n < g_evlog_cap) {
            double *row = g_evlog + 6 * g_evlog_n++;
            row[0] = x; row[1] = y[0]; row[2] = y[1]; row[3] = p_before; row[4] = y[2]; row[5] = sol->event;
        }
        memcpy(sol->y_last, y, sizeof y); /* y_intermediate is a view of sol.y[:, -1] */
        if (n_surface + n_bottom >= max_bounces) { status = ORC_MAX_BOUNCES; break; }
    }
    if (status == ORC_OK && n_sols > 0) {
        memcpy(final, sols[n_sols - 1].y_last, 3 * sizeof(double));
        if (out_traj) {
            /* _interpolate_ray */
            for (int k = 0; k < 3; ++k)
                for (int i = 0; i < n_save; ++i) out_traj[(size_t)k * n_save + i] = NAN;
            for (int s = 0; s < n_sols; ++s) {
                segment_t *sol = &sols[s];
                int idx1 = argmin_abs_diff(range_save, n_save, sol->ts[0]);
                int idx2 = argmin_abs_diff(range_save, n_save, sol->ts[sol->n_ts - 1]);
                if (idx1 == idx2) continue;
                if (idx2 > n_save - 1) idx2 = n_save - 1; /* full_ray has n_save-1 columns */
                for (int i = idx1; i < idx2; ++i) {
                    double v[3];
                    odesolution_call(sol, range_save[i], v);
                    for (int k = 0; k < 3; ++k) out_traj[(size_t)k * n_save + i] = v[k];
                }
            }
            for (int k = 0; k < 3; ++k) out_traj[(size_t)k * n_save + n_save - 1] = final[k];
        }
    } else if (status == ORC_OK) {
        /* while loop never entered: source_range >= receiver_range (not reachable via public API) */
        memcpy(final, y, sizeof y);
    } else {
        final[0] = final[1] = final[2] = NAN;
        if (out_traj)
            for (size_t i = 0; i < (size_t)3 * n_save; ++i) out_traj[i] = NAN;
    }
    for (int s = 0; s < n_sols; ++s) segment_free(&sols[s]);
    free(sols);
    *n_bottom_out = n_bottom;
    *n_surface_out = n_surface;
    return status;
}

/* y0 = [0, z_src, sin(radians(angle))/c(src)]   ref:launch_rays.py:140-144, 284-285 */
ORC_API void orc_initial_state(const orc_env *E, double source_depth, double source_range,
                               double angle_deg, double *y0)
{
    double c = bilinear_interp(source_range, source_depth, E->rin, E->n_r, E->zin, E->n_z, E->cin);
    y0[0] = 0.0;
    y0[1] = source_depth;
    y0[2] = orc_sin(angle_deg * DEG2RAD) / c;
}

/* ------------------------------------------------------------------------------------------ */
/* pthread work queue standing in for the reference's multiprocessing Pool.imap                 */
static int default_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static void run_threads(void *(*fn)(void *), void *arg, int n_threads)
{
    if (n_threads <= 0) n_threads = default_threads();
    if (n_threads == 1) { fn(arg); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, fn, arg);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
}

typedef struct {
    const orc_env *E; int64_t n_rays; const double *angles_deg, *y0s;
    double source_depth, source_range, receiver_range, rtol, atol;
    int terminate_backwards, n_save; const double *range_save;
    double *out_traj, *out_final; int32_t *out_nb, *out_ns, *out_status; int64_t *counters;
    int max_bounces; int64_t next;
} fan_job;

static void *fan_worker(void *arg)
{
    fan_job *J = (fan_job *)arg;
    for (;;) {
        int64_t i0 = __atomic_fetch_add(&J->next, 4, __ATOMIC_RELAXED);
        if (i0 >= J->n_rays) break;
        int64_t i1 = i0 + 4 < J->n_rays ? i0 + 4 : J->n_rays;
        for (int64_t i = i0; i < i1; ++i) {
            double y0[3];
            int64_t cnt[CNT_N] = {0};
            if (J->y0s) memcpy(y0, J->y0s + 3 * i, sizeof y0);
            else orc_initial_state(J->E, J->source_depth, J->source_range, J->angles_deg[i], y0);
            J->out_status[i] = orc_shoot_ray(J->E, y0, J->source_range, J->receiver_range, J->rtol, J->atol,
                                             J->terminate_backwards, J->n_save, J->range_save,
                                             J->out_traj ? J->out_traj + (size_t)i * 3 * J->n_save : NULL,
                                             J->out_final + 3 * i, J->out_nb + i, J->out_ns + i, cnt,
                                             J->max_bounces);
            if (J->counters) memcpy(J->counters + i * CNT_N, cnt, sizeof cnt);
        }
    }
    return NULL;
}

/*
 * Fan: the data-parallel loop of ref:launch_rays.py:94-117 / 157-164 (sequential loop or
 * Pool.imap over launch angles) -- here a pthread work queue.  angles are INTERNAL angles (the
 * sign flips of shoot_rays/shoot_ray are host-layer business).  If y0s != NULL it is used instead.
 *   out_traj : NULL or [n_rays][3][n_save];  out_final : [n_rays][3]
 *   counters : NULL or [n_rays][CNT_N];  n_threads <= 0 -> all online cores
 */
ORC_API int orc_shoot_fan(const orc_env *E, int64_t n_rays, const double *angles_deg, const double *y0s,
                          double source_depth, double source_range, double receiver_range, double rtol,
                          double atol, int terminate_backwards, int n_save, const double *range_save,
                          double *out_traj, double *out_final, int32_t *out_nb, int32_t *out_ns,
                          int32_t *out_status, int64_t *counters, int max_bounces, int n_threads)
{
    fan_job job = {E, n_rays, angles_deg, y0s, source_depth, source_range, receiver_range, rtol, atol,
                   terminate_backwards, n_save, range_save, out_traj, out_final, out_nb, out_ns,
                   out_status, counters, max_bounces, 0};
    run_threads(fan_worker, &job, n_threads);
    return 0;
}

typedef struct {
    const orc_env *E; int64_t n_br; const double *z1s, *z2s, *th1s, *th2s, *rds;
    double source_depth, source_range, receiver_range, rtol, atol;
    int terminate_backwards; double ztol; int max_iter;
    double *out_theta, *out_final; int32_t *out_nb, *out_ns, *out_status, *out_shots, *out_ray_status;
    int max_bounces; int64_t next;
} refine_job;

/* ref:eigenrays.py:206-268 _find_single_eigenray for one bracket */
static void find_single_eigenray(refine_job *J, int64_t b)
{
    const orc_env *E = J->E;
    double z1 = J->z1s[b], z2 = J->z2s[b], theta1 = J->th1s[b], theta2 = J->th2s[b], rd = J->rds[b];
    double th = theta1 - (z1 + rd) * (theta2 - theta1) / (z2 - z1); /* ref:eigenrays.py:118-120 */
    int iter_count = 0, shots = 0;
    J->out_theta[b] = NAN;
    J->out_final[3 * b] = J->out_final[3 * b + 1] = J->out_final[3 * b + 2] = NAN;
    J->out_nb[b] = J->out_ns[b] = 0;
    J->out_ray_status[b] = 0;
    for (;;) {
        double y0[3], fin[3];
        int64_t cnt[CNT_N] = {0};
        int32_t nb, ns;
        /* shoot_ray negates the user angle, ref:launch_rays.py:251 */
        orc_initial_state(E, J->source_depth, J->source_range, -th, y0);
        int st = orc_shoot_ray(E, y0, J->source_range, J->receiver_range, J->rtol, J->atol,
                               J->terminate_backwards, 0, NULL, NULL, fin, &nb, &ns, cnt, J->max_bounces);
        shots++;
        if (st != ORC_OK) { J->out_status[b] = 1; J->out_ray_status[b] = st; break; }
        double zend = -fin[1]; /* ray.z[-1], negative-z convention ref:ray_objects.py:50 */
        if (fabs(zend + rd) < J->ztol) {
            J->out_status[b] = 0; J->out_theta[b] = th;
            memcpy(J->out_final + 3 * b, fin, sizeof fin);
            J->out_nb[b] = nb; J->out_ns[b] = ns;
            break;
        }
        /* np.sign comparison, ref:eigenrays.py:253 */
        double s_ray = (zend + rd > 0) - (zend + rd < 0), s_1 = (z1 + rd > 0) - (z1 + rd < 0);
        if (s_ray == s_1) { z1 = zend; theta1 = th; }
        else { z2 = zend; theta2 = th; }
        th = theta1 - (z1 + rd) * (theta2 - theta1) / (z2 - z1);
        if (iter_count > J->max_iter) { J->out_status[b] = 2; break; }
        iter_count++;
    }
    J->out_shots[b] = shots;
}

static void *refine_worker(void *arg)
{
    refine_job *J = (refine_job *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (b >= J->n_br) break;
        find_single_eigenray(J, b);
    }
    return NULL;
}

/*
 * ref:eigenrays.py:122-190: the loop (Pool.imap or sequential) over brackets.
 * Sign conventions are the reference's: theta* are USER launch angles, z1/z2 and ray.z[-1] are
 * NEGATIVE-z (Ray stores -y[1], ref:ray_objects.py:49-52), receiver depths are positive.
 *   out_theta[b]  : converged user launch angle (NaN if failed)
 *   out_final[b]  : [3] final ODE state of the converged ray
 *   out_status[b] : 0 converged, 1 ray dropped (shoot_ray -> None), 2 max_iter exceeded
 *   out_shots[b]  : number of shoot_ray calls
 */
ORC_API int orc_refine_brackets(const orc_env *E, int64_t n_br, const double *z1s, const double *z2s,
                                const double *th1s, const double *th2s, const double *rds,
                                double source_depth, double source_range, double receiver_range,
                                double rtol, double atol, int terminate_backwards, double ztol,
                                int max_iter, double *out_theta, double *out_final, int32_t *out_nb,
                                int32_t *out_ns, int32_t *out_status, int32_t *out_shots,
                                int32_t *out_ray_status, int max_bounces, int n_threads)
{
    refine_job job = {E, n_br, z1s, z2s, th1s, th2s, rds, source_depth, source_range, receiver_range,
                      rtol, atol, terminate_backwards, ztol, max_iter, out_theta, out_final, out_nb,
                      out_ns, out_status, out_shots, out_ray_status, max_bounces, 0};
    run_threads(refine_worker, &job, n_threads);
    return 0;
}

/* thin exports for unit tests */
ORC_API double orc_bilinear(const orc_env *E, double x, double z, int which)
{
    return bilinear_interp(x, z, E->rin, E->n_r, E->zin, E->n_z, which ? E->cpin : E->cin);
}
ORC_API void orc_derivsrd(const orc_env *E, double x, const double *y, double *dydx)
{
    int64_t cnt[CNT_N] = {0};
    derivsrd(E, x, y, dydx, cnt);
}
ORC_API double orc_event(const orc_env *E, int which, double x, const double *y) { return EVENTS[which](E, x, y); }
ORC_API double orc_bottom_angle(const orc_env *E, double x, int *oob) { return bottom_angle_eval(E, x, oob); }
static double step_fn(double t, void *p) { return (t >= *(double *)p) ? 1.0 : -1.0; }
/* brentq on a -1/+1 step function with the jump at `jump` (unit test against scipy.optimize.brentq) */
ORC_API double orc_brentq_step(double jump, double a, double b, int *err)
{
    return brentq(step_fn, a, b, 4 * 2.220446049250313e-16, 4 * 2.220446049250313e-16, 100, &jump, err);
}
ORC_API int orc_num_threads(void) { return default_threads(); }

/* debug/test export: integrate ONE segment and return its last interpolant + event info */
ORC_API int orc_debug_segment(const orc_env *E, double x0, const double *y0, double receiver_range, double rtol,
                              double atol, double *dense17, double *t_event, int *event, int *n_ts, double *y_last)
{
    int64_t cnt[CNT_N] = {0};
    segment_t s;
    shoot_ray_segment(E, x0, y0, receiver_range, rtol, atol, cnt, &s);
    if (s.n_interp > 0) {
        const dense_t *d = &s.interp[s.n_interp - 1];
        dense17[0] = d->t_old; dense17[1] = d->h;
        memcpy(dense17 + 2, d->y_old, 3 * sizeof(double));
        memcpy(dense17 + 5, d->Q, 12 * sizeof(double));
    }
    *t_event = s.t_event; *event = s.event; *n_ts = s.n_ts;
    memcpy(y_last, s.y_last, sizeof s.y_last);
    int st = s.status;
    segment_free(&s);
    return st;
}

ORC_API int orc_debug_segment_ts(const orc_env *E, double x0, const double *y0, double receiver_range, double rtol,
                                 double atol, double *ts, int cap)
{
    int64_t cnt[CNT_N] = {0};
    segment_t s;
    shoot_ray_segment(E, x0, y0, receiver_range, rtol, atol, cnt, &s);
    int n = s.n_ts < cap ? s.n_ts : cap;
    memcpy(ts, s.ts, n * sizeof(double));
    segment_free(&s);
    return n;
}
