/*
 * prt.h -- C ABI of libprt.so, the B200 (sm_100a) ray-fan / eigenray engine.
 *
 * This is the drop-in boundary for pygenray's hot path.  pygenray is pure Python and has no FFI
 * of its own; the seam these entry points replace is the ARRAY-LEVEL interface of
 * /root/reference/src/pygenray/launch_rays.py:
 *
 *   _unpack_envi(environment, flatearth)                 :717-742  -> prt_env_create
 *   _shoot_ray_array(y0, ..., cin, cpin, rin, zin, depths, depth_ranges, bottom_angles,
 *                    rtol, terminate_backwards, debug)   :325-484  \
 *   _interpolate_ray(sols, range_save)                   :745-784   > prt_shoot_fan (batched over
 *   Pool.imap(_shoot_single_ray_process, y0s)            :157-164  /   launch angles)
 *   find_eigenrays / _find_single_eigenray               eigenrays.py:62-120, 206-268
 *                                                                  -> prt_count_brackets + prt_fill_brackets,
 *                                                                     prt_refine_brackets
 *   EigenRays.__init__ (received angles, ray ids)         ray_objects.py:505-548 -> prt_eigen_table,
 *                                                                     prt_received_angles
 *   eflat / flat_earth_transform / flat_earth_c           environment.py:121-173, 239-401 -> prt_flat_earth
 *
 * Conventions
 *   - plain C types only; every array is a raw pointer + explicit sizes; FP64 everywhere.
 *   - `*_dev` pointers are DEVICE pointers owned by the caller (e.g. torch CUDA tensors);
 *     calls taking a `stream` are asynchronous on that CUDA stream (cudaStream_t passed as void*;
 *     NULL = the legacy default stream).  The library allocates nothing persistent except the
 *     environment handle.
 *   - return value: PRT_SUCCESS or a PRT_ERR_* code for argument / CUDA errors only
 *     (prt_last_error() has the text).  Per-ray failures are DATA: the `status` output holds a
 *     PRT_RAY_* code per launch angle, mirroring the reference's "dropped ray -> None"
 *     convention (launch_rays.py:305-306, 430, 448, 456, 477).
 *   - angles passed to the library are INTERNAL launch angles in degrees (positive = down-going,
 *     i.e. y0[2] = sin(radians(angle)) / c); the sign flips of shoot_rays / shoot_ray
 *     (launch_rays.py:64-67, 251) belong to the host layer.
 *   - ODE state sign convention: z positive downwards, exactly as inside _shoot_ray_array.
 */
#ifndef PRT_H
#define PRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRT_ABI_VERSION 2

#if defined(__GNUC__)
#define PRT_API __attribute__((visibility("default")))
#else
#define PRT_API
#endif

/* return codes */
#define PRT_SUCCESS 0
#define PRT_ERR_ARG 1
#define PRT_ERR_CUDA 2
#define PRT_ERR_NO_DEVICE 3

/* per-ray status */
#define PRT_RAY_OK 0             /* reached receiver_range */
#define PRT_RAY_STEP_TOO_SMALL 1 /* solve_ivp status -1 (launch_rays.py:427-430) */
#define PRT_RAY_VERTICAL 2       /* vertical_ray event (launch_rays.py:443-448) */
#define PRT_RAY_BBOX 3           /* ray_bounding_box_event (launch_rays.py:451-456) */
#define PRT_RAY_BACKWARDS 4      /* |theta_bounce| > 90 with terminate_backwards (launch_rays.py:474-477) */
#define PRT_RAY_BOTTOM_OOB 5     /* bounce outside the bathymetry grid: interp1d raises ValueError (:469) */
#define PRT_RAY_BRENTQ_SIGN 6    /* scipy brentq ValueError inside solve_ivp */
#define PRT_RAY_BRENTQ_ITER 7    /* scipy brentq RuntimeError inside solve_ivp */
#define PRT_RAY_MAX_BOUNCES 9    /* safety cap `max_bounces` hit (no reference equivalent) */

/* per-ray work counters, out_counters[k][ray] */
#define PRT_CNT_NFEV 0     /* right-hand-side evaluations (== sum of sol.nfev) */
#define PRT_CNT_ATTEMPTS 1 /* RK45 step attempts */
#define PRT_CNT_ACCEPTED 2 /* accepted steps */
#define PRT_CNT_SEGMENTS 3 /* solve_ivp segments (1 + bounces) */
#define PRT_CNT_BISECT 4   /* event-function evaluations inside root searches */
#define PRT_CNT_DENSE 5    /* dense-output constructions */
#define PRT_CNT_N 6

/* eigenray bracket status */
#define PRT_EIG_CONVERGED 0 /* |z_end + receiver_depth| < ztol */
#define PRT_EIG_RAY_DROPPED 1 /* a trial ray was dropped (eigenrays.py:241-245) */
#define PRT_EIG_MAX_ITER 2  /* iteration limit (eigenrays.py:265-266) */

typedef struct prt_env prt_env_t;

PRT_API int prt_abi_version(void);
PRT_API const char *prt_last_error(void);

/* Launch hint: the caller is about to run `n_fans` similar fans concurrently on different streams (e.g. a
 * fan split into chunks whose device-to-host copies overlap the remaining chunks); affects only how small
 * fans are spread over warps, never results.  Default 1. */
PRT_API int prt_set_concurrency(int n_fans);

/* Launch hint: rays per warp (1 .. 32) of the launches that follow on this host thread's behalf;
 * 0 (default) lets the library choose from the fan size.  Fewer rays per warp shorten the critical path of
 * reflecting (steep) rays, whose lanes otherwise wait for one another's root searches; affects only speed. */
PRT_API int prt_set_lanes(int lanes);

/* number of visible CUDA devices (0 if none / no driver) */
PRT_API int prt_device_count(void);

/*
 * Stage one environment on `device`.  Inputs are HOST arrays exactly as returned by the
 * reference's _unpack_envi (launch_rays.py:717-742), plus the not-a-knot cubic spline of
 * bottom_angles over depth_ranges (the interp1d(kind="cubic") that _shoot_ray_array rebuilds per
 * ray, launch_rays.py:397-399) as piecewise-polynomial coefficients:
 *   cin, cpin     [n_r][n_z]   sound speed and d c / d z, C order [range][depth]
 *   rin [n_r], zin [n_z]       grid coordinates (ascending)
 *   depth_ranges, depths [n_b] bathymetry
 *   ba_x [n_ba]                spline breakpoints; ba_c [4][n_ba-1] coefficients, highest power
 *                              first, local variable (x - ba_x[i])
 */
PRT_API int prt_env_create(int device, int32_t n_r, int32_t n_z, const double *rin, const double *zin,
                   const double *cin, const double *cpin, int32_t n_b, const double *depth_ranges,
                   const double *depths, int32_t n_ba, const double *ba_x, const double *ba_c,
                   prt_env_t **out);
/*
 * Tomography ensembles (BASELINE config 5): n_env realisations of the sound-speed field on ONE grid
 * geometry and bathymetry.  cin, cpin are [n_env][n_r][n_z]; everything else as prt_env_create.
 * The handle is used with prt_shoot_ensemble and destroyed with prt_env_destroy.
 */
PRT_API int prt_ensemble_create(int device, int32_t n_env, int32_t n_r, int32_t n_z, const double *rin,
                                const double *zin, const double *cin, const double *cpin, int32_t n_b,
                                const double *depth_ranges, const double *depths, int32_t n_ba,
                                const double *ba_x, const double *ba_c, prt_env_t **out);
/*
 * The same for sound-speed grids that already live in device memory (c_dev: [n_env][n_r][n_z] doubles on
 * `device`), e.g. perturbed realisations generated on the GPU: the depth derivative the reference takes in
 * _unpack_envi (sound_speed.differentiate("depth") = np.gradient(c, z, axis=depth, edge_order=1),
 * launch_rays.py:728-733) is formed by a kernel with numpy's operations in numpy's order (bit-identical to the
 * host route), interleaved with c, and nothing but the coordinate tables crosses PCIe.  n_env = 1 stages a
 * single environment.  Synchronises `stream` before returning.
 */
PRT_API int prt_ensemble_create_dev(int device, int32_t n_env, int32_t n_r, int32_t n_z, const double *rin,
                                    const double *zin, const double *c_dev, int32_t n_b,
                                    const double *depth_ranges, const double *depths, int32_t n_ba,
                                    const double *ba_x, const double *ba_c, void *stream, prt_env_t **out);
PRT_API int prt_env_destroy(prt_env_t *env);
/* bytes of device memory held by the environment */
PRT_API int64_t prt_env_device_bytes(const prt_env_t *env);

/*
 * Shoot a fan of n_rays launch angles from (source_range, source_depth) to receiver_range
 * (source_range < receiver_range; backwards shots are mirrored by the caller as the reference
 * does, launch_rays.py:124-131).
 *
 *   angles_dev [n_rays]        internal launch angles [deg], or NULL if y0_dev is given
 *   y0_dev     [n_rays][3]     explicit initial states [T, z, p] (as _shoot_ray_array's y0), or NULL
 *   range_save_dev [n_save]    the save grid np.linspace(source_range, receiver_range, n_save);
 *                              NULL / n_save = 0 -> arrival table only
 *   out_traj_dev [3][n_rays][n_save]  T, z, p on the save grid (_interpolate_ray); NULL if n_save = 0
 *   out_final_dev [3][n_rays]  exact final state (sols[-1].y[:, -1]); NaN for dropped rays
 *   out_n_bottom_dev, out_n_surface_dev, out_status_dev [n_rays] int32
 *   out_counters_dev [PRT_CNT_N][n_rays] uint32, or NULL
 *   stored_convention          0: trajectory holds z, p as integrated (z positive down);
 *                              1: holds -z, -p, the convention Ray/RayFan store (ray_objects.py:49-52)
 *   out_p_sign_changes_dev [n_rays] int32 or NULL: number of sign changes of p along the saved
 *                              grid, i.e. sum(diff(sign(ps)) != 0) of RayFan.compute_rayids
 *                              (ray_objects.py:142-144); needs n_save > 0
 */
PRT_API int prt_shoot_fan(const prt_env_t *env, int64_t n_rays, const double *angles_dev, const double *y0_dev,
                  double source_depth, double source_range, double receiver_range, double rtol, double atol,
                  int32_t terminate_backwards, int32_t n_save, const double *range_save_dev,
                  double *out_traj_dev, double *out_final_dev, int32_t *out_n_bottom_dev,
                  int32_t *out_n_surface_dev, int32_t *out_status_dev, uint32_t *out_counters_dev,
                  int32_t stored_convention, int32_t *out_p_sign_changes_dev, int32_t max_bounces,
                  void *stream);

/*
 * One launch over the first n_env realisations of an ensemble handle: the same rays_per_env launch
 * angles (angles_dev[rays_per_env]) are shot through every realisation.  Outputs are laid out as for a
 * fan of n_env * rays_per_env rays, realisation-major (ray = env * rays_per_env + k).  The reference
 * would loop shoot_rays over environments (one process pool each).
 */
PRT_API int prt_shoot_ensemble(const prt_env_t *env, int32_t n_env, int64_t rays_per_env, const double *angles_dev,
                               double source_depth, double source_range, double receiver_range, double rtol,
                               double atol, int32_t terminate_backwards, int32_t n_save,
                               const double *range_save_dev, double *out_traj_dev, double *out_final_dev,
                               int32_t *out_n_bottom_dev, int32_t *out_n_surface_dev, int32_t *out_status_dev,
                               uint32_t *out_counters_dev, int32_t stored_convention,
                               int32_t *out_p_sign_changes_dev, int32_t max_bounces, void *stream);

/*
 * Same, HOST pointers in and out: stages the launch angles, runs the fan and copies the results
 * back (synchronous).  This is the call a non-Python host (C, Fortran, Julia via ccall) binds.
 */
PRT_API int prt_shoot_fan_host(const prt_env_t *env, int64_t n_rays, const double *angles, const double *y0,
                       double source_depth, double source_range, double receiver_range, double rtol,
                       double atol, int32_t terminate_backwards, int32_t n_save, const double *range_save,
                       double *out_traj, double *out_final, int32_t *out_n_bottom, int32_t *out_n_surface,
                       int32_t *out_status, uint32_t *out_counters, int32_t stored_convention,
                       int32_t *out_p_sign_changes, int32_t max_bounces);

/*
 * Pack the arrival table of a fan into records [n_rays][6] = T, z, p, n_bottom, n_surface, status
 * (doubles): the unit exchanged between GPUs (one all-gather; the reference gathers pickled Ray
 * objects from its pool workers, launch_rays.py:157-186).
 */
PRT_API int prt_pack_arrivals(int64_t n_rays, const double *final_dev, const int32_t *n_bottom_dev,
                              const int32_t *n_surface_dev, const int32_t *status_dev, double *records_dev,
                              void *stream);

/*
 * Eigenray bracketing (eigenrays.py:62-79): for every receiver depth, find adjacent surviving fan
 * rays whose final depths straddle it (sign change of zs[:, -1] + receiver_depth, zs negative-z).
 *   z_end_dev [n_rays]   final depths of the surviving fan rays in STORED convention (= -z)
 *   thetas_dev [n_rays]  their launch angles in the fan's stored (user) convention
 *   receiver_depths_dev [n_rcv]
 * Two passes over a caller-provided scratch of prt_bracket_workspace_bytes(n_rays, n_rcv) bytes:
 * prt_count_brackets counts the sign changes per (receiver, tile of 2048 rays), scans them into output
 * positions (kept in work_dev) and fills counts_dev[n_rcv]; the caller reads the counts, allocates
 * n_brackets = sum(counts) entries, and prt_fill_brackets writes, in (receiver, ray) order,
 *   br_z1/z2/th1/th2/rd [n_brackets], br_rcv [n_brackets] (receiver index), br_k [n_brackets] (ray index k
 *   of the pair (k, k + 1)).
 * Both passes must see the same inputs and the same work_dev.
 */
PRT_API int64_t prt_bracket_workspace_bytes(int64_t n_rays, int32_t n_rcv);
PRT_API int prt_count_brackets(int64_t n_rays, const double *z_end_dev, int32_t n_rcv,
                       const double *receiver_depths_dev, int32_t *counts_dev, void *work_dev, void *stream);
PRT_API int prt_fill_brackets(int64_t n_rays, const double *z_end_dev, const double *thetas_dev, int32_t n_rcv,
                      const double *receiver_depths_dev, const void *work_dev, double *br_z1_dev,
                      double *br_z2_dev, double *br_th1_dev, double *br_th2_dev, double *br_rd_dev,
                      int32_t *br_rcv_dev, int64_t *br_k_dev, void *stream);

/*
 * Regula-falsi refinement of n_br brackets (eigenrays.py:206-268), all brackets advanced together
 * one trial shot per round; each trial is a full ray (no trajectory).  theta are USER angles: the
 * trial ray is launched with internal angle -theta (launch_rays.py:251).
 *   work_dev: caller-provided scratch of prt_refine_workspace_bytes(n_br) bytes
 *   order_dev [n_br] int32 or NULL: a permutation of 0..n_br-1, the order in which the brackets are laid out over
 *             warps (results do not depend on it).  Brackets sorted by launch angle put 32 nearly identical trial
 *             rays in a warp -- same steps, same grid cells, same reflections -- which the (receiver, ray) order of
 *             prt_fill_brackets does not (neighbours there are ~80 fan rays apart): 1.9x faster trial shots on config 4
 *   out_theta_dev [n_br]  converged user launch angle (NaN otherwise)
 *   out_final_dev [3][n_br], out_n_bottom_dev, out_n_surface_dev, out_status_dev (PRT_EIG_*),
 *   out_shots_dev [n_br], out_ray_status_dev [n_br] (PRT_RAY_* of the dropped trial)
 * Asynchronous on `stream`: max_iter + 2 rounds are enqueued (trial shots of the brackets still active, then an
 * order-preserving compaction of the active list); the active count never leaves the device.
 */
PRT_API int64_t prt_refine_workspace_bytes(int64_t n_br);
PRT_API int prt_refine_brackets(const prt_env_t *env, int64_t n_br, const double *z1_dev, const double *z2_dev,
                        const double *th1_dev, const double *th2_dev, const double *rd_dev,
                        double source_depth, double source_range, double receiver_range, double rtol,
                        double atol, int32_t terminate_backwards, double ztol, int32_t max_iter,
                        double *out_theta_dev, double *out_final_dev, int32_t *out_n_bottom_dev,
                        int32_t *out_n_surface_dev, int32_t *out_status_dev, int32_t *out_shots_dev,
                        int32_t *out_ray_status_dev, void *work_dev, const int32_t *order_dev, int32_t max_bounces,
                        void *stream);

/*
 * Earth flattening of sound-speed grids in device memory (environment.py:121-151, 239-401) -- the step right before
 * the path, per realisation in ensemble mode.  c_dev / out_dev: [n_env][n_r][n_z]; depth_dev [n_z] the UN-flattened
 * depth coordinate; re_dev [n_r] the local earth radius of every range column (WGS-84, eflat :382-395; one value
 * repeated for the single-latitude transform).
 *   resample = 0: out = c (1 + E (1 + E)), E = depth / re  (the values of flat_earth_transform; the caller also
 *                 stretches the depth coordinate: z_f = z (1 + E (1/2 + E / 3)))
 *   resample = 1: flat_earth_c: the flattened column on its stretched grid interpolated back onto depth_dev
 *                 (linear, NaN outside), i.e. sound_speed_fe of flat_earth_transform_rd.  Not in place.
 */
PRT_API int prt_flat_earth(int32_t n_env, int32_t n_r, int32_t n_z, const double *depth_dev, const double *re_dev,
                           const double *c_dev, int32_t resample, double *out_dev, void *stream);

/*
 * Perturbed sound-speed realisations generated in device memory for tomography ensembles (BASELINE config 5; the
 * reference has no generator -- SURVEY.md section 8f rank 3).  Separable internal-wave-like field
 *   out_e(r, z) = c_base(r, z) + s_e * sum_j vert[j][z] * sum_k cos(kappa[k] * r + phases[e][j][k])
 * with s_e chosen so that the rms of the perturbation over range at depth index iax equals rms_axis.
 *   r_dev [n_r], vert_dev [J][n_z] (amplitude-weighted vertical modes), kappa_dev [K], phases_dev [n_env][J][K],
 *   c_base_dev [n_r][n_z], work_dev: prt_iw_workspace_bytes(n_env, n_r, J) bytes, out_dev [n_env][n_r][n_z]
 * The result feeds prt_ensemble_create_dev directly (nothing crosses PCIe but the phase table).
 */
PRT_API int64_t prt_iw_workspace_bytes(int32_t n_env, int32_t n_r, int32_t J);
PRT_API int prt_iw_perturb(int32_t n_env, int32_t n_r, int32_t n_z, int32_t J, int32_t K, const double *r_dev,
                           const double *vert_dev, const double *kappa_dev, const double *phases_dev,
                           const double *c_base_dev, int32_t iax, double rms_axis, void *work_dev, double *out_dev,
                           void *stream);

/*
 * Eigenray post-processing on the device (EigenRays.__init__, ray_objects.py:505-548).
 *
 * prt_received_angles: ray_angle (integration_processes.py:306-334) the way that constructor calls it -- on the
 * STORED end state (z_dev = -depth, p_dev sign-flipped) at range x and on the UN-flattened sound speed
 * environment.sound_speed (cin_dev [n_r][n_z], rin_dev, zin_dev: device copies made by the caller; this is not the
 * staged environment, which may be flat-earth transformed or mirrored): literal bilinear_interp, degrees(arcsin(p c)).
 *
 * prt_eigen_table: one row of PRT_EIG_TABLE_WIDTH doubles per converged eigenray i = 0..n-1, bracket b = idx_dev[i]
 * of a prt_refine_brackets result (rcv_dev from prt_fill_brackets; theta/final/n_bottom/n_surface its outputs):
 *   [receiver index, arrival time T, z (stored, = -depth), p (stored), launch angle, received angle,
 *    ray id number, n_bottom, n_surface, bracket index]
 * ray id number = p_sign_changes_dev[i] * sign(launch angle) (ray_objects.py:533-535) with p_sign_changes_dev the
 * out_p_sign_changes_dev of the re-shot of the n converged rays on the save grid, or NaN when it is NULL.
 */
#define PRT_EIG_TABLE_WIDTH 10
PRT_API int prt_received_angles(int64_t n, double x, const double *z_dev, const double *p_dev, int32_t n_r,
                                int32_t n_z, const double *rin_dev, const double *zin_dev, const double *cin_dev,
                                double *out_theta_dev, void *stream);
PRT_API int prt_eigen_table(int64_t n, int64_t n_br, const int64_t *idx_dev, const int32_t *rcv_dev,
                            const double *theta_dev, const double *final_dev, const int32_t *n_bottom_dev,
                            const int32_t *n_surface_dev, const int32_t *p_sign_changes_dev, double receiver_range,
                            int32_t n_r, int32_t n_z, const double *rin_dev, const double *zin_dev,
                            const double *cin_dev, double *out_table_dev, void *stream);

/* number of kernels this library has launched in the calling process (monotonic; for bench.py) */
PRT_API int64_t prt_launch_count(void);

/* FP64 FMA-pipe microbenchmark: runs `iters` dependent-chain DFMA rounds on every SM and returns
 * the achieved FLOP/s (2 flops per FMA), or a negative value on error.  Used as the roofline
 * denominator of bench.py (MEASURED_PEAKS.json has no FP64 figure). */
PRT_API double prt_measure_fp64_peak(int device, int32_t iters);

#ifdef __cplusplus
}
#endif
#endif /* PRT_H */
