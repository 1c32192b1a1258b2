// prt_kernels.cu -- sm_100a kernels + the C ABI of libprt.so (include/prt.h).
//
// K1  prt_fan_kernel<SAVE, ZM, COUNT>  one FP64 thread per launch angle: the whole ray (all solve_ivp
//                               segments, events, reflections, save-grid resampling) in registers.
// K2  prt_refine_round_kernel   one regula-falsi trial shot per active eigenray bracket.
// K3  prt_count/fill_brackets   sign-change bracketing of the arrival table.
//     prt_fp64_peak_kernel      DFMA-pipe microbenchmark (roofline denominator).
//
// No tensor cores on purpose: the path is a scalar adaptive ODE, not a contraction.  The
// environment grid is read through the read-only (LDG.NC) path as 16-byte {c, dc/dz} nodes and
// stays L1/L2 resident (DESIGN.md sections 2 and 5).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/prt.h"
#include "prt_core.h"
#include "prt_env_layout.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}
#define PRT_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(PRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

// entry points run on the environment's device and leave the caller's current device as they found it
struct prt_device_guard {
    int prev = -1;
    cudaError_t err;
    explicit prt_device_guard(int device)
    {
        cudaGetDevice(&prev);
        err = (prev == device) ? cudaSuccess : cudaSetDevice(device);
    }
    ~prt_device_guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define PRT_ON_DEVICE(dev)                                                                          \
    prt_device_guard guard__(dev);                                                                  \
    if (guard__.err != cudaSuccess)                                                                 \
        return fail(PRT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard__.err))

struct prt_env {
    int device;
    void* dev_blob;
    size_t bytes;
    prt_env_view view;  // device pointers
};

// ---------------------------------------------------------------------------------------------
// K1: the fan
struct prt_fan_params {
    int64_t n_rays;
    const double* angles;
    const double* y0;
    double source_depth, source_range, receiver_range, rtol, atol;
    int32_t terminate_backwards, n_save;
    const double* range_save;
    double* out_traj;
    double* out_final;
    int32_t* out_nb;
    int32_t* out_ns;
    int32_t* out_status;
    uint32_t* out_counters;
    int32_t max_bounces;
    int32_t stored_convention;
    int32_t park_wait, park_wait_long, park_long_num;
    int32_t lanes;  // rays per warp (32, 16, 8, 4, 2 or 1): small fans use partly filled warps, see pick_lanes()
    int64_t rays_per_env;  // ensembles: ray r belongs to realisation r / rays_per_env and uses angles[r % rays_per_env]
    int32_t reorder;       // 1: second half of the blocks takes its chunks in reverse order (see prt_fan_kernel)
};

// Lanes that hit a boundary wait (at most this many step attempts of the other lanes) so that the
// event root searches of a warp run together; tunable for experiments via PRT_PARK_WAIT.
static int g_park_wait = -1;
// The eigenray refinement waits longer (PRT_PARK_WAIT_REFINE, default 128): its trial rays are mostly steep, multiply
// reflecting ones, and a root-search pass costs a warp about three step attempts whether one lane or twenty take part,
// so batching more lanes per pass pays -- measured on config 4 (816 408 brackets, 22 rounds): 16 -> 3.70 s, 32 -> 3.10,
// 64 -> 3.03, 128 -> 2.89, 256 -> 3.05 s; the fans prefer 16 (10k-ray config 2: 4.85 ms against 5.27-5.67 at 64-128;
// 1M-ray fans within 2 % either way; profiles/r02_variants.md).  Interleaving only: results do not depend on it.
static int park_wait_refine_setting()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PRT_PARK_WAIT_REFINE");
        v = e ? atoi(e) : 128;
        if (v < 0) v = 0;
    }
    return v;
}
static int park_wait_setting()
{
    if (g_park_wait < 0) {
        const char* e = getenv("PRT_PARK_WAIT");
        g_park_wait = e ? atoi(e) : 16;  // measured best on B200 (profiles/r01_variants.md)
        if (g_park_wait < 0) g_park_wait = 0;
    }
    return g_park_wait;
}

// A lane whose accepted step fired an event parks until the warp's root searches run.  Only the step's end
// abscissa and its event mask are parked (2 doubles per thread in shared memory): the ray still holds the
// state at the START of that step, so the step record (stage values -> dense-output rows) is rebuilt by
// repeating the trial step -- same inputs, same arithmetic, same bits -- once per reflection.  Keeping the
// parked record this small leaves the kernel with 2 KB of shared memory per block, i.e. practically the
// whole 256 KB of an SM as L1 for the sound-speed grid (the 2001 x 1201 internal-wave grid of config 3 runs
// 3.6x slower when a third of the L1 is taken away; profiles/r01_variants.md).
#define PRT_PARK_FIELDS 2
struct prt_park_slot {
    double* base;  // &park[0][threadIdx.x]
    __device__ __forceinline__ void store(const prt_step& st) const
    {
        base[0] = st.t_new;
        base[blockDim.x] = __longlong_as_double((long long)st.active);
    }
    // rebuild the step record of the parked step of ray r (state still at the start of that step)
    template <bool SAVE, bool COUNT>
    __device__ __forceinline__ void rebuild(const prt_envp& E, const prt_shot& S, const prt_ray<SAVE, COUNT>& r, prt_step& st) const
    {
        const double tn = base[0];
        st.active = (int32_t)__double_as_longlong(base[blockDim.x]);
        st.t_old = r.t; st.t_new = tn; st.h = tn - r.t;
        st.finished = (tn - S.x_end >= 0.0);
        prt_try v;
        prt_rk_try_exact<true>(&E.v, E.grid, S.rtol, S.atol, r.t, st.h, r.T, r.z, r.p, r.f0, r.f1, r.f2, &v);
        st.qT = v.qT; st.qz = v.qz; st.qp = v.qp; st.inv_h = v.inv_h;
        st.yT = v.nT; st.yz = v.nz; st.yp = v.np_;
        st.nf0 = v.k7T; st.nf1 = v.k7z; st.nf2 = v.k7p;
        st.root = 0.0;
        st.event = -1;
    }
};

// Warp-synchronous driver of 32 rays (semantics = prt_integrate_ray for each lane; only the
// interleaving differs, so results are bit-identical):
//   * every iteration all stepping lanes make exactly ONE RK45 attempt together, accepted or not
//     (no lane re-runs the stage code alone after a rejection);
//   * a lane whose accepted step fired an event parks its step record in shared memory; the bisection
//     root search, the reflection and the next segment's start-up run once for all parked lanes, when
//     no lane is left stepping or the oldest parked lane has waited `park_wait` iterations
//     (neighbouring launch angles reach a boundary within a few steps of each other).
//   * the wait adapts: after a pass in which at least three quarters of the lanes still in play took part (a warp of
//     steep rays that reflect together) the next wait is `park_wait_long`, otherwise `park_wait` (mixed warps, where
//     parked lanes would idle while their neighbours step on).  Measured (profiles/r02_variants.md): half -> +2.2 % on
//     the config-5 ensemble, three quarters -> +1.3 %, all lanes -> 0 % but the steep fans lose most of their -32 %.
template <bool SAVE, bool COUNT>
__device__ __forceinline__ void prt_integrate_warp(const prt_envp& E, prt_ray<SAVE, COUNT>& r, const prt_shot& S,
                                                   const prt_traj_sink& sink, int park_wait, int park_wait_long,
                                                   const prt_park_slot& park, int park_long_num = 3)
{
    const unsigned full = 0xffffffffu;
    int waited = 0, limit = park_wait;
    for (;;) {
        if (r.phase == PRT_PH_INIT) r.seg_init(E, S);  // INIT -> STEP, always
        const unsigned stepping_now = __ballot_sync(full, r.phase == PRT_PH_STEP);
        if (r.phase == PRT_PH_STEP) {
            prt_step st;
            if (r.template attempt<SAVE>(E, S, st, stepping_now)) {
                if (st.active) {
                    park.store(st);
                    r.phase = PRT_PH_ROOT;
                } else {
                    r.advance(S, sink, st);
                }
            }
        }
        const unsigned parked = __ballot_sync(full, r.phase == PRT_PH_ROOT);  // (also re-converges the warp)
        if (!parked) {
            if (!stepping_now) break;  // nobody stepped, nobody is parked, no segment was waiting to start: all done
        } else {
            const unsigned stepping = __ballot_sync(full, r.phase == PRT_PH_STEP);
            ++waited;
            if (!stepping || waited > limit) {
                limit = (4 * __popc(parked) >= park_long_num * __popc(parked | stepping)) ? park_wait_long : park_wait;
                if (r.phase == PRT_PH_ROOT) {
                    prt_step st;
                    park.rebuild(E, S, r, st);
                    r.solve_roots(E, st);
                    r.finish_terminal(E, S, sink, st);
                }
                waited = 0;
            }
        }
    }
}

// Resident blocks of 128 threads per SM: 4 for every kernel (128 registers per thread; measured on B200, 1M-ray fan:
// 3 blocks / 142 regs 160 ms, 4 blocks 100 ms).  The arrival-table kernels keep no stage values across the trial step
// (prt_rk_try) and would fit a fifth block at 96 registers, but that build spills and puts 25 % more scattered loads on
// the L1 data pipe: 12-16 % SLOWER, measured twice (profiles/r02_variants.md sections 1 and 5c).
#ifndef PRT_MIN_BLOCKS
#define PRT_MIN_BLOCKS 4
#endif
#ifndef PRT_MIN_BLOCKS_ARR
#define PRT_MIN_BLOCKS_ARR 4
#endif
#ifndef PRT_BLOCK
#define PRT_BLOCK 128  // threads per block, trajectory-saving kernels
#endif
#ifndef PRT_BLOCK_ARR
#define PRT_BLOCK_ARR 128  // threads per block, arrival-table kernels (fan without save grid, eigenray trial shots)
#endif
#define PRT_BLOCK_MAX (PRT_BLOCK > PRT_BLOCK_ARR ? PRT_BLOCK : PRT_BLOCK_ARR)
template <bool SAVE, int ZM, bool COUNT>
__global__ void __launch_bounds__(SAVE ? PRT_BLOCK : PRT_BLOCK_ARR, SAVE ? PRT_MIN_BLOCKS : PRT_MIN_BLOCKS_ARR) prt_fan_kernel(const __grid_constant__ prt_env_view E0,
                                                      const __grid_constant__ prt_fan_params P)
{
    // P.lanes rays per warp (lanes >= P.lanes idle); whole warps stay in the kernel: the driver is warp-synchronous
    const int lane = threadIdx.x & 31;
    // Blocks start roughly in index order and the cost of a ray grows with |launch angle| (reflections), so a fan
    // sorted by angle would start its most expensive chunks LAST and end on a long, poorly filled tail.  The second
    // half of the blocks therefore walks the fan backwards: both halves go from steep to shallow and the launch
    // ends on the cheapest rays.  (Any order is correct; unsorted fans are no worse off than before.)
    const unsigned nb = gridDim.x, half = nb >> 1;
    const unsigned blk = (blockIdx.x < half || P.reorder == 0) ? blockIdx.x : nb - 1u - (blockIdx.x - half);
    const int64_t warp = ((int64_t)blk * blockDim.x + threadIdx.x) >> 5;
    const int64_t ray = warp * P.lanes + lane;
    const bool live = (lane < P.lanes) && (ray < P.n_rays);
    // ensembles: same tables, this ray's realisation of the sound-speed grid
    const prt_d2* grid = E0.grid;
    int64_t angle_index = ray;
    if (P.rays_per_env > 0 && live) {
        const int64_t env = ray / P.rays_per_env;
        angle_index = ray - env * P.rays_per_env;
        grid = E0.grid + env * E0.grid_stride;
    }
    const prt_envp E{E0, grid, ZM};
    prt_shot S;
    S.x_end = P.receiver_range; S.rtol = P.rtol; S.atol = P.atol;
    S.terminate_backwards = P.terminate_backwards; S.n_save = P.n_save; S.max_bounces = P.max_bounces;
    S.range_save = P.range_save;
    double T = 0.0, z = 0.0, p = 0.0;
    if (live) {
        if (P.y0) {
            T = P.y0[3 * ray];
            z = P.y0[3 * ray + 1];
            p = P.y0[3 * ray + 2];
        } else {
            prt_initial_state(E, P.source_depth, P.source_range, P.angles[angle_index], T, z, p);
        }
    }
    prt_traj_sink sink;
    sink.stride = 1;
#if defined(PRT_BOUNDS_CHECK)
    sink.n_save = P.n_save;
#endif
    sink.sgn = P.stored_convention ? -1.0 : 1.0;
    sink.T = sink.z = sink.p = nullptr;
    if (SAVE && live) {
        const size_t plane = (size_t)P.n_rays * (size_t)P.n_save;
        sink.T = P.out_traj + (size_t)ray * (size_t)P.n_save;
        sink.z = sink.T + plane;
        sink.p = sink.z + plane;
    }
    __shared__ double park_mem[PRT_PARK_FIELDS * (SAVE ? PRT_BLOCK : PRT_BLOCK_ARR)];
    prt_park_slot park;
    park.base = park_mem + threadIdx.x;
    prt_ray<SAVE, COUNT> r;
    r.start(T, z, p, P.source_range, S);
    if (!live) r.phase = PRT_PH_DONE;
    prt_integrate_warp(E, r, S, sink, P.park_wait, P.park_wait_long, park, P.park_long_num);
    if (!live) return;
    prt_ray_result res;
    uint32_t cnt[PRT_CNT_N];
    r.result(S, sink, res, cnt);
    if (SAVE && res.status != PRT_RAY_OK) {
        // the reference returns None for a dropped ray; leave no half-written trajectory behind
        for (int i = 0; i < P.n_save; ++i) sink.put(i, NAN, NAN, NAN);
    }
    P.out_final[ray] = res.T;
    P.out_final[P.n_rays + ray] = res.z;
    P.out_final[2 * P.n_rays + ray] = res.p;
    P.out_nb[ray] = res.n_bottom;
    P.out_ns[ray] = res.n_surface;
    P.out_status[ray] = res.status;
    if (COUNT) {
#pragma unroll
        for (int k = 0; k < PRT_CNT_N; ++k) P.out_counters[(size_t)k * P.n_rays + ray] = cnt[k];
    }
}

// sum(diff(sign(p)) != 0) along the saved grid, one warp per ray (ray_objects.py:142-144)
__global__ void prt_sign_changes_kernel(int64_t n_rays, int32_t n_save, const double* __restrict__ p_plane,
                                        int32_t* __restrict__ out)
{
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const double* p = p_plane + (size_t)ray * (size_t)n_save;
    int local = 0;
    for (int i = lane; i < n_save - 1; i += 32) {
        double a = p[i], b = p[i + 1];
        double sa = (double)((a > 0.0) - (a < 0.0)), sb = (double)((b > 0.0) - (b < 0.0));
        local += ((a != a) || (b != b) || (sb - sa != 0.0)) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if (lane == 0) out[ray] = local;
}

// Rays per warp.  The ray ODE is one long dependent FP64 chain (~20 cycles per DFMA on B200), so a
// warp alone on a scheduler keeps the FP64 pipe ~5 % busy whatever its lane count.  A fan too small
// to give every scheduler several full warps is therefore spread over MORE, partly filled warps
// (fewer rays per warp) until the resident-warp capacity of the chip is used: latency hiding by
// thread-level parallelism instead of idle schedulers.  Large fans use full warps.
static std::atomic<int> g_concurrency{1};

static std::atomic<int> g_lanes_override{0};

static int lanes_forced()
{
    const int ov = g_lanes_override.load();
    if (ov >= 1 && ov <= 32) return ov;
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("PRT_LANES");
        forced = e ? atoi(e) : 0;
    }
    if (forced >= 1 && forced <= 32) return forced;  // any count works: ray = warp * lanes + lane, lanes >= count idle
    return 0;
}

// warps_per_sm: resident warps per SM of the kernel about to be launched
static int pick_lanes(int64_t n_units, int device, int warps_per_sm, int64_t* capacity_out = nullptr)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t capacity = (int64_t)sms * warps_per_sm;
    if (capacity_out) *capacity_out = capacity;
    const int f = lanes_forced();
    if (f) return f;
    n_units *= g_concurrency.load();  // several fans in flight on different streams share the chip
    int lanes = 32;
    while (lanes > 2 && (n_units + lanes / 2 - 1) / (lanes / 2) <= capacity) lanes /= 2;
    return lanes;
}

// ---------------------------------------------------------------------------------------------
// K3: bracketing (eigenrays.py:62-79)
__device__ __forceinline__ double prt_sign(double v) { return (double)((v > 0.0) - (v < 0.0)); }  // np.sign (NaN -> 0 here)

// Two-level ordered compaction.  A tile is PRT_BR_TILE consecutive ray pairs (k, k + 1); one block of 256 threads
// per (receiver, tile), thread t owning the PRT_BR_PER_THREAD consecutive pairs from k0 + t * PRT_BR_PER_THREAD.
//   count:  tile_cnt[rcv][tile] = number of sign changes of the tile            (no atomics: one writer per entry)
//   scan:   one block turns tile_cnt, flattened in (receiver, tile) order, into exclusive offsets in place -- the
//           output position of every tile's first bracket, which is the (receiver, ray) order the host layer
//           promises -- and writes the per-receiver counts
//   fill:   the same tiling again; a block-wide exclusive scan of the per-thread counts places every bracket
#define PRT_BR_PER_THREAD 8
#define PRT_BR_TILE (256 * PRT_BR_PER_THREAD)

// np.sign propagates NaN: NaN - x != 0 is True, so a NaN neighbour counts as a change (eigenrays.py:66-70)
__device__ __forceinline__ bool prt_bracket_here(double za, double zb, double rd)
{
    const double a = za + rd, b = zb + rd;
    return (a != a) || (b != b) || (prt_sign(b) - prt_sign(a) != 0.0);
}

// exclusive prefix of `v` over the 256 threads of a block (returns it; `total` = block sum); scratch: 8 ints
__device__ __forceinline__ int prt_block_exscan256(int v, int* warp_tot, int& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int c = warp_tot[w];
        before += (w < warp) ? c : 0;
        tot += c;
    }
    total = tot;
    return before + incl - v;
}

__global__ void __launch_bounds__(256) prt_count_brackets_kernel(int64_t n_rays, const double* __restrict__ z_end,
                                                                 const double* __restrict__ rds, int32_t n_tiles,
                                                                 int64_t* __restrict__ tile_cnt)
{
    const int rcv = blockIdx.x, tile = blockIdx.y;
    const double rd = rds[rcv];
    const int64_t k0 = (int64_t)tile * PRT_BR_TILE + (int64_t)threadIdx.x * PRT_BR_PER_THREAD;
    int local = 0;
    if (k0 < n_rays - 1) {
        double za = z_end[k0];
#pragma unroll
        for (int j = 0; j < PRT_BR_PER_THREAD; ++j) {
            const int64_t k = k0 + j;
            if (k < n_rays - 1) {
                const double zb = z_end[k + 1];
                local += prt_bracket_here(za, zb, rd) ? 1 : 0;
                za = zb;
            }
        }
    }
    __shared__ int warp_tot[8];
    int total;
    (void)prt_block_exscan256(local, warp_tot, total);
    if (threadIdx.x == 0) tile_cnt[(int64_t)rcv * n_tiles + tile] = total;
}

__global__ void __launch_bounds__(1024) prt_scan_brackets_kernel(int64_t n_items, int32_t n_tiles, int32_t n_rcv,
                                                                 int64_t* __restrict__ tile_cnt,
                                                                 int32_t* __restrict__ counts)
{
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_items; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long v = (i < n_items) ? tile_cnt[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = carry, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const long long c = warp_tot[w];
            before += (w < warp) ? c : 0;
            tot += c;
        }
        if (i < n_items) tile_cnt[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    // per-receiver counts = difference of the offsets of consecutive receivers' first tiles
    for (int r = threadIdx.x; r < n_rcv; r += 1024) {
        const long long lo = tile_cnt[(int64_t)r * n_tiles];
        const long long hi = (r + 1 < n_rcv) ? tile_cnt[(int64_t)(r + 1) * n_tiles] : carry;
        counts[r] = (int32_t)(hi - lo);
    }
}

__global__ void __launch_bounds__(256) prt_fill_brackets_kernel(int64_t n_rays, const double* __restrict__ z_end,
                                                                const double* __restrict__ thetas,
                                                                const double* __restrict__ rds, int32_t n_tiles,
                                                                const int64_t* __restrict__ tile_off,
                                                                double* __restrict__ z1, double* __restrict__ z2,
                                                                double* __restrict__ th1, double* __restrict__ th2,
                                                                double* __restrict__ brd, int32_t* __restrict__ brcv,
                                                                int64_t* __restrict__ bk)
{
    const int rcv = blockIdx.x, tile = blockIdx.y;
    const double rd = rds[rcv];
    const int64_t k0 = (int64_t)tile * PRT_BR_TILE + (int64_t)threadIdx.x * PRT_BR_PER_THREAD;
    double z[PRT_BR_PER_THREAD + 1];
    unsigned mask = 0;
    if (k0 < n_rays - 1) {
        z[0] = z_end[k0];
#pragma unroll
        for (int j = 0; j < PRT_BR_PER_THREAD; ++j) {
            const int64_t k = k0 + j;
            z[j + 1] = 0.0;
            if (k < n_rays - 1) {
                z[j + 1] = z_end[k + 1];
                mask |= prt_bracket_here(z[j], z[j + 1], rd) ? (1u << j) : 0u;
            }
        }
    }
    __shared__ int warp_tot[8];
    int total;
    const int before = prt_block_exscan256(__popc(mask), warp_tot, total);
    if (!mask) return;
    int64_t o = tile_off[(int64_t)rcv * n_tiles + tile] + before;
#pragma unroll
    for (int j = 0; j < PRT_BR_PER_THREAD; ++j) {
        if (mask & (1u << j)) {
            const int64_t k = k0 + j;
            PRT_CHECK_IDX(k + 1, n_rays);
            z1[o] = z[j];
            z2[o] = z[j + 1];
            th1[o] = thetas[k];
            th2[o] = thetas[k + 1];
            brd[o] = rd;
            brcv[o] = rcv;
            bk[o] = k;
            ++o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: regula falsi, one trial shot per active bracket per round (eigenrays.py:206-268)
struct prt_refine_state {
    double *z1, *z2, *th1, *th2, *th;  // [n_br]
    int32_t *iter, *shots;             // [n_br]
    int32_t *list_a, *list_b;          // active bracket ids (ping-pong), always in ascending bracket order
    int32_t* keep;                     // [n_br] per list slot: 1 = the bracket goes on to the next round
    int32_t* blk;                      // [ceil(n_br / PRT_CP_TILE)] survivors per tile of list slots, then their offsets
    int32_t* counters;                 // [2] active counts (ping-pong)
};

struct prt_refine_params {
    int64_t n_br;
    const double* rd;
    double source_depth, source_range, receiver_range, rtol, atol, ztol;
    int32_t terminate_backwards, max_iter, max_bounces, park_wait;
    int32_t lanes;        // > 0: rays per warp fixed by the caller (prt_set_lanes); 0: chosen from the active count
    int32_t warp_budget;  // warps the round may spread the active brackets over (resident capacity, launched warps)
    double* out_theta;
    double* out_final;
    int32_t *out_nb, *out_ns, *out_status, *out_shots, *out_ray_status;
};

__global__ void prt_refine_init_kernel(prt_refine_state S, int64_t n_br, const double* __restrict__ z1,
                                       const double* __restrict__ z2, const double* __restrict__ th1,
                                       const double* __restrict__ th2, const double* __restrict__ rd,
                                       const int32_t* __restrict__ order, prt_refine_params P)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) {
        S.counters[0] = (int32_t)n_br;
        S.counters[1] = 0;
    }
    if (b >= n_br) return;
    S.z1[b] = z1[b];
    S.z2[b] = z2[b];
    S.th1[b] = th1[b];
    S.th2[b] = th2[b];
    // regula_falsi_thetas, eigenrays.py:118-120
    S.th[b] = th1[b] - (z1[b] + rd[b]) * (th2[b] - th1[b]) / (z2[b] - z1[b]);
    S.iter[b] = 0;
    S.shots[b] = 0;
    S.list_a[b] = order ? order[b] : (int32_t)b;  // processing order: slot b of the first round
    P.out_theta[b] = NAN;
    P.out_final[b] = NAN;
    P.out_final[n_br + b] = NAN;
    P.out_final[2 * n_br + b] = NAN;
    P.out_nb[b] = 0;
    P.out_ns[b] = 0;
    P.out_status[b] = PRT_EIG_MAX_ITER;
    P.out_shots[b] = 0;
    P.out_ray_status[b] = 0;
}

template <int ZM>
__global__ void __launch_bounds__(PRT_BLOCK_ARR, PRT_MIN_BLOCKS_ARR) prt_refine_round_kernel(const __grid_constant__ prt_env_view E0,
                                                               const __grid_constant__ prt_refine_params P,
                                                               prt_refine_state S, const int32_t* __restrict__ list_in,
                                                               const int32_t* n_in)
{
    const prt_envp E{E0, E0.grid, ZM};
    const int n_active = *n_in;
    // rays per warp: the rule of pick_lanes() on the number of brackets still active, which only the device knows
    // (the host never waits for a round): few survivors are spread over more, partly filled warps
    int lanes = P.lanes;
    if (lanes == 0) {
        lanes = 32;
        while (lanes > 2 && (n_active + (lanes >> 1) - 1) / (lanes >> 1) <= P.warp_budget) lanes >>= 1;
    }
    const int lane = threadIdx.x & 31;
    const int idx = (int)(((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * lanes + lane);
    const bool live = (lane < lanes) && (idx < n_active);
    const int b = live ? list_in[idx] : 0;
    PRT_CHECK_IDX(b, P.n_br);
    const double rd = live ? P.rd[b] : 0.0;
    const double th = live ? S.th[b] : 0.0;
    prt_shot shot;
    shot.x_end = P.receiver_range; shot.rtol = P.rtol; shot.atol = P.atol;
    shot.terminate_backwards = P.terminate_backwards; shot.n_save = 0; shot.max_bounces = P.max_bounces;
    shot.range_save = nullptr;
    double T = 0.0, z = 0.0, p = 0.0;
    if (live) prt_initial_state(E, P.source_depth, P.source_range, -th, T, z, p);  // shoot_ray negates, launch_rays.py:251
    prt_traj_sink sink;
    sink.T = sink.z = sink.p = nullptr;
    sink.stride = 1;
#if defined(PRT_BOUNDS_CHECK)
    sink.n_save = 0;
#endif
    sink.sgn = 1.0;
    __shared__ double park_mem[PRT_PARK_FIELDS * PRT_BLOCK_ARR];
    prt_park_slot park;
    park.base = park_mem + threadIdx.x;
    prt_ray<false, false> r;
    r.start(T, z, p, P.source_range, shot);
    if (!live) r.phase = PRT_PH_DONE;
    prt_integrate_warp(E, r, shot, sink, P.park_wait, P.park_wait, park);
    if (!live) return;
    prt_ray_result res;
    r.result(shot, sink, res, nullptr);
    const int shots = ++S.shots[b];
    P.out_shots[b] = shots;
    S.keep[idx] = 0;  // list slot idx: set to 1 below iff the bracket needs another trial shot
    if (res.status != PRT_RAY_OK) {  // ray is None, eigenrays.py:241-245
        P.out_status[b] = PRT_EIG_RAY_DROPPED;
        P.out_ray_status[b] = res.status;
        return;
    }
    const double zend = -res.z;  // ray.z[-1] (negative-z convention)
    if (fabs(zend + rd) < P.ztol) {  // eigenrays.py:247-250
        P.out_status[b] = PRT_EIG_CONVERGED;
        P.out_theta[b] = th;
        P.out_final[b] = res.T;
        P.out_final[P.n_br + b] = res.z;
        P.out_final[2 * P.n_br + b] = res.p;
        P.out_nb[b] = res.n_bottom;
        P.out_ns[b] = res.n_surface;
        return;
    }
    double z1 = S.z1[b], z2 = S.z2[b], th1 = S.th1[b], th2 = S.th2[b];
    if (prt_sign(zend + rd) == prt_sign(z1 + rd)) {  // eigenrays.py:253-259
        z1 = zend;
        th1 = th;
        S.z1[b] = z1;
        S.th1[b] = th1;
    } else {
        z2 = zend;
        th2 = th;
        S.z2[b] = z2;
        S.th2[b] = th2;
    }
    S.th[b] = th1 - (z1 + rd) * (th2 - th1) / (z2 - z1);  // eigenrays.py:261-263
    const int it = S.iter[b];
    if (it > P.max_iter) {  // eigenrays.py:265-266
        P.out_status[b] = PRT_EIG_MAX_ITER;
        return;
    }
    S.iter[b] = it + 1;
    S.keep[idx] = 1;
}

// Order-preserving compaction of the active list between rounds (count per tile of list slots -> one-block scan ->
// scatter): the survivors keep their (receiver, ray) order, so a warp of the next round still holds neighbouring
// launch angles of one receiver -- rays that take the same steps through the same grid cells.  (An atomic append
// shuffled the list after round 0: warps of unrelated angles, 2x slower trial shots; profiles/r02_variants.md.)
#define PRT_CP_PER_THREAD 4
#define PRT_CP_TILE (256 * PRT_CP_PER_THREAD)
__global__ void __launch_bounds__(256) prt_compact_count_kernel(const int32_t* __restrict__ keep, const int32_t* n_in,
                                                                int32_t* __restrict__ blk)
{
    const int n = *n_in;
    const int s0 = blockIdx.x * PRT_CP_TILE + threadIdx.x * PRT_CP_PER_THREAD;
    int local = 0;
#pragma unroll
    for (int j = 0; j < PRT_CP_PER_THREAD; ++j) local += (s0 + j < n && keep[s0 + j]) ? 1 : 0;
    __shared__ int warp_tot[8];
    int total;
    (void)prt_block_exscan256(local, warp_tot, total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) prt_compact_scan_kernel(int32_t n_blk, int32_t* __restrict__ blk, int32_t* n_out)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n_blk) ? blk[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = carry, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = warp_tot[w];
            before += (w < warp) ? c : 0;
            tot += c;
        }
        if (i < n_blk) blk[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

__global__ void __launch_bounds__(256) prt_compact_scatter_kernel(const int32_t* __restrict__ keep, const int32_t* n_in,
                                                                  const int32_t* __restrict__ blk,
                                                                  const int32_t* __restrict__ list_in,
                                                                  int32_t* __restrict__ list_out)
{
    const int n = *n_in;
    const int s0 = blockIdx.x * PRT_CP_TILE + threadIdx.x * PRT_CP_PER_THREAD;
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < PRT_CP_PER_THREAD; ++j) mask |= (s0 + j < n && keep[s0 + j]) ? (1u << j) : 0u;
    __shared__ int warp_tot[8];
    int total;
    int o = blk[blockIdx.x] + prt_block_exscan256(__popc(mask), warp_tot, total);
#pragma unroll
    for (int j = 0; j < PRT_CP_PER_THREAD; ++j)
        if (mask & (1u << j)) list_out[o++] = list_in[s0 + j];
}

// arrival records [n][6] = T, z, p, n_bottom, n_surface, status (the unit of the multi-GPU all-gather)
__global__ void prt_pack_arrivals_kernel(int64_t n, const double* __restrict__ fin, const int32_t* __restrict__ nb,
                                         const int32_t* __restrict__ ns, const int32_t* __restrict__ st,
                                         double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* o = out + 6 * i;
    o[0] = fin[i];
    o[1] = fin[n + i];
    o[2] = fin[2 * n + i];
    o[3] = (double)nb[i];
    o[4] = (double)ns[i];
    o[5] = (double)st[i];
}

// ---------------------------------------------------------------------------------------------
// Eigenray post-processing: the per-receiver arrival ("time front") table of the converged eigenrays, built on the
// device from the refinement output -- EigenRays.__init__ of the reference loops over rays on the host
// (ray_objects.py:505-548).  The received angle is ray_angle (integration_processes.py:306-334) exactly as that
// constructor calls it: on the STORED end state (z = -depth, p sign-flipped) and on the UN-flattened sound speed
// `environment.sound_speed`, i.e. literal bilinear_interp (:152-174: cell = searchsorted - 1 clamped
// to [0, n - 2], unclamped weights, four-term sum), theta = degrees(arcsin(p c)).
__device__ __forceinline__ int prt_searchsorted_cell(const double* __restrict__ a, int n, double v)
{
    int lo = 0, hi = n;  // np.searchsorted(a, v, side="left"); NaN sorts last
    if (v != v) lo = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    int k = lo - 1;
    k = k < 0 ? 0 : k;
    return k > n - 2 ? n - 2 : k;
}

__device__ __forceinline__ double prt_received_angle(const double* __restrict__ rin, const double* __restrict__ zin,
                                                     const double* __restrict__ cin, int n_r, int n_z, double x,
                                                     double z, double p)
{
    const int i = prt_searchsorted_cell(rin, n_r, x), j = prt_searchsorted_cell(zin, n_z, z);
    const double wx = (x - rin[i]) / (rin[i + 1] - rin[i]);
    const double wy = (z - zin[j]) / (zin[j + 1] - zin[j]);
    const double* c0 = cin + (size_t)i * n_z + j;
    const double* c1 = c0 + n_z;
    const double c = (1.0 - wx) * (1.0 - wy) * c0[0] + wx * (1.0 - wy) * c1[0] + (1.0 - wx) * wy * c0[1] + wx * wy * c1[1];
    return prt_asin(p * c) * PRT_RAD2DEG;
}

__global__ void prt_received_angles_kernel(int64_t n, double x, const double* __restrict__ z, const double* __restrict__ p,
                                           int32_t n_r, int32_t n_z, const double* __restrict__ rin,
                                           const double* __restrict__ zin, const double* __restrict__ cin,
                                           double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = prt_received_angle(rin, zin, cin, n_r, n_z, x, z[i], p[i]);
}

// one row per converged eigenray: PRT_EIG_TABLE_WIDTH doubles (include/prt.h)
__global__ void prt_eigen_table_kernel(int64_t n, int64_t n_br, const int64_t* __restrict__ idx,
                                       const int32_t* __restrict__ rcv, const double* __restrict__ theta,
                                       const double* __restrict__ fin, const int32_t* __restrict__ nb,
                                       const int32_t* __restrict__ ns, const int32_t* __restrict__ flips, double x,
                                       int32_t n_r, int32_t n_z, const double* __restrict__ rin,
                                       const double* __restrict__ zin, const double* __restrict__ cin,
                                       double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t b = idx[i];
    const double th = theta[b];
    const double zs = -fin[n_br + b], ps = -fin[2 * n_br + b];  // stored convention (ray_objects.py:49-52)
    double* o = out + PRT_EIG_TABLE_WIDTH * i;
    o[0] = (double)rcv[b];
    o[1] = fin[b];
    o[2] = zs;
    o[3] = ps;
    o[4] = th;
    o[5] = prt_received_angle(rin, zin, cin, n_r, n_z, x, zs, ps);
    // ray id number = (# sign changes of p on the save grid) * sign(launch angle)  (ray_objects.py:533-535)
    o[6] = flips ? (double)flips[i] * prt_sign(th) : (double)NAN;
    o[7] = (double)nb[b];
    o[8] = (double)ns[b];
    o[9] = (double)b;
}

// ---------------------------------------------------------------------------------------------
// Device-side staging of sound-speed grids that already live in HBM (perturbed realisations of a tomography
// ensemble generated on the GPU): grid[k] = {c, dc/dz} with dc/dz = np.gradient(c, z, axis=depth, edge_order=1),
// the derivative _unpack_envi takes (launch_rays.py:728-733); same operations in numpy's order, hence its bits.
__global__ void prt_stage_grid_kernel(int64_t n_nodes, int32_t n_z, const double* __restrict__ c, int uniform,
                                      double two_dz, double dz_first, double dz_last, const double* __restrict__ ga,
                                      const double* __restrict__ gb, const double* __restrict__ gc,
                                      prt_d2* __restrict__ grid)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    const int j = (int)(k % n_z);
    const double f = c[k];
    double d;
    if (j == 0) {
        d = __ddiv_rn(__dsub_rn(c[k + 1], f), dz_first);
    } else if (j == n_z - 1) {
        d = __ddiv_rn(__dsub_rn(f, c[k - 1]), dz_last);
    } else if (uniform) {
        d = __ddiv_rn(__dsub_rn(c[k + 1], c[k - 1]), two_dz);
    } else {
        d = __dadd_rn(__dadd_rn(__dmul_rn(ga[j], c[k - 1]), __dmul_rn(gb[j], f)), __dmul_rn(gc[j], c[k + 1]));
    }
    prt_d2 v;
    v.x = f;
    v.y = d;
    grid[k] = v;
}

// ---------------------------------------------------------------------------------------------
// Earth flattening of sound-speed grids that live in HBM (the preprocessing step right before the path; per
// realisation in ensemble mode).  eflat (environment.py:371-401): E = z / R_e(lat), z_f = z (1 + E (1/2 + E / 3)),
// c_f = c (1 + E (1 + E)).
//   resample = 0: single-latitude transform (flat_earth_transform, :121-151): out = c_f on the stretched grid z_f
//                 (the caller stretches the depth coordinate once).
//   resample = 1: range-dependent transform (flat_earth_c, :239-368): every range column has its own earth radius
//                 re[r]; c_f on z_f is interpolated back onto the ORIGINAL depth grid, linearly, as
//                 xarray.DataArray.interp does through scipy interp1d (slope = (y_hi - y_lo) / (x_hi - x_lo),
//                 y = slope (x - x_lo) + y_lo, interval by searchsorted clipped to [1, n - 1], NaN outside).
__global__ void prt_flat_earth_kernel(int64_t n_nodes, int32_t n_r, int32_t n_z, const double* __restrict__ depth,
                                      const double* __restrict__ re, const double* __restrict__ c, int resample,
                                      double* __restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    const int j = (int)(k % n_z);
    const int r = (int)((k / n_z) % n_r);
    const double R = re[r];
    const double* col = c + (k - j);
    auto zf = [&](int i) { const double E = depth[i] / R; return depth[i] * (1.0 + E * (0.50 + E / 3.0)); };
    auto cf = [&](int i) { const double E = depth[i] / R; return col[i] * (1.0 + E * (1.0 + E)); };
    if (!resample) {
        out[k] = cf(j);
        return;
    }
    const double x = depth[j];
    // hi = searchsorted(z_f, x) clipped to [1, n_z - 1]; z_f[j] >= depth[j] for depth >= 0, so scan down from j
    // (general input: scan up first)
    int hi = j;
    while (hi < n_z - 1 && zf(hi) < x) ++hi;
    while (hi > 1 && !(zf(hi - 1) < x)) --hi;
    hi = hi < 1 ? 1 : hi;
    const int lo = hi - 1;
    const double x_lo = zf(lo), x_hi = zf(hi), y_lo = cf(lo), y_hi = cf(hi);
    const double slope = (y_hi - y_lo) / (x_hi - x_lo);
    double y = slope * (x - x_lo) + y_lo;
    if (!(x >= zf(0)) || !(x <= zf(n_z - 1))) y = NAN;  // bounds_error=False, fill_value=nan
    out[k] = y;
}

// ---------------------------------------------------------------------------------------------
// Tomography ensembles: perturbed sound-speed realisations generated in HBM (the
// reference has no generator).  Separable internal-wave-like field
//   dc_e(r, z) = s_e * sum_j V[j][z] * H_e[j][r],   H_e[j][r] = sum_k cos(kappa_k r + phi_e[j][k]),
// V = amplitude-weighted vertical modes (host, [J][n_z]); s_e normalises the rms over range at depth index iax to
// rms_axis.  out_e = c_base + dc_e.  (Plain CUDA libm cos: this defines inputs, it is not on the parity path.)
__global__ void prt_iw_horizontal_kernel(int64_t n, int32_t n_r, int32_t K, const double* __restrict__ r,
                                         const double* __restrict__ kappa, const double* __restrict__ phases,
                                         double* __restrict__ H)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (e * J + j) * n_r + ir
    if (i >= n) return;
    const int ir = (int)(i % n_r);
    const double* ph = phases + (i / n_r) * K;
    const double x = r[ir];
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc += cos(kappa[k] * x + ph[k]);
    H[i] = acc;
}

__global__ void prt_iw_scale_kernel(int32_t n_r, int32_t n_z, int32_t J, int32_t iax, double rms_axis,
                                    const double* __restrict__ V, const double* __restrict__ H, double* __restrict__ scale)
{
    const int e = blockIdx.x;
    const double* He = H + (size_t)e * J * n_r;
    double acc = 0.0;
    for (int ir = threadIdx.x; ir < n_r; ir += blockDim.x) {
        double v = 0.0;
        for (int j = 0; j < J; ++j) v += V[(size_t)j * n_z + iax] * He[(size_t)j * n_r + ir];
        acc += v * v;
    }
    __shared__ double part[256];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) scale[e] = rms_axis / sqrt(part[0] / (double)n_r);
}

__global__ void prt_iw_compose_kernel(int64_t n_nodes, int32_t n_r, int32_t n_z, int32_t J, const double* __restrict__ V,
                                      const double* __restrict__ H, const double* __restrict__ scale,
                                      const double* __restrict__ c_base, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (e * n_r + ir) * n_z + iz
    if (i >= n_nodes) return;
    const int iz = (int)(i % n_z);
    const int64_t er = i / n_z;
    const int ir = (int)(er % n_r);
    const int64_t e = er / n_r;
    const double* He = H + (size_t)e * J * n_r + ir;
    double acc = 0.0;
    for (int j = 0; j < J; ++j) acc += V[(size_t)j * n_z + iz] * He[(size_t)j * n_r];
    out[i] = c_base[(size_t)ir * n_z + iz] + scale[e] * acc;
}

// ---------------------------------------------------------------------------------------------
// FP64 FMA-pipe peak: 8 independent DFMA chains per thread
__global__ void prt_fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// C ABI
extern "C" {

int prt_abi_version(void) { return PRT_ABI_VERSION; }
#if defined(PRT_BOUNDS_CHECK)
PRT_API long long prt_debug_oob(void)
{
    unsigned long long v = 0;
    cudaMemcpyFromSymbol(&v, prt_dbg_oob_count, sizeof v);
    return (long long)v;
}
#endif
#if defined(PRT_COUNT_COLD)
PRT_API long long prt_debug_cold(void)
{
    unsigned long long v = 0;
    cudaMemcpyFromSymbol(&v, prt_dbg_cold, sizeof v);
    return (long long)v;
}
#endif
int prt_set_concurrency(int n_fans)
{
    g_concurrency.store(n_fans > 0 ? n_fans : 1);
    return PRT_SUCCESS;
}
int prt_set_lanes(int lanes)
{
    g_lanes_override.store(lanes);
    return PRT_SUCCESS;
}
const char* prt_last_error(void) { return g_last_error.c_str(); }
int64_t prt_launch_count(void) { return g_launches.load(); }

int prt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int prt_env_create(int device, int32_t n_r, int32_t n_z, const double* rin, const double* zin, const double* cin,
                   const double* cpin, int32_t n_b, const double* depth_ranges, const double* depths, int32_t n_ba,
                   const double* ba_x, const double* ba_c, prt_env_t** out)
{
    if (!out || !rin || !zin || !cin || !cpin || !depth_ranges || !depths || !ba_x || !ba_c)
        return fail(PRT_ERR_ARG, "prt_env_create: null pointer");
    if (prt_device_count() <= device || device < 0)
        return fail(PRT_ERR_NO_DEVICE, "prt_env_create: no such CUDA device (this library has no CPU path)");
    prt_env_blob blob;
    if (prt_env_pack(n_r, n_z, rin, zin, cin, cpin, n_b, depth_ranges, depths, n_ba, ba_x, ba_c, blob))
        return fail(PRT_ERR_ARG, "prt_env_create: every grid needs at least 2 points");
    PRT_ON_DEVICE(device);
    void* dev = nullptr;
    PRT_CUDA(cudaMalloc(&dev, blob.bytes.size()));
    cudaError_t e = cudaMemcpy(dev, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return fail(PRT_ERR_CUDA, std::string("cudaMemcpy(env): ") + cudaGetErrorString(e));
    }
    prt_env* env = new prt_env;
    env->device = device;
    env->dev_blob = dev;
    env->bytes = blob.bytes.size();
    env->view = prt_env_rebase(blob.view, dev);
    *out = env;
    return PRT_SUCCESS;
}

int prt_ensemble_create(int device, int32_t n_env, int32_t n_r, int32_t n_z, const double* rin, const double* zin,
                        const double* cin, const double* cpin, int32_t n_b, const double* depth_ranges,
                        const double* depths, int32_t n_ba, const double* ba_x, const double* ba_c, prt_env_t** out)
{
    if (!out || !rin || !zin || !cin || !cpin || !depth_ranges || !depths || !ba_x || !ba_c)
        return fail(PRT_ERR_ARG, "prt_ensemble_create: null pointer");
    if (prt_device_count() <= device || device < 0)
        return fail(PRT_ERR_NO_DEVICE, "prt_ensemble_create: no such CUDA device (this library has no CPU path)");
    prt_env_blob blob;
    if (prt_env_pack(n_r, n_z, rin, zin, cin, cpin, n_b, depth_ranges, depths, n_ba, ba_x, ba_c, blob, n_env))
        return fail(PRT_ERR_ARG, "prt_ensemble_create: n_env >= 1 and every grid needs at least 2 points");
    PRT_ON_DEVICE(device);
    void* dev = nullptr;
    PRT_CUDA(cudaMalloc(&dev, blob.bytes.size()));
    cudaError_t e = cudaMemcpy(dev, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return fail(PRT_ERR_CUDA, std::string("cudaMemcpy(ensemble): ") + cudaGetErrorString(e));
    }
    prt_env* env = new prt_env;
    env->device = device;
    env->dev_blob = dev;
    env->bytes = blob.bytes.size();
    env->view = prt_env_rebase(blob.view, dev);
    *out = env;
    return PRT_SUCCESS;
}

int prt_ensemble_create_dev(int device, int32_t n_env, int32_t n_r, int32_t n_z, const double* rin, const double* zin,
                            const double* c_dev, int32_t n_b, const double* depth_ranges, const double* depths,
                            int32_t n_ba, const double* ba_x, const double* ba_c, void* stream, prt_env_t** out)
{
    if (!out || !rin || !zin || !c_dev || !depth_ranges || !depths || !ba_x || !ba_c)
        return fail(PRT_ERR_ARG, "prt_ensemble_create_dev: null pointer");
    if (prt_device_count() <= device || device < 0)
        return fail(PRT_ERR_NO_DEVICE, "prt_ensemble_create_dev: no such CUDA device (this library has no CPU path)");
    prt_env_blob blob;
    if (prt_env_pack(n_r, n_z, rin, zin, nullptr, nullptr, n_b, depth_ranges, depths, n_ba, ba_x, ba_c, blob, n_env))
        return fail(PRT_ERR_ARG, "prt_ensemble_create_dev: n_env >= 1 and every grid needs at least 2 points");
    PRT_ON_DEVICE(device);
    const prt_gradient_tables g = prt_gradient_coefficients(zin, n_z);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char* dev = nullptr;
    double* coef = nullptr;
    PRT_CUDA(cudaMalloc(&dev, blob.total_bytes));
    cudaError_t e = cudaMalloc(&coef, sizeof(double) * 3 * (size_t)n_z);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dev + blob.grid_bytes, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(coef, g.a.data(), sizeof(double) * n_z, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(coef + n_z, g.b.data(), sizeof(double) * n_z, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(coef + 2 * (size_t)n_z, g.c.data(), sizeof(double) * n_z, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        const int64_t n_nodes = (int64_t)n_env * n_r * n_z;
        prt_stage_grid_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(
            n_nodes, n_z, c_dev, g.uniform, g.two_dz, g.dz_first, g.dz_last, coef, coef + n_z, coef + 2 * (size_t)n_z,
            reinterpret_cast<prt_d2*>(dev));
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the host tables and `coef` go out of scope
    cudaFree(coef);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return fail(PRT_ERR_CUDA, std::string("prt_ensemble_create_dev: ") + cudaGetErrorString(e));
    }
    prt_env* env = new prt_env;
    env->device = device;
    env->dev_blob = dev;
    env->bytes = blob.total_bytes;
    env->view = prt_env_rebase(blob.view, dev);
    *out = env;
    return PRT_SUCCESS;
}

int prt_env_destroy(prt_env_t* env)
{
    if (!env) return PRT_SUCCESS;
    prt_device_guard guard(env->device);
    cudaFree(env->dev_blob);
    delete env;
    return PRT_SUCCESS;
}

int64_t prt_env_device_bytes(const prt_env_t* env) { return env ? (int64_t)env->bytes : 0; }

}  // extern "C"

static int shoot_fan_impl(const prt_env_t* env, int64_t n_rays, int64_t rays_per_env, const double* angles_dev,
                          const double* y0_dev, double source_depth, double source_range, double receiver_range,
                          double rtol, double atol, int32_t terminate_backwards, int32_t n_save,
                          const double* range_save_dev, double* out_traj_dev, double* out_final_dev,
                          int32_t* out_n_bottom_dev, int32_t* out_n_surface_dev, int32_t* out_status_dev,
                          uint32_t* out_counters_dev, int32_t stored_convention, int32_t* out_p_sign_changes_dev,
                          int32_t max_bounces, void* stream)
{
    if (!env) return fail(PRT_ERR_ARG, "prt_shoot_fan: null environment");
    if (n_rays < 0) return fail(PRT_ERR_ARG, "prt_shoot_fan: n_rays < 0");
    if (n_rays == 0) return PRT_SUCCESS;
    if (!angles_dev && !y0_dev) return fail(PRT_ERR_ARG, "prt_shoot_fan: need angles_dev or y0_dev");
    if (rays_per_env == 0 && env->view.n_env > 1)
        return fail(PRT_ERR_ARG, "prt_shoot_fan: this handle is an ensemble, use prt_shoot_ensemble");
    if (!out_final_dev || !out_n_bottom_dev || !out_n_surface_dev || !out_status_dev)
        return fail(PRT_ERR_ARG, "prt_shoot_fan: null output pointer");
    const bool save = (n_save > 0);
    if (save && (!range_save_dev || !out_traj_dev))
        return fail(PRT_ERR_ARG, "prt_shoot_fan: range_save_dev and out_traj_dev required to save");
    if (!(source_range < receiver_range))
        return fail(PRT_ERR_ARG, "prt_shoot_fan: source_range must be < receiver_range (mirror backwards shots)");
    PRT_ON_DEVICE(env->device);
    prt_fan_params P;
    P.n_rays = n_rays; P.angles = angles_dev; P.y0 = y0_dev;
    P.source_depth = source_depth; P.source_range = source_range; P.receiver_range = receiver_range;
    P.rtol = rtol; P.atol = atol; P.terminate_backwards = terminate_backwards; P.n_save = save ? n_save : 0;
    P.range_save = range_save_dev; P.out_traj = out_traj_dev; P.out_final = out_final_dev;
    P.out_nb = out_n_bottom_dev; P.out_ns = out_n_surface_dev; P.out_status = out_status_dev;
    P.out_counters = out_counters_dev; P.max_bounces = max_bounces > 0 ? max_bounces : (1 << 20);
    P.stored_convention = stored_convention;
    P.park_wait = park_wait_setting();
    {
        // wait of warps whose lanes reflect together (PRT_PARK_WAIT_LONG, default 128; < 0: same as park_wait).
        // Measured (profiles/r02_variants.md): a +-40 deg 10k-ray fan 26.9 -> 18.3 ms, the config-5 ensemble +1.3 %,
        // everything else within 0.5 %.
        static int lw = -2;
        if (lw == -2) {
            const char* e = getenv("PRT_PARK_WAIT_LONG");
            lw = e ? atoi(e) : 128;
        }
        P.park_wait_long = lw >= 0 ? lw : P.park_wait;
        static int num = -1;  // quarters of the lanes in play that must have taken part (PRT_PARK_LONG_NUM, 1..4)
        if (num < 0) {
            const char* e = getenv("PRT_PARK_LONG_NUM");
            num = e ? atoi(e) : 3;
            if (num < 1 || num > 4) num = 3;
        }
        P.park_long_num = num;
    }
    P.rays_per_env = rays_per_env;
    {
        static int reorder = -1;  // PRT_REORDER=0 switches the block order off (experiments)
        if (reorder < 0) {
            const char* e = getenv("PRT_REORDER");
            reorder = e ? atoi(e) : 1;
        }
        P.reorder = reorder;
    }
    if (P.rtol < 100.0 * PRT_EPS) P.rtol = 100.0 * PRT_EPS;  // validate_tol (scipy common.py:44-60)
    const int block = save ? PRT_BLOCK : PRT_BLOCK_ARR;
    P.lanes = pick_lanes(n_rays, env->device, (save ? PRT_MIN_BLOCKS : PRT_MIN_BLOCKS_ARR) * (block / 32));
    const int64_t n_warps = (n_rays + P.lanes - 1) / P.lanes;
    const unsigned grid = (unsigned)((n_warps * 32 + block - 1) / block);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int zm = prt_z_mode(env->view);  // depth-axis variant compiled in: plain / refined guess / exactly uniform grid
    const bool cnt = out_counters_dev != nullptr;
    // experiment hook: PRT_SMEM_CARVEOUT=<percent> passes a shared-memory carve-out preference for the fan kernels
    // (they use 3 KB per block; the driver's default choice is a 32 KB configuration, profiles/r02_fan_cfg3_1M_final_raw.csv)
    static int carve = -2;
    if (carve == -2) {
        const char* e = getenv("PRT_SMEM_CARVEOUT");
        carve = e ? atoi(e) : -1;
    }
#define PRT_FAN_LAUNCH(SV, Z, CN)                                                                                      \
    do {                                                                                                               \
        if (carve >= 0)                                                                                                \
            cudaFuncSetAttribute(prt_fan_kernel<SV, Z, CN>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);    \
        prt_fan_kernel<SV, Z, CN><<<grid, block, 0, s>>>(env->view, P);                                               \
    } while (0)
#define PRT_FAN_PICK_CN(SV, Z) do { if (cnt) PRT_FAN_LAUNCH(SV, Z, true); else PRT_FAN_LAUNCH(SV, Z, false); } while (0)
#define PRT_FAN_PICK_Z(SV) do { if (zm == 3) PRT_FAN_PICK_CN(SV, 3); else if (zm == 2) PRT_FAN_PICK_CN(SV, 2); else if (zm == 1) PRT_FAN_PICK_CN(SV, 1); else PRT_FAN_PICK_CN(SV, 0); } while (0)
    if (save) PRT_FAN_PICK_Z(true);
    else PRT_FAN_PICK_Z(false);
#undef PRT_FAN_PICK_Z
#undef PRT_FAN_PICK_CN
#undef PRT_FAN_LAUNCH
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    if (out_p_sign_changes_dev) {
        if (!save) return fail(PRT_ERR_ARG, "prt_shoot_fan: out_p_sign_changes_dev needs n_save > 0");
        const double* p_plane = out_traj_dev + 2 * (size_t)n_rays * (size_t)n_save;
        const unsigned g2 = (unsigned)((n_rays * 32 + 255) / 256);
        prt_sign_changes_kernel<<<g2, 256, 0, s>>>(n_rays, n_save, p_plane, out_p_sign_changes_dev);
        g_launches.fetch_add(1);
        PRT_CUDA(cudaGetLastError());
    }
    return PRT_SUCCESS;
}

extern "C" {

int prt_shoot_fan(const prt_env_t* env, int64_t n_rays, const double* angles_dev, const double* y0_dev,
                  double source_depth, double source_range, double receiver_range, double rtol, double atol,
                  int32_t terminate_backwards, int32_t n_save, const double* range_save_dev, double* out_traj_dev,
                  double* out_final_dev, int32_t* out_n_bottom_dev, int32_t* out_n_surface_dev,
                  int32_t* out_status_dev, uint32_t* out_counters_dev, int32_t stored_convention,
                  int32_t* out_p_sign_changes_dev, int32_t max_bounces, void* stream)
{
    return shoot_fan_impl(env, n_rays, 0, angles_dev, y0_dev, source_depth, source_range, receiver_range, rtol, atol,
                          terminate_backwards, n_save, range_save_dev, out_traj_dev, out_final_dev, out_n_bottom_dev,
                          out_n_surface_dev, out_status_dev, out_counters_dev, stored_convention,
                          out_p_sign_changes_dev, max_bounces, stream);
}

int prt_shoot_ensemble(const prt_env_t* env, int32_t n_env, int64_t rays_per_env, const double* angles_dev,
                       double source_depth, double source_range, double receiver_range, double rtol, double atol,
                       int32_t terminate_backwards, int32_t n_save, const double* range_save_dev,
                       double* out_traj_dev, double* out_final_dev, int32_t* out_n_bottom_dev,
                       int32_t* out_n_surface_dev, int32_t* out_status_dev, uint32_t* out_counters_dev,
                       int32_t stored_convention, int32_t* out_p_sign_changes_dev, int32_t max_bounces, void* stream)
{
    if (!env) return fail(PRT_ERR_ARG, "prt_shoot_ensemble: null environment");
    if (n_env < 1 || n_env > env->view.n_env) return fail(PRT_ERR_ARG, "prt_shoot_ensemble: n_env out of range");
    if (rays_per_env < 1 || !angles_dev) return fail(PRT_ERR_ARG, "prt_shoot_ensemble: need rays_per_env >= 1 angles");
    return shoot_fan_impl(env, (int64_t)n_env * rays_per_env, rays_per_env, angles_dev, nullptr, source_depth,
                          source_range, receiver_range, rtol, atol, terminate_backwards, n_save, range_save_dev,
                          out_traj_dev, out_final_dev, out_n_bottom_dev, out_n_surface_dev, out_status_dev,
                          out_counters_dev, stored_convention, out_p_sign_changes_dev, max_bounces, stream);
}

// Scratch of prt_shoot_fan_host, kept between calls (grow-only, one per device): device memory for inputs and
// outputs plus a page-locked staging block for everything but the trajectories, so that a call costs no
// cudaMalloc / cudaFree, ONE staged host-to-device copy and TWO device-to-host copies (small outputs; trajectories).
struct prt_host_scratch {
    std::mutex lock;
    unsigned char* dev = nullptr;
    size_t dev_bytes = 0;
    unsigned char* pin = nullptr;
    size_t pin_bytes = 0;
};
static prt_host_scratch g_host_scratch[64];

static cudaError_t host_scratch_reserve(prt_host_scratch& S, size_t dev_bytes, size_t pin_bytes)
{
    if (dev_bytes > S.dev_bytes) {
        if (S.dev) cudaFree(S.dev);
        S.dev = nullptr;
        S.dev_bytes = 0;
        const size_t want = dev_bytes + dev_bytes / 4;
        cudaError_t e = cudaMalloc(&S.dev, want);
        if (e != cudaSuccess) return e;
        S.dev_bytes = want;
    }
    if (pin_bytes > S.pin_bytes) {
        if (S.pin) cudaFreeHost(S.pin);
        S.pin = nullptr;
        S.pin_bytes = 0;
        const size_t want = pin_bytes + pin_bytes / 4;
        cudaError_t e = cudaMallocHost(&S.pin, want);
        if (e != cudaSuccess) return e;
        S.pin_bytes = want;
    }
    return cudaSuccess;
}

int prt_shoot_fan_host(const prt_env_t* env, int64_t n_rays, const double* angles, const double* y0,
                       double source_depth, double source_range, double receiver_range, double rtol, double atol,
                       int32_t terminate_backwards, int32_t n_save, const double* range_save, double* out_traj,
                       double* out_final, int32_t* out_n_bottom, int32_t* out_n_surface, int32_t* out_status,
                       uint32_t* out_counters, int32_t stored_convention, int32_t* out_p_sign_changes,
                       int32_t max_bounces)
{
    if (!env) return fail(PRT_ERR_ARG, "prt_shoot_fan_host: null environment");
    if (n_rays <= 0) return n_rays == 0 ? PRT_SUCCESS : fail(PRT_ERR_ARG, "prt_shoot_fan_host: n_rays < 0");
    if (!angles && !y0) return fail(PRT_ERR_ARG, "prt_shoot_fan_host: need angles or y0");
    if (!out_final || !out_n_bottom || !out_n_surface || !out_status)
        return fail(PRT_ERR_ARG, "prt_shoot_fan_host: null output pointer");
    if (env->device < 0 || env->device >= 64) return fail(PRT_ERR_ARG, "prt_shoot_fan_host: device index out of range");
    PRT_ON_DEVICE(env->device);
    const bool save = n_save > 0 && out_traj && range_save;
    const bool flips = save && out_p_sign_changes;
    const size_t n = (size_t)n_rays;
    // staged block (identical layout on the device and in page-locked host memory): inputs | small outputs
    const size_t in_bytes = (y0 ? 3 * n : n) * sizeof(double);
    const size_t o_rs = prt_align256(in_bytes);
    const size_t o_final = prt_align256(o_rs + (save ? (size_t)n_save * sizeof(double) : 0));
    const size_t o_nb = prt_align256(o_final + 3 * n * sizeof(double));
    const size_t o_ns = prt_align256(o_nb + n * sizeof(int32_t));
    const size_t o_st = prt_align256(o_ns + n * sizeof(int32_t));
    const size_t o_cnt = prt_align256(o_st + n * sizeof(int32_t));
    const size_t o_flip = prt_align256(o_cnt + (out_counters ? PRT_CNT_N * n * sizeof(uint32_t) : 0));
    const size_t staged = prt_align256(o_flip + (flips ? n * sizeof(int32_t) : 0));
    const size_t total = staged + (save ? 3 * n * (size_t)n_save * sizeof(double) : 0);
    prt_host_scratch& S = g_host_scratch[env->device];
    std::lock_guard<std::mutex> hold(S.lock);
    cudaError_t e = host_scratch_reserve(S, total, staged);
    if (e != cudaSuccess) return fail(PRT_ERR_CUDA, std::string("prt_shoot_fan_host: scratch: ") + cudaGetErrorString(e));
    unsigned char *d = S.dev, *h = S.pin;
    memcpy(h, y0 ? y0 : angles, in_bytes);
    if (save) memcpy(h + o_rs, range_save, (size_t)n_save * sizeof(double));
    cudaStream_t s = nullptr;  // legacy default stream
    e = cudaMemcpyAsync(d, h, o_final, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return fail(PRT_ERR_CUDA, std::string("prt_shoot_fan_host: H2D: ") + cudaGetErrorString(e));
    int rc = prt_shoot_fan(env, n_rays, y0 ? nullptr : (const double*)d, y0 ? (const double*)d : nullptr, source_depth,
                           source_range, receiver_range, rtol, atol, terminate_backwards, save ? n_save : 0,
                           save ? (const double*)(d + o_rs) : nullptr, save ? (double*)(d + staged) : nullptr,
                           (double*)(d + o_final), (int32_t*)(d + o_nb), (int32_t*)(d + o_ns), (int32_t*)(d + o_st),
                           out_counters ? (uint32_t*)(d + o_cnt) : nullptr, stored_convention,
                           flips ? (int32_t*)(d + o_flip) : nullptr, max_bounces, s);
    if (rc != PRT_SUCCESS) return rc;
    e = cudaMemcpyAsync(h + o_final, d + o_final, staged - o_final, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && save)  // the big one goes straight to the caller's (pageable) array
        e = cudaMemcpy(out_traj, d + staged, 3 * n * (size_t)n_save * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(PRT_ERR_CUDA, std::string("prt_shoot_fan_host: ") + cudaGetErrorString(e));
    memcpy(out_final, h + o_final, 3 * n * sizeof(double));
    memcpy(out_n_bottom, h + o_nb, n * sizeof(int32_t));
    memcpy(out_n_surface, h + o_ns, n * sizeof(int32_t));
    memcpy(out_status, h + o_st, n * sizeof(int32_t));
    if (out_counters) memcpy(out_counters, h + o_cnt, PRT_CNT_N * n * sizeof(uint32_t));
    if (flips) memcpy(out_p_sign_changes, h + o_flip, n * sizeof(int32_t));
    return PRT_SUCCESS;
}

static int32_t bracket_tiles(int64_t n_rays)
{
    return n_rays < 2 ? 0 : (int32_t)((n_rays - 1 + PRT_BR_TILE - 1) / PRT_BR_TILE);
}

int64_t prt_bracket_workspace_bytes(int64_t n_rays, int32_t n_rcv)
{
    if (n_rcv <= 0) return 256;
    return (int64_t)prt_align256(sizeof(int64_t) * (size_t)n_rcv * (size_t)(bracket_tiles(n_rays) + 1));
}

int prt_count_brackets(int64_t n_rays, const double* z_end_dev, int32_t n_rcv, const double* receiver_depths_dev,
                       int32_t* counts_dev, void* work_dev, void* stream)
{
    if (n_rcv <= 0) return PRT_SUCCESS;
    if (!counts_dev) return fail(PRT_ERR_ARG, "prt_count_brackets: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int32_t n_tiles = bracket_tiles(n_rays);
    if (n_tiles == 0) {  // fewer than two rays: no pair, no bracket (the table pointers may be null)
        PRT_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(int32_t) * (size_t)n_rcv, s));
        return PRT_SUCCESS;
    }
    if (!z_end_dev || !receiver_depths_dev || !work_dev) return fail(PRT_ERR_ARG, "prt_count_brackets: null pointer");
    if (n_tiles > 65535) return fail(PRT_ERR_ARG, "prt_count_brackets: more than 134 million rays in one table");
    int64_t* tile_cnt = static_cast<int64_t*>(work_dev);
    prt_count_brackets_kernel<<<dim3((unsigned)n_rcv, (unsigned)n_tiles), 256, 0, s>>>(n_rays, z_end_dev,
                                                                                     receiver_depths_dev, n_tiles, tile_cnt);
    prt_scan_brackets_kernel<<<1, 1024, 0, s>>>((int64_t)n_rcv * n_tiles, n_tiles, n_rcv, tile_cnt, counts_dev);
    g_launches.fetch_add(2);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int prt_fill_brackets(int64_t n_rays, const double* z_end_dev, const double* thetas_dev, int32_t n_rcv,
                      const double* receiver_depths_dev, const void* work_dev, double* br_z1_dev,
                      double* br_z2_dev, double* br_th1_dev, double* br_th2_dev, double* br_rd_dev,
                      int32_t* br_rcv_dev, int64_t* br_k_dev, void* stream)
{
    const int32_t n_tiles = bracket_tiles(n_rays);
    if (n_rcv <= 0 || n_tiles == 0) return PRT_SUCCESS;
    if (!z_end_dev || !thetas_dev || !receiver_depths_dev || !work_dev || !br_z1_dev || !br_z2_dev || !br_th1_dev ||
        !br_th2_dev || !br_rd_dev || !br_rcv_dev || !br_k_dev)
        return fail(PRT_ERR_ARG, "prt_fill_brackets: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    prt_fill_brackets_kernel<<<dim3((unsigned)n_rcv, (unsigned)n_tiles), 256, 0, s>>>(
        n_rays, z_end_dev, thetas_dev, receiver_depths_dev, n_tiles, static_cast<const int64_t*>(work_dev), br_z1_dev,
        br_z2_dev, br_th1_dev, br_th2_dev, br_rd_dev, br_rcv_dev, br_k_dev);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int prt_pack_arrivals(int64_t n_rays, const double* final_dev, const int32_t* n_bottom_dev,
                      const int32_t* n_surface_dev, const int32_t* status_dev, double* records_dev, void* stream)
{
    if (n_rays <= 0) return PRT_SUCCESS;
    if (!final_dev || !n_bottom_dev || !n_surface_dev || !status_dev || !records_dev)
        return fail(PRT_ERR_ARG, "prt_pack_arrivals: null pointer");
    prt_pack_arrivals_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_rays, final_dev, n_bottom_dev, n_surface_dev, status_dev, records_dev);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int prt_received_angles(int64_t n, double x, const double* z_dev, const double* p_dev, int32_t n_r, int32_t n_z,
                        const double* rin_dev, const double* zin_dev, const double* cin_dev, double* out_theta_dev,
                        void* stream)
{
    if (n <= 0) return n == 0 ? PRT_SUCCESS : fail(PRT_ERR_ARG, "prt_received_angles: n < 0");
    if (!z_dev || !p_dev || !rin_dev || !zin_dev || !cin_dev || !out_theta_dev)
        return fail(PRT_ERR_ARG, "prt_received_angles: null pointer");
    if (n_r < 2 || n_z < 2) return fail(PRT_ERR_ARG, "prt_received_angles: the sound-speed grid needs 2 x 2 nodes");
    prt_received_angles_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        n, x, z_dev, p_dev, n_r, n_z, rin_dev, zin_dev, cin_dev, out_theta_dev);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int prt_eigen_table(int64_t n, int64_t n_br, const int64_t* idx_dev, const int32_t* rcv_dev, const double* theta_dev,
                    const double* final_dev, const int32_t* n_bottom_dev, const int32_t* n_surface_dev,
                    const int32_t* p_sign_changes_dev, double receiver_range, int32_t n_r, int32_t n_z,
                    const double* rin_dev, const double* zin_dev, const double* cin_dev, double* out_table_dev,
                    void* stream)
{
    if (n <= 0) return n == 0 ? PRT_SUCCESS : fail(PRT_ERR_ARG, "prt_eigen_table: n < 0");
    if (!idx_dev || !rcv_dev || !theta_dev || !final_dev || !n_bottom_dev || !n_surface_dev || !rin_dev || !zin_dev ||
        !cin_dev || !out_table_dev)
        return fail(PRT_ERR_ARG, "prt_eigen_table: null pointer");
    if (n_r < 2 || n_z < 2) return fail(PRT_ERR_ARG, "prt_eigen_table: the sound-speed grid needs 2 x 2 nodes");
    prt_eigen_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        n, n_br, idx_dev, rcv_dev, theta_dev, final_dev, n_bottom_dev, n_surface_dev, p_sign_changes_dev, receiver_range,
        n_r, n_z, rin_dev, zin_dev, cin_dev, out_table_dev);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int prt_flat_earth(int32_t n_env, int32_t n_r, int32_t n_z, const double* depth_dev, const double* re_dev,
                   const double* c_dev, int32_t resample, double* out_dev, void* stream)
{
    if (n_env < 1 || n_r < 1 || n_z < 2) return fail(PRT_ERR_ARG, "prt_flat_earth: need n_env, n_r >= 1 and n_z >= 2");
    if (!depth_dev || !re_dev || !c_dev || !out_dev) return fail(PRT_ERR_ARG, "prt_flat_earth: null pointer");
    if (c_dev == out_dev && resample) return fail(PRT_ERR_ARG, "prt_flat_earth: resampling cannot run in place");
    const int64_t n_nodes = (int64_t)n_env * n_r * n_z;
    prt_flat_earth_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_nodes, n_r, n_z, depth_dev, re_dev, c_dev, resample, out_dev);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int64_t prt_iw_workspace_bytes(int32_t n_env, int32_t n_r, int32_t J)
{
    if (n_env < 0 || n_r < 0 || J < 0) return 0;
    return (int64_t)(prt_align256(sizeof(double) * (size_t)n_env * (size_t)J * (size_t)n_r) + prt_align256(sizeof(double) * (size_t)n_env));
}

int prt_iw_perturb(int32_t n_env, int32_t n_r, int32_t n_z, int32_t J, int32_t K, const double* r_dev,
                   const double* vert_dev, const double* kappa_dev, const double* phases_dev, const double* c_base_dev,
                   int32_t iax, double rms_axis, void* work_dev, double* out_dev, void* stream)
{
    if (n_env < 1 || n_r < 1 || n_z < 1 || J < 1 || K < 1 || iax < 0 || iax >= n_z)
        return fail(PRT_ERR_ARG, "prt_iw_perturb: bad sizes");
    if (!r_dev || !vert_dev || !kappa_dev || !phases_dev || !c_base_dev || !work_dev || !out_dev)
        return fail(PRT_ERR_ARG, "prt_iw_perturb: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* H = static_cast<double*>(work_dev);
    double* scale = reinterpret_cast<double*>(static_cast<unsigned char*>(work_dev) +
                                              prt_align256(sizeof(double) * (size_t)n_env * (size_t)J * (size_t)n_r));
    const int64_t nH = (int64_t)n_env * J * n_r, n_nodes = (int64_t)n_env * n_r * n_z;
    prt_iw_horizontal_kernel<<<(unsigned)((nH + 255) / 256), 256, 0, s>>>(nH, n_r, K, r_dev, kappa_dev, phases_dev, H);
    prt_iw_scale_kernel<<<(unsigned)n_env, 256, 0, s>>>(n_r, n_z, J, iax, rms_axis, vert_dev, H, scale);
    prt_iw_compose_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(n_nodes, n_r, n_z, J, vert_dev, H, scale,
                                                                            c_base_dev, out_dev);
    g_launches.fetch_add(3);
    PRT_CUDA(cudaGetLastError());
    return PRT_SUCCESS;
}

int64_t prt_refine_workspace_bytes(int64_t n_br)
{
    if (n_br < 0) return 0;
    size_t n = (size_t)n_br;
    return (int64_t)(5 * prt_align256(n * sizeof(double)) + 6 * prt_align256(n * sizeof(int32_t)) + 256);
}

int prt_refine_brackets(const prt_env_t* env, int64_t n_br, const double* z1_dev, const double* z2_dev,
                        const double* th1_dev, const double* th2_dev, const double* rd_dev, double source_depth,
                        double source_range, double receiver_range, double rtol, double atol,
                        int32_t terminate_backwards, double ztol, int32_t max_iter, double* out_theta_dev,
                        double* out_final_dev, int32_t* out_n_bottom_dev, int32_t* out_n_surface_dev,
                        int32_t* out_status_dev, int32_t* out_shots_dev, int32_t* out_ray_status_dev, void* work_dev,
                        const int32_t* order_dev, int32_t max_bounces, void* stream)
{
    if (!env) return fail(PRT_ERR_ARG, "prt_refine_brackets: null environment");
    if (n_br < 0 || n_br > 0x7fffffff) return fail(PRT_ERR_ARG, "prt_refine_brackets: bad n_br");
    if (n_br == 0) return PRT_SUCCESS;
    if (!z1_dev || !z2_dev || !th1_dev || !th2_dev || !rd_dev || !out_theta_dev || !out_final_dev ||
        !out_n_bottom_dev || !out_n_surface_dev || !out_status_dev || !out_shots_dev || !out_ray_status_dev || !work_dev)
        return fail(PRT_ERR_ARG, "prt_refine_brackets: null pointer");
    if (!(source_range < receiver_range))
        return fail(PRT_ERR_ARG, "prt_refine_brackets: source_range must be < receiver_range");
    if (max_iter < 0) return fail(PRT_ERR_ARG, "prt_refine_brackets: max_iter < 0");
    PRT_ON_DEVICE(env->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)n_br;
    unsigned char* w = static_cast<unsigned char*>(work_dev);
    const size_t dsz = prt_align256(n * sizeof(double)), isz = prt_align256(n * sizeof(int32_t));
    prt_refine_state S;
    S.z1 = (double*)w; S.z2 = (double*)(w + dsz); S.th1 = (double*)(w + 2 * dsz); S.th2 = (double*)(w + 3 * dsz);
    S.th = (double*)(w + 4 * dsz);
    unsigned char* wi = w + 5 * dsz;
    S.iter = (int32_t*)wi; S.shots = (int32_t*)(wi + isz); S.list_a = (int32_t*)(wi + 2 * isz);
    S.list_b = (int32_t*)(wi + 3 * isz); S.keep = (int32_t*)(wi + 4 * isz); S.blk = (int32_t*)(wi + 5 * isz);
    S.counters = (int32_t*)(wi + 6 * isz);
    prt_refine_params P;
    P.n_br = n_br; P.rd = rd_dev; P.source_depth = source_depth; P.source_range = source_range;
    P.receiver_range = receiver_range; P.rtol = rtol; P.atol = atol; P.ztol = ztol;
    P.terminate_backwards = terminate_backwards; P.max_iter = max_iter;
    P.max_bounces = max_bounces > 0 ? max_bounces : (1 << 20);
    P.park_wait = park_wait_refine_setting();
    if (P.rtol < 100.0 * PRT_EPS) P.rtol = 100.0 * PRT_EPS;
    P.out_theta = out_theta_dev; P.out_final = out_final_dev; P.out_nb = out_n_bottom_dev; P.out_ns = out_n_surface_dev;
    P.out_status = out_status_dev; P.out_shots = out_shots_dev; P.out_ray_status = out_ray_status_dev;
    prt_refine_init_kernel<<<(unsigned)((n_br + 255) / 256), 256, 0, s>>>(S, n_br, z1_dev, z2_dev, th1_dev, th2_dev,
                                                                          rd_dev, order_dev, P);
    g_launches.fetch_add(1);
    PRT_CUDA(cudaGetLastError());
    // Every round is launched for the first round's size: the number of brackets still active lives on the device
    // only (no read-back, no host synchronisation between rounds), surplus warps find nothing and exit, and the
    // rays-per-warp rule is applied by the kernel itself.  At most max_iter + 2 trial shots per bracket
    // (eigenrays.py:265-268).
    int64_t capacity = 0;
    const int lanes0 = pick_lanes(n_br, env->device, PRT_MIN_BLOCKS_ARR * (PRT_BLOCK_ARR / 32), &capacity);
    const int block = PRT_BLOCK_ARR;
    const int64_t n_warps = (n_br + lanes0 - 1) / lanes0;
    const unsigned grid = (unsigned)((n_warps * 32 + block - 1) / block);
    const bool fixed = (lanes_forced() != 0);
    P.lanes = fixed ? lanes0 : 0;
    P.warp_budget = (int32_t)(n_warps < capacity ? n_warps : capacity);
    const int n_blk = (int)((n_br + PRT_CP_TILE - 1) / PRT_CP_TILE);
    int flip = 0;
    for (int round = 0; round < max_iter + 2; ++round) {
        int32_t* list_in = flip ? S.list_b : S.list_a;
        int32_t* list_out = flip ? S.list_a : S.list_b;
        int32_t* n_in = S.counters + flip;
        int32_t* n_out = S.counters + (1 - flip);
        const int zm = prt_z_mode(env->view);
        if (zm == 3) prt_refine_round_kernel<3><<<grid, block, 0, s>>>(env->view, P, S, list_in, n_in);
        else if (zm == 2) prt_refine_round_kernel<2><<<grid, block, 0, s>>>(env->view, P, S, list_in, n_in);
        else if (zm == 1) prt_refine_round_kernel<1><<<grid, block, 0, s>>>(env->view, P, S, list_in, n_in);
        else prt_refine_round_kernel<0><<<grid, block, 0, s>>>(env->view, P, S, list_in, n_in);
        prt_compact_count_kernel<<<n_blk, 256, 0, s>>>(S.keep, n_in, S.blk);
        prt_compact_scan_kernel<<<1, 1024, 0, s>>>(n_blk, S.blk, n_out);
        prt_compact_scatter_kernel<<<n_blk, 256, 0, s>>>(S.keep, n_in, S.blk, list_in, list_out);
        g_launches.fetch_add(4);
        PRT_CUDA(cudaGetLastError());
        flip = 1 - flip;
    }
    return PRT_SUCCESS;
}

double prt_measure_fp64_peak(int device, int32_t iters)
{
    if (prt_device_count() <= device || device < 0) return -1.0;
    prt_device_guard guard(device);
    if (guard.err != cudaSuccess) return -1.0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
    const int threads = 512, blocks = sms * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    prt_fp64_peak_kernel<<<blocks, threads>>>(out, 16, 1.0000001, 1e-9);  // warm-up
    cudaEventRecord(a);
    prt_fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    g_launches.fetch_add(2);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    if (e != cudaSuccess || ms <= 0.f) return -1.0;
    double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
    return flops / (ms * 1e-3);
}

}  // extern "C"
