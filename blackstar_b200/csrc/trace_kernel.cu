// trace_kernel.cu -- K1 (geodesic trace) + K2 (sky lookup) + K3 (2x2 supersample), fused.
//
// One thread per ray (src/Raytracer.hs:66 evaluates traceRay at every index of the
// (h' x w') array; massiv's Par splits the index range over CPU threads -- here the index
// range is split over 148 SMs x resident warps).  FP64 throughout.
//
// Two schedules, same arithmetic (trace_core.cuh):
//   trace_tiles_kernel   a warp owns a tile of 32 rays, runs them to completion, looks the
//                        escaped rays up in the star tree together, reduces 2x2 quads with
//                        shuffles and stores one float4 per output pixel.  Tiles are handed
//                        out by a global atomic counter to persistent warps.
//   trace_refill_kernel  persistent warps advance their 32 ray slots in blocks of
//                        kBlockSteps RK4 steps; between blocks a warp ballot finds the slots
//                        whose ray has terminated, finishes them and compacts new rays from
//                        the global queue into exactly those lanes, so no lane idles while
//                        a neighbour is still integrating (north_star: "warp-ballot
//                        compaction of live rays between step blocks").
//
// HBM traffic: 16 B per OUTPUT pixel (one float4 store) + the star tree (L2 resident).
#include "bsb_common.cuh"
#include "trace_core.cuh"

#include <cuda_runtime.h>

namespace bsb {

#ifndef BSB_TRACE_THREADS
#define BSB_TRACE_THREADS 256
#endif
constexpr int kTraceThreads = BSB_TRACE_THREADS;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void stage_tree_top(const FrameParams &P, float *s_top)
{
    if (P.tree.n_stars > 0) {
        const int n_top = (1 << P.tree.top_levels) - 1;
        for (int i = threadIdx.x; i < n_top; i += blockDim.x) s_top[i] = P.tree.top[i];
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// The ray's orbital-plane frame is needed again when the ray is finished (3-D exit velocity
// for the sky lookup).  It is parked in shared memory meanwhile instead of being recomputed
// (two divisions, three square roots and a fifth root) or held in registers.
struct FrameStore { double v[8][kTraceThreads]; };   // f1[3], f2[3], L, (unused)

__device__ __forceinline__ void park_frame(FrameStore &fs, const RayFrame &F)
{
    const int t = threadIdx.x;
    fs.v[0][t] = F.f1[0]; fs.v[1][t] = F.f1[1]; fs.v[2][t] = F.f1[2];
    fs.v[3][t] = F.f2[0]; fs.v[4][t] = F.f2[1]; fs.v[5][t] = F.f2[2];
    fs.v[6][t] = F.L;
}
__device__ __forceinline__ void fetch_frame(const FrameStore &fs, RayFrame &F)
{
    const int t = threadIdx.x;
    F.f1[0] = fs.v[0][t]; F.f1[1] = fs.v[1][t]; F.f1[2] = fs.v[2][t];
    F.f2[0] = fs.v[3][t]; F.f2[1] = fs.v[4][t]; F.f2[2] = fs.v[5][t];
    F.L = fs.v[6][t];
}

// lane -> ray coordinates inside tile `tile`
template <bool SS>
__device__ __forceinline__ void tile_coords(const FrameParams &P, unsigned tile, int lane, int &ox, int &oy,
                                            int &gx, int &gy)
{
    const int ty = (int)(tile / (unsigned)P.tiles_x), tx = (int)(tile - (unsigned)ty * (unsigned)P.tiles_x);
    if (SS) {
        const int qd = lane >> 2, sub = lane & 3;
        ox = tx * 4 + (qd & 3);
        oy = P.row0 + ty * 2 + (qd >> 2);
        gx = 2 * ox + (sub >> 1);  // sub: 0=(2y,2x) 1=(2y+1,2x) 2=(2y,2x+1) 3=(2y+1,2x+1)
        gy = 2 * oy + (sub & 1);
    } else {
        ox = tx * 8 + (lane & 7);
        oy = P.row0 + ty * 4 + (lane >> 3);
        gx = ox;
        gy = oy;
    }
}

// supersample (src/ImageFilters.hs:88-97): 0.25 * (((a + b) + c) + d) over the quad held by
// lanes 4k..4k+3; result valid in lane 4k.
__device__ __forceinline__ double quad_mean(double v)
{
    const double b = __shfl_down_sync(kFull, v, 1);
    const double c = __shfl_down_sync(kFull, v, 2);
    const double d = __shfl_down_sync(kFull, v, 3);
    return __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(v, b), c), d));
}

template <bool SS, int MINB>
__global__ void __launch_bounds__(kTraceThreads, MINB)
trace_tiles_kernel(const __grid_constant__ FrameParams P, float4 *__restrict__ out, TraceCounters *ctr)
{
    __shared__ float s_top[kSmemTreeNodes + 1];
    __shared__ FrameStore s_frames;
    stage_tree_top(P, s_top);

    const int lane = threadIdx.x & 31;
    unsigned long long my_steps = 0, my_capped = 0, my_hits = 0;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(&ctr->next_tile, 1u);
        tile = __shfl_sync(kFull, tile, 0);
        if (tile >= (unsigned)P.n_tiles) break;
        int ox, oy, gx, gy;
        tile_coords<SS>(P, tile, lane, ox, oy, gx, gy);
        const bool valid = ox < P.W && oy < P.row1;
        double rgb[3] = { 0.0, 0.0, 0.0 };
        if (valid) {
            RayState s;
            RayFrame F;
            ray_init(P, gx, gy, s, F);
            park_frame(s_frames, F);
            ray_advance(P, s, 0xffffffffu);
            fetch_frame(s_frames, F);
            my_hits += ray_finish(P, s_top, F, s, rgb);
            my_steps += s.steps;
            my_capped += (s.status == kCapped);
        }
        if (SS) {
            rgb[0] = quad_mean(rgb[0]);
            rgb[1] = quad_mean(rgb[1]);
            rgb[2] = quad_mean(rgb[2]);
            if (valid && (lane & 3) == 0)
                out[(size_t)(oy - P.row0) * P.W + ox] = make_float4((float)rgb[0], (float)rgb[1], (float)rgb[2], 1.0f);
        } else if (valid) {
            out[(size_t)(oy - P.row0) * P.W + ox] = make_float4((float)rgb[0], (float)rgb[1], (float)rgb[2], 1.0f);
        }
    }
    my_steps = warp_sum(my_steps);
    my_capped = warp_sum(my_capped);
    my_hits = warp_sum(my_hits);
    if (lane == 0) {
        atomicAdd(&ctr->steps, my_steps);
        if (my_capped) atomicAdd(&ctr->capped, my_capped);
        if (my_hits) atomicAdd(&ctr->star_hits, my_hits);
    }
}

// ---------------------------------------------------------------------------------------
// Persistent warps with ballot compaction.  The unit of work ("job") is one ray (SS off)
// or one 2x2 quad on four adjacent lanes (SS on); job j lives in tile j / U at slot j % U
// (U = 32 or 8 units per tile), so consecutive jobs are neighbours on the image and in
// the star tree.
template <bool SS, int kBlockSteps, int kRefillMin>
__global__ void __launch_bounds__(kTraceThreads, 2)
trace_refill_kernel(const __grid_constant__ FrameParams P, float4 *__restrict__ out, TraceCounters *ctr)
{
    __shared__ float s_top[kSmemTreeNodes + 1];
    __shared__ FrameStore s_frames;
    stage_tree_top(P, s_top);

    constexpr int U = SS ? 8 : 32;           // units per warp
    constexpr int LPU = 32 / U;              // lanes per unit
    const int lane = threadIdx.x & 31;
    const int unit = lane / LPU;
    const unsigned n_jobs = (unsigned)P.n_tiles * (unsigned)U;

    RayState s;
    s.status = kIdle;
    s.steps = 0;
    int ox = 0, oy = 0, gx = 0, gy = 0;
    bool occupied = false;   // this lane's unit holds a job (valid or padding)
    bool valid = false;      // ... and the job is a real pixel
    bool drained = false;    // the global queue is empty
    unsigned long long my_steps = 0, my_capped = 0, my_hits = 0;

    for (;;) {
        // ---- which units are finished (every lane of the unit has terminated)?
        const bool lane_done = !(occupied && valid && s.status == kAlive);
        const unsigned done_mask = __ballot_sync(kFull, lane_done);
        // unit_done: all LPU lanes of my unit are done
        const unsigned umask = ((1u << LPU) - 1u) << (unit * LPU);
        const bool unit_done = (done_mask & umask) == umask;
        const unsigned unit_done_mask = __ballot_sync(kFull, unit_done);
        const int n_done_lanes = __popc(unit_done_mask);
        // lanes of finished units that actually hold a job (idle slots of a drained queue do not count)
        const int n_fin_lanes = __popc(__ballot_sync(kFull, unit_done && occupied));
        const bool any_alive = done_mask != kFull;
        const bool refill = n_fin_lanes >= kRefillMin * LPU || !any_alive;

        if (refill && n_done_lanes > 0) {
            // ---- finish the terminated rays of finished units and store their pixels
            double rgb[3] = { 0.0, 0.0, 0.0 };
            const bool fin = unit_done && occupied && valid;
            if (fin) {
                RayFrame F;
                fetch_frame(s_frames, F);
                my_hits += ray_finish(P, s_top, F, s, rgb);
                my_steps += s.steps;
                my_capped += (s.status == kCapped);
            }
            if (SS) {
                rgb[0] = quad_mean(rgb[0]);
                rgb[1] = quad_mean(rgb[1]);
                rgb[2] = quad_mean(rgb[2]);
                if (fin && (lane & 3) == 0)
                    out[(size_t)(oy - P.row0) * P.W + ox] = make_float4((float)rgb[0], (float)rgb[1], (float)rgb[2], 1.0f);
            } else if (fin) {
                out[(size_t)(oy - P.row0) * P.W + ox] = make_float4((float)rgb[0], (float)rgb[1], (float)rgb[2], 1.0f);
            }
            if (unit_done) { occupied = false; valid = false; s.status = kIdle; }
            // ---- compact new jobs into exactly the freed units
            if (!drained) {
                const int n_units = n_done_lanes / LPU;
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&ctr->next_tile, (unsigned)n_units);
                base = __shfl_sync(kFull, base, 0);
                if (base + (unsigned)n_units >= n_jobs) drained = true;
                if (unit_done) {
                    // rank of my unit among the freed units (leader-lane bits below mine)
                    const unsigned leaders = (LPU == 1) ? unit_done_mask : (unit_done_mask & 0x11111111u);
                    const unsigned below = leaders & ((1u << (unit * LPU)) - 1u);
                    const unsigned job = base + (unsigned)__popc(below);
                    if (job < n_jobs) {
                        const unsigned tile = job / (unsigned)U;
                        const int slot = (int)(job - tile * (unsigned)U);
                        tile_coords<SS>(P, tile, slot * LPU + (lane & (LPU - 1)), ox, oy, gx, gy);
                        occupied = true;
                        valid = ox < P.W && oy < P.row1;
                        if (valid) {
                            RayFrame F;
                            ray_init(P, gx, gy, s, F);
                            park_frame(s_frames, F);
                        }
                    }
                }
            }
        }
        if (__ballot_sync(kFull, occupied && valid && s.status == kAlive) == 0u) {
            // nothing left to integrate in this warp; if the queue is drained too we are done,
            // otherwise loop once more to fetch work (padding-only units count as done).
            if (drained && __ballot_sync(kFull, occupied) == 0u) break;
            continue;  // finish the remaining (terminated or padding) units / fetch more
        }
        // ---- one block of RK4 steps for every live lane
        if (occupied && valid && s.status == kAlive) ray_advance(P, s, kBlockSteps);
    }
    my_steps = warp_sum(my_steps);
    my_capped = warp_sum(my_capped);
    my_hits = warp_sum(my_hits);
    if (lane == 0) {
        atomicAdd(&ctr->steps, my_steps);
        if (my_capped) atomicAdd(&ctr->capped, my_capped);
        if (my_hits) atomicAdd(&ctr->star_hits, my_hits);
    }
}

// ---- per-frame ray tables: vx[x] (x < W2) and vy[y] (y < H2), exact reference op order ----
__global__ void ray_tables_kernel(const __grid_constant__ FrameParams P, double *vx, double *vy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.W2) vx[i] = pixel_vx(P, i);
    if (i < P.H2) vy[i] = pixel_vy(P, i);
}

cudaError_t launch_ray_tables(const FrameParams &P_in, double *vx, double *vy, cudaStream_t stream)
{
    FrameParams P = P_in;
    P.vx_tab = nullptr;
    P.vy_tab = nullptr;
    const int n = P.W2 > P.H2 ? P.W2 : P.H2;
    ray_tables_kernel<<<(n + 255) / 256, 256, 0, stream>>>(P, vx, vy);
    return cudaGetLastError();
}

// ---- numerics self-test of the |pos|^-5 kernel primitive: max relative error of
// rinv5(q) against 0.4 pow(q, -2.5) over n log-spaced q in [q_lo, q_hi]
__global__ void rinv5_selftest_kernel(double q_lo, double q_hi, int n, double *max_rel, double *max_seed_err)
{
    double worst = 0.0, worst_seed = 0.0;
    const double lr = log(q_hi / q_lo);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double q = q_lo * exp(lr * ((double)i + 0.5) / (double)n);
        const double ref = 0.4 * pow(q, -2.5);
        const double y0 = rsqrt_seed(q);
        const double got = rinv5_seeded(q, y0, 1.4);
        worst = fmax(worst, fabs(got - ref) / ref);
        worst_seed = fmax(worst_seed, fabs(fma(-q * y0, y0, 1.0)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        worst = fmax(worst, __shfl_xor_sync(kFull, worst, o));
        worst_seed = fmax(worst_seed, __shfl_xor_sync(kFull, worst_seed, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // non-negative doubles order like their bit patterns
        atomicMax((unsigned long long *)max_rel, (unsigned long long)__double_as_longlong(worst));
        atomicMax((unsigned long long *)max_seed_err, (unsigned long long)__double_as_longlong(worst_seed));
    }
}

// ------------------------------------------------------------------------------ launch
template <typename K>
static int persistent_grid(K kernel, int n_sms)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceThreads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return n_sms * per_sm;
}

// variant: 0 = tiles, 1 = refill(block 16, min 8 lanes), 2 = refill(block 8, min 4), 3 = refill(block 32, min 8),
//          4 = tiles compiled for 4 CTAs/SM (64 registers),
//          6 = tiles compiled for 2 CTAs/SM (<= 128 registers: every loop constant stays in a register; default)
cudaError_t launch_trace(const FrameParams &P_in, float4 *out, TraceCounters *ctr, int n_sms, int variant,
                         cudaStream_t stream)
{
    if (P_in.n_tiles <= 0) return cudaSuccess;
    const FrameParams &P = P_in;
#define BSB_LAUNCH(KERNEL)                                                                        \
    do {                                                                                          \
        int grid = persistent_grid(KERNEL, n_sms);                                                \
        const int max_useful = (P.n_tiles + (kTraceThreads / 32) - 1) / (kTraceThreads / 32);     \
        if (grid > max_useful) grid = max_useful;                                                 \
        KERNEL<<<grid, kTraceThreads, 0, stream>>>(P, out, ctr);                                  \
    } while (0)
    if (P.ss) {
        switch (variant) {
        case 1: BSB_LAUNCH((trace_refill_kernel<true, 16, 2>)); break;
        case 2: BSB_LAUNCH((trace_refill_kernel<true, 8, 1>)); break;
        case 3: BSB_LAUNCH((trace_refill_kernel<true, 32, 2>)); break;
        case 4: BSB_LAUNCH((trace_tiles_kernel<true, 4>)); break;
        case 6: BSB_LAUNCH((trace_tiles_kernel<true, 2>)); break;
        default: BSB_LAUNCH((trace_tiles_kernel<true, 3>)); break;
        }
    } else {
        switch (variant) {
        case 1: BSB_LAUNCH((trace_refill_kernel<false, 16, 8>)); break;
        case 2: BSB_LAUNCH((trace_refill_kernel<false, 8, 4>)); break;
        case 3: BSB_LAUNCH((trace_refill_kernel<false, 32, 8>)); break;
        case 4: BSB_LAUNCH((trace_tiles_kernel<false, 4>)); break;
        case 6: BSB_LAUNCH((trace_tiles_kernel<false, 2>)); break;
        default: BSB_LAUNCH((trace_tiles_kernel<false, 3>)); break;
        }
    }
#undef BSB_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_rinv5_selftest(double q_lo, double q_hi, int n, double *d_out2, cudaStream_t stream)
{
    rinv5_selftest_kernel<<<64, 256, 0, stream>>>(q_lo, q_hi, n, d_out2, d_out2 + 1);
    return cudaGetLastError();
}

}  // namespace bsb
