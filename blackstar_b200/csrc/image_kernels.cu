// image_kernels.cu -- K4 bloom (ImageFilters.boxBlur/bloom, src/ImageFilters.hs:28-86),
// writeImg's pixel map (src/Raytracer.hs:23-32) and the roofline micro-benchmarks.
//
// Bloom.  The reference runs 3 x [horizontal sweep, vertical sweep] of a running-sum box
// filter whose window is [x-r+1, x+r] (2r taps) divided by 2r+1, with zero padding
// re-applied between sweeps (:41-46, :51, :59-64).  A horizontal sweep is I (x) T_w and a
// vertical one T_h (x) I on the image, so they commute exactly: HVHVHV = H^3 V^3 up to
// rounding.  box3_kernel does all three 1-D sweeps of a line while the line sits on the SM and
// writes the result TRANSPOSED; running it twice gives H^3 then V^3 and restores the
// orientation.  `img + strength * blurred` (:85-86) and, optionally, writeImg's sRGB + toWord8
// map are fused into the second launch.
//   traffic: launch 1 reads 16 B/px, writes 16 B/px; launch 2 reads 2 x 16 B/px, writes
//   16 B/px (+3 B/px RGB8)  => 5 x 16 B/px  (the algorithmic minimum is 2 x 16 B/px).
//
// A sweep is a difference of prefix sums, window(x) = S(x+r) - S(x-r).  S is kept in two parts:
// an FP64 base per thread chunk (C pixels) and a FLOAT prefix inside the chunk (at most C terms,
// exact to 1e-7 of C), so the shared-memory traffic of a sweep is 4 bytes per pixel and channel
// written + 8 read (the first version kept S in FP64: 3x the wavefronts, one CTA per SM).
// Window sums are formed as (float)(base_hi - base_lo) + (loc_hi - loc_lo): the FP64 subtraction
// removes the large common part, what is left is at most 2r+C terms, so the float rounding is
// <= 1.2e-7 of the output scale per sweep (the parity tests hold 1e-5 against the FP64 oracle).
#include "bsb_common.cuh"

#include <cuda_runtime.h>

namespace bsb {

constexpr unsigned kFullMask = 0xffffffffu;
#ifndef BSB_BLOOM_CTAS
#define BSB_BLOOM_CTAS 3   // resident 512-thread CTAs per SM the bloom kernel is compiled for
#endif

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256); the address must be 32-byte aligned.
__device__ __forceinline__ void ld256(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(float4 *p, const float4 &a, const float4 &b)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// ---- writeImg's map: sRGB (Raytracer.hs:23-27) then toWord8 = round-half-even(255*clamp01).
// The map float -> u8 is monotone, so it is a count of thresholds: thr[k] (k = 1..255) is the
// smallest float whose level is >= k, found on the host by bisection over the SAME double
// arithmetic the reference uses (host_setup.cpp: srgb8_thresholds).  The device guesses the level
// with MUFU-grade arithmetic and corrects it against the table: exact, and no FP64 pow.
__device__ __forceinline__ unsigned srgb8_level(float x, const float *__restrict__ thr)
{
    if (!(x > 0.0f)) return 0u;                    // negative, zero, NaN -> 0 (toWord8 clamps; NaN -> 0)
    const float s = x < 0.0031308f ? 12.92f * x : 1.055f * __powf(x, 1.0f / 2.4f) - 0.055f;
    int g = __float2int_rn(255.0f * fminf(s, 1.0f));
    g = g < 0 ? 0 : (g > 255 ? 255 : g);
    while (g < 255 && x >= __ldg(&thr[g + 1])) g++;
    while (g > 0 && x < __ldg(&thr[g])) g--;
    return (unsigned)g;
}

// One CTA per PAIR of lines of n <= C*T pixels; thread t owns pixels [tC, tC+C) of both lines.
// A line is read COALESCED into a conflict-free permuted staging layout and redistributed; the
// two lines of the pair are filtered one after the other and written together: in the transposed
// image they are adjacent, so each thread moves 32 contiguous bytes per pixel (one 256-bit store,
// and one 256-bit load of `img` in the combine launch) -- a full sector.
//
// Instruction count is what bounds this kernel (ncu: issue slots and L1TEX wavefronts, not DRAM), so
// the sweep is written to need no per-pixel index arithmetic, clamps or selects:
//  * DELTA = r mod C is a template parameter: with r = RHO*C + DELTA, pixel j of thread t looks at
//    element (j+DELTA) mod C of chunk t+RHO (+1 if j+DELTA >= C) and element (j-DELTA) mod C of chunk
//    t-RHO (-1 if j < DELTA) -- which plane and which of the two chunks is known at compile time;
//  * the prefix arrays carry one sentinel chunk on either side (index -1: sum 0, index T: the line
//    total), and chunk indices are clamped ONCE per thread, so windows that stick out of the line
//    need no special case; pixels beyond the line hold zeros, so the prefix is flat there;
//  * the scan of the chunk totals is a FLOAT warp scan (32 chunks = 32 C pixels: exact to 2e-5 of a
//    2r-pixel window sum) and FP64 only across the T/32 warp totals.
template <int C, int T, int DELTA>
__global__ void __launch_bounds__(T, (BSB_BLOOM_CTAS * 512 / T) > 4 ? 4 : (BSB_BLOOM_CTAS * 512 / T)) box3_kernel(const __grid_constant__ BoxArgs A)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int NW = T / 32;
    constexpr int TP = T + 32 / C;                    // plane pitch of the staging layout: conflict free
    constexpr int TT = T + 2;                         // plane pitch of the prefix arrays: chunks -1 .. T
    float *L = reinterpret_cast<float *>(s_raw);      // [3][C][TT] float prefix inside the chunk
    double *B = reinterpret_cast<double *>(s_raw + sizeof(float) * 3 * C * TT);  // [3][TT] FP64 base of the chunk (exclusive)
    double *WT = B + 3 * TT;                          // [3][NW] warp totals
    float *S = L;                                     // staging [3][C*TP] floats, aliases L (+ part of B): dead by then
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n = A.n;
    const int line0 = 2 * blockIdx.x;
    const int n_here = (line0 + 1 < A.lines) ? 2 : 1;
    const int x0 = t * C;
    const int rho = A.r / C;                          // A.r = rho * C + DELTA (the host picks the instantiation)
    // chunk indices (+1 for the front sentinel) of the four chunks my windows touch
    const int ih0 = min(t + rho, T) + 1, ih1 = min(t + rho + 1, T) + 1;
    const int il0 = max(t - rho - 1, -1) + 1, il1 = max(t - rho, -1) + 1;
    const int nvalid = min(max(n - x0, 0), C);        // how many of my pixels are inside the line
    const bool full = n == C * T;

    float res[C][3];
    float v[C][3];
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        if (half == 1) {
#pragma unroll
            for (int j = 0; j < C; j++) { res[j][0] = v[j][0]; res[j][1] = v[j][1]; res[j][2] = v[j][2]; }
        }
        if (half >= n_here) {
#pragma unroll
            for (int j = 0; j < C; j++) v[j][0] = v[j][1] = v[j][2] = 0.0f;
            break;
        }
        {
#pragma unroll
            for (int i = 0; i < C; i++) {
                const int x = i * T + t;
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (x < n) {
                    int sg = 0;
                    while (sg + 1 < A.nseg && x >= A.seg_start[sg + 1]) sg++;
                    p = A.seg_in[sg][(size_t)(line0 + half) * A.seg_pitch[sg] + (size_t)(x - A.seg_start[sg])];
                }
                const int o = (x % C) * TP + x / C;
                S[o] = p.x; S[C * TP + o] = p.y; S[2 * C * TP + o] = p.z;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < C; j++) {
                v[j][0] = S[j * TP + t]; v[j][1] = S[C * TP + j * TP + t]; v[j][2] = S[2 * C * TP + j * TP + t];
            }
            __syncthreads();
            if (t < 3 * C) { L[t * TT] = 0.0f; L[t * TT + T + 1] = 0.0f; }   // sentinels: nothing inside the chunk
            if (t < 3) B[t * TT] = 0.0;
        }
#pragma unroll 1
        for (int pass = 0; pass < 3; pass++) {
            // inclusive float prefix inside the chunk; float scan of the chunk totals over the warp
            float wex[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
#pragma unroll
                for (int j = 1; j < C; j++) v[j][c] += v[j - 1][c];
                float x = v[C - 1][c];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float y = __shfl_up_sync(kFullMask, x, o);
                    if (lane >= o) x += y;
                }
                wex[c] = x - v[C - 1][c];
                if (lane == 31) WT[c * NW + warp] = (double)x;
#pragma unroll
                for (int j = 0; j < C; j++) L[(c * C + j) * TT + t + 1] = v[j][c];
            }
            __syncthreads();
            // FP64 across the warps: warp w adds the totals of warps 0 .. w-1 (every warp does the
            // small scan itself: no second barrier).  For NW = 16 two channels share one scan.
            double base[3];
            if (NW == 16) {
                double x = WT[lane];                             // channels 0 and 1: 2 x 16 totals
                double z = WT[2 * NW + (lane & 15)];
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const double y = __shfl_up_sync(kFullMask, x, o), y2 = __shfl_up_sync(kFullMask, z, o);
                    if ((lane & 15) >= o) { x += y; z += y2; }
                }
                base[0] = __shfl_sync(kFullMask, x, (warp + 15) & 15);
                base[1] = __shfl_sync(kFullMask, x, 16 + ((warp + 15) & 15));
                base[2] = __shfl_sync(kFullMask, z, (warp + 15) & 15);
            } else if (NW == 8) {
                double x = lane < 24 ? WT[lane] : 0.0;           // all three channels: 3 x 8 totals
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const double y = __shfl_up_sync(kFullMask, x, o);
                    if ((lane & 7) >= o) x += y;
                }
                base[0] = __shfl_sync(kFullMask, x, (warp + 7) & 7);
                base[1] = __shfl_sync(kFullMask, x, 8 + ((warp + 7) & 7));
                base[2] = __shfl_sync(kFullMask, x, 16 + ((warp + 7) & 7));
            } else {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double x = WT[c * NW + lane];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double y = __shfl_up_sync(kFullMask, x, o);
                        if (lane >= o) x += y;
                    }
                    base[c] = __shfl_sync(kFullMask, x, (warp + 31) & 31);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double b = (warp > 0 ? base[c] : 0.0) + (double)wex[c];
                B[c * TT + t + 1] = b;
                if (t == T - 1) B[c * TT + T + 1] = b + (double)v[C - 1][c];   // back sentinel: the line total
            }
            __syncthreads();
            // window [x-r+1, x+r] = S(x+r) - S(x-r); S(k) = B[chunk] + L[k]
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double ref = B[c * TT + il0];
                const float fh0 = (float)(B[c * TT + ih0] - ref), fh1 = (float)(B[c * TT + ih1] - ref);
                const float fl1 = (float)(B[c * TT + il1] - ref);
                const float *Lc = L + c * C * TT;
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int ph = (j + DELTA) % C, oh = (j + DELTA) / C;        // compile time
                    const int pl = (j - DELTA + C) % C, ol = j >= DELTA ? 1 : 0;
                    const float s_hi = (oh ? fh1 : fh0) + Lc[ph * TT + (oh ? ih1 : ih0)];
                    const float s_lo = (ol ? fl1 : 0.0f) + Lc[pl * TT + (ol ? il1 : il0)];
                    v[j][c] = A.norm * (s_hi - s_lo);
                }
                if (!full) {
#pragma unroll
                    for (int j = 0; j < C; j++) v[j][c] = j < nvalid ? v[j][c] : 0.0f;
                }
            }
            __syncthreads();
        }
    }

    // 32-byte aligned pairs: even pitch and 32-byte aligned buffers
    const bool wide = n_here == 2 && (A.out_pitch & 1) == 0 &&
                      ((reinterpret_cast<size_t>(A.out) | (A.combine == 1 ? reinterpret_cast<size_t>(A.img) : 0)) & 31) == 0;
#pragma unroll
    for (int j = 0; j < C; j++) {
        const int x = x0 + j;
        if (x < A.x_lo || x >= A.x_hi || x >= n) continue;
        const size_t o = (size_t)(x - A.x_lo) * A.out_pitch + line0;
        float4 a = make_float4(res[j][0], res[j][1], res[j][2], 1.0f);
        float4 b = make_float4(v[j][0], v[j][1], v[j][2], 1.0f);
        if (A.combine) {
            float4 pa, pb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A.combine == 2) {
                int sg = 0;
                while (sg + 1 < A.nseg && x >= A.seg_start[sg + 1]) sg++;
                const float4 *q = A.seg_img[sg] + (size_t)line0 * A.seg_pitch[sg] + (size_t)(x - A.seg_start[sg]);
                pa = *q;
                if (n_here == 2) pb = q[A.seg_pitch[sg]];
            } else if (wide) ld256(&A.img[o], pa, pb);
            else { pa = A.img[o]; if (n_here == 2) pb = A.img[o + 1]; }
            a = make_float4(fmaf(A.strength, a.x, pa.x), fmaf(A.strength, a.y, pa.y), fmaf(A.strength, a.z, pa.z), pa.w);
            b = make_float4(fmaf(A.strength, b.x, pb.x), fmaf(A.strength, b.y, pb.y), fmaf(A.strength, b.z, pb.z), pb.w);
        }
        if (A.out) {
            if (wide) st256(&A.out[o], a, b);
            else { A.out[o] = a; if (n_here == 2) A.out[o + 1] = b; }
        }
        if (A.rgb8) {
            uint8_t *q = A.rgb8 + (size_t)(x - A.x_lo) * A.rgb8_pitch + 3 * (size_t)line0;
            const unsigned a0 = srgb8_level(a.x, A.thr), a1 = srgb8_level(a.y, A.thr), a2 = srgb8_level(a.z, A.thr);
            if (n_here == 2) {
                const unsigned b0 = srgb8_level(b.x, A.thr), b1 = srgb8_level(b.y, A.thr), b2 = srgb8_level(b.z, A.thr);
                if ((reinterpret_cast<size_t>(q) & 1) == 0) {   // 3 * line0 is even: three 16-bit stores
                    unsigned short *q2 = reinterpret_cast<unsigned short *>(q);
                    q2[0] = (unsigned short)(a0 | a1 << 8); q2[1] = (unsigned short)(a2 | b0 << 8); q2[2] = (unsigned short)(b1 | b2 << 8);
                } else {
                    q[0] = (uint8_t)a0; q[1] = (uint8_t)a1; q[2] = (uint8_t)a2; q[3] = (uint8_t)b0; q[4] = (uint8_t)b1; q[5] = (uint8_t)b2;
                }
            } else {
                q[0] = (uint8_t)a0; q[1] = (uint8_t)a1; q[2] = (uint8_t)a2;
            }
        }
    }
}

template <int C, int T, int DELTA>
static cudaError_t launch_box3_ctd(const BoxArgs &A, cudaStream_t stream)
{
    constexpr size_t smem = sizeof(float) * 3 * C * (T + 2) + sizeof(double) * (3 * (T + 2) + 3 * (T / 32));
    static_assert(sizeof(float) * 3 * C * (T + 32 / C) <= smem, "the staging layout must fit");
    auto kern = box3_kernel<C, T, DELTA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(A.lines + 1) / 2, T, smem, stream>>>(A);
    return cudaGetLastError();
}

template <int C, int T>
static cudaError_t launch_box3_ct(const BoxArgs &A, cudaStream_t stream)
{
    switch (A.r % C) {
#define BSB_D(D) case D: return launch_box3_ctd<C, T, (D < C ? D : 0)>(A, stream)
        BSB_D(0); BSB_D(1); BSB_D(2); BSB_D(3); BSB_D(4); BSB_D(5); BSB_D(6); BSB_D(7);
#undef BSB_D
    }
    return cudaErrorInvalidValue;
}

// longest line the shared-memory kernel takes; longer lines go through launch_bloom_long (below)
int bloom_max_line() { return 8 * 1024; }

cudaError_t launch_box3(const BoxArgs &A, cudaStream_t stream)
{
    if (A.lines <= 0 || A.n <= 0) return cudaSuccess;
    const int n = A.n;   // every thread of a CTA works whether its pixels are inside the line or not: size it
    if (n <= 512) return launch_box3_ct<2, 256>(A, stream);
    if (n <= 1024) return launch_box3_ct<4, 256>(A, stream);
    if (n <= 2048) return launch_box3_ct<8, 256>(A, stream);
    if (n <= 4096) return launch_box3_ct<8, 512>(A, stream);
    if (n <= 8192) return launch_box3_ct<8, 1024>(A, stream);
    return cudaErrorInvalidValue;
}

// ---- float4 transpose: in [rows][cols] (pitch in_pitch) -> out [cols][rows] (pitch out_pitch)
__global__ void __launch_bounds__(256) transpose_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int rows, int cols,
                                                        size_t in_pitch, size_t out_pitch)
{
    __shared__ float4 tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int y = by + ty + k, x = bx + tx;
        if (y < rows && x < cols) tile[ty + k][tx] = in[(size_t)y * in_pitch + x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int x = bx + ty + k, y = by + tx;
        if (x < cols && y < rows) out[(size_t)x * out_pitch + y] = tile[tx][ty + k];
    }
}

cudaError_t launch_transpose(const float4 *in, float4 *out, int rows, int cols, size_t in_pitch, size_t out_pitch, cudaStream_t stream)
{
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, stream>>>(in, out, rows, cols, in_pitch, out_pitch);
    return cudaGetLastError();
}

// ---- lines longer than the shared-memory kernel takes: the reference's own algorithm, one thread
// per (line, channel) walking along the line with an FP64 running sum (src/ImageFilters.hs:59-64).
// dir = 0: lines are rows (walk along x); dir = 1: lines are columns (walk along y, coalesced).
__global__ void __launch_bounds__(256) box_pass_seq_kernel(const float4 *in, float4 *out, int W, int H, int r, int dir, double norm)
{
    const int n = dir == 0 ? W : H, lines = dir == 0 ? H : W;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(g & 3), l = (int)(g >> 2);
    if (l >= lines) return;
    const float *src = reinterpret_cast<const float *>(in) + c;
    float *dst = reinterpret_cast<float *>(out) + c;
    const size_t step = dir == 0 ? 4 : (size_t)W * 4;
    const size_t base = dir == 0 ? (size_t)l * W * 4 : (size_t)l * 4;
    if (c == 3) {
        for (int x = 0; x < n; x++) dst[base + x * step] = 1.0f;
        return;
    }
    double acc = 0.0;                                    // :59 start = sum of the first r pixels
    for (int x = 0; x < r && x < n; x++) acc += (double)src[base + x * step];
    for (int x = 0; x < n; x++) {                        // :61-63 acc + p(x+r) - p(x-r)
        const double add = x + r < n ? (double)src[base + (size_t)(x + r) * step] : 0.0;
        const double sub = x - r >= 0 ? (double)src[base + (size_t)(x - r) * step] : 0.0;
        acc = acc + add - sub;
        dst[base + x * step] = (float)(acc * norm);
    }
}

__global__ void __launch_bounds__(256) combine_kernel(const float4 *blur, const float4 *img, float4 *out, uint8_t *rgb8,
                                                      const float *thr, size_t npix, float strength)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 b = blur[i], p = img[i];
    const float4 o = make_float4(fmaf(strength, b.x, p.x), fmaf(strength, b.y, p.y), fmaf(strength, b.z, p.z), p.w);
    if (out) out[i] = o;
    if (rgb8) {
        rgb8[3 * i + 0] = (uint8_t)srgb8_level(o.x, thr); rgb8[3 * i + 1] = (uint8_t)srgb8_level(o.y, thr);
        rgb8[3 * i + 2] = (uint8_t)srgb8_level(o.z, thr);
    }
}

// img (H x W) -> out = img + strength * box^3(img), any size; tmp_a / tmp_b are two H x W scratch frames
cudaError_t launch_bloom_long(const float4 *img, float4 *out, uint8_t *rgb8, const float *thr, float4 *tmp_a, float4 *tmp_b,
                              int W, int H, int r, float strength, cudaStream_t stream)
{
    const double norm = 1.0 / (2.0 * (double)r + 1.0);
    const float4 *src = img;
    float4 *bufs[2] = { tmp_a, tmp_b };
    int k = 0;
    for (int dir = 0; dir < 2; dir++)
        for (int pass = 0; pass < 3; pass++) {
            const size_t threads = (size_t)(dir == 0 ? H : W) * 4;
            box_pass_seq_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, bufs[k], W, H, r, dir, norm);
            src = bufs[k];
            k ^= 1;
        }
    const size_t npix = (size_t)W * H;
    combine_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(src, img, out, rgb8, thr, npix, strength);
    return cudaGetLastError();
}

// 4 pixels per thread: 64 B in, 12 B out
__global__ void __launch_bounds__(256) srgb8_kernel(const float4 *__restrict__ in, uint8_t *__restrict__ out,
                                                    const float *__restrict__ thr, size_t npix)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p0 = g * 4;
    if (p0 >= npix) return;
    if (p0 + 4 <= npix) {
        unsigned b[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 p = __ldg(&in[p0 + k]);
            b[3 * k + 0] = srgb8_level(p.x, thr); b[3 * k + 1] = srgb8_level(p.y, thr); b[3 * k + 2] = srgb8_level(p.z, thr);
        }
        uint32_t *o = reinterpret_cast<uint32_t *>(out + p0 * 3);  // p0*3 is a multiple of 12
        o[0] = b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24;
        o[1] = b[4] | b[5] << 8 | b[6] << 16 | b[7] << 24;
        o[2] = b[8] | b[9] << 8 | b[10] << 16 | b[11] << 24;
    } else {
        for (size_t p = p0; p < npix; p++) {
            const float4 q = in[p];
            out[p * 3 + 0] = (uint8_t)srgb8_level(q.x, thr); out[p * 3 + 1] = (uint8_t)srgb8_level(q.y, thr);
            out[p * 3 + 2] = (uint8_t)srgb8_level(q.z, thr);
        }
    }
}

cudaError_t launch_srgb8(const float4 *in, uint8_t *out, const float *thr, size_t npix, cudaStream_t stream)
{
    if (npix == 0) return cudaSuccess;
    const size_t threads = (npix + 3) / 4;
    srgb8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(in, out, thr, npix);
    return cudaGetLastError();
}

// ---- FP64 roofline denominator: 8 independent DFMA chains per thread, every SM full
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t launch_dfma_peak(double *sink, int blocks, int iters, cudaStream_t stream)
{
    dfma_peak_kernel<<<blocks, 256, 0, stream>>>(sink, iters, 0.999999, 1e-9);
    return cudaGetLastError();
}

}  // namespace bsb
