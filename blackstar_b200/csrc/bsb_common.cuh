// bsb_common.cuh -- shared declarations for libblackstar_b200 (sm_100a only).
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define BSB_HD __host__ __device__ __forceinline__
#define BSB_D __device__ __forceinline__
#else
#define BSB_HD inline
#define BSB_D inline
#endif

namespace bsb {

// Star record in device memory, in k-d leaf order (64 bytes, 16-byte aligned).  Besides the
// position and magnitude of bsb_star it carries the star's colour as three per-channel
// coefficients: massiv-io's HSI->RGB is linear in the saturation for a fixed hue,
//   channel = I * (1 + S * k_c),   k_c in {K, -1, 1-K} permuted by the hue sector,
//   K = cos(a)/cos(b) of the sector (oracle/oracle_thirdparty.c),
// so the two cosines and the division are paid once per star at upload, not once per hit.
// kr/kg/kb hold  sat_star * k_c ; the frame's starSaturation multiplies them in the kernel.
// (fx, fy, fz) is the position rounded to float: a leaf slot is rejected with six FP32
// operations unless it is within radius + 1e-6 of the query, and only then tested exactly.
struct StarRec {
    double x, y, z;
    double kr, kg, kb;
    // one 16-byte word for the leaf scan's single-precision pre-filter: magnitude + float position
    int32_t mag;
    float fx, fy, fz;
};
static_assert(sizeof(StarRec) == 64, "StarRec is 64 bytes");

// Bucketed k-d tree over the unit-sphere star positions (DESIGN.md "star map").
//   2^depth leaves of exactly kLeafSlots star records each (padded with records that can never
//   be in range), so a leaf is one unrolled burst of independent loads and needs no offsets.
//   Split planes of an implicit complete binary tree; node (d, i) = i-th node of depth d,
//   children (d+1, 2i) and (d+1, 2i+1); points with coord <= split go left, >= split go right.
//   Levels 0 .. top_levels-1: `top`, heap order, FLOAT with the split axis in the two lowest
//     mantissa bits -- 2^13 - 1 nodes = 32 KB, staged in shared memory by every CTA.  On these
//     levels the axis cycles with the level (x, y, z, x, ... as kdt's own build does), so a walk from
//     the root is a fully unrolled chain of 13 {load, subtract, compare} steps with no decode.
//   Deeper levels in groups of three: one 64-byte record per 3-level subtree (local heap order
//     1..7, doubles with the axis in the two lowest mantissa bits), so that three levels cost ONE
//     dependent L2 access instead of three.  Group g starts at record rec_off[g].
struct StarTreeDev {
    const float *top;
    const double *rec;
    const StarRec *stars;
    int32_t depth;        // leaves live at this depth
    int32_t top_levels;   // T; (depth - T) is a multiple of 3
    int32_t n_stars;
    int32_t pad;
    uint32_t rec_off[4];
};

constexpr int kLeafSlots = 8;
constexpr int kSmemTreeLevels = 13;                        // top levels staged in shared memory
constexpr int kSmemTreeNodes = (1 << kSmemTreeLevels) - 1; // 8191 floats = 32 KB

// Everything a ray needs that is constant over the frame.  Passed by value as a
// __grid_constant__ kernel parameter (constant bank, warp-uniform loads).
struct FrameParams {
    // camera: src/Raytracer.hs:40-51
    double cam[3];            // position
    double xa[3], ya[3], za[3]; // rows of lookAt's 3x3 are xa, ya, -za
    double fov;
    double e1[3];             // cam / |cam| : first axis of every ray's orbital plane
    double r0;                // |cam|
    double q0;                // quadrance cam (as the reference sums it)
    // integrator: src/Raytracer.hs:113-134
    double h;                 // stepSize
    double hh;                // h/2
    double h6;                // h/6
    double hh2;               // (h/2)^2
    double hhh;               // h*(h/2)
    double hsq6;              // h*h/6
    double h3;                // h/3
    double k13, k23;          // 1/3, 2/3: the step constants in the kernel's units (time in half steps)
    double k14;               // 1.4, the constant of the |pos|^-5 correction (the kernel now uses the literal:
                              // ptxas keeps it in a register either way, and the literal build is 0.9 % faster)
    // termination / disk: src/Raytracer.hs:58-65, 88-111
    double safe2, din2, dout2;
    double r_in, r_out;       // sqrt of din2, dout2 (diskColor' recomputes them per hit)
    double disk_rgb[3];
    double disk_opacity;
    int32_t disk_on;          // disk_opacity != 0 (src/Raytracer.hs:96)
    int32_t pad0_;
    // sky: src/StarMap.hs:93-115
    double star_intensity, star_saturation;
    StarTreeDev tree;
    // per-column / per-row ray offsets vx[x], vy[y] (src/Raytracer.hs:49-50), filled by
    // ray_tables_kernel once per frame; nullptr = compute them per ray
    const double *vx_tab;
    const double *vy_tab;
    // grid
    int32_t W2, H2;           // traced grid (doubled under supersampling)
    int32_t W, H;             // final image
    int32_t ss;               // supersampling 0/1
    int32_t row0, row1;       // final-image rows rendered by this launch
    int32_t tiles_x, tiles_y, n_tiles;
    uint32_t step_cap;        // the reference has no cap (Raytracer.hs:80-85); ours is a safety net
    uint32_t pad1_;
};

// Arguments of the bloom line kernel (image_kernels.cu: box3_kernel).  A "line" is a row of the
// image in the H launch and a column in the V launch.  A line may be assembled from up to
// kMaxSegments pieces (multi-GPU: the column of an H^3-filtered frame whose row tiles came from
// different GPUs), and only positions [x_lo, x_hi) of it are written (the rest is halo).
constexpr int kMaxSegments = 8;
#if defined(__CUDACC__)
struct BoxArgs {
    const float4 *seg_in[kMaxSegments];   // piece s of line l = seg_in[s] + l * seg_pitch[s]; it holds
    size_t seg_pitch[kMaxSegments];       //   positions [seg_start[s], seg_start[s+1]) of the line
    int seg_start[kMaxSegments + 1];
    int nseg;
    int n, lines;         // line length, number of lines
    int x_lo, x_hi;       // positions of each line that are written
    float4 *out;          // result of (line l, position x) -> out[(x - x_lo) * out_pitch + l]  (may be null)
    const float4 *img;    // combine == 1: indexed like out; may alias out
    const float4 *seg_img[kMaxSegments];  // combine == 2: the image TRANSPOSED, in pieces laid out like seg_in
    uint8_t *rgb8;        // optional: sRGB8 of the result at rgb8[(x - x_lo) * rgb8_pitch + 3 * l]
    const float *thr;     // 256 floats: sRGB8 thresholds (host_setup.cpp: srgb8_thresholds)
    size_t out_pitch, rgb8_pitch;
    int r;                // box radius (src/ImageFilters.hs:83)
    int combine;          // out = img + strength * blur (src/ImageFilters.hs:85-86); 0 = just the blur
    float norm;           // 1 / (2r+1)   (src/ImageFilters.hs:51)
    float strength;
};
#endif

// Per-launch counters (device memory, zeroed before the launch).
struct TraceCounters {
    unsigned long long steps;
    unsigned long long capped;
    unsigned long long star_hits;
    unsigned int next_tile;
    unsigned int pad;
};

}  // namespace bsb
