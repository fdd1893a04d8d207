/*
 * blackstar_b200.h -- C ABI of libblackstar_b200.so
 *
 * Drop-in boundary for the hot path of flannelhead/blackstar (a Haskell program with no
 * FFI of its own).  The seam is the export lists
 *     module Raytracer    (render, writeImg)    src/Raytracer.hs:4
 *     module ImageFilters (bloom, supersample)  src/ImageFilters.hs:5
 * whose only caller is Main.doRender (app/Main.hs:105-125).  Each entry point below names
 * the reference function it replaces.  INTEGRATION.md shows the `foreign import ccall`
 * shim a maintainer adds on the Haskell side.
 *
 * Conventions
 *  - plain C types only; every int return is a status code, 0 = BSB_OK;
 *    bsb_last_error() returns a human-readable message for the last non-zero status.
 *  - the caller owns every host buffer; the library owns device memory and the ctx.
 *  - a ctx is not re-entrant: one call at a time per ctx, from any OS thread
 *    (GHC -threaded `safe` calls may migrate threads; the library calls cudaSetDevice itself).
 *  - there is no CPU fallback: without a usable sm_100 device bsb_create returns NULL and
 *    bsb_last_error(NULL) says why.
 *  - framebuffers are row-major float RGBA (x fastest), 16 bytes per pixel, linear light
 *    (pre-sRGB), alpha = 1.  This is the reference's `Image U RGB Double` narrowed to
 *    float32 (error <= 6e-8 relative, far inside the 1e-4 parity bar).
 */
#ifndef BLACKSTAR_B200_H
#define BLACKSTAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSB_OK 0
#define BSB_ERR_INVALID 1   /* bad argument (NULL, negative size, rows out of range ...) */
#define BSB_ERR_CUDA 2      /* a CUDA runtime call failed */
#define BSB_ERR_NCCL 3      /* NCCL missing or a NCCL call failed (multi-GPU only) */
#define BSB_ERR_UNSUPPORTED 4 /* e.g. an image row above 8 MB in a staged host copy */
#define BSB_ERR_STEPCAP 5   /* a ray hit the step cap (the reference would loop forever) */

typedef struct bsb_ctx bsb_ctx;

/* Camera of src/ConfigFile.hs:34-37 */
typedef struct bsb_camera {
    double pos[3];      /* position */
    double look_at[3];  /* lookAt   */
    double up[3];       /* upVec    */
    double fov;         /* fov (tangent of the view angle) */
} bsb_camera;

/* Scene of src/ConfigFile.hs:20-31.  safeDistance is not a field: render derives it
 * (src/Raytracer.hs:59-60).  disk_inner/disk_outer are RADII (render squares them, :61-62). */
typedef struct bsb_scene {
    double step_size;        /* stepSize        (default 0.3)  */
    double bloom_strength;   /* bloomStrength   (default 0.4)  */
    double star_intensity;   /* starIntensity   (default 0.7)  */
    double star_saturation;  /* starSaturation  (default 0.7)  */
    double disk_hsi[3];      /* diskColor, hue ALREADY divided by 360 (src/ConfigFile.hs:51) */
    double disk_opacity;     /* diskOpacity     (default 0)    */
    double disk_inner;       /* diskInner       (default 3)    */
    double disk_outer;       /* diskOuter       (default 12)   */
    int32_t bloom_divider;   /* bloomDivider    (default 25)   */
    int32_t width;           /* resolution fst  (final image)  */
    int32_t height;          /* resolution snd                 */
    int32_t supersampling;   /* supersampling   (0/1)          */
} bsb_scene;

/* Star of src/StarMap.hs:25 = (V3 Double, (Int, Double, Double)): unit vector, catalogue
 * magnitude x100, hue and saturation as produced by starColor' (src/StarMap.hs:60-72). */
typedef struct bsb_star {
    double pos[3];
    double hue;
    double sat;
    int32_t mag;
    int32_t pad_;
} bsb_star;

/* What the reference prints with timeAction (src/Util.hs:33-41), at usable resolution. */
typedef struct bsb_stats {
    uint64_t rays;        /* traced rays (includes the x4 of supersampling) */
    uint64_t steps;       /* RK4 steps taken over all rays */
    uint64_t capped;      /* rays stopped by the step cap (0 in any sane scene) */
    uint64_t star_hits;   /* stars splatted over all sky lookups */
    double trace_ms;      /* geodesic kernel, CUDA events */
    double bloom_ms;      /* both bloom kernels, CUDA events (0 if bloom is off) */
    double gather_ms;     /* multi-GPU tile gather (0 on one GPU) */
    double d2h_ms;        /* device->host copy of the framebuffer */
    double total_ms;      /* host wall clock of the call */
    int32_t n_gpus;
    int32_t launches;     /* kernels launched by the call */
} bsb_stats;

/* ---- lifetime ---------------------------------------------------------------------- */

/* n_gpus = 0: every visible device; otherwise devices 0..n_gpus-1.  NULL on failure. */
bsb_ctx *bsb_create(int n_gpus);
/* Explicit device list (one-process-per-GPU launchers pass their LOCAL_RANK). */
bsb_ctx *bsb_create_on(const int *devices, int n);
void bsb_destroy(bsb_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed bsb_create on this thread. */
const char *bsb_last_error(const bsb_ctx *ctx);
/* "blackstar_b200 <version> sm_100a" */
const char *bsb_version(void);
/* Use an existing cudaStream_t (as void*) for device 0's work instead of the ctx's own
 * stream, so a caller that owns the stream can bracket calls with its own events.
 * NULL restores the ctx's stream; pass cudaStreamLegacy ((void*)1) for the default stream. */
int bsb_set_stream(bsb_ctx *ctx, void *cuda_stream);

/* Tuning knobs.
 *   "copy_threads"  host threads that move staged chunks into pageable output buffers (default 8;
 *                   0 = leave copies into pageable memory to the driver)
 *   "step_cap"      RK4 steps after which a ray is abandoned (default 1e6; the reference has no cap and
 *                   would not terminate; a capped ray makes the render return BSB_ERR_STEPCAP)
 * "trace_variant" selects the schedule / build of the trace kernel; every variant
 * computes the same image bit for bit:
 *   6 (default) one tile of 32 rays per warp, <=128 registers (2 CTAs/SM)
 *   0 / 4       same schedule compiled for 3 / 4 CTAs per SM (80 / 64 registers)
 *   1 / 2 / 3   persistent warps with ballot compaction of live rays between blocks of 16 / 8 / 32 steps */
int bsb_set_option(bsb_ctx *ctx, const char *key, double value);

/* ---- star map: replaces StarMap.readTreeFromFile + the StarTree argument --------------
 * (src/StarMap.hs:82-85, src/Raytracer.hs:53).  The flat star list is copied, a
 * bucketed k-d tree is built on the host and uploaded once to every GPU of the ctx.
 * n = 0 clears the map: every sky lookup then returns black (an empty KdMap). */
int bsb_set_stars(bsb_ctx *ctx, const bsb_star *stars, size_t n);
/* Same, straight from a PPM-format binary catalogue (28-byte header + 28-byte records) as
 * StarMap.readMap parses it (src/StarMap.hs:45-58), applying starColor' (:60-72). */
int bsb_set_stars_ppm(bsb_ctx *ctx, const uint8_t *bytes, size_t len);
/* The file `--starmap` names (app/Main.hs:36,46-50), whichever it is: the reference's tree file stars.kdt
 * (what generate-tree writes: cereal's encoding of kdt's KdMap, src/StarMap.hs:28-41,87-88 -- the layout is
 * recalled from the packages, so every structural invariant is checked and anything else is refused) or the
 * PPM catalogue it was generated from.  Replaces StarMap.readTreeFromFile (src/StarMap.hs:82-85). */
int bsb_set_stars_file(bsb_ctx *ctx, const uint8_t *bytes, size_t len);
size_t bsb_star_count(const bsb_ctx *ctx);

/* ---- render: replaces Raytracer.render (src/Raytracer.hs:53-67) including the
 * supersample it applies (src/ImageFilters.hs:88-97) -------------------------------------
 * Renders rows [row0,row1) of the FINAL image (post-supersample, pre-bloom) on the ctx's
 * first GPU into a HOST buffer of (row1-row0)*width float4.  stats may be NULL. */
int bsb_render(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, int row0, int row1,
               float *out_rgba, bsb_stats *stats);
/* Same, into DEVICE memory on the ctx's first GPU, asynchronously on the ctx stream (no
 * host synchronisation unless stats != NULL). */
int bsb_render_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, int row0,
                      int row1, void *dev_out_rgba, bsb_stats *stats);

/* ---- bloom: replaces ImageFilters.bloom (src/ImageFilters.hs:80-86) ---------------------
 * out = img + strength * boxBlur (w `div` divider) 3 img.  in and out may alias. */
int bsb_bloom(bsb_ctx *ctx, double strength, int divider, int width, int height,
              const float *in_rgba, float *out_rgba);
int bsb_bloom_device(bsb_ctx *ctx, double strength, int divider, int width, int height,
                     const void *dev_in_rgba, void *dev_out_rgba);

/* ---- what Main.doRender does between reading the scene and writeImg
 * (app/Main.hs:105-118): render, bloom if bloom_strength != 0, copy to the host.
 * On a ctx with N > 1 GPUs: GPU k traces a tile of rows and runs the horizontal half of the bloom
 * on it; ONE all-to-all over NVLink re-cuts the frame into column bands; GPU k runs the vertical
 * half + `img + strength * blur` on its band, and the N bands are copied into out_rgba in parallel
 * (N PCIe links).  out_rgba may be pageable (plain malloc) or page-locked; see "copy_threads". */
int bsb_render_full(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn,
                    float *out_rgba, bsb_stats *stats);
/* Same work, frame left on the GPU(s); asynchronous (bsb_synchronize waits for it). */
int bsb_render_full_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn,
                           int want_float, int want_rgb8);
/* Waits for everything queued on the ctx's GPUs; BSB_ERR_STEPCAP if the last trace launch capped a ray
 * (the asynchronous entry points cannot report that when they return). */
int bsb_synchronize(bsb_ctx *ctx);

/* ---- writeImg's pixel map (src/Raytracer.hs:23-32): sRGB then toWord8, RGB8 out ------- */
int bsb_to_srgb8(bsb_ctx *ctx, int width, int height, const float *in_rgba, uint8_t *out_rgb8);
int bsb_to_srgb8_device(bsb_ctx *ctx, int width, int height, const void *dev_in_rgba,
                        void *dev_out_rgb8);
/* doRender + writeImg's map in one call: the sRGB + toWord8 map runs in the epilogue of the last
 * bloom launch, the float frame is never written, and only width*height*3 bytes cross PCIe.
 * This is what the reference-side shim calls before its PNG encoder. */
int bsb_render_full_srgb8(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn,
                          uint8_t *out_rgb8, bsb_stats *stats);
/* Both images of one render; either pointer may be NULL. */
int bsb_render_full_both(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn,
                         float *out_rgba, uint8_t *out_rgb8, bsb_stats *stats);

/* bloom with writeImg's map fused: writes the float frame and / or its sRGB8 image (either may be NULL) */
int bsb_bloom_to_device(bsb_ctx *ctx, double strength, int divider, int width, int height,
                        const void *dev_in_rgba, void *dev_out_rgba, void *dev_out_rgb8);
/* rows x width_bytes of dense device rows into a host frame with row pitch host_pitch, ordered after
 * the work queued on the ctx stream (asynchronous for page-locked memory, staged for pageable). */
int bsb_download_2d(bsb_ctx *ctx, void *host_dst, size_t host_pitch, const void *dev_src,
                    size_t dev_pitch, size_t width_bytes, int rows);

/* ---- the two halves of the bloom as separate device-side steps, for launchers that run ONE
 * PROCESS PER GPU and own the exchange between them (blackstar_b200/dist.py under torchrun):
 *   h: rows x width tile -> its box^3-filtered rows, TRANSPOSED ([width][rows] float4), and
 *      optionally the tile itself transposed;  no communication (boxBlur's horizontal sweeps,
 *      src/ImageFilters.hs:72-73, commute with the vertical ones);
 *   v: `cols` columns of the full-height frame, each assembled from nseg row-tile pieces
 *      (piece s of column l = seg_midT[s] + l * seg_rows[s]); writes img + strength * blur for
 *      the band as [height][cols] float4 and / or its sRGB8 image [height][cols*3]. */
int bsb_bloom_h_device(bsb_ctx *ctx, int radius, int width, int rows, const void *dev_rows_rgba,
                       void *dev_midT, void *dev_imgT);
int bsb_bloom_v_device(bsb_ctx *ctx, double strength, int radius, int height, int cols, int nseg,
                       const void *const *seg_midT, const void *const *seg_imgT,
                       const int *seg_rows, void *dev_out_rgba, void *dev_out_rgb8);

/* ---- device micro-benchmarks used for the roofline denominators ------------------------
 * Dependent-free DFMA streams on every SM: returns achieved FP64 TFLOP/s (2 flops/FMA). */
int bsb_measure_fp64_peak(bsb_ctx *ctx, double *tflops);
/* device-to-device copy of `bytes` bytes, best of `reps`: returns GB/s (read+write). */
int bsb_measure_hbm_copy(bsb_ctx *ctx, size_t bytes, int reps, double *gbs);

/* Numerics self-test of the geodesic kernel's |pos|^-5 primitive (MUFU.RSQ64H seed +
 * first-order correction, returns 0.4 q^-5/2): max relative error against 0.4 pow(q, -2.5)
 * and the largest seed residual e = |1 - q*y0^2| over n log-spaced q in [q_lo, q_hi].  The
 * error bound of the primitive is 4.375 e^2 (1.4e-11 for the measured e = 2^-19.1). */
int bsb_selftest_rinv5(bsb_ctx *ctx, double q_lo, double q_hi, int n, double *max_rel_err,
                       double *max_seed_residual);

#ifdef __cplusplus
}
#endif
#endif /* BLACKSTAR_B200_H */
