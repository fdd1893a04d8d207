/*
 * blackstar_oracle.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A line-by-line double-precision restatement (plain C11, -O2 -ffp-contract=off)
 * of the hot path of flannelhead/blackstar:
 *     src/Raytracer.hs:23-134, src/StarMap.hs:45-115, src/ImageFilters.hs:28-97,
 *     src/ConfigFile.hs:48-51,66-79, app/Main.hs:93-103,113-118.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (blackstar_b200/, libblackstar_b200.so) never links, imports or calls it.
 *
 * PARITY UNPINNED for four third-party behaviours whose source is not under
 * /root/reference (stack.yaml:1 pins them only through lts-13.16) and for which
 * the reference holds no test vector: massiv-io HSI->RGB (toPixelRGB),
 * massiv-io toWord8 rounding, kdt inRadius boundary inclusivity / result order,
 * linear normalize/lookAt.  They are restated from the published algorithms in
 * oracle_thirdparty.c so anyone with GHC can falsify them in one place.
 * The reference cannot be built here (no ghc/stack/cabal), so there is no
 * oracle/_ref.  What pins the oracle instead: closed-form known-answer tests
 * (capture threshold b_c, conserved E and |L|^2, straight radial ray, HSI mean,
 * box-blur impulse response / DC gain) in tests/test_oracle_kat.py.
 */
#ifndef BLACKSTAR_ORACLE_H
#define BLACKSTAR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/ConfigFile.hs:34-37 */
typedef struct {
    double pos[3];
    double look_at[3];
    double up[3];
    double fov;
} orc_camera;

/* src/ConfigFile.hs:20-31 (safeDistance is derived in render, Raytracer.hs:59-60) */
typedef struct {
    double step_size;
    double bloom_strength;
    double star_intensity;
    double star_saturation;
    double disk_hsi[3]; /* hue already divided by 360 (ConfigFile.hs:51) */
    double disk_opacity;
    double disk_inner;
    double disk_outer;
    int32_t bloom_divider;
    int32_t width;
    int32_t height;
    int32_t supersampling;
} orc_scene;

/* src/StarMap.hs:25: (V3 Double, (Int, Double, Double)) */
typedef struct {
    double pos[3];
    double hue;
    double sat;
    int32_t mag;
    int32_t pad_;
} orc_star;

typedef struct orc_tree orc_tree;

/* ---- third-party restatements (oracle_thirdparty.c) ---- */
void orc_hsi_to_rgb(double h, double s, double i, double rgb[3]); /* returns NaNs on out-of-range hue */
void orc_normalize(const double v[3], double out[3]);
void orc_look_at_rows(const double eye[3], const double center[3], const double up[3],
                      double xa[3], double ya[3], double za[3]);
uint8_t orc_to_word8(double x);
orc_tree *orc_tree_build(const orc_star *stars, size_t n);
void orc_tree_free(orc_tree *t);
size_t orc_tree_size(const orc_tree *t);
/* indices (into the array given to orc_tree_build) of all stars with qd <= r*r, in kdt's
 * result-list order; returns the count (at most cap are written). */
size_t orc_in_radius(const orc_tree *t, double radius, const double q[3], uint32_t *idx, size_t cap);

/* ---- StarMap.hs ---- */
void orc_star_color(int spectral_char, double *hue, double *sat);     /* :60-72 */
void orc_ra_dec_to_cartesian(double ra, double dec, double out[3]);    /* :74-75 */
/* :45-58 -- parse PPM binary catalogue (28-byte header, 28-byte records); returns #stars written */
size_t orc_read_ppm(const uint8_t *bytes, size_t len, orc_star *out, size_t cap);
void orc_star_lookup(const orc_tree *t, double intensity, double saturation, const double vel[3],
                     double rgb[3]);                                   /* :93-115 */

/* ---- Raytracer.hs ---- */
double orc_srgb(double x);                                             /* :23-27 */
void orc_blend(const double top[4], const double bottom[4], double out[4]); /* :34-37 */
void orc_generate_ray(const orc_camera *cam, int w, int h, int x, int y, double vel[3], double pos[3]); /* :40-51 */
void orc_rk4(double h, double h2, const double vel[3], const double pos[3], double nvel[3], double npos[3]); /* :113-134 */
/* traceRay for pixel (x,y) of a w x h grid (already doubled under SS). rgb out, returns #rk4 steps taken. :69-111 */
uint32_t orc_trace_ray(const orc_camera *cam, const orc_scene *scn, const orc_tree *t, int w, int h,
                       int x, int y, double rgb[3]);
/* render rows [row0,row1) of the FINAL image (post-supersample, pre-bloom), RGB f64, row-major.
 * nthreads<=0 -> all online cores.  total_steps may be NULL.  :53-67 */
int orc_render(const orc_camera *cam, const orc_scene *scn, const orc_tree *t, int row0, int row1,
               int nthreads, double *out_rgb, uint64_t *total_steps);

/* ---- ImageFilters.hs ---- */
void orc_supersample(const double *in_rgb, int h, int w, double *out_rgb); /* :88-97, in is h x w, out (h/2)x(w/2) */
void orc_box_blur(int r, int passes, double *img_rgb, int h, int w);       /* :28-78, in place */
void orc_bloom(double strength, int divider, const double *in_rgb, int h, int w, double *out_rgb); /* :80-86 */

/* writeImg's per-pixel map (Raytracer.hs:29-32): sRGB then toWord8 */
void orc_to_srgb8(const double *in_rgb, size_t npix, uint8_t *out_rgb8);

#ifdef __cplusplus
}
#endif
#endif
