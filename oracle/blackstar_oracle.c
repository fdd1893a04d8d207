/*
 * blackstar_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see blackstar_oracle.h).
 *
 * Double-precision restatement of the reference's own code for the hot path.
 * Compile with -O2 -ffp-contract=off (no FMA contraction, no fast-math) so every
 * operation rounds exactly where the Haskell source rounds.  Operation ORDER follows
 * the Haskell fixity rules; each function cites the lines it follows.
 */
#define _GNU_SOURCE
#include "blackstar_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

const orc_star *orc_tree_star(const orc_tree *t, uint32_t i); /* oracle_thirdparty.c */

/* ---------------------------------------------------------------- StarMap.hs */

/* src/StarMap.hs:60-72 starColor */
void orc_star_color(int ch, double *hue, double *sat)
{
    switch (ch) {
    case 'O': *hue = 0.631; *sat = 0.39; break;
    case 'B': *hue = 0.628; *sat = 0.33; break;
    case 'A': *hue = 0.622; *sat = 0.21; break;
    case 'F': *hue = 0.650; *sat = 0.03; break;
    case 'G': *hue = 0.089; *sat = 0.09; break;
    case 'K': *hue = 0.094; *sat = 0.29; break;
    case 'M': *hue = 0.094; *sat = 0.56; break;
    default:  *hue = 0;     *sat = 0;    break;
    }
}

/* src/StarMap.hs:74-75 raDecToCartesian ra dec = V3 (cos dec*cos ra) (cos dec*sin ra) (sin dec) */
void orc_ra_dec_to_cartesian(double ra, double dec, double out[3])
{
    out[0] = cos(dec) * cos(ra);
    out[1] = cos(dec) * sin(ra);
    out[2] = sin(dec);
}

static double get_f64be(const uint8_t *p)
{
    uint64_t u = 0;
    for (int k = 0; k < 8; k++) u = (u << 8) | p[k];
    double d;
    memcpy(&d, &u, 8);
    return d;
}

/* src/StarMap.hs:45-58 readMap: skip 28; n = remaining `div` 28 records of
 *   f64be ra, f64be dec, u8 spectral, skip 1, i16be mag, skip 8;
 * then starColor' (:60-61) as readTreeFromFile applies it (:85). */
size_t orc_read_ppm(const uint8_t *bytes, size_t len, orc_star *out, size_t cap)
{
    if (len < 28) return 0;
    size_t n = (len - 28) / 28;
    const uint8_t *p = bytes + 28;
    size_t k;
    for (k = 0; k < n && k < cap; k++, p += 28) {
        double ra = get_f64be(p), dec = get_f64be(p + 8);
        int spectral = p[16];
        int16_t mag = (int16_t)((uint16_t)p[18] << 8 | p[19]);
        orc_ra_dec_to_cartesian(ra, dec, out[k].pos);
        out[k].mag = mag;
        out[k].pad_ = 0;
        orc_star_color(spectral, &out[k].hue, &out[k].sat);
    }
    return k;
}

/* src/StarMap.hs:93-115 starLookup */
void orc_star_lookup(const orc_tree *t, double intensity, double saturation, const double vel[3],
                     double rgb[3])
{
    const double max_brightness = 950; /* :99 */
    const double dynamic = 50;         /* :100 */
    const double w = 0.0005;           /* :101 */
    double nvel[3];
    uint32_t idx[256];
    orc_normalize(vel, nvel);                                   /* :103 */
    size_t n = orc_in_radius(t, 3 * w, nvel, idx, 256);         /* :104 */
    if (n > 256) n = 256;
    double acc[3] = { 0, 0, 0 };                                /* :115 foldl' (liftA2 (+)) (PixelRGB 0 0 0) */
    for (size_t k = 0; k < n; k++) {
        const orc_star *s = orc_tree_star(t, idx[k]);
        double dx = s->pos[0] - nvel[0], dy = s->pos[1] - nvel[1], dz = s->pos[2] - nvel[2];
        double d2 = dx * dx + dy * dy + dz * dz;                /* :107 qd pos nvel */
        double a = log(2) / dynamic;                            /* :108 */
        double e = exp(a * (max_brightness - (double)s->mag) - d2 / (2 * (w * w))); /* :113 */
        double val = (e < 1 ? e : 1) * intensity;               /* :112 (* intensity) . min 1 */
        double c[3];
        orc_hsi_to_rgb(s->hue, saturation * s->sat, val, c);    /* :114 */
        acc[0] = acc[0] + c[0]; acc[1] = acc[1] + c[1]; acc[2] = acc[2] + c[2];
    }
    for (int k = 0; k < 3; k++) rgb[k] = acc[k] < 1 ? acc[k] : 1; /* :115 fmap (min 1) */
}

/* ------------------------------------------------------------- Raytracer.hs */

/* src/Raytracer.hs:23-27 */
double orc_srgb(double x)
{
    const double a = 0.055;
    if (x < 0.0031308) return 12.92 * x;
    return (1 + a) * pow(x, 1.0 / 2.4) - a;
}

/* src/Raytracer.hs:34-37 blend src dst: comp tc bc = tc + bc * (1 - ta), on all four channels */
void orc_blend(const double top[4], const double bottom[4], double out[4])
{
    double ta = top[3];
    for (int k = 0; k < 4; k++) out[k] = top[k] + bottom[k] * (1 - ta);
}

/* src/Raytracer.hs:40-51 generateRay; (w,h) is the (possibly doubled) resolution of cfg' (:63) */
void orc_generate_ray(const orc_camera *cam, int wi, int hi, int xi, int yi, double vel[3], double pos[3])
{
    double w = (double)wi, h = (double)hi;
    double xa[3], ya[3], za[3];
    orc_look_at_rows(cam->pos, cam->look_at, cam->up, xa, ya, za); /* :47 */
    double vx = cam->fov * ((double)xi / w - 0.5);                 /* :49 */
    double vy = cam->fov * (0.5 - (double)yi / h) * h / w;         /* :50 ((fov*(..))*h)/w */
    double vz = -1;                                                /* :51 */
    /* :48 transpose matr !* v : component i = (xa_i*vx + ya_i*vy) + (-za_i)*vz */
    double u[3];
    for (int k = 0; k < 3; k++) u[k] = (xa[k] * vx + ya[k] * vy) + (-za[k]) * vz;
    orc_normalize(u, vel);
    pos[0] = cam->pos[0]; pos[1] = cam->pos[1]; pos[2] = cam->pos[2];
}

/* f of src/Raytracer.hs:124-127: (vel,pos) -> (-1.5*h2 / (norm pos ^ 5) *^ pos, vel) */
static void rk_f(double h2, const double vel[3], const double pos[3], double dvel[3], double dpos[3])
{
    double n = sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]); /* norm = sqrt . quadrance */
    double n2 = n * n;      /* x^5 by GHC's (^): f x 5 -> g (x*x) 2 x -> g (x2*x2) 1 x -> x4*x */
    double n5 = (n2 * n2) * n;
    double c = -(1.5 * h2 / n5); /* prefix minus binds looser than * and /; negation is exact */
    dvel[0] = c * pos[0]; dvel[1] = c * pos[1]; dvel[2] = c * pos[2];
    dpos[0] = vel[0]; dpos[1] = vel[1]; dpos[2] = vel[2];
}

/* src/Raytracer.hs:113-134 rk4 */
void orc_rk4(double h, double h2, const double vel[3], const double pos[3], double nvel[3], double npos[3])
{
    double k1v[3], k1p[3], k2v[3], k2p[3], k3v[3], k3p[3], k4v[3], k4p[3], tv[3], tp[3];
    double hh = h / 2;
    rk_f(h2, vel, pos, k1v, k1p);                                            /* :129 */
    for (int k = 0; k < 3; k++) { tv[k] = vel[k] + k1v[k] * hh; tp[k] = pos[k] + k1p[k] * hh; }
    rk_f(h2, tv, tp, k2v, k2p);                                              /* :130 */
    for (int k = 0; k < 3; k++) { tv[k] = vel[k] + k2v[k] * hh; tp[k] = pos[k] + k2p[k] * hh; }
    rk_f(h2, tv, tp, k3v, k3p);                                              /* :131 */
    for (int k = 0; k < 3; k++) { tv[k] = vel[k] + k3v[k] * h; tp[k] = pos[k] + k3p[k] * h; }
    rk_f(h2, tv, tp, k4v, k4p);                                              /* :132 */
    double h6 = h / 6;
    for (int k = 0; k < 3; k++) {
        /* :133 sumK = ((k1 + 2*k2) + 2*k3) + k4 ; :134 y + (h/6)*sumK */
        double sv = ((k1v[k] + k2v[k] * 2) + k3v[k] * 2) + k4v[k];
        double sp = ((k1p[k] + k2p[k] * 2) + k3p[k] * 2) + k4p[k];
        nvel[k] = vel[k] + sv * h6;
        npos[k] = pos[k] + sp * h6;
    }
}

static double signum(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }

typedef struct {
    double safe2, din2, dout2; /* :59-62, squared radii */
    double disk_rgb[3];        /* :65 */
    int w, h;                  /* :58,63 */
} frame_consts;

static void frame_setup(const orc_camera *cam, const orc_scene *scn, frame_consts *fc)
{
    double q = cam->pos[0] * cam->pos[0] + cam->pos[1] * cam->pos[1] + cam->pos[2] * cam->pos[2];
    double s = 2 * q;
    fc->safe2 = 2500.0 > s ? 2500.0 : s;                  /* :60 max (50^2) (2 * quadrance pos) */
    fc->din2 = scn->disk_inner * scn->disk_inner;         /* :61 */
    fc->dout2 = scn->disk_outer * scn->disk_outer;        /* :62 */
    fc->w = scn->supersampling ? 2 * scn->width : scn->width;   /* :58 */
    fc->h = scn->supersampling ? 2 * scn->height : scn->height;
    orc_hsi_to_rgb(scn->disk_hsi[0], scn->disk_hsi[1], scn->disk_hsi[2], fc->disk_rgb); /* :65 */
}

/* src/Raytracer.hs:104-111 diskColor' (diskInner/diskOuter hold SQUARED radii here, :61-62) */
static void disk_color(const orc_scene *scn, const frame_consts *fc, double r, double rgba[4])
{
    const double pi = 3.141592653589793;
    double r_inner = sqrt(fc->din2), r_outer = sqrt(fc->dout2);
    double q = (r_outer - r) / (r_outer - r_inner);
    double intensity = sin(pi * (q * q));
    rgba[0] = fc->disk_rgb[0] * intensity;
    rgba[1] = fc->disk_rgb[1] * intensity;
    rgba[2] = fc->disk_rgb[2] * intensity;
    rgba[3] = intensity * scn->disk_opacity;
}

static uint32_t trace_with(const orc_camera *cam, const orc_scene *scn, const frame_consts *fc,
                           const orc_tree *t, int x, int y, double rgb[3])
{
    double vel[3], pos[3];
    orc_generate_ray(cam, fc->w, fc->h, x, y, vel, pos);                 /* :72 */
    double cr[3] = { pos[1] * vel[2] - pos[2] * vel[1], pos[2] * vel[0] - pos[0] * vel[2],
                     pos[0] * vel[1] - pos[1] * vel[0] };
    double h2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];           /* :73 */
    double acc[4] = { 0, 0, 0, 0 };                                      /* :86 */
    uint32_t steps = 0;
    for (;;) {                                                           /* :80-85 colorize' */
        double nvel[3], npos[3];
        orc_rk4(scn->step_size, h2, vel, pos, nvel, npos);               /* :81 */
        steps++;
        /* :88-102 findColor on (vel,pos) and newPos */
        double yo = pos[1], yn = npos[1];
        double r2 = pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2];
        double r2n = npos[0] * npos[0] + npos[1] * npos[1] + npos[2] * npos[2];
        if (r2 < 1) {                                                    /* :93 */
            double c[4] = { 0, 0, 0, 1 }, o[4];
            orc_blend(acc, c, o);
            memcpy(acc, o, sizeof o);
            break;
        }
        if (r2 > fc->safe2) {                                            /* :94-95 */
            double c[4], o[4];
            if (t && orc_tree_size(t) > 0)
                orc_star_lookup(t, scn->star_intensity, scn->star_saturation, vel, c);
            else
                c[0] = c[1] = c[2] = 0; /* empty tree: inRadius returns [] -> PixelRGB 0 0 0 */
            c[3] = 1.0;
            orc_blend(acc, c, o);
            memcpy(acc, o, sizeof o);
            break;
        }
        if (scn->disk_opacity != 0 && signum(yn) != signum(yo)) {        /* :96 */
            double r2ave = (yn * r2 - yo * r2n) / (yn - yo);             /* :102 */
            if (r2ave > fc->din2 && r2ave < fc->dout2) {                 /* :97 */
                double c[4], o[4];
                disk_color(scn, fc, sqrt(r2ave), c);                     /* :98 */
                orc_blend(acc, c, o);                                    /* :83 */
                memcpy(acc, o, sizeof o);
            }
        }
        memcpy(vel, nvel, sizeof vel);
        memcpy(pos, npos, sizeof pos);
        if (steps > 100000000u) break; /* the reference has no cap (:80-85); never reached */
    }
    rgb[0] = acc[0]; rgb[1] = acc[1]; rgb[2] = acc[2];                   /* :75 dropAlpha */
    return steps;
}

uint32_t orc_trace_ray(const orc_camera *cam, const orc_scene *scn, const orc_tree *t, int w, int h,
                       int x, int y, double rgb[3])
{
    frame_consts fc;
    frame_setup(cam, scn, &fc);
    fc.w = w; fc.h = h;
    return trace_with(cam, scn, &fc, t, x, y, rgb);
}

/* ---- render: massiv makeArrayR U Par (src/Raytracer.hs:66) = all pixels, split over the
 *      RTS capabilities.  Here: pthreads, dynamic scheduling over final-image rows. ---- */
typedef struct {
    const orc_camera *cam;
    const orc_scene *scn;
    const orc_tree *t;
    frame_consts fc;
    int row0, row1;
    double *out;
    atomic_int next;
    atomic_ullong steps;
} render_job;

static void *render_worker(void *arg)
{
    render_job *j = (render_job *)arg;
    const int W = j->scn->width;
    unsigned long long steps = 0;
    for (;;) {
        int row = atomic_fetch_add(&j->next, 1);
        if (row >= j->row1) break;
        double *o = j->out + (size_t)(row - j->row0) * W * 3;
        if (!j->scn->supersampling) {
            for (int x = 0; x < W; x++) steps += trace_with(j->cam, j->scn, &j->fc, j->t, x, row, o + 3 * x);
        } else {
            /* src/ImageFilters.hs:88-97: 0.25 * (((p(2y,2x) + p(2y+1,2x)) + p(2y,2x+1)) + p(2y+1,2x+1)) */
            for (int x = 0; x < W; x++) {
                double a[3], b[3], c[3], d[3];
                steps += trace_with(j->cam, j->scn, &j->fc, j->t, 2 * x, 2 * row, a);
                steps += trace_with(j->cam, j->scn, &j->fc, j->t, 2 * x, 2 * row + 1, b);
                steps += trace_with(j->cam, j->scn, &j->fc, j->t, 2 * x + 1, 2 * row, c);
                steps += trace_with(j->cam, j->scn, &j->fc, j->t, 2 * x + 1, 2 * row + 1, d);
                for (int k = 0; k < 3; k++) o[3 * x + k] = 0.25 * (((a[k] + b[k]) + c[k]) + d[k]);
            }
        }
    }
    atomic_fetch_add(&j->steps, steps);
    return NULL;
}

int orc_render(const orc_camera *cam, const orc_scene *scn, const orc_tree *t, int row0, int row1,
               int nthreads, double *out_rgb, uint64_t *total_steps)
{
    if (row0 < 0 || row1 > scn->height || row0 > row1) return 1;
    render_job j;
    j.cam = cam; j.scn = scn; j.t = t; j.row0 = row0; j.row1 = row1; j.out = out_rgb;
    frame_setup(cam, scn, &j.fc);
    atomic_init(&j.next, row0);
    atomic_init(&j.steps, 0);
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t th[1024];
    int started = 0;
    for (int k = 1; k < nthreads; k++)
        if (pthread_create(&th[started], NULL, render_worker, &j) == 0) started++;
    render_worker(&j);
    for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
    if (total_steps) *total_steps = atomic_load(&j.steps);
    return 0;
}

/* ---------------------------------------------------------- ImageFilters.hs */

/* src/ImageFilters.hs:88-97 */
void orc_supersample(const double *in, int h, int w, double *out)
{
    int oh = h / 2, ow = w / 2;
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++)
            for (int k = 0; k < 3; k++) {
                double a = in[((size_t)(2 * y) * w + 2 * x) * 3 + k];
                double b = in[((size_t)(2 * y + 1) * w + 2 * x) * 3 + k];
                double c = in[((size_t)(2 * y) * w + 2 * x + 1) * 3 + k];
                double d = in[((size_t)(2 * y + 1) * w + 2 * x + 1) * 3 + k];
                out[((size_t)y * ow + x) * 3 + k] = 0.25 * (((a + b) + c) + d);
            }
}

/* one 1-D sweep of src/ImageFilters.hs:53-65 along a line of n pixels with the given stride:
 *   startVal = foldl1' add (map pix (take r crds))          -- sum of p(0..r-1), out-of-range = 0
 *   accumulate rgb x = (rgb + pix (x+r)) - pix (x-r); write (normFactor * new); carry new */
static void sweep(const double *src, double *dst, int n, size_t stride, int r, double norm)
{
    for (int k = 0; k < 3; k++) {
        double acc = 0;
        int first = 1;
        for (int x = 0; x < r; x++) { /* foldl1': first element then left fold */
            double p = x < n ? src[(size_t)x * stride + k] : 0.0;
            if (first) { acc = p; first = 0; } else acc = acc + p;
        }
        for (int x = 0; x < n; x++) {
            double pin = (x + r >= 0 && x + r < n) ? src[(size_t)(x + r) * stride + k] : 0.0;
            double pout = (x - r >= 0 && x - r < n) ? src[(size_t)(x - r) * stride + k] : 0.0;
            acc = (acc + pin) - pout;
            dst[(size_t)x * stride + k] = norm * acc;
        }
    }
}

/* src/ImageFilters.hs:28-78 boxBlur r passes (in place; each sweep reads a frozen copy :72,75) */
void orc_box_blur(int r, int passes, double *img, int h, int w)
{
    size_t n = (size_t)h * w * 3;
    double *tmp = (double *)malloc(n * sizeof(double));
    double norm = 1 / (2 * (double)r + 1);                /* :51 */
    for (int p = 0; p < passes; p++) {
        memcpy(tmp, img, n * sizeof(double));             /* :72 tmp1 <- freeze mv */
        for (int y = 0; y < h; y++)                       /* :73 horizontal */
            sweep(tmp + (size_t)y * w * 3, img + (size_t)y * w * 3, w, 3, r, norm);
        memcpy(tmp, img, n * sizeof(double));             /* :75 tmp2 <- freeze mv */
        for (int x = 0; x < w; x++)                       /* :76 vertical */
            sweep(tmp + (size_t)x * 3, img + (size_t)x * 3, h, (size_t)w * 3, r, norm);
    }
    free(tmp);
}

/* src/ImageFilters.hs:80-86 bloom: r = w `div` divider; out = img + strength * boxBlur r 3 img */
void orc_bloom(double strength, int divider, const double *in, int h, int w, double *out)
{
    size_t n = (size_t)h * w * 3;
    double *bl = (double *)malloc(n * sizeof(double));
    memcpy(bl, in, n * sizeof(double));
    orc_box_blur(w / divider, 3, bl, h, w);
    for (size_t k = 0; k < n; k++) out[k] = in[k] + strength * bl[k];
    free(bl);
}

/* src/Raytracer.hs:29-32 writeImg: A.map (toWord8 . fmap sRGB) */
void orc_to_srgb8(const double *in, size_t npix, uint8_t *out)
{
    for (size_t k = 0; k < npix * 3; k++) out[k] = orc_to_word8(orc_srgb(in[k]));
}
