/*
 * oracle_thirdparty.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see blackstar_oracle.h).
 *
 * Restatements of the third-party arithmetic the reference's hot path calls into.
 * None of these packages is under /root/reference; they are pinned only by
 * stack.yaml:1 (resolver lts-13.16): linear-1.20.8, massiv-io-0.1.6, kdt-0.2.4.
 * The formulas below restate their published algorithms.  PARITY UNPINNED: the
 * reference has no test vector at these boundaries; they are isolated here so a
 * GHC user can falsify them in one place.
 */
#include "blackstar_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- *
 * linear: Linear.Metric.normalize, Linear.Epsilon.nearZero (Double: |a| <= 1e-12)
 *   normalize v = if nearZero l || nearZero (1-l) then v else fmap (/sqrt l) v
 *     where l = quadrance v
 * call sites: src/Raytracer.hs:48, src/StarMap.hs:103
 * ------------------------------------------------------------------------- */
static int near_zero(double a) { return fabs(a) <= 1e-12; }

void orc_normalize(const double v[3], double out[3])
{
    double l = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (near_zero(l) || near_zero(1 - l)) {
        out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    } else {
        double s = sqrt(l);
        out[0] = v[0] / s; out[1] = v[1] / s; out[2] = v[2] / s;
    }
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    /* Linear.V3.cross (V3 a b c) (V3 d e f) = V3 (b*f-c*e) (c*d-a*f) (a*e-b*d) */
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* linear: Linear.Projection.lookAt eye center up; rows of its _m33 are xa, ya, -za with
 *   za = normalize (center - eye); xa = normalize (cross za up); ya = cross xa za
 * call site: src/Raytracer.hs:47 */
void orc_look_at_rows(const double eye[3], const double center[3], const double up[3],
                      double xa[3], double ya[3], double za[3])
{
    double d[3] = { center[0] - eye[0], center[1] - eye[1], center[2] - eye[2] };
    double c[3];
    orc_normalize(d, za);
    cross3(za, up, c);
    orc_normalize(c, xa);
    cross3(xa, za, ya);
}

/* ------------------------------------------------------------------------- *
 * massiv-io: Graphics.ColorSpace.HSI  instance ToRGB HSI
 *   toPixelRGB (PixelHSI h' s i) = getRGB (h' * 2 * pi)
 *     is = i*s; second = i - is
 *     getFirst a b = i + is * cos a / cos b
 *     getThird v1 v2 = i + 2*is + v1 - v2
 *     h < 2pi/3 : r = getFirst h (pi/3 - h);            b = second; g = getThird b r
 *     h < 4pi/3 : g = getFirst (h - 2pi/3) (h + pi);    r = second; b = getThird r g
 *     h < 2pi   : b = getFirst (h - 4pi/3) (2pi-pi/3-h); g = second; r = getThird g b
 *     h < 0 or h >= 2pi: error
 * call sites: src/Raytracer.hs:65, src/StarMap.hs:114
 * ------------------------------------------------------------------------- */
void orc_hsi_to_rgb(double hp, double s, double i, double rgb[3])
{
    const double pi = 3.141592653589793; /* Haskell's pi :: Double */
    double h = hp * 2 * pi;
    double is = i * s;
    double second = i - is;
    double r, g, b;
    if (h < 0) {
        r = g = b = NAN;
    } else if (h < 2 * pi / 3) {
        r = i + is * cos(h) / cos(pi / 3 - h);
        b = second;
        g = i + 2 * is + b - r;
    } else if (h < 4 * pi / 3) {
        g = i + is * cos(h - 2 * pi / 3) / cos(h + pi);
        r = second;
        b = i + 2 * is + r - g;
    } else if (h < 2 * pi) {
        b = i + is * cos(h - 4 * pi / 3) / cos(2 * pi - pi / 3 - h);
        g = second;
        r = i + 2 * is + g - b;
    } else {
        r = g = b = NAN;
    }
    rgb[0] = r; rgb[1] = g; rgb[2] = b;
}

/* massiv-io: Graphics.ColorSpace.Elevator, instance Elevator Double:
 *   toWord8 = round . (255 *) . clamp01 ; Haskell `round` is half-to-even.
 * call site: src/Raytracer.hs:32 */
uint8_t orc_to_word8(double x)
{
    double c = x < 0 ? 0 : (x > 1 ? 1 : x);
    if (x != x) c = 0;
    return (uint8_t)nearbyint(255 * c); /* default rounding mode = to nearest even */
}

/* ------------------------------------------------------------------------- *
 * kdt: Data.KdMap.Static.build / inRadius  (call sites src/StarMap.hs:91,104)
 *   build: sort by the current axis (axes cycle x,y,z with depth); the element at
 *          index n `div` 2 becomes the node, the ones before it the left subtree,
 *          the ones after it the right subtree.
 *   inRadius r q: at a node with axis value a:
 *          onTheLeft = q_axis <= a; recurse into the on-side child;
 *          recurse into the off-side child iff |q_axis - a| < r;
 *          keep the node's own point iff distSqr p q <= r*r (inclusive);
 *          list order: node, off-side results, on-side results.
 *   distSqr = defaultSqrDist = sum (zipWith (\a b -> (a-b)^2)) = ((0+dx^2)+dy^2)+dz^2
 * ------------------------------------------------------------------------- */
typedef struct {
    int32_t left, right; /* node indices or -1 */
    uint32_t star;       /* index into the caller's array */
    double p[3];
} kd_node;

struct orc_tree {
    kd_node *nodes;
    size_t n;
    int32_t root;
    const orc_star *stars; /* copy */
    orc_star *owned;
};

typedef struct { double key; uint32_t idx; } sort_item;

static int cmp_item(const void *a, const void *b)
{
    const sort_item *x = (const sort_item *)a, *y = (const sort_item *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    /* Haskell's sortBy is stable: ties keep input order */
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static int32_t build_rec(orc_tree *t, uint32_t *ids, size_t n, int axis, size_t *next, sort_item *scratch)
{
    if (n == 0) return -1;
    for (size_t k = 0; k < n; k++) {
        scratch[k].key = t->stars[ids[k]].pos[axis];
        scratch[k].idx = (uint32_t)k; /* position in the current list, for stability */
    }
    qsort(scratch, n, sizeof(sort_item), cmp_item);
    uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
    for (size_t k = 0; k < n; k++) tmp[k] = ids[scratch[k].idx];
    memcpy(ids, tmp, n * sizeof(uint32_t));
    free(tmp);
    size_t m = n / 2;
    int32_t me = (int32_t)(*next)++;
    kd_node *nd = &t->nodes[me];
    nd->star = ids[m];
    memcpy(nd->p, t->stars[ids[m]].pos, sizeof nd->p);
    int nax = (axis + 1) % 3;
    int32_t l = build_rec(t, ids, m, nax, next, scratch);
    int32_t r = build_rec(t, ids + m + 1, n - m - 1, nax, next, scratch);
    t->nodes[me].left = l;
    t->nodes[me].right = r;
    return me;
}

orc_tree *orc_tree_build(const orc_star *stars, size_t n)
{
    orc_tree *t = (orc_tree *)calloc(1, sizeof *t);
    t->n = n;
    t->root = -1;
    if (n == 0) return t;
    t->owned = (orc_star *)malloc(n * sizeof(orc_star));
    memcpy(t->owned, stars, n * sizeof(orc_star));
    t->stars = t->owned;
    t->nodes = (kd_node *)malloc(n * sizeof(kd_node));
    uint32_t *ids = (uint32_t *)malloc(n * sizeof(uint32_t));
    sort_item *scratch = (sort_item *)malloc(n * sizeof(sort_item));
    for (size_t k = 0; k < n; k++) ids[k] = (uint32_t)k;
    size_t next = 0;
    t->root = build_rec(t, ids, n, 0, &next, scratch);
    free(ids);
    free(scratch);
    return t;
}

void orc_tree_free(orc_tree *t)
{
    if (!t) return;
    free(t->nodes);
    free(t->owned);
    free(t);
}

size_t orc_tree_size(const orc_tree *t) { return t ? t->n : 0; }

static double dist_sqr(const double a[3], const double b[3])
{
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return ((0 + dx * dx) + dy * dy) + dz * dz;
}

static void in_radius_rec(const orc_tree *t, int32_t ni, int axis, double radius, const double q[3],
                          uint32_t *idx, size_t cap, size_t *cnt)
{
    if (ni < 0) return;
    const kd_node *nd = &t->nodes[ni];
    double qa = q[axis], a = nd->p[axis];
    int on_left = qa <= a;
    int nax = (axis + 1) % 3;
    if (dist_sqr(nd->p, q) <= radius * radius) {
        if (*cnt < cap) idx[*cnt] = nd->star;
        (*cnt)++;
    }
    if (fabs(qa - a) < radius)
        in_radius_rec(t, on_left ? nd->right : nd->left, nax, radius, q, idx, cap, cnt);
    in_radius_rec(t, on_left ? nd->left : nd->right, nax, radius, q, idx, cap, cnt);
}

size_t orc_in_radius(const orc_tree *t, double radius, const double q[3], uint32_t *idx, size_t cap)
{
    size_t cnt = 0;
    if (t && t->root >= 0) in_radius_rec(t, t->root, 0, radius, q, idx, cap, &cnt);
    return cnt;
}

/* internal accessor for blackstar_oracle.c */
const orc_star *orc_tree_star(const orc_tree *t, uint32_t i) { return &t->stars[i]; }
