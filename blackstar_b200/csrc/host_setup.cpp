// host_setup.cpp -- see host_setup.hpp.  Compiled with -ffp-contract=off: the per-frame
// constants are formed with the reference's operation order so the device starts from
// bit-identical inputs.
#include "host_setup.hpp"

#include "trace_core.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <numeric>

namespace bsb {

namespace {

struct V3 { double x, y, z; };

inline V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline double quadrance(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
// Linear.V3.cross
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
// Linear.Metric.normalize (nearZero = |a| <= 1e-12)
inline V3 normalize(V3 v)
{
    const double l = quadrance(v);
    if (std::fabs(l) <= 1e-12 || std::fabs(1 - l) <= 1e-12) return v;
    const double s = std::sqrt(l);
    return { v.x / s, v.y / s, v.z / s };
}

}  // namespace

void host_hsi_to_rgb(double h, double s, double i, double rgb[3]) { hsi_to_rgb(h, s, i, rgb); }

// HSI -> RGB as channel = I * (1 + S * k_c) (massiv-io's formula, linear in S for a fixed hue)
void hue_coefficients(double hp, double k[3])
{
    const double pi = 3.141592653589793;
    const double h = hp * 2 * pi;
    if (h < 0 || !(h < 2 * pi)) { k[0] = k[1] = k[2] = NAN; return; }  // massiv-io: `error`
    if (h < 2 * pi / 3) {
        const double K = std::cos(h) / std::cos(pi / 3 - h);
        k[0] = K; k[2] = -1.0; k[1] = 1.0 - K;
    } else if (h < 4 * pi / 3) {
        const double K = std::cos(h - 2 * pi / 3) / std::cos(h + pi);
        k[1] = K; k[0] = -1.0; k[2] = 1.0 - K;
    } else {
        const double K = std::cos(h - 4 * pi / 3) / std::cos(2 * pi - pi / 3 - h);
        k[2] = K; k[1] = -1.0; k[0] = 1.0 - K;
    }
}

// src/Raytracer.hs:23-27 sRGB, then massiv-io toWord8 = round-half-even (255 * clamp01 x)
int srgb8_level_host(float xf)
{
    const double x = (double)xf;
    const double a = 0.055;
    const double s = x < 0.0031308 ? 12.92 * x : (1 + a) * std::pow(x, 1.0 / 2.4) - a;
    double c = s < 0 ? 0 : (s > 1 ? 1 : s);
    if (s != s) c = 0;
    return (int)std::nearbyint(255 * c);
}

void srgb8_thresholds(float thr[256])
{
    thr[0] = 0.0f;
    for (int k = 1; k < 256; k++) {
        // non-negative floats order like their bit patterns; level(0) = 0 < k <= 255 = level(1)
        uint32_t lo = 0u, hi = 0x3f800000u;  // level(lo) < k <= level(hi)
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float f;
            std::memcpy(&f, &mid, 4);
            if (srgb8_level_host(f) >= k) hi = mid; else lo = mid;
        }
        std::memcpy(&thr[k], &hi, 4);
    }
}

std::string make_frame_params(const bsb_camera &cam, const bsb_scene &scn, int row0, int row1, FrameParams &P)
{
    if (scn.width <= 0 || scn.height <= 0) return "resolution must be positive";
    if (scn.width > 32768 || scn.height > 32768) return "resolution above 32768 is not supported";
    if (row0 < 0 || row1 > scn.height || row0 > row1) return "row range outside the image";
    if (!(scn.step_size > 0) || !std::isfinite(scn.step_size)) return "stepSize must be a positive finite number";
    const double all[] = { cam.pos[0], cam.pos[1], cam.pos[2], cam.look_at[0], cam.look_at[1], cam.look_at[2],
                           cam.up[0], cam.up[1], cam.up[2], cam.fov, scn.disk_opacity, scn.disk_inner,
                           scn.disk_outer, scn.star_intensity, scn.star_saturation, scn.disk_hsi[0],
                           scn.disk_hsi[1], scn.disk_hsi[2] };
    for (double v : all)
        if (!std::isfinite(v)) return "non-finite camera/scene value";

    std::memset(&P, 0, sizeof P);
    const V3 eye = { cam.pos[0], cam.pos[1], cam.pos[2] };
    const V3 center = { cam.look_at[0], cam.look_at[1], cam.look_at[2] };
    const V3 up = { cam.up[0], cam.up[1], cam.up[2] };
    // Linear.Projection.lookAt (src/Raytracer.hs:47)
    const V3 za = normalize(sub(center, eye));
    const V3 xa = normalize(cross(za, up));
    const V3 ya = cross(xa, za);
    P.cam[0] = eye.x; P.cam[1] = eye.y; P.cam[2] = eye.z;
    P.xa[0] = xa.x; P.xa[1] = xa.y; P.xa[2] = xa.z;
    P.ya[0] = ya.x; P.ya[1] = ya.y; P.ya[2] = ya.z;
    P.za[0] = za.x; P.za[1] = za.y; P.za[2] = za.z;
    P.fov = cam.fov;
    P.q0 = quadrance(eye);
    P.r0 = std::sqrt(P.q0);
    if (P.r0 > 0) {
        P.e1[0] = eye.x / P.r0; P.e1[1] = eye.y / P.r0; P.e1[2] = eye.z / P.r0;
    } else {
        P.e1[0] = 1; P.e1[1] = 0; P.e1[2] = 0;  // camera at the singularity: every ray is black (r2 < 1)
    }
    const double h = scn.step_size;
    P.h = h;
    P.hh = h / 2;
    P.h6 = h / 6;
    P.hh2 = (h / 2) * (h / 2);
    P.hhh = h * (h / 2);
    P.hsq6 = h * h / 6;
    P.h3 = h / 3;
    P.k14 = 1.4;
    P.k13 = 1.0 / 3.0;
    P.k23 = 2.0 / 3.0;
    // src/Raytracer.hs:59-62
    const double twoq = 2 * P.q0;
    P.safe2 = 2500.0 > twoq ? 2500.0 : twoq;
    P.din2 = scn.disk_inner * scn.disk_inner;
    P.dout2 = scn.disk_outer * scn.disk_outer;
    P.r_in = std::sqrt(P.din2);   // diskColor' takes sqrt of the squared radii (:107-108)
    P.r_out = std::sqrt(P.dout2);
    hsi_to_rgb(scn.disk_hsi[0], scn.disk_hsi[1], scn.disk_hsi[2], P.disk_rgb);  // :65
    if (scn.disk_opacity != 0 && !(std::isfinite(P.disk_rgb[0])))
        return "diskColor hue outside [0, 360)";  // massiv-io raises `error` here
    P.disk_opacity = scn.disk_opacity;
    P.disk_on = scn.disk_opacity != 0 ? 1 : 0;
    P.star_intensity = scn.star_intensity;
    P.star_saturation = scn.star_saturation;
    P.ss = scn.supersampling ? 1 : 0;
    P.W = scn.width; P.H = scn.height;
    P.W2 = P.ss ? 2 * scn.width : scn.width;   // :58
    P.H2 = P.ss ? 2 * scn.height : scn.height;
    P.row0 = row0; P.row1 = row1;
    const int rows = row1 - row0;
    if (P.ss) {  // a warp = 4x2 output pixels = 8x4 rays
        P.tiles_x = (P.W + 3) / 4;
        P.tiles_y = (rows + 1) / 2;
    } else {     // a warp = 8x4 pixels
        P.tiles_x = (P.W + 7) / 8;
        P.tiles_y = (rows + 3) / 4;
    }
    P.n_tiles = P.tiles_x * P.tiles_y;
    P.step_cap = 1000000u;
    return "";
}

// ------------------------------------------------------------------ k-d tree
namespace {

struct Builder {
    const bsb_star *s;
    std::vector<uint32_t> idx;
    HostStarTree *t;
    int depth, T;

    void store_split(int d, size_t i, double sv, int axis)
    {
        if (d < T) {
            float f = (float)sv;
            uint32_t bits;
            std::memcpy(&bits, &f, 4);
            bits = (bits & ~uint32_t(3)) | uint32_t(axis);
            std::memcpy(&t->top[(size_t(1) << d) - 1 + i], &bits, 4);
        } else {
            const int g = (d - T) / 3, l = (d - T) - 3 * g;
            const size_t blk = i >> l, local = (size_t(1) << l) + (i & ((size_t(1) << l) - 1));
            uint64_t bits;
            std::memcpy(&bits, &sv, 8);
            bits = (bits & ~uint64_t(3)) | uint64_t(axis);
            std::memcpy(&t->rec[(size_t(t->rec_off[g]) + blk) * 8 + local], &bits, 8);
        }
    }

    void rec(int d, size_t i, size_t lo, size_t hi)
    {
        if (d == depth) {
            StarRec *slot = &t->stars[i * kLeafSlots];
            for (size_t k = lo; k < hi; k++) {
                const bsb_star &st = s[idx[k]];
                double k3[3];
                hue_coefficients(st.hue, k3);
                slot[k - lo] = StarRec{ st.pos[0], st.pos[1], st.pos[2], st.sat * k3[0], st.sat * k3[1], st.sat * k3[2], st.mag,
                                        (float)st.pos[0], (float)st.pos[1], (float)st.pos[2] };
            }
            return;
        }
        int axis = d < T ? d % 3 : 0;
        double sv = 0.0;
        size_t mid = lo;
        if (hi > lo) {
            double mn[3] = { 2, 2, 2 }, mx[3] = { -2, -2, -2 };
            for (size_t k = lo; k < hi; k++)
                for (int a = 0; a < 3; a++) {
                    const double c = s[idx[k]].pos[a];
                    mn[a] = std::min(mn[a], c);
                    mx[a] = std::max(mx[a], c);
                }
            const double ex[3] = { mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2] };
            axis = (ex[1] > ex[0]) ? 1 : 0;
            if (ex[2] > ex[axis]) axis = 2;
            if (d < T) axis = d % 3;   // shared-memory levels: the axis is a function of the level (see below)
            mid = lo + (hi - lo) / 2;
            const bsb_star *sp = s;
            const int ax = axis;
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                             [sp, ax](uint32_t a, uint32_t b) { return sp[a].pos[ax] < sp[b].pos[ax]; });
            sv = s[idx[mid]].pos[axis];
        }
        store_split(d, i, sv, axis);
        if (d < 3 && hi - lo > 40000) {
            // the two subtrees touch disjoint parts of every array: build the left one on another thread
            // (8 threads at depth 3; the full 468 861-star catalogue takes ~50 ms instead of ~240)
            std::thread left([this, d, i, lo, mid] { rec(d + 1, 2 * i, lo, mid); });
            rec(d + 1, 2 * i + 1, mid, hi);
            left.join();
        } else {
            rec(d + 1, 2 * i, lo, mid);
            rec(d + 1, 2 * i + 1, mid, hi);
        }
    }
};

}  // namespace

void build_star_tree(const bsb_star *stars, size_t n, HostStarTree &out)
{
    int depth = 0;
    while ((size_t(kLeafSlots) << depth) < n && depth < 25) depth++;
    // top levels (shared memory) + groups of three levels below them
    int T = depth;
    if (depth > kSmemTreeLevels) {
        const int groups = (depth - kSmemTreeLevels + 2) / 3;
        T = depth - 3 * groups;
    }
    out.depth = depth;
    out.top_levels = T;
    out.top.assign((size_t(1) << T) - 1 + 1, 0.0f);  // +1: never zero-sized
    size_t n_rec = 0;
    for (int g = 0; T + 3 * g < depth; g++) {
        out.rec_off[g] = (uint32_t)n_rec;
        n_rec += size_t(1) << (T + 3 * g);
    }
    out.rec.assign(n_rec * 8 + 8, 0.0);
    // padding record: far outside the unit sphere, never within the lookup radius
    out.stars.assign((size_t(1) << depth) * kLeafSlots, StarRec{ 4.0, 4.0, 4.0, 0.0, 0.0, 0.0, 0, 4.0f, 4.0f, 4.0f });
    Builder b{ stars, std::vector<uint32_t>(n), &out, depth, T };
    std::iota(b.idx.begin(), b.idx.end(), 0u);
    b.rec(0, 0, 0, n);
}

// ------------------------------------------------------------------ PPM catalogue
namespace {

double f64be(const uint8_t *p)
{
    uint64_t u = 0;
    for (int k = 0; k < 8; k++) u = (u << 8) | p[k];
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}

// starColor (src/StarMap.hs:63-72)
void spectral_colour(int ch, double &hue, double &sat)
{
    switch (ch) {
    case 'O': hue = 0.631; sat = 0.39; break;
    case 'B': hue = 0.628; sat = 0.33; break;
    case 'A': hue = 0.622; sat = 0.21; break;
    case 'F': hue = 0.650; sat = 0.03; break;
    case 'G': hue = 0.089; sat = 0.09; break;
    case 'K': hue = 0.094; sat = 0.29; break;
    case 'M': hue = 0.094; sat = 0.56; break;
    default: hue = 0; sat = 0; break;
    }
}

}  // namespace

bool parse_ppm(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err)
{
    if (len < 28) { err = "too few bytes"; return false; }  // cereal's message for a failed `skip 28`
    // readMap itself silently drops a ragged tail (:48-49 remaining `div` 28); a catalogue written in
    // this layout never has one, and a file in another format (e.g. the reference's cereal-encoded
    // stars.kdt) almost always does -- refuse it instead of rendering noise
    if ((len - 28) % 28 != 0) {
        err = "not a PPM-format star catalogue (28-byte header + 28-byte records); note that this library reads "
              "the catalogue itself, not the stars.kdt tree file generate-tree writes";
        return false;
    }
    const size_t n = (len - 28) / 28;
    out.resize(n);
    const uint8_t *p = bytes + 28;
    for (size_t k = 0; k < n; k++, p += 28) {
        const double ra = f64be(p), dec = f64be(p + 8);      // :50-51
        const int spectral = p[16];                          // :52, then skip 1
        const int16_t mag = (int16_t)(uint16_t(p[18]) << 8 | p[19]);  // :54, then skip 8
        // right ascension / declination are angles: anything non-finite or far outside [-2pi, 2pi] is
        // not a catalogue record (random bytes decode to 1e+-300 or NaN)
        if (!std::isfinite(ra) || !std::isfinite(dec) || std::fabs(ra) > 7.0 || std::fabs(dec) > 7.0) {
            char buf[160];
            std::snprintf(buf, sizeof buf, "record %zu is not a star (RA %g, Dec %g): not a PPM-format catalogue", k, ra, dec);
            err = buf;
            out.clear();
            return false;
        }
        bsb_star &s = out[k];
        s.pos[0] = std::cos(dec) * std::cos(ra);             // :74-75 raDecToCartesian
        s.pos[1] = std::cos(dec) * std::sin(ra);
        s.pos[2] = std::sin(dec);
        s.mag = mag;
        s.pad_ = 0;
        spectral_colour(spectral, s.hue, s.sat);
    }
    return true;
}

// ------------------------------------------------------------------ stars.kdt (the reference's tree file)
// What `generate-tree` writes (app/GenerateTree.hs:11-29): Data.Serialize.encode of a
//     KdMap Double (V3 Double) (Int, Char)                                     (src/StarMap.hs:28-41, 87-88)
// through cereal's Generic instances.  NEITHER cereal NOR kdt is under /root/reference and there is no GHC in
// this image, so the byte layout below is RECALLED from the packages' sources (cereal-0.5.8, kdt-0.2.4), not
// verified -- which is why the reader checks everything the format lets it check and refuses the file
// otherwise, rather than rendering a wrong sky:
//   KdMap    = pointAsList fn (Word8 0, :33-35) ++ distSqr fn (Word8 0, :38-40) ++ root TreeNode ++ size (Int = Int64 BE)
//   TreeNode = Word8 tag: 0 -> left TreeNode ++ point (3 x Float64 BE) ++ (mag Int64 BE, spectral Char as UTF-8)
//                              ++ axisValue (Float64 BE) ++ right TreeNode
//                         1 -> Empty
// Checks: tags are 0/1; every point is a finite unit vector; axisValue is, bit for bit, the point's
// coordinate on axis (depth mod 3) (kdt cycles the axes and stores the node's own coordinate); left/right
// points respect the split; the node count equals `size`; the input is consumed exactly.
namespace {

struct KdtReader {
    const uint8_t *p, *end;
    std::vector<bsb_star> *out;
    std::string err;
    size_t nodes = 0;

    bool need(size_t n) { if ((size_t)(end - p) < n) { err = "truncated"; return false; } return true; }
    bool f64(double &d) { if (!need(8)) return false; d = f64be(p); p += 8; return true; }
    bool i64(int64_t &v) { if (!need(8)) return false; uint64_t u = 0; for (int k = 0; k < 8; k++) u = (u << 8) | p[k]; p += 8; v = (int64_t)u; return true; }
    bool chr(int &c)
    {
        if (!need(1)) return false;
        const int b = *p++;
        if (b < 0x80) { c = b; return true; }
        const int extra = b >= 0xf0 ? 3 : (b >= 0xe0 ? 2 : 1);
        if (!need((size_t)extra)) return false;
        c = b & (0x3f >> extra);
        for (int k = 0; k < extra; k++) c = (c << 6) | (*p++ & 0x3f);
        return true;
    }
    // lo/hi: bounds the ancestors' splits put on each coordinate
    bool node(int depth, double lo[3], double hi[3])
    {
        if (depth > 200) { err = "tree deeper than 200 levels"; return false; }
        if (!need(1)) return false;
        const int tag = *p++;
        if (tag == 1) return true;
        if (tag != 0) { err = "constructor tag is neither TreeNode (0) nor Empty (1)"; return false; }
        // the left subtree comes first, but its bound is this node's axis value, which comes after it:
        // remember where the left subtree starts, skip ahead structurally by parsing it with open bounds,
        // then check it against the split afterwards via the recorded extent
        const size_t first = out->size();
        double l_lo[3] = { lo[0], lo[1], lo[2] }, l_hi[3] = { hi[0], hi[1], hi[2] };
        if (!node(depth + 1, l_lo, l_hi)) return false;
        const size_t mid = out->size();
        bsb_star s;
        int64_t mag;
        int ch;
        double axis_value;
        if (!f64(s.pos[0]) || !f64(s.pos[1]) || !f64(s.pos[2]) || !i64(mag) || !chr(ch) || !f64(axis_value)) return false;
        const double q = s.pos[0] * s.pos[0] + s.pos[1] * s.pos[1] + s.pos[2] * s.pos[2];
        if (!std::isfinite(q) || std::fabs(q - 1.0) > 1e-9) { err = "a point is not a unit vector"; return false; }
        const int ax = depth % 3;
        if (std::memcmp(&axis_value, &s.pos[ax], 8) != 0) { err = "axisValue is not the point's coordinate on axis (depth mod 3)"; return false; }
        if (mag < -32768 || mag > 32767) { err = "magnitude outside the catalogue's int16 range"; return false; }
        for (int a = 0; a < 3; a++)
            if (s.pos[a] < lo[a] || s.pos[a] > hi[a]) { err = "a point lies on the wrong side of an ancestor's split"; return false; }
        for (size_t k = first; k < mid; k++)
            if ((*out)[k].pos[ax] > axis_value) { err = "a left-subtree point lies right of its node's split"; return false; }
        s.mag = (int32_t)mag;
        s.pad_ = 0;
        spectral_colour(ch, s.hue, s.sat);       // starColor' as readTreeFromFile applies it (:85)
        out->push_back(s);
        nodes++;
        double r_lo[3] = { lo[0], lo[1], lo[2] }, r_hi[3] = { hi[0], hi[1], hi[2] };
        r_lo[ax] = axis_value;
        return node(depth + 1, r_lo, r_hi);
    }
};

}  // namespace

bool parse_kdt(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err)
{
    out.clear();
    KdtReader r{ bytes, bytes + len, &out, "" };
    if (len < 11 || bytes[0] != 0 || bytes[1] != 0) { err = "not a stars.kdt tree file (it starts with the two placeholder bytes 00 00)"; return false; }
    r.p += 2;
    double lo[3] = { -2, -2, -2 }, hi[3] = { 2, 2, 2 };
    int64_t size = 0;
    if (!r.node(0, lo, hi) || !r.i64(size)) { err = "stars.kdt: " + (r.err.empty() ? std::string("malformed") : r.err); out.clear(); return false; }
    if ((uint64_t)size != r.nodes) { err = "stars.kdt: the stored size does not match the number of tree nodes"; out.clear(); return false; }
    if (r.p != r.end) { err = "stars.kdt: bytes left over after the tree"; out.clear(); return false; }
    return true;
}

// A star map file as --starmap names it: the reference's tree file (stars.kdt) or the PPM catalogue it was made from.
bool parse_star_file(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err)
{
    std::string e1, e2;
    if (parse_kdt(bytes, len, out, e1)) return true;
    if (parse_ppm(bytes, len, out, e2)) return true;
    err = "neither a stars.kdt tree file (" + e1 + ") nor a PPM catalogue (" + e2 + ")";
    return false;
}

// The star list handed to bsb_set_stars: unit vectors, finite colours, hue in [0, 1) (massiv-io's
// toPixelRGB raises `error` outside it).  NaNs would also break the strict weak ordering nth_element needs.
std::string validate_stars(const bsb_star *stars, size_t n)
{
    for (size_t k = 0; k < n; k++) {
        const bsb_star &s = stars[k];
        const double q = s.pos[0] * s.pos[0] + s.pos[1] * s.pos[1] + s.pos[2] * s.pos[2];
        char buf[160];
        if (!std::isfinite(q) || std::fabs(q - 1.0) > 1e-6) {
            std::snprintf(buf, sizeof buf, "star %zu: position is not a unit vector (|p|^2 = %g)", k, q);
            return buf;
        }
        if (!std::isfinite(s.hue) || !std::isfinite(s.sat) || s.hue < 0.0 || !(s.hue < 1.0)) {
            std::snprintf(buf, sizeof buf, "star %zu: hue %g / saturation %g outside the HSI domain (hue in [0,1))", k, s.hue, s.sat);
            return buf;
        }
    }
    return "";
}

}  // namespace bsb
