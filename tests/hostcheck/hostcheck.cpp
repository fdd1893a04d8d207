// hostcheck.cpp -- TEST-ONLY harness: instantiates blackstar_b200/csrc/trace_core.cuh (the
// exact per-ray arithmetic the sm_100a kernels inline) for the HOST so that tests can diff
// it against the oracle in a container without a GPU.  Never linked into
// libblackstar_b200.so; the product has no CPU path.
#include "../../blackstar_b200/csrc/host_setup.hpp"
#include "../../blackstar_b200/csrc/trace_core.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace bsb;

struct HcCtx {
    HostStarTree tree;
    size_t n = 0;
};

extern "C" {

void *hc_create(const bsb_star *stars, size_t n, int /*unused*/)
{
    HcCtx *c = new HcCtx();
    c->n = n;
    if (n) build_star_tree(stars, n, c->tree);
    return c;
}

int hc_tree_top_levels(void *p) { return static_cast<HcCtx *>(p)->tree.top_levels; }

void hc_destroy(void *p) { delete static_cast<HcCtx *>(p); }

int hc_tree_depth(void *p) { return static_cast<HcCtx *>(p)->tree.depth; }

static void attach(HcCtx *c, FrameParams &P)
{
    P.tree.top = c->tree.top.data();
    P.tree.rec = c->tree.rec.data();
    P.tree.stars = c->tree.stars.data();
    P.tree.depth = c->tree.depth;
    P.tree.top_levels = c->tree.top_levels;
    P.tree.n_stars = (int)c->n;
    for (int g = 0; g < 4; g++) P.tree.rec_off[g] = c->tree.rec_off[g];
}

// Renders rows [row0,row1) of the final image exactly as the tiles kernel does per lane.
// block_steps > 0 advances in blocks (the refill kernel's schedule) instead of one call.
int hc_render(void *p, const bsb_camera *cam, const bsb_scene *scn, int row0, int row1, int block_steps,
              double *out_rgb, unsigned long long *steps_out, unsigned long long *hits_out)
{
    HcCtx *c = static_cast<HcCtx *>(p);
    FrameParams P;
    if (!make_frame_params(*cam, *scn, row0, row1, P).empty()) return 1;
    attach(c, P);
    unsigned long long steps = 0, hits = 0;
    for (int oy = row0; oy < row1; oy++)
        for (int ox = 0; ox < P.W; ox++) {
            double px[4][3];
            const int nsub = P.ss ? 4 : 1;
            for (int sub = 0; sub < nsub; sub++) {
                const int gx = P.ss ? 2 * ox + (sub >> 1) : ox;
                const int gy = P.ss ? 2 * oy + (sub & 1) : oy;
                RayState s;
                RayFrame F;
                ray_init(P, gx, gy, s, F);
                if (block_steps > 0) {
                    while (s.status == kAlive) ray_advance(P, s, (uint32_t)block_steps);
                } else {
                    ray_advance(P, s, 0xffffffffu);
                }
                hits += ray_finish(P, P.tree.top, F, s, px[sub]);
                steps += s.steps;
            }
            double *o = out_rgb + ((size_t)(oy - row0) * P.W + ox) * 3;
            for (int k = 0; k < 3; k++)
                o[k] = P.ss ? 0.25 * (((px[0][k] + px[1][k]) + px[2][k]) + px[3][k]) : px[0][k];
        }
    if (steps_out) *steps_out = steps;
    if (hits_out) *hits_out = hits;
    return 0;
}

void hc_star_lookup(void *p, double intensity, double saturation, const double vel[3], double rgb[3], unsigned *hits)
{
    HcCtx *c = static_cast<HcCtx *>(p);
    FrameParams P;
    std::memset(&P, 0, sizeof P);
    attach(c, P);
    P.star_intensity = intensity;
    P.star_saturation = saturation;
    *hits = star_lookup(P, P.tree.top, vel, rgb);
}

// one ray of the traced grid (gx, gy): colour, RK4 steps taken, final status
int hc_trace_ray(void *p, const bsb_camera *cam, const bsb_scene *scn, int gx, int gy, double rgb[3], unsigned *steps, int *status)
{
    HcCtx *c = static_cast<HcCtx *>(p);
    FrameParams P;
    if (!make_frame_params(*cam, *scn, 0, scn->height, P).empty()) return 1;
    attach(c, P);
    RayState s;
    RayFrame F;
    ray_init(P, gx, gy, s, F);
    ray_advance(P, s, 0xffffffffu);
    ray_finish(P, P.tree.top, F, s, rgb);
    *steps = s.steps;
    *status = s.status;
    return 0;
}

// StarMap.readMap + starColor' as the library parses a PPM catalogue: n stars, or -1 with the message in err
long hc_parse_ppm(const uint8_t *bytes, size_t len, bsb_star *out, size_t cap, char *err, size_t errcap)
{
    std::vector<bsb_star> v;
    std::string e;
    if (!parse_ppm(bytes, len, v, e)) {
        std::snprintf(err, errcap, "%s", e.c_str());
        return -1;
    }
    for (size_t k = 0; k < v.size() && k < cap; k++) out[k] = v[k];
    return (long)v.size();
}

// a star map file as --starmap names it (stars.kdt or PPM): n stars, or -1 with the message in err
long hc_parse_star_file(const uint8_t *bytes, size_t len, bsb_star *out, size_t cap, char *err, size_t errcap)
{
    std::vector<bsb_star> v;
    std::string e;
    if (!parse_star_file(bytes, len, v, e)) {
        std::snprintf(err, errcap, "%s", e.c_str());
        return -1;
    }
    for (size_t k = 0; k < v.size() && k < cap; k++) out[k] = v[k];
    return (long)v.size();
}

// wall-clock milliseconds of build_star_tree (what bsb_set_stars spends on the host before the upload)
double hc_build_tree_ms(const bsb_star *stars, size_t n, int reps)
{
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        HostStarTree t;
        const auto t0 = std::chrono::steady_clock::now();
        build_star_tree(stars, n, t);
        best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    return best;
}

const char *hc_validate_stars(const bsb_star *stars, size_t n)
{
    static thread_local std::string msg;
    msg = validate_stars(stars, n);
    return msg.c_str();
}

double hc_rinv5(double q) { return rinv5_seeded(q, rsqrt_seed(q), 1.4); }

}  // extern "C"
