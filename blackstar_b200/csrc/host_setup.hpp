// host_setup.hpp -- host-side preparation for the kernels: frame constants (the part of
// Raytracer.render that runs once per frame, src/Raytracer.hs:53-65), the bucketed k-d tree
// over the star list (replaces kdt's `build`, src/StarMap.hs:91), and the PPM catalogue
// reader (StarMap.readMap, src/StarMap.hs:45-58).
#pragma once

#include "../../include/blackstar_b200.h"
#include "bsb_common.cuh"

#include <string>
#include <vector>

namespace bsb {

// Fills every field of FrameParams except `tree` (device pointers).  Returns "" or an
// error message (invalid scene).
std::string make_frame_params(const bsb_camera &cam, const bsb_scene &scn, int row0, int row1,
                              FrameParams &P);

struct HostStarTree {
    std::vector<float> top;          // 2^top_levels - 1 entries (heap order), axis in the 2 low mantissa bits
    std::vector<double> rec;         // 8 doubles per 3-level subtree below the top
    std::vector<StarRec> stars;      // kLeafSlots records per leaf, padded
    uint32_t rec_off[4] = { 0, 0, 0, 0 };
    int depth = 0;
    int top_levels = 0;
};

// Median-split k-d tree (widest-extent axis) with 2^depth leaves of <= kLeafSlots stars.
void build_star_tree(const bsb_star *stars, size_t n, HostStarTree &out);

// StarMap.readMap + starColor' : PPM binary catalogue -> flat star list.
bool parse_ppm(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err);

// The reference's tree file (stars.kdt = cereal encoding of kdt's KdMap; layout recalled, every structural
// invariant checked) -> flat star list with starColor' applied; parse_star_file tries it, then the PPM layout.
bool parse_kdt(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err);
bool parse_star_file(const uint8_t *bytes, size_t len, std::vector<bsb_star> &out, std::string &err);

// "" if every star is a finite unit vector with a colour inside the HSI domain, else what is wrong
std::string validate_stars(const bsb_star *stars, size_t n);

// massiv-io HSI -> RGB on the host (disk colour, once per frame; src/Raytracer.hs:65)
void host_hsi_to_rgb(double h, double s, double i, double rgb[3]);
void hue_coefficients(double hue, double k[3]);

// writeImg's map (src/Raytracer.hs:23-32) as a table: thr[k], k = 1..255, is the smallest float x
// with toWord8(sRGB(x)) >= k (thr[0] = 0).  The map is monotone, so level(x) = #{k >= 1 : thr[k] <= x}.
// Found by bisection over float bit patterns on the reference's double arithmetic.
void srgb8_thresholds(float thr[256]);
int srgb8_level_host(float x);

}  // namespace bsb
