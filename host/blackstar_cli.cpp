// blackstar_cli.cpp -- host side above the C ABI, in C++: a mirror of the reference's
// `blackstar` executable (app/Main.hs) with the same flags, batch-directory behaviour, preview
// override, file naming and messages.  Scene loading mirrors src/ConfigFile.hs:40-79.
// The hot path (render + supersample + bloom + sRGB/8-bit) runs on the GPU(s) through
// libblackstar_b200.so; there is no CPU rendering path in this program.
//
//   blackstar [-p|--preview] [-o|--output PATH] [-f|--force] [-s|--starmap PATH] INPUTFILE
//
// The star map is a PPM-format binary catalogue (the input of `generate-tree`,
// src/StarMap.hs:45-58), not a cereal-encoded stars.kdt.
// Test hooks (no GPU needed): --dump-config FILE prints the parsed scene as JSON;
// --selftest-png FILE writes a test pattern.
#include "../include/blackstar_b200.h"
#include "png_writer.hpp"
#include "yaml_lite.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Options {
    bool preview = false;
    bool force = false;
    std::string output;
    std::string starmap = "stars.kdt";  // app/Main.hs:36
    std::string inputfile;
};

struct Config {
    bsb_camera cam;
    bsb_scene scn;
};

std::string read_file(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(path + ": openBinaryFile: does not exist (No such file or directory)");
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

void vec3(const yamllite::Node *n, const char *what, double out[3])
{
    if (!n || n->kind != yamllite::Node::Seq || n->seq.size() != 3)
        throw std::runtime_error(std::string(what) + ": expected [x, y, z]");
    for (int k = 0; k < 3; k++) out[k] = n->seq[k]->as_double(what);
}

// FromJSON Scene (src/ConfigFile.hs:66-79) with its defaults; unknown keys are ignored
void scene_from_yaml(const yamllite::Node *v, bsb_scene &s)
{
    if (!v || v->kind != yamllite::Node::Map) throw std::runtime_error("scene: expected Object");
    s.step_size = 0.3; s.bloom_strength = 0.4; s.bloom_divider = 25;
    s.star_intensity = 0.7; s.star_saturation = 0.7;
    s.disk_hsi[0] = 0.16; s.disk_hsi[1] = 0.1; s.disk_hsi[2] = 0.95;
    s.disk_opacity = 0; s.disk_inner = 3; s.disk_outer = 12;
    s.width = 1280; s.height = 720; s.supersampling = 0;
    auto opt = [&](const char *k) -> const yamllite::Node * {
        const yamllite::Node *n = v->get(k);
        return (n && !n->is_null()) ? n : nullptr;
    };
    if (auto n = opt("stepSize")) s.step_size = n->as_double("stepSize");
    if (auto n = opt("bloomStrength")) s.bloom_strength = n->as_double("bloomStrength");
    if (auto n = opt("bloomDivider")) s.bloom_divider = (int32_t)n->as_int("bloomDivider");
    if (auto n = opt("starIntensity")) s.star_intensity = n->as_double("starIntensity");
    if (auto n = opt("starSaturation")) s.star_saturation = n->as_double("starSaturation");
    if (auto n = opt("diskColor")) {
        vec3(n, "diskColor", s.disk_hsi);
        s.disk_hsi[0] = s.disk_hsi[0] / 360;  // src/ConfigFile.hs:51
    }
    if (auto n = opt("diskOpacity")) s.disk_opacity = n->as_double("diskOpacity");
    if (auto n = opt("diskInner")) s.disk_inner = n->as_double("diskInner");
    if (auto n = opt("diskOuter")) s.disk_outer = n->as_double("diskOuter");
    if (auto n = opt("resolution")) {
        if (n->kind != yamllite::Node::Seq || n->seq.size() != 2) throw std::runtime_error("resolution: expected [w, h]");
        s.width = (int32_t)n->seq[0]->as_int("resolution");
        s.height = (int32_t)n->seq[1]->as_int("resolution");
    }
    if (auto n = opt("supersampling")) s.supersampling = n->as_bool("supersampling") ? 1 : 0;
}

// Generic FromJSON Camera (src/ConfigFile.hs:61): all four fields are mandatory
void camera_from_yaml(const yamllite::Node *v, bsb_camera &c)
{
    if (!v || v->kind != yamllite::Node::Map) throw std::runtime_error("camera: expected Object");
    for (const char *k : { "position", "lookAt", "upVec", "fov" })
        if (!v->get(k)) throw std::runtime_error(std::string("camera: key \"") + k + "\" not present");
    vec3(v->get("position"), "position", c.pos);
    vec3(v->get("lookAt"), "lookAt", c.look_at);
    vec3(v->get("upVec"), "upVec", c.up);
    c.fov = v->get("fov")->as_double("fov");
}

Config load_config(const std::string &path)
{
    const yamllite::NodeP root = yamllite::parse(read_file(path));
    if (root->kind != yamllite::Node::Map || !root->get("scene") || !root->get("camera"))
        throw std::runtime_error("config needs 'scene' and 'camera'");
    Config c;
    std::memset(&c, 0, sizeof c);
    scene_from_yaml(root->get("scene"), c.scn);
    camera_from_yaml(root->get("camera"), c.cam);
    return c;
}

// app/Main.hs:93-103
void prepare_scene(Config &c, bool preview)
{
    if (!preview) return;
    const int w = c.scn.width, h = c.scn.height, res = 300;
    if (w >= h) { c.scn.width = res; c.scn.height = res * h / w; }
    else { c.scn.width = res * w / h; c.scn.height = res; }
    c.scn.supersampling = 0;
    c.scn.bloom_strength = 0;
}

void dump_config(const Config &c)
{
    std::printf("{\"camera\": {\"position\": [%.17g, %.17g, %.17g], \"lookAt\": [%.17g, %.17g, %.17g], "
                "\"upVec\": [%.17g, %.17g, %.17g], \"fov\": %.17g}, ",
                c.cam.pos[0], c.cam.pos[1], c.cam.pos[2], c.cam.look_at[0], c.cam.look_at[1], c.cam.look_at[2],
                c.cam.up[0], c.cam.up[1], c.cam.up[2], c.cam.fov);
    std::printf("\"scene\": {\"stepSize\": %.17g, \"bloomStrength\": %.17g, \"bloomDivider\": %d, \"starIntensity\": %.17g, "
                "\"starSaturation\": %.17g, \"diskColor\": [%.17g, %.17g, %.17g], \"diskOpacity\": %.17g, "
                "\"diskInner\": %.17g, \"diskOuter\": %.17g, \"resolution\": [%d, %d], \"supersampling\": %s}}\n",
                c.scn.step_size, c.scn.bloom_strength, c.scn.bloom_divider, c.scn.star_intensity, c.scn.star_saturation,
                c.scn.disk_hsi[0], c.scn.disk_hsi[1], c.scn.disk_hsi[2], c.scn.disk_opacity, c.scn.disk_inner,
                c.scn.disk_outer, c.scn.width, c.scn.height, c.scn.supersampling ? "true" : "false");
}

bool is_dir(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
bool exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

void mkdirs(const std::string &p)
{
    std::string cur;
    for (size_t i = 0; i <= p.size(); i++) {
        if (i == p.size() || p[i] == '/') {
            if (!cur.empty() && !exists(cur)) mkdir(cur.c_str(), 0777);
        }
        if (i < p.size()) cur.push_back(p[i]);
    }
}

std::string base_name(const std::string &path)  // takeBaseName
{
    const size_t s = path.find_last_of('/');
    std::string f = s == std::string::npos ? path : path.substr(s + 1);
    const size_t d = f.find_last_of('.');
    return d == std::string::npos ? f : f.substr(0, d);
}

std::string extension(const std::string &path)
{
    const size_t s = path.find_last_of('/');
    const std::string f = s == std::string::npos ? path : path.substr(s + 1);
    const size_t d = f.find_last_of('.');
    return d == std::string::npos ? "" : f.substr(d);
}

// Util.promptOverwriteFile (src/Util.hs:18-27)
bool prompt_overwrite(const std::string &path)
{
    if (!exists(path)) return true;
    std::cout << "Overwrite " << path << "? [y/N] " << std::flush;
    std::string ans;
    std::getline(std::cin, ans);
    return !ans.empty() && (ans[0] == 'y' || ans[0] == 'Y');
}

// Batch mode (app/Main.hs:64-77) renders the scenes strictly one after the other; here the PNG of
// scene k is deflated and written by a background thread while the GPU traces scene k+1 (row N4 of
// SURVEY.md section 8f).  Only with --force: the overwrite prompt needs the terminal.
class PngWriter {
public:
    explicit PngWriter(bool background) : background_(background)
    {
        if (background_) th_ = std::thread([this] { run(); });
    }
    ~PngWriter() { finish(); }
    void submit(std::string path, std::vector<uint8_t> rgb, int w, int h)
    {
        if (!background_) { write(path, rgb, w, h); return; }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return q_.size() < 2; });   // at most two frames in flight
        q_.push_back(Job{ std::move(path), std::move(rgb), w, h });
        cv_.notify_all();
    }
    void finish()
    {
        if (!background_ || !th_.joinable()) return;
        { std::lock_guard<std::mutex> lk(m_); done_ = true; }
        cv_.notify_all();
        th_.join();
    }

private:
    struct Job { std::string path; std::vector<uint8_t> rgb; int w, h; };
    static void write(const std::string &path, const std::vector<uint8_t> &rgb, int w, int h)
    {
        // zlib level and deflate threads of the PNG encoder: the defaults reproduce what the reference's encoder
        // spends (level 6); a batch that is bound by the encoder (600-frame animations: profiles/) can trade
        // file size for speed with BSB_PNG_LEVEL=1
        static const int level = [] { const char *e = std::getenv("BSB_PNG_LEVEL"); const int v = e ? std::atoi(e) : 6; return v < 0 || v > 9 ? 6 : v; }();
        static const int threads = [] { const char *e = std::getenv("BSB_PNG_THREADS"); return e ? std::max(0, std::atoi(e)) : 0; }();
        const std::string err = pngw::write_rgb8_parallel(path, rgb.data(), w, h, threads, level);
        if (!err.empty()) std::cout << err << std::endl;
    }
    void run()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return done_ || !q_.empty(); });
                if (q_.empty()) return;
                j = std::move(q_.front());
                q_.pop_front();
                cv_.notify_all();
            }
            write(j.path, j.rgb, j.w, j.h);
        }
    }
    bool background_, done_ = false;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job> q_;
};

// Main.handleScene + doRender (app/Main.hs:80-125)
void handle_scene(bsb_ctx *ctx, const Options &o, const std::string &outdir, const std::string &filename, PngWriter &writer)
{
    std::string name = base_name(filename);
    std::cout << "Reading " << filename << "..." << std::endl;
    Config cfg;
    try {
        cfg = load_config(filename);
    } catch (const std::exception &e) {  // prettyPrintParseException, then carry on (app/Main.hs:91)
        std::cout << e.what() << std::endl;
        return;
    }
    std::cout << "Scene successfully read." << std::endl;
    if (o.preview) name = "prev-" + name;
    prepare_scene(cfg, o.preview);
    std::cout << "Rendering " << name << "..." << std::endl;
    if (cfg.scn.width <= 0 || cfg.scn.height <= 0) { std::cout << "resolution must be positive" << std::endl; return; }
    std::vector<uint8_t> rgb((size_t)cfg.scn.width * cfg.scn.height * 3);
    bsb_stats st;
    const auto t0 = std::chrono::steady_clock::now();
    if (cfg.scn.bloom_strength != 0) std::cout << "Applying bloom..." << std::endl;
    const int rc = bsb_render_full_srgb8(ctx, &cfg.cam, &cfg.scn, rgb.data(), &st);
    if (rc != BSB_OK) {
        std::cout << "Rendering failed: " << bsb_last_error(ctx) << std::endl;
        return;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Rendering completed in %.3f seconds (%llu rays, %.1f Mrays/s in the trace kernel on %d GPU(s)).\n", secs,
                (unsigned long long)st.rays, st.trace_ms > 0 ? st.rays / st.trace_ms / 1e3 : 0.0, st.n_gpus);
    const std::string out_name = outdir + "/" + name + ".png";
    std::cout << "Saving to " << out_name << "..." << std::endl;
    if (o.force || prompt_overwrite(out_name)) writer.submit(out_name, std::move(rgb), cfg.scn.width, cfg.scn.height);
    std::cout << "Everything done. Thank you!" << std::endl;
}

int usage(int rc)
{
    std::cout << "Blackstar v0.1 (B200)\n\nblackstar [OPTIONS] INPUTFILE\n\nCommon flags:\n"
                 "  -p --preview         preview render (small size)\n"
                 "  -o --output=PATH     output directory\n"
                 "  -f --force           overwrite images without asking\n"
                 "  -s --starmap=PATH    path to starmap (stars.kdt tree file or PPM-format catalogue)\n"
                 "  -? --help            Display help message\n";
    return rc;
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&](const std::string &flag) -> std::string {
            const size_t eq = a.find('=');
            if (eq != std::string::npos) return a.substr(eq + 1);
            if (i + 1 >= argc) { std::cerr << "Missing value for " << flag << "\n"; std::exit(1); }
            return argv[++i];
        };
        if (a == "-p" || a == "--preview") o.preview = true;
        else if (a == "-f" || a == "--force") o.force = true;
        else if (a == "-o" || a.rfind("--output", 0) == 0) o.output = value("--output");
        else if (a == "-s" || a.rfind("--starmap", 0) == 0) o.starmap = value("--starmap");
        else if (a == "-?" || a == "--help" || a == "-h") return usage(0);
        else if (a == "--dump-config") {
            if (i + 1 >= argc) return usage(1);
            try {
                Config c = load_config(argv[++i]);
                prepare_scene(c, o.preview);
                dump_config(c);
                return 0;
            } catch (const std::exception &e) { std::cout << e.what() << std::endl; return 2; }
        } else if (a == "--selftest-png") {
            if (i + 1 >= argc) return usage(1);
            const int w = 67, h = 31;
            std::vector<uint8_t> px((size_t)w * h * 3);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    px[(y * w + x) * 3 + 0] = uint8_t(x * 255 / (w - 1));
                    px[(y * w + x) * 3 + 1] = uint8_t(y * 255 / (h - 1));
                    px[(y * w + x) * 3 + 2] = uint8_t((x * 7 + y * 13) & 255);
                }
            const std::string err = pngw::write_rgb8(argv[++i], px.data(), w, h);
            if (!err.empty()) { std::cout << err << std::endl; return 2; }
            return 0;
        } else if (a == "--selftest-png-parallel") {
            // a frame large enough for several deflate bands; pixels follow a fixed formula
            if (i + 2 >= argc) return usage(1);
            const std::string path = argv[++i];
            const int threads = std::atoi(argv[++i]);
            const int w = 1031, h = 517;
            std::vector<uint8_t> px((size_t)w * h * 3);
            uint32_t lcg = 12345u;
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    lcg = lcg * 1664525u + 1013904223u;
                    px[((size_t)y * w + x) * 3 + 0] = uint8_t((x * 3 + y) & 255);
                    px[((size_t)y * w + x) * 3 + 1] = uint8_t((lcg >> 24) & 255);
                    px[((size_t)y * w + x) * 3 + 2] = uint8_t(((x ^ y) * 5) & 255);
                }
            const std::string err = pngw::write_rgb8_parallel(path, px.data(), w, h, threads);
            if (!err.empty()) { std::cout << err << std::endl; return 2; }
            return 0;
        } else if (!a.empty() && a[0] == '-') { std::cerr << "Unknown flag: " << a << "\n"; return usage(1); }
        else pos.push_back(a);
    }
    if (pos.size() != 1) { std::cerr << "Requires at least 1 arguments, got " << pos.size() << "\n"; return usage(1); }
    o.inputfile = pos[0];

    // app/Main.hs:43-50: read the star map first; refuse to start without it
    std::string catalogue;
    try {
        catalogue = read_file(o.starmap);
    } catch (const std::exception &e) {
        std::cout << "Error decoding star tree: \n" << e.what() << std::endl;
        return 1;
    }
    bsb_ctx *ctx = bsb_create(0);
    if (!ctx) { std::cout << bsb_last_error(nullptr) << std::endl; return 1; }
    if (bsb_set_stars_file(ctx, reinterpret_cast<const uint8_t *>(catalogue.data()), catalogue.size()) != BSB_OK) {
        std::cout << "Error decoding star tree: \n" << bsb_last_error(ctx) << std::endl;
        bsb_destroy(ctx);
        return 1;
    }
    std::cout << "Starmap successfully read." << std::endl;

    // app/Main.hs:52-78 doStart
    char cwd[4096];
    std::string outdir = o.output.empty() ? std::string(getcwd(cwd, sizeof cwd) ? cwd : ".") : o.output;
    mkdirs(outdir);
    const std::string filename = o.inputfile;
    if (is_dir(filename)) {
        std::cout << filename << " is a directory. Rendering all scenes inside it..." << std::endl;
        std::vector<std::string> files;
        if (DIR *d = opendir(filename.c_str())) {
            while (dirent *e = readdir(d))
                if (extension(e->d_name) == ".yaml") files.push_back(e->d_name);
            closedir(d);
        }
        std::sort(files.begin(), files.end());
        PngWriter writer(o.force);
        for (size_t k = 0; k < files.size(); k++) {
            std::cout << "Batch mode progress: " << (k + 1) << "/" << files.size() << std::endl;
            handle_scene(ctx, o, outdir, filename + "/" + files[k], writer);
        }
        writer.finish();
    } else {
        PngWriter writer(false);
        handle_scene(ctx, o, outdir, filename, writer);
    }
    bsb_destroy(ctx);
    return 0;
}
