// bsb_api.cu -- the C ABI of include/blackstar_b200.h over the sm_100a kernels.
//
// Replaces, for Main.doRender (app/Main.hs:105-118):
//   Raytracer.render (+ supersample)  -> bsb_render / bsb_render_device
//   ImageFilters.bloom                -> bsb_bloom / bsb_bloom_device
//   render -> bloom on every GPU      -> bsb_render_full
//   writeImg's sRGB + toWord8 map     -> bsb_to_srgb8 (N1 of SURVEY.md section 8f)
// There is no CPU fallback anywhere in this file.
#include "../../include/blackstar_b200.h"

#include "bsb_common.cuh"
#include "host_setup.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace bsb {
cudaError_t launch_trace(const FrameParams &P, float4 *out, TraceCounters *ctr, int n_sms, int variant,
                         cudaStream_t stream);
cudaError_t launch_ray_tables(const FrameParams &P, double *vx, double *vy, cudaStream_t stream);
cudaError_t launch_rinv5_selftest(double q_lo, double q_hi, int n, double *d_out2, cudaStream_t stream);
cudaError_t launch_box3(const BoxArgs &A, cudaStream_t stream);
cudaError_t launch_transpose(const float4 *in, float4 *out, int rows, int cols, size_t in_pitch, size_t out_pitch, cudaStream_t stream);
cudaError_t launch_bloom_long(const float4 *img, float4 *out, uint8_t *rgb8, const float *thr, float4 *tmp_a, float4 *tmp_b,
                              int W, int H, int r, float strength, cudaStream_t stream);
int bloom_max_line();
cudaError_t launch_srgb8(const float4 *in, uint8_t *out, const float *thr, size_t npix, cudaStream_t stream);
cudaError_t launch_dfma_peak(double *sink, int blocks, int iters, cudaStream_t stream);
}  // namespace bsb

using namespace bsb;

namespace {

thread_local std::string g_create_error;

// ---- NCCL, loaded lazily: only a multi-GPU ctx needs it -------------------------------
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *handle = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string &err)
    {
        if (handle) return true;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define BSB_SYM(field, name)                                             \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));      \
    if (!field) { err = std::string("libnccl lacks ") + name; return false; }
        BSB_SYM(CommInitAll, "ncclCommInitAll");
        BSB_SYM(CommDestroy, "ncclCommDestroy");
        BSB_SYM(GroupStart, "ncclGroupStart");
        BSB_SYM(GroupEnd, "ncclGroupEnd");
        BSB_SYM(Send, "ncclSend");
        BSB_SYM(Recv, "ncclRecv");
        BSB_SYM(GetErrorString, "ncclGetErrorString");
#undef BSB_SYM
        return true;
    }
};
constexpr int kNcclFloat = 7;  // ncclFloat32 in nccl.h

// A few parked host threads that move staged chunks into the caller's (pageable) buffer.
class CopyPool {
public:
    explicit CopyPool(int n) : n_(n)
    {
        for (int k = 0; k < n; k++) th_.emplace_back([this, k] { loop(k); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void begin(std::function<void(int)> f)
    {
        std::lock_guard<std::mutex> g(m_);
        job_ = std::move(f); pending_ = n_; gen_++;
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
    }

private:
    void loop(int k)
    {
        unsigned long seen = 0;
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = job_;
            }
            f(k);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(int)> job_;
    unsigned long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

constexpr int kStageSlots = 4;
constexpr size_t kStageChunk = (size_t)8 << 20;

struct DeviceState {
    int dev = -1;
    int n_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // own_stream, or the caller's (device 0 only)
    // star map
    float *d_top = nullptr;
    double *d_rec = nullptr;
    StarRec *d_stars = nullptr;
    uint32_t rec_off[4] = { 0, 0, 0, 0 };
    int depth = 0;
    int top_levels = 0;
    int n_stars = 0;
    // per-launch counters
    TraceCounters *d_ctr = nullptr;
    TraceCounters *h_ctr = nullptr;  // pinned
    // scratch framebuffers
    float4 *d_frame = nullptr; size_t frame_cap = 0;  // this GPU's tile / the full frame on GPU 0
    float4 *d_tmp = nullptr;   size_t tmp_cap = 0;    // bloom's transposed intermediate
    float4 *d_tmp2 = nullptr;  size_t tmp2_cap = 0;   // second scratch frame of the long-line bloom
    float *d_thr = nullptr;                           // sRGB8 thresholds (256 floats)
    cudaEvent_t stage_ev[kStageSlots] = {};           // "my part of staging slot s has landed"
    // multi-GPU bloom: H^3 of my row tile and the tile itself, both transposed ([W][rows]); what the
    // other GPUs sent me for my column band ([cols][H] in row-tile pieces); my band of the result
    float4 *d_mid = nullptr;   size_t mid_cap = 0;
    float4 *d_imgT = nullptr;  size_t imgT_cap = 0;
    float4 *d_rmid = nullptr;  size_t rmid_cap = 0;
    float4 *d_rimg = nullptr;  size_t rimg_cap = 0;
    float4 *d_band = nullptr;  size_t band_cap = 0;
    float4 *d_aux = nullptr;   size_t aux_cap = 0;    // staging for host-buffer bloom / srgb
    uint8_t *d_u8 = nullptr;   size_t u8_cap = 0;
    double *d_vx = nullptr;    size_t vx_cap = 0;     // per-frame ray tables
    double *d_vy = nullptr;    size_t vy_cap = 0;
    double *d_misc = nullptr;                         // 2 doubles for self-tests
    cudaEvent_t ev[6] = {};
    double rows_per_ms = 0.0;                         // measured trace rate of the previous bsb_render_full tile
    int last_rows = 0;
};

}  // namespace

struct bsb_ctx {
    std::vector<DeviceState> devs;
    std::string err;
    int trace_variant = 6;  // tiles schedule, <=128-register build (2 CTAs/SM): fastest measured (profiles/)
    size_t n_stars = 0;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    std::unique_ptr<CopyPool> pool;   // created on the first copy into pageable memory
    uint8_t *h_stage = nullptr;       // kStageSlots pinned (portable) chunks for copies into pageable memory
    int copy_threads = 8;
    uint32_t step_cap = 0;            // 0 = the default of make_frame_params
};

namespace {

int fail(bsb_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

#define BSB_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(ctx, BSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <typename T>
int ensure(bsb_ctx *ctx, T *&ptr, size_t &cap, size_t need)
{
    if (need <= cap) return BSB_OK;
    if (ptr) BSB_CUDA(ctx, cudaFree(ptr));
    ptr = nullptr; cap = 0;
    // 1/8 of slack: the row tiles of a multi-GPU ctx are re-cut from measured rates every frame, and a tile that
    // grows by a few rows must not cost a cudaFree + cudaMalloc (both synchronise the device)
    const size_t want = need + need / 8;
    if (cudaMalloc(reinterpret_cast<void **>(&ptr), want * sizeof(T)) == cudaSuccess) { cap = want; return BSB_OK; }
    (void)cudaGetLastError();
    BSB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ptr), need * sizeof(T)));
    cap = need;
    return BSB_OK;
}

void free_tree(DeviceState &d)
{
    if (d.d_top) cudaFree(d.d_top);
    if (d.d_rec) cudaFree(d.d_rec);
    if (d.d_stars) cudaFree(d.d_stars);
    d.d_top = nullptr; d.d_rec = nullptr; d.d_stars = nullptr;
    d.depth = 0; d.top_levels = 0; d.n_stars = 0;
}

double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// Launch the trace of rows [row0,row1) on device d into dst (device memory on d), async.
int trace_async(bsb_ctx *ctx, DeviceState &d, const bsb_camera *cam, const bsb_scene *scn, int row0, int row1,
                float4 *dst, cudaEvent_t ev_begin, cudaEvent_t ev_end)
{
    FrameParams P;
    const std::string msg = make_frame_params(*cam, *scn, row0, row1, P);
    if (!msg.empty()) return fail(ctx, BSB_ERR_INVALID, msg);
    if (ctx->step_cap) P.step_cap = ctx->step_cap;
    P.tree.top = d.d_top;
    P.tree.rec = d.d_rec;
    P.tree.stars = d.d_stars;
    P.tree.depth = d.depth;
    P.tree.top_levels = d.top_levels;
    P.tree.n_stars = d.n_stars;
    for (int g = 0; g < 4; g++) P.tree.rec_off[g] = d.rec_off[g];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    int rc = ensure(ctx, d.d_vx, d.vx_cap, (size_t)P.W2);
    if (rc) return rc;
    rc = ensure(ctx, d.d_vy, d.vy_cap, (size_t)P.H2);
    if (rc) return rc;
    BSB_CUDA(ctx, cudaMemsetAsync(d.d_ctr, 0, sizeof(TraceCounters), d.stream));
    if (ev_begin) BSB_CUDA(ctx, cudaEventRecord(ev_begin, d.stream));
    if (row1 > row0) {
        BSB_CUDA(ctx, launch_ray_tables(P, d.d_vx, d.d_vy, d.stream));
        P.vx_tab = d.d_vx;
        P.vy_tab = d.d_vy;
    }
    BSB_CUDA(ctx, launch_trace(P, dst, d.d_ctr, d.n_sms, ctx->trace_variant, d.stream));
    if (ev_end) BSB_CUDA(ctx, cudaEventRecord(ev_end, d.stream));
    BSB_CUDA(ctx, cudaMemcpyAsync(d.h_ctr, d.d_ctr, sizeof(TraceCounters), cudaMemcpyDeviceToHost, d.stream));
    return BSB_OK;
}

// bloom on device d: src (h x w) -> dst (h x w) and / or its sRGB8 image; src may equal dst.
// 2 launches (7 for lines too long for the shared-memory kernel).
int bloom_async(bsb_ctx *ctx, DeviceState &d, double strength, int divider, int w, int h, const float4 *src,
                float4 *dst, uint8_t *rgb8, int *launches = nullptr)
{
    if (w <= 0 || h <= 0) return fail(ctx, BSB_ERR_INVALID, "bloom: image size must be positive");
    if (divider <= 0) return fail(ctx, BSB_ERR_INVALID, "bloom: bloomDivider must be positive (`div` by zero in the reference)");
    const int r = w / divider;  // src/ImageFilters.hs:83
    if (r < 1)
        return fail(ctx, BSB_ERR_INVALID,
                    "bloom: radius 0 (width < bloomDivider); the reference's boxBlur fails here (foldl1' of an empty window)");
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    int rc = ensure(ctx, d.d_tmp, d.tmp_cap, (size_t)w * h);
    if (rc) return rc;
    if (w > bloom_max_line() || h > bloom_max_line()) {
        // the reference has no size limit: sequential running sums, one thread per line and channel
        rc = ensure(ctx, d.d_tmp2, d.tmp2_cap, (size_t)w * h);
        if (rc) return rc;
        BSB_CUDA(ctx, launch_bloom_long(src, dst, rgb8, d.d_thr, d.d_tmp, d.d_tmp2, w, h, r, (float)strength, d.stream));
        if (launches) *launches += 7;
        return BSB_OK;
    }
    BoxArgs A;
    std::memset(&A, 0, sizeof A);
    A.r = r;
    A.norm = (float)(1.0 / (2.0 * (double)r + 1.0));  // src/ImageFilters.hs:51
    A.thr = d.d_thr;
    // H^3: lines = rows of the image, written transposed (w x h)
    A.nseg = 1; A.seg_in[0] = src; A.seg_pitch[0] = (size_t)w; A.seg_start[0] = 0; A.seg_start[1] = w;
    A.out = d.d_tmp; A.out_pitch = (size_t)h;
    A.n = w; A.lines = h; A.x_lo = 0; A.x_hi = w;
    BSB_CUDA(ctx, launch_box3(A, d.stream));
    // V^3 on the transposed image: lines = w, n = h; transposes back, adds the original, maps to sRGB8
    A.seg_in[0] = d.d_tmp; A.seg_pitch[0] = (size_t)h; A.seg_start[1] = h;
    A.out = dst; A.out_pitch = (size_t)w; A.img = src;
    A.rgb8 = rgb8; A.rgb8_pitch = (size_t)w * 3;
    A.n = h; A.lines = w; A.x_lo = 0; A.x_hi = h; A.combine = 1; A.strength = (float)strength;
    BSB_CUDA(ctx, launch_box3(A, d.stream));
    if (launches) *launches += 2;
    return BSB_OK;
}

// One GPU's contribution to a host frame: rows [row0, row0+rows) x bytes [x_off, x_off+band_bytes)
// of every such row, read from device memory laid out [rows][band_bytes].
struct Band {
    DeviceState *d;
    const uint8_t *src;
    size_t x_off, band_bytes;
    int row0, rows;
};

// Device(s) -> caller-owned host frame of `rows` rows of `row_bytes` bytes; every band is ordered
// after the work already queued on its GPU's stream.  Returns once the copies are QUEUED for a
// page-locked target and once the bytes are in `dst` for a pageable one.
//  * page-locked target (cudaHostAlloc / cudaHostRegister, e.g. torch's pinned tensors or a buffer
//    the Haskell shim registered): every GPU's DMA engine writes its band directly, in parallel;
//  * pageable target (plain malloc, what mallocForeignPtrBytes hands to the FFI): a cudaMemcpy to
//    pageable memory is staged by the driver through one small buffer and runs at a fraction of the
//    link rate, so the library stages it itself -- the frame is cut into ~8 MB groups of rows that go
//    through a ring of pinned buffers while a few parked host threads move finished groups into the
//    caller's buffer, so the PCIe transfers and the host copy overlap.
int copy_bands_to_host(bsb_ctx *ctx, void *dst, size_t row_bytes, int rows, const std::vector<Band> &bands)
{
    if (rows <= 0 || row_bytes == 0 || bands.empty()) return BSB_OK;
    uint8_t *out = static_cast<uint8_t *>(dst);
    const size_t bytes = row_bytes * (size_t)rows;
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, dst);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    const bool pinned = pe == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    // small frames (a 1080p RGB8 image is 6 MB) are not worth waking the copy threads for: the driver's own
    // staged copy takes well under a millisecond, and batch jobs run one such process per GPU.  That holds for
    // full-width bands only: a 2-D copy into PAGEABLE memory is staged by the driver row by row (measured:
    // ~70 us per row, 150 ms for the two column bands of a 1080p frame), so those always go through the ring.
    bool full_width = true;
    for (const Band &b : bands) full_width = full_width && b.band_bytes == row_bytes;
    if (pinned || ctx->copy_threads <= 0 || (full_width && bytes < 2 * kStageChunk)) {
        for (const Band &b : bands) {
            BSB_CUDA(ctx, cudaSetDevice(b.d->dev));
            uint8_t *row = out + (size_t)b.row0 * row_bytes;
            if (b.band_bytes == row_bytes)
                BSB_CUDA(ctx, cudaMemcpyAsync(row, b.src, row_bytes * (size_t)b.rows, cudaMemcpyDeviceToHost, b.d->stream));
            else
                BSB_CUDA(ctx, cudaMemcpy2DAsync(row + b.x_off, row_bytes, b.src, b.band_bytes, b.band_bytes, (size_t)b.rows,
                                                cudaMemcpyDeviceToHost, b.d->stream));
        }
        return BSB_OK;
    }
    if (!ctx->h_stage)
        BSB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_stage), kStageSlots * kStageChunk, cudaHostAllocPortable));
    for (const Band &b : bands)
        if (!b.d->stage_ev[0]) {
            BSB_CUDA(ctx, cudaSetDevice(b.d->dev));
            for (auto &e : b.d->stage_ev) BSB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    if (!ctx->pool || ctx->pool->size() != ctx->copy_threads) ctx->pool.reset(new CopyPool(ctx->copy_threads));
    if (row_bytes > kStageChunk) return fail(ctx, BSB_ERR_UNSUPPORTED, "image row above 8 MB");
    const int rows_per_chunk = (int)std::max<size_t>(1, kStageChunk / row_bytes);
    const int nchunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    const int nt = ctx->pool->size();
    std::vector<std::atomic<int>> copied(nchunks);
    for (auto &c : copied) c.store(0, std::memory_order_relaxed);
    std::atomic<int> enqueued{ 0 };
    std::atomic<int> failed{ 0 };
    uint8_t *stage = ctx->h_stage;
    ctx->pool->begin([&, nt](int k) {
        for (int i = 0; i < nchunks; i++) {
            while (enqueued.load(std::memory_order_acquire) <= i) {
                if (failed.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            for (const Band &b : bands)
                if (cudaEventSynchronize(b.d->stage_ev[i % kStageSlots]) != cudaSuccess) failed.store(1);
            const int ya = i * rows_per_chunk, yb = std::min(rows, ya + rows_per_chunk);
            const size_t len = (size_t)(yb - ya) * row_bytes;
            const size_t a = len * k / nt, e = len * (k + 1) / nt;   // my slice of the group of rows
            std::memcpy(out + (size_t)ya * row_bytes + a, stage + (size_t)(i % kStageSlots) * kStageChunk + a, e - a);
            copied[i].fetch_add(1, std::memory_order_release);
        }
    });
    cudaError_t err = cudaSuccess;
    for (int i = 0; i < nchunks && err == cudaSuccess; i++) {
        if (i >= kStageSlots)   // the slot is free once every thread has copied its slice of group i - slots
            while (copied[i - kStageSlots].load(std::memory_order_acquire) < nt) std::this_thread::yield();
        const int ya = i * rows_per_chunk, yb = std::min(rows, ya + rows_per_chunk);
        uint8_t *slot = stage + (size_t)(i % kStageSlots) * kStageChunk;
        for (const Band &b : bands) {
            if (err != cudaSuccess) break;
            err = cudaSetDevice(b.d->dev);
            const int y0 = std::max(ya, b.row0), y1 = std::min(yb, b.row0 + b.rows);
            if (err == cudaSuccess && y1 > y0)
                err = cudaMemcpy2DAsync(slot + (size_t)(y0 - ya) * row_bytes + b.x_off, row_bytes,
                                        b.src + (size_t)(y0 - b.row0) * b.band_bytes, b.band_bytes, b.band_bytes, (size_t)(y1 - y0),
                                        cudaMemcpyDeviceToHost, b.d->stream);
            if (err == cudaSuccess) err = cudaEventRecord(b.d->stage_ev[i % kStageSlots], b.d->stream);
        }
        if (err == cudaSuccess) enqueued.store(i + 1, std::memory_order_release);
    }
    if (err != cudaSuccess) failed.store(1);
    ctx->pool->wait();
    if (err != cudaSuccess) return fail(ctx, BSB_ERR_CUDA, std::string("staged device-to-host copy: ") + cudaGetErrorString(err));
    if (failed.load()) return fail(ctx, BSB_ERR_CUDA, "staged device-to-host copy failed");
    return BSB_OK;
}

int copy_to_host(bsb_ctx *ctx, DeviceState &d, void *dst, const void *src_dev, size_t bytes)
{
    if (bytes == 0) return BSB_OK;
    // a flat buffer is a frame of 1 MB rows (the last one ragged: copied as its own band)
    const size_t rb = (size_t)1 << 20;
    const int full = (int)(bytes / rb);
    const uint8_t *src = static_cast<const uint8_t *>(src_dev);
    if (full > 0) {
        std::vector<Band> bands{ Band{ &d, src, 0, rb, 0, full } };
        const int rc = copy_bands_to_host(ctx, dst, rb, full, bands);
        if (rc) return rc;
    }
    if (bytes > (size_t)full * rb) {
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaMemcpyAsync(static_cast<uint8_t *>(dst) + (size_t)full * rb, src + (size_t)full * rb, bytes - (size_t)full * rb,
                                      cudaMemcpyDeviceToHost, d.stream));
    }
    return BSB_OK;
}

void fill_counter_stats(const DeviceState &d, bsb_stats *st)
{
    st->steps += d.h_ctr->steps;
    st->capped += d.h_ctr->capped;
    st->star_hits += d.h_ctr->star_hits;
}

uint64_t rays_of(const bsb_scene *scn, int rows)
{
    return (uint64_t)rows * (uint64_t)scn->width * (scn->supersampling ? 4u : 1u);
}

}  // namespace

// ======================================================================== lifetime
extern "C" const char *bsb_version(void) { return "blackstar_b200 0.1.0 sm_100a"; }

extern "C" const char *bsb_last_error(const bsb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" bsb_ctx *bsb_create_on(const int *devices, int n)
{
    g_create_error.clear();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (blackstar_b200 has no CPU fallback)";
        return nullptr;
    }
    if (n <= 0 || !devices) { g_create_error = "empty device list"; return nullptr; }
    bsb_ctx *ctx = new bsb_ctx();
    if (const char *v = std::getenv("BSB_TRACE_VARIANT")) {   // same range check as bsb_set_option
        char *end = nullptr;
        const long x = std::strtol(v, &end, 10);
        if (end == v || *end != 0 || x < 0 || x > 6 || x == 5) {
            g_create_error = "BSB_TRACE_VARIANT must be 0..4 or 6";
            delete ctx;
            return nullptr;
        }
        ctx->trace_variant = (int)x;
    }
    if (n > kMaxSegments) {
        g_create_error = "at most 8 GPUs per context (one NVSwitch box)";
        delete ctx;
        return nullptr;
    }
    float thr[256];
    srgb8_thresholds(thr);
    for (int k = 0; k < n; k++) {
        const int dev = devices[k];
        if (dev < 0 || dev >= count) {
            g_create_error = "device index out of range";
            bsb_destroy(ctx);
            return nullptr;
        }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) {
            g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
            bsb_destroy(ctx);
            return nullptr;
        }
        if (prop.major != 10) {
            char buf[160];
            std::snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library contains sm_100a code only", dev, prop.major, prop.minor);
            g_create_error = buf;
            bsb_destroy(ctx);
            return nullptr;
        }
        DeviceState d;
        d.dev = dev;
        d.n_sms = prop.multiProcessorCount;
        bool ok = cudaSetDevice(dev) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&d.own_stream, cudaStreamNonBlocking) == cudaSuccess;
        d.stream = d.own_stream;
        ok = ok && cudaMalloc(reinterpret_cast<void **>(&d.d_ctr), sizeof(TraceCounters)) == cudaSuccess;
        ok = ok && cudaMallocHost(reinterpret_cast<void **>(&d.h_ctr), sizeof(TraceCounters)) == cudaSuccess;
        ok = ok && cudaMalloc(reinterpret_cast<void **>(&d.d_misc), 4 * sizeof(double)) == cudaSuccess;
        ok = ok && cudaMalloc(reinterpret_cast<void **>(&d.d_thr), 256 * sizeof(float)) == cudaSuccess;
        ok = ok && cudaMemcpy(d.d_thr, thr, sizeof thr, cudaMemcpyHostToDevice) == cudaSuccess;
        for (auto &ev : d.ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        ctx->devs.push_back(d);
        if (!ok) {
            g_create_error = std::string("device setup failed: ") + cudaGetErrorString(cudaGetLastError());
            bsb_destroy(ctx);
            return nullptr;
        }
        std::memset(d.h_ctr, 0, sizeof(TraceCounters));
    }
    if (n > 1) {
        std::string err;
        if (!ctx->nccl.load(err)) { g_create_error = err; bsb_destroy(ctx); return nullptr; }
        ctx->comms.assign(n, nullptr);
        const int rc = ctx->nccl.CommInitAll(ctx->comms.data(), n, devices);
        if (rc != 0) {
            g_create_error = std::string("ncclCommInitAll: ") + ctx->nccl.GetErrorString(rc);
            ctx->comms.clear();
            bsb_destroy(ctx);
            return nullptr;
        }
    }
    return ctx;
}

extern "C" bsb_ctx *bsb_create(int n_gpus)
{
    g_create_error.clear();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (blackstar_b200 has no CPU fallback)";
        return nullptr;
    }
    if (n_gpus < 0 || n_gpus > count) { g_create_error = "n_gpus exceeds the visible devices"; return nullptr; }
    const bool all = n_gpus == 0;
    if (all) n_gpus = std::min(count, kMaxSegments);
    std::vector<int> devs(n_gpus);
    for (int k = 0; k < n_gpus; k++) devs[k] = k;
    bsb_ctx *ctx = bsb_create_on(devs.data(), n_gpus);
    if (!ctx && all && n_gpus > 1) {
        // "every visible device" is a wish, not a requirement: without a loadable NCCL fall back to one GPU
        const std::string why = g_create_error;
        ctx = bsb_create_on(devs.data(), 1);
        if (ctx) std::fprintf(stderr, "blackstar_b200: multi-GPU context failed (%s); using device 0 only\n", why.c_str());
        else g_create_error = why;
    }
    return ctx;
}

extern "C" void bsb_destroy(bsb_ctx *ctx)
{
    if (!ctx) return;
    for (size_t k = 0; k < ctx->comms.size(); k++)
        if (ctx->comms[k]) ctx->nccl.CommDestroy(ctx->comms[k]);
    for (DeviceState &d : ctx->devs) {
        cudaSetDevice(d.dev);
        if (d.own_stream) cudaStreamSynchronize(d.own_stream);
        free_tree(d);
        if (d.d_ctr) cudaFree(d.d_ctr);
        if (d.h_ctr) cudaFreeHost(d.h_ctr);
        if (d.d_frame) cudaFree(d.d_frame);
        if (d.d_tmp) cudaFree(d.d_tmp);
        if (d.d_tmp2) cudaFree(d.d_tmp2);
        if (d.d_thr) cudaFree(d.d_thr);
        for (auto &e : d.stage_ev) if (e) cudaEventDestroy(e);
        if (d.d_mid) cudaFree(d.d_mid);
        if (d.d_imgT) cudaFree(d.d_imgT);
        if (d.d_rmid) cudaFree(d.d_rmid);
        if (d.d_rimg) cudaFree(d.d_rimg);
        if (d.d_band) cudaFree(d.d_band);
        if (d.d_aux) cudaFree(d.d_aux);
        if (d.d_u8) cudaFree(d.d_u8);
        if (d.d_vx) cudaFree(d.d_vx);
        if (d.d_vy) cudaFree(d.d_vy);
        if (d.d_misc) cudaFree(d.d_misc);
        for (auto &ev : d.ev) if (ev) cudaEventDestroy(ev);
        if (d.own_stream) cudaStreamDestroy(d.own_stream);
    }
    ctx->pool.reset();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    delete ctx;
}

extern "C" int bsb_set_stream(bsb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return BSB_ERR_INVALID;
    DeviceState &d = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    d.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : d.own_stream;
    return BSB_OK;
}

extern "C" int bsb_set_option(bsb_ctx *ctx, const char *key, double value)
{
    if (!ctx || !key) return BSB_ERR_INVALID;
    if (std::strcmp(key, "trace_variant") == 0) {
        if (value < 0 || value > 6 || (int)value == 5) return fail(ctx, BSB_ERR_INVALID, "trace_variant must be 0..4 or 6");
        ctx->trace_variant = (int)value;
        return BSB_OK;
    }
    if (std::strcmp(key, "copy_threads") == 0) {
        if (value < 0 || value > 64) return fail(ctx, BSB_ERR_INVALID, "copy_threads must be 0..64 (0 = leave pageable copies to the driver)");
        ctx->copy_threads = (int)value;
        return BSB_OK;
    }
    if (std::strcmp(key, "step_cap") == 0) {
        if (value < 1 || value > 4294967295.0) return fail(ctx, BSB_ERR_INVALID, "step_cap must be 1..2^32-1");
        ctx->step_cap = (uint32_t)value;
        return BSB_OK;
    }
    return fail(ctx, BSB_ERR_INVALID, std::string("unknown option ") + key);
}

// ======================================================================== star map
extern "C" int bsb_set_stars(bsb_ctx *ctx, const bsb_star *stars, size_t n)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (n > 0 && !stars) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars: NULL star list");
    if (n > (size_t)1 << 28) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars: too many stars");
    if (n > 0) {
        const std::string bad = validate_stars(stars, n);
        if (!bad.empty()) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars: " + bad);
    }
    HostStarTree t;
    if (n > 0) build_star_tree(stars, n, t);
    for (DeviceState &d : ctx->devs) {
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        free_tree(d);
        if (n == 0) continue;
        BSB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d.d_top), t.top.size() * sizeof(float)));
        BSB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d.d_rec), t.rec.size() * sizeof(double)));
        BSB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d.d_stars), t.stars.size() * sizeof(StarRec)));
        BSB_CUDA(ctx, cudaMemcpy(d.d_top, t.top.data(), t.top.size() * sizeof(float), cudaMemcpyHostToDevice));
        BSB_CUDA(ctx, cudaMemcpy(d.d_rec, t.rec.data(), t.rec.size() * sizeof(double), cudaMemcpyHostToDevice));
        BSB_CUDA(ctx, cudaMemcpy(d.d_stars, t.stars.data(), t.stars.size() * sizeof(StarRec), cudaMemcpyHostToDevice));
        for (int g = 0; g < 4; g++) d.rec_off[g] = t.rec_off[g];
        d.top_levels = t.top_levels;
        d.depth = t.depth;
        d.n_stars = (int)n;
    }
    ctx->n_stars = n;
    return BSB_OK;
}

extern "C" int bsb_set_stars_ppm(bsb_ctx *ctx, const uint8_t *bytes, size_t len)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!bytes) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars_ppm: NULL buffer");
    std::vector<bsb_star> stars;
    std::string err;
    if (!parse_ppm(bytes, len, stars, err)) return fail(ctx, BSB_ERR_INVALID, "Error decoding star map: " + err);
    return bsb_set_stars(ctx, stars.data(), stars.size());
}

extern "C" int bsb_set_stars_file(bsb_ctx *ctx, const uint8_t *bytes, size_t len)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!bytes) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars_file: NULL buffer");
    std::vector<bsb_star> stars;
    std::string err;
    if (!parse_star_file(bytes, len, stars, err)) return fail(ctx, BSB_ERR_INVALID, "Error decoding star map: " + err);
    return bsb_set_stars(ctx, stars.data(), stars.size());
}

extern "C" size_t bsb_star_count(const bsb_ctx *ctx) { return ctx ? ctx->n_stars : 0; }

// ======================================================================== render
extern "C" int bsb_render_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, int row0, int row1,
                                 void *dev_out, bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!cam || !scn || (!dev_out && row1 > row0)) return fail(ctx, BSB_ERR_INVALID, "bsb_render_device: NULL argument");
    const auto t0 = std::chrono::steady_clock::now();
    DeviceState &d = ctx->devs[0];
    int rc = trace_async(ctx, d, cam, scn, row0, row1, static_cast<float4 *>(dev_out), d.ev[0], d.ev[1]);
    if (rc) return rc;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0;
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        stats->trace_ms = ms;
        stats->rays = rays_of(scn, row1 - row0);
        fill_counter_stats(d, stats);
        stats->n_gpus = 1;
        stats->launches = row1 > row0 ? 2 : 0;
        stats->total_ms = ms_since(t0);
        if (stats->capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    }
    return BSB_OK;
}

extern "C" int bsb_render(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, int row0, int row1,
                          float *out_rgba, bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!cam || !scn) return fail(ctx, BSB_ERR_INVALID, "bsb_render: NULL argument");
    if (row0 < 0 || row1 > scn->height || row0 > row1 || scn->width <= 0)
        return fail(ctx, BSB_ERR_INVALID, "row range outside the image");
    if (!out_rgba && row1 > row0) return fail(ctx, BSB_ERR_INVALID, "bsb_render: NULL output buffer");
    const auto t0 = std::chrono::steady_clock::now();
    DeviceState &d = ctx->devs[0];
    const size_t npix = (size_t)(row1 - row0) * scn->width;
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    int rc = ensure(ctx, d.d_frame, d.frame_cap, npix ? npix : 1);
    if (rc) return rc;
    rc = trace_async(ctx, d, cam, scn, row0, row1, d.d_frame, d.ev[0], d.ev[1]);
    if (rc) return rc;
    if (npix) {
        rc = copy_to_host(ctx, d, out_rgba, d.d_frame, npix * sizeof(float4));
        if (rc) return rc;
    }
    BSB_CUDA(ctx, cudaEventRecord(d.ev[2], d.stream));
    BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    bsb_stats st;
    std::memset(&st, 0, sizeof st);
    float ms = 0;
    BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
    st.trace_ms = ms;
    BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[1], d.ev[2]));
    st.d2h_ms = ms;
    st.rays = rays_of(scn, row1 - row0);
    fill_counter_stats(d, &st);
    st.n_gpus = 1;
    st.launches = npix ? 2 : 0;
    st.total_ms = ms_since(t0);
    if (stats) *stats = st;
    if (st.capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    return BSB_OK;
}

// ======================================================================== bloom
extern "C" int bsb_bloom_device(bsb_ctx *ctx, double strength, int divider, int width, int height,
                                const void *dev_in, void *dev_out)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!dev_in || !dev_out) return fail(ctx, BSB_ERR_INVALID, "bsb_bloom_device: NULL argument");
    return bloom_async(ctx, ctx->devs[0], strength, divider, width, height, static_cast<const float4 *>(dev_in),
                       static_cast<float4 *>(dev_out), nullptr);
}

extern "C" int bsb_bloom(bsb_ctx *ctx, double strength, int divider, int width, int height, const float *in_rgba,
                         float *out_rgba)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!in_rgba || !out_rgba) return fail(ctx, BSB_ERR_INVALID, "bsb_bloom: NULL argument");
    if (width <= 0 || height <= 0) return fail(ctx, BSB_ERR_INVALID, "bloom: image size must be positive");
    DeviceState &d = ctx->devs[0];
    const size_t npix = (size_t)width * height;
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    int rc = ensure(ctx, d.d_aux, d.aux_cap, npix);
    if (rc) return rc;
    BSB_CUDA(ctx, cudaMemcpyAsync(d.d_aux, in_rgba, npix * sizeof(float4), cudaMemcpyHostToDevice, d.stream));
    rc = bloom_async(ctx, d, strength, divider, width, height, d.d_aux, d.d_aux, nullptr);
    if (rc) return rc;
    rc = copy_to_host(ctx, d, out_rgba, d.d_aux, npix * sizeof(float4));
    if (rc) return rc;
    BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    return BSB_OK;
}

// ======================================================================== sRGB / 8 bit
extern "C" int bsb_to_srgb8_device(bsb_ctx *ctx, int width, int height, const void *dev_in, void *dev_out)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (width <= 0 || height <= 0 || !dev_in || !dev_out) return fail(ctx, BSB_ERR_INVALID, "bsb_to_srgb8_device: bad argument");
    DeviceState &d = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    BSB_CUDA(ctx, launch_srgb8(static_cast<const float4 *>(dev_in), static_cast<uint8_t *>(dev_out), d.d_thr, (size_t)width * height, d.stream));
    return BSB_OK;
}

extern "C" int bsb_to_srgb8(bsb_ctx *ctx, int width, int height, const float *in_rgba, uint8_t *out_rgb8)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (width <= 0 || height <= 0 || !in_rgba || !out_rgb8) return fail(ctx, BSB_ERR_INVALID, "bsb_to_srgb8: bad argument");
    DeviceState &d = ctx->devs[0];
    const size_t npix = (size_t)width * height;
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    int rc = ensure(ctx, d.d_aux, d.aux_cap, npix);
    if (rc) return rc;
    rc = ensure(ctx, d.d_u8, d.u8_cap, npix * 3 + 16);
    if (rc) return rc;
    BSB_CUDA(ctx, cudaMemcpyAsync(d.d_aux, in_rgba, npix * sizeof(float4), cudaMemcpyHostToDevice, d.stream));
    BSB_CUDA(ctx, launch_srgb8(d.d_aux, d.d_u8, d.d_thr, npix, d.stream));
    rc = copy_to_host(ctx, d, out_rgb8, d.d_u8, npix * 3);
    if (rc) return rc;
    BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    return BSB_OK;
}

// ======================================================================== doRender
namespace {

// Contiguous row tiles.  First frame: equal shares.  Afterwards: proportional to the rate every GPU
// achieved on its previous tile (rows per millisecond), so that tiles with fewer RK4 steps per ray
// (the hole, the far sky) get more rows and all GPUs finish together.  Boundaries are even.
void plan_row_tiles(bsb_ctx *ctx, int H, std::vector<int> &r0, std::vector<int> &r1)
{
    const int n = (int)ctx->devs.size();
    r0.assign(n, 0); r1.assign(n, 0);
    double total_rate = 0;
    bool known = n > 1;
    for (int k = 0; k < n; k++) { known = known && ctx->devs[k].rows_per_ms > 0; total_rate += ctx->devs[k].rows_per_ms; }
    double acc = 0;
    for (int k = 0; k < n; k++) {
        r0[k] = k == 0 ? 0 : r1[k - 1];
        int e;
        if (known) {
            acc += ctx->devs[k].rows_per_ms / total_rate;
            e = (int)(acc * H + 0.5);
        } else {
            e = (int)((long long)H * (k + 1) / n);
        }
        e &= ~1;
        r1[k] = k == n - 1 ? H : std::max(r0[k], std::min(H, e));
    }
}

int bloom_radius(bsb_ctx *ctx, int w, int h, int divider, int *r)
{
    if (w <= 0 || h <= 0) return fail(ctx, BSB_ERR_INVALID, "bloom: image size must be positive");
    if (divider <= 0) return fail(ctx, BSB_ERR_INVALID, "bloom: bloomDivider must be positive (`div` by zero in the reference)");
    *r = w / divider;  // src/ImageFilters.hs:83
    if (*r < 1)
        return fail(ctx, BSB_ERR_INVALID,
                    "bloom: radius 0 (width < bloomDivider); the reference's boxBlur fails here (foldl1' of an empty window)");
    return BSB_OK;
}

// H^3 of `rows` rows of width w on device d: src [rows][w] -> midT [w][rows]; imgT (optional) = src transposed
int bloom_h_async(bsb_ctx *ctx, DeviceState &d, int r, int w, int rows, const float4 *src, float4 *midT, float4 *imgT)
{
    if (rows <= 0) return BSB_OK;
    if (w > bloom_max_line()) return fail(ctx, BSB_ERR_UNSUPPORTED, "bloom: rows longer than 8192 pixels take the single-GPU path");
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    BoxArgs A;
    std::memset(&A, 0, sizeof A);
    A.r = r;
    A.norm = (float)(1.0 / (2.0 * (double)r + 1.0));
    A.nseg = 1; A.seg_in[0] = src; A.seg_pitch[0] = (size_t)w; A.seg_start[0] = 0; A.seg_start[1] = w;
    A.out = midT; A.out_pitch = (size_t)rows;
    A.n = w; A.lines = rows; A.x_lo = 0; A.x_hi = w;
    BSB_CUDA(ctx, launch_box3(A, d.stream));
    if (imgT) BSB_CUDA(ctx, launch_transpose(src, imgT, rows, w, (size_t)w, (size_t)rows, d.stream));
    return BSB_OK;
}

// V^3 + combine (+ sRGB8) of `cols` columns of height h on device d.  Column l is assembled from nseg
// pieces: piece s = seg_mid[s] + l * seg_rows[s] holds rows [sum of seg_rows before s, ...) of it;
// seg_img is the unfiltered image laid out the same way.  out [h][cols] float4 and / or rgb8 [h][cols*3].
int bloom_v_async(bsb_ctx *ctx, DeviceState &d, double strength, int r, int h, int cols, int nseg, const float4 *const *seg_mid,
                  const float4 *const *seg_img, const int *seg_rows, float4 *out, uint8_t *rgb8)
{
    if (cols <= 0 || h <= 0) return BSB_OK;
    if (nseg < 1 || nseg > kMaxSegments) return fail(ctx, BSB_ERR_INVALID, "bloom: 1..8 row segments");
    if (h > bloom_max_line()) return fail(ctx, BSB_ERR_UNSUPPORTED, "bloom: columns longer than 8192 pixels take the single-GPU path");
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    BoxArgs A;
    std::memset(&A, 0, sizeof A);
    A.r = r;
    A.norm = (float)(1.0 / (2.0 * (double)r + 1.0));
    A.thr = d.d_thr;
    int y = 0, m = 0;
    for (int s = 0; s < nseg; s++) {
        if (seg_rows[s] < 0) return fail(ctx, BSB_ERR_INVALID, "bloom: negative segment");
        if (seg_rows[s] == 0) continue;
        A.seg_in[m] = seg_mid[s]; A.seg_img[m] = seg_img[s]; A.seg_pitch[m] = (size_t)seg_rows[s]; A.seg_start[m] = y;
        y += seg_rows[s];
        m++;
    }
    if (y != h) return fail(ctx, BSB_ERR_INVALID, "bloom: the row segments do not add up to the image height");
    A.seg_start[m] = h;
    A.nseg = m;
    A.out = out; A.out_pitch = (size_t)cols; A.rgb8 = rgb8; A.rgb8_pitch = (size_t)cols * 3;
    A.n = h; A.lines = cols; A.x_lo = 0; A.x_hi = h; A.combine = 2; A.strength = (float)strength;
    BSB_CUDA(ctx, launch_box3(A, d.stream));
    return BSB_OK;
}

struct FullResult {
    std::vector<Band> f32, u8;   // where the frame is (device memory), as bands of the host frame
};

// One GPU: trace -> bloom (2 launches, sRGB8 fused) -> frame in d_frame / d_u8.
int render_full_single(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, bsb_stats *st, bool want_float, bool want_rgb8,
                       FullResult &res)
{
    const int W = scn->width, H = scn->height;
    const size_t npix = (size_t)W * H;
    DeviceState &d0 = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d0.dev));
    int rc = ensure(ctx, d0.d_frame, d0.frame_cap, npix);
    if (rc) return rc;
    rc = trace_async(ctx, d0, cam, scn, 0, H, d0.d_frame, d0.ev[0], d0.ev[1]);
    if (rc) return rc;
    d0.last_rows = H;
    int launches = 2;
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[5], d0.stream));
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[2], d0.stream));
    if (want_rgb8) {
        rc = ensure(ctx, d0.d_u8, d0.u8_cap, npix * 3 + 16);
        if (rc) return rc;
    }
    if (scn->bloom_strength != 0) {   // app/Main.hs:113: bloom only if bloomStrength /= 0
        // the sRGB + toWord8 map of writeImg rides in the epilogue of the second bloom launch; the
        // float frame is only written if somebody wants it
        rc = bloom_async(ctx, d0, scn->bloom_strength, scn->bloom_divider, W, H, d0.d_frame, want_float ? d0.d_frame : nullptr,
                         want_rgb8 ? d0.d_u8 : nullptr, &launches);
        if (rc) return rc;
    } else if (want_rgb8) {
        BSB_CUDA(ctx, launch_srgb8(d0.d_frame, d0.d_u8, d0.d_thr, npix, d0.stream));
        launches += 1;
    }
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
    if (want_float) res.f32.push_back(Band{ &d0, reinterpret_cast<const uint8_t *>(d0.d_frame), 0, (size_t)W * 16, 0, H });
    if (want_rgb8) res.u8.push_back(Band{ &d0, d0.d_u8, 0, (size_t)W * 3, 0, H });
    st->launches = launches; st->n_gpus = 1; st->rays = rays_of(scn, H);
    return BSB_OK;
}

// N GPUs.  Rays are independent, so GPU k traces a contiguous tile of rows.  The bloom is separable:
//   H^3 needs whole rows  -> every GPU filters its own row tile (no communication);
//   V^3 needs whole columns -> ONE all-to-all re-cuts the frame from row tiles into column bands
//       (the H^3-filtered tile and the tile itself, both already transposed, go out in contiguous
//       pieces); every GPU then filters its band, adds the original and maps it to sRGB8.
// The finished frame is a set of column bands, one per GPU; N DMA engines copy them into the caller's
// host frame in parallel.  (Without bloom the row tiles go straight to the host.)  Everything is
// queued asynchronously; the only host synchronisation is at the end.
int render_full_multi(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, bsb_stats *st, bool want_float, bool want_rgb8,
                      FullResult &res)
{
    const int n = (int)ctx->devs.size();
    const int W = scn->width, H = scn->height;
    const bool bloom = scn->bloom_strength != 0;
    int r = 0;
    if (bloom) {
        const int rc = bloom_radius(ctx, W, H, scn->bloom_divider, &r);
        if (rc) return rc;
    }
    // sides above 8192 pixels: the bloom's long-line path runs on one GPU (the reference has no size limit, so
    // neither may this): tiles are traced on all GPUs, copied to the first one over NVLink, bloomed there
    const bool gather_bloom = bloom && (W > bloom_max_line() || H > bloom_max_line());
    std::vector<int> r0, r1, c0(n), c1(n);
    plan_row_tiles(ctx, H, r0, r1);
    for (int k = 0; k < n; k++) {   // column bands: equal, even boundaries
        c0[k] = k == 0 ? 0 : c1[k - 1];
        c1[k] = k == n - 1 ? W : std::max(c0[k], std::min(W, (int)((long long)W * (k + 1) / n) & ~1));
    }
    int launches = 0, rc;
    if (gather_bloom) {
        DeviceState &d0 = ctx->devs[0];
        const size_t npix = (size_t)W * H;
        BSB_CUDA(ctx, cudaSetDevice(d0.dev));
        if ((rc = ensure(ctx, d0.d_aux, d0.aux_cap, npix))) return rc;
        if (want_rgb8 && (rc = ensure(ctx, d0.d_u8, d0.u8_cap, npix * 3 + 16))) return rc;
        for (int k = 0; k < n; k++) {
            DeviceState &d = ctx->devs[k];
            const int hk = r1[k] - r0[k];
            BSB_CUDA(ctx, cudaSetDevice(d.dev));
            if ((rc = ensure(ctx, d.d_frame, d.frame_cap, (size_t)hk * W + 1))) return rc;
            if ((rc = trace_async(ctx, d, cam, scn, r0[k], r1[k], d.d_frame, d.ev[0], d.ev[1]))) return rc;
            d.last_rows = hk;
            launches += hk > 0 ? 2 : 0;
            BSB_CUDA(ctx, cudaEventRecord(d.ev[5], d.stream));
            BSB_CUDA(ctx, cudaEventRecord(d.ev[2], d.stream));
            BSB_CUDA(ctx, cudaEventRecord(d.ev[3], d.stream));
        }
        BSB_CUDA(ctx, cudaSetDevice(d0.dev));
        for (int k = 0; k < n; k++) {
            DeviceState &d = ctx->devs[k];
            const size_t cnt = (size_t)(r1[k] - r0[k]) * W;
            if (cnt == 0) continue;
            BSB_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d.ev[1], 0));
            BSB_CUDA(ctx, cudaMemcpyPeerAsync(d0.d_aux + (size_t)r0[k] * W, d0.dev, d.d_frame, d.dev, cnt * sizeof(float4), d0.stream));
        }
        BSB_CUDA(ctx, cudaEventRecord(d0.ev[2], d0.stream));
        if ((rc = bloom_async(ctx, d0, scn->bloom_strength, scn->bloom_divider, W, H, d0.d_aux, want_float ? d0.d_aux : nullptr,
                              want_rgb8 ? d0.d_u8 : nullptr, &launches)))
            return rc;
        BSB_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
        if (want_float) res.f32.push_back(Band{ &d0, reinterpret_cast<const uint8_t *>(d0.d_aux), 0, (size_t)W * 16, 0, H });
        if (want_rgb8) res.u8.push_back(Band{ &d0, d0.d_u8, 0, (size_t)W * 3, 0, H });
        st->launches = launches; st->n_gpus = n; st->rays = rays_of(scn, H);
        return BSB_OK;
    }
    for (int k = 0; k < n; k++) {
        DeviceState &d = ctx->devs[k];
        const int hk = r1[k] - r0[k], wk = c1[k] - c0[k];
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        if ((rc = ensure(ctx, d.d_frame, d.frame_cap, (size_t)hk * W + 1))) return rc;
        if (bloom) {
            if ((rc = ensure(ctx, d.d_mid, d.mid_cap, (size_t)hk * W + 1))) return rc;
            if ((rc = ensure(ctx, d.d_imgT, d.imgT_cap, (size_t)hk * W + 1))) return rc;
            if ((rc = ensure(ctx, d.d_rmid, d.rmid_cap, (size_t)wk * H + 1))) return rc;
            if ((rc = ensure(ctx, d.d_rimg, d.rimg_cap, (size_t)wk * H + 1))) return rc;
            if (want_float && (rc = ensure(ctx, d.d_band, d.band_cap, (size_t)wk * H + 1))) return rc;
            if (want_rgb8 && (rc = ensure(ctx, d.d_u8, d.u8_cap, (size_t)wk * H * 3 + 16))) return rc;
        } else if (want_rgb8) {
            if ((rc = ensure(ctx, d.d_u8, d.u8_cap, (size_t)hk * W * 3 + 16))) return rc;
        }
        if ((rc = trace_async(ctx, d, cam, scn, r0[k], r1[k], d.d_frame, d.ev[0], d.ev[1]))) return rc;
        d.last_rows = hk;
        launches += hk > 0 ? 2 : 0;
        if (bloom) {
            if ((rc = bloom_h_async(ctx, d, r, W, hk, d.d_frame, d.d_mid, d.d_imgT))) return rc;
            launches += hk > 0 ? 2 : 0;
        } else if (want_rgb8 && hk > 0) {
            BSB_CUDA(ctx, launch_srgb8(d.d_frame, d.d_u8, d.d_thr, (size_t)hk * W, d.stream));
            launches += 1;
        }
        BSB_CUDA(ctx, cudaEventRecord(d.ev[5], d.stream));
    }
    if (!bloom) {
        for (int k = 0; k < n; k++) {
            DeviceState &d = ctx->devs[k];
            BSB_CUDA(ctx, cudaSetDevice(d.dev));
            BSB_CUDA(ctx, cudaEventRecord(d.ev[2], d.stream));
            BSB_CUDA(ctx, cudaEventRecord(d.ev[3], d.stream));
            const int hk = r1[k] - r0[k];
            if (hk <= 0) continue;
            if (want_float) res.f32.push_back(Band{ &d, reinterpret_cast<const uint8_t *>(d.d_frame), 0, (size_t)W * 16, r0[k], hk });
            if (want_rgb8) res.u8.push_back(Band{ &d, d.d_u8, 0, (size_t)W * 3, r0[k], hk });
        }
        st->launches = launches; st->n_gpus = n; st->rays = rays_of(scn, H);
        return BSB_OK;
    }
    // the single collective of the path: all-to-all of the transposed tiles (grouped send/recv).
    // GPU i's tile, transposed, is [W][h_i]: the rows c0_j..c1_j of it are what GPU j needs -- contiguous.
    {
        int nrc = ctx->nccl.GroupStart();
        for (int i = 0; i < n && nrc == 0; i++) {
            const int hi = r1[i] - r0[i];
            if (hi <= 0) continue;
            DeviceState &di = ctx->devs[i];
            for (int j = 0; j < n && nrc == 0; j++) {
                const int wj = c1[j] - c0[j];
                if (wj <= 0) continue;
                DeviceState &dj = ctx->devs[j];
                const size_t cnt = (size_t)wj * hi * 4;                   // floats
                const size_t src_off = (size_t)c0[j] * hi;                // float4
                const size_t dst_off = (size_t)wj * r0[i];                // float4: pieces in tile order
                nrc = ctx->nccl.Send(di.d_mid + src_off, cnt, kNcclFloat, j, ctx->comms[i], di.stream);
                if (nrc == 0) nrc = ctx->nccl.Recv(dj.d_rmid + dst_off, cnt, kNcclFloat, i, ctx->comms[j], dj.stream);
                if (nrc == 0) nrc = ctx->nccl.Send(di.d_imgT + src_off, cnt, kNcclFloat, j, ctx->comms[i], di.stream);
                if (nrc == 0) nrc = ctx->nccl.Recv(dj.d_rimg + dst_off, cnt, kNcclFloat, i, ctx->comms[j], dj.stream);
            }
        }
        const int erc = ctx->nccl.GroupEnd();
        if (nrc == 0) nrc = erc;
        if (nrc != 0) return fail(ctx, BSB_ERR_NCCL, std::string("NCCL all-to-all: ") + ctx->nccl.GetErrorString(nrc));
    }
    for (int j = 0; j < n; j++) {
        DeviceState &d = ctx->devs[j];
        const int wj = c1[j] - c0[j];
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaEventRecord(d.ev[2], d.stream));
        const float4 *seg_mid[kMaxSegments], *seg_img[kMaxSegments];
        int seg_rows[kMaxSegments];
        for (int i = 0; i < n; i++) {
            seg_mid[i] = d.d_rmid + (size_t)wj * r0[i];
            seg_img[i] = d.d_rimg + (size_t)wj * r0[i];
            seg_rows[i] = r1[i] - r0[i];
        }
        if ((rc = bloom_v_async(ctx, d, scn->bloom_strength, r, H, wj, n, seg_mid, seg_img, seg_rows, want_float ? d.d_band : nullptr,
                                want_rgb8 ? d.d_u8 : nullptr)))
            return rc;
        launches += wj > 0 ? 1 : 0;
        BSB_CUDA(ctx, cudaEventRecord(d.ev[3], d.stream));
        if (wj <= 0) continue;
        if (want_float) res.f32.push_back(Band{ &d, reinterpret_cast<const uint8_t *>(d.d_band), (size_t)c0[j] * 16, (size_t)wj * 16, 0, H });
        if (want_rgb8) res.u8.push_back(Band{ &d, d.d_u8, (size_t)c0[j] * 3, (size_t)wj * 3, 0, H });
    }
    st->launches = launches; st->n_gpus = n; st->rays = rays_of(scn, H);
    return BSB_OK;
}

int render_full_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, bsb_stats *st, bool want_float, bool want_rgb8,
                       FullResult &res)
{
    if (!cam || !scn) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full: NULL argument");
    if (scn->width <= 0 || scn->height <= 0) return fail(ctx, BSB_ERR_INVALID, "resolution must be positive");
    if (ctx->devs.size() == 1) return render_full_single(ctx, cam, scn, st, want_float, want_rgb8, res);
    return render_full_multi(ctx, cam, scn, st, want_float, want_rgb8, res);
}

// wait for every GPU, read the events and counters
int collect_full_stats(bsb_ctx *ctx, bsb_stats *st)
{
    const int n = (int)ctx->devs.size();
    double trace_ms = 0, post_ms = 0, xchg_ms = 0, d2h_ms = 0;
    for (int k = 0; k < n; k++) {
        DeviceState &d = ctx->devs[k];
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0, a = 0, b = 0;
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        trace_ms = std::max(trace_ms, (double)ms);
        d.rows_per_ms = (d.last_rows >= 8 && ms > 1e-3f) ? d.last_rows / (double)ms : 0.0;
        BSB_CUDA(ctx, cudaEventElapsedTime(&a, d.ev[1], d.ev[5]));   // H^3 (+ transpose)
        BSB_CUDA(ctx, cudaEventElapsedTime(&b, d.ev[2], d.ev[3]));   // V^3 + combine (or the whole bloom on one GPU)
        post_ms = std::max(post_ms, (double)a + b);
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[5], d.ev[2]));  // all-to-all (includes waiting for the slowest tile)
        xchg_ms = std::max(xchg_ms, (double)ms);
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[3], d.ev[4]));
        d2h_ms = std::max(d2h_ms, (double)ms);
        fill_counter_stats(d, st);
    }
    st->trace_ms = trace_ms;
    st->gather_ms = n > 1 ? xchg_ms : 0.0;
    st->bloom_ms = post_ms;
    st->d2h_ms = d2h_ms;
    return BSB_OK;
}

int render_full_host(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, float *out_rgba, uint8_t *out_rgb8, bsb_stats *stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    bsb_stats st;
    std::memset(&st, 0, sizeof st);
    FullResult res;
    int rc = render_full_device(ctx, cam, scn, &st, out_rgba != nullptr, out_rgb8 != nullptr, res);
    if (rc) return rc;
    const int W = scn->width, H = scn->height;
    if (out_rgba && (rc = copy_bands_to_host(ctx, out_rgba, (size_t)W * 16, H, res.f32))) return rc;
    if (out_rgb8 && (rc = copy_bands_to_host(ctx, out_rgb8, (size_t)W * 3, H, res.u8))) return rc;
    for (DeviceState &d : ctx->devs) {
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaEventRecord(d.ev[4], d.stream));
    }
    rc = collect_full_stats(ctx, &st);
    if (rc) return rc;
    st.total_ms = ms_since(t0);
    if (stats) *stats = st;
    if (st.capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    return BSB_OK;
}

}  // namespace

extern "C" int bsb_render_full(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, float *out_rgba,
                               bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!out_rgba) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full: NULL output buffer");
    return render_full_host(ctx, cam, scn, out_rgba, nullptr, stats);
}

extern "C" int bsb_render_full_srgb8(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, uint8_t *out_rgb8,
                                     bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!out_rgb8) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full_srgb8: NULL output buffer");
    return render_full_host(ctx, cam, scn, nullptr, out_rgb8, stats);
}

// Both images of one render (the float frame and its sRGB8 map) -- either pointer may be NULL.
extern "C" int bsb_render_full_both(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, float *out_rgba,
                                    uint8_t *out_rgb8, bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!out_rgba && !out_rgb8) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full_both: no output buffer");
    return render_full_host(ctx, cam, scn, out_rgba, out_rgb8, stats);
}

// The same work with the frame left on the GPUs (no host copy, no host synchronisation): what a
// caller that keeps post-processing on the device uses, and what bench.py times as `value`.
extern "C" int bsb_render_full_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, int want_float, int want_rgb8)
{
    if (!ctx) return BSB_ERR_INVALID;
    bsb_stats st;
    std::memset(&st, 0, sizeof st);
    FullResult res;
    return render_full_device(ctx, cam, scn, &st, want_float != 0, want_rgb8 != 0, res);
}

// Blocks until everything queued on the ctx's GPUs has finished.  Returns BSB_ERR_STEPCAP if a ray of the
// LAST trace launch on any of them was stopped by the step cap -- the one thing the asynchronous entry
// points (bsb_render_device without stats, bsb_render_full_device) cannot report when they return.
extern "C" int bsb_synchronize(bsb_ctx *ctx)
{
    if (!ctx) return BSB_ERR_INVALID;
    unsigned long long capped = 0;
    for (DeviceState &d : ctx->devs) {
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        capped += d.h_ctr->capped;
    }
    if (capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    return BSB_OK;
}

// bloom + writeImg's map on device memory: out_rgba and / or out_rgb8 (either may be NULL)
extern "C" int bsb_bloom_to_device(bsb_ctx *ctx, double strength, int divider, int width, int height, const void *dev_in,
                                   void *dev_out_rgba, void *dev_out_rgb8)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!dev_in || (!dev_out_rgba && !dev_out_rgb8)) return fail(ctx, BSB_ERR_INVALID, "bsb_bloom_to_device: NULL argument");
    return bloom_async(ctx, ctx->devs[0], strength, divider, width, height, static_cast<const float4 *>(dev_in),
                       static_cast<float4 *>(dev_out_rgba), static_cast<uint8_t *>(dev_out_rgb8));
}

// rows x width_bytes from device memory (pitch dev_pitch) into a host frame (pitch host_pitch), ordered
// after the work queued on the ctx stream.  Page-locked host memory: asynchronous DMA; pageable: staged.
extern "C" int bsb_download_2d(bsb_ctx *ctx, void *host_dst, size_t host_pitch, const void *dev_src, size_t dev_pitch,
                               size_t width_bytes, int rows)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (rows < 0 || (rows > 0 && (!host_dst || !dev_src)) || width_bytes > host_pitch || dev_pitch != width_bytes)
        return fail(ctx, BSB_ERR_INVALID, "bsb_download_2d: bad argument (the device rows must be dense)");
    if (rows == 0 || width_bytes == 0) return BSB_OK;
    std::vector<Band> bands{ Band{ &ctx->devs[0], static_cast<const uint8_t *>(dev_src), 0, width_bytes, 0, rows } };
    return copy_bands_to_host(ctx, host_dst, host_pitch, rows, bands);
}

// ---- building blocks of the distributed bloom for one-process-per-GPU launchers (blackstar_b200/dist.py):
// the launcher owns the all-to-all between them.
extern "C" int bsb_bloom_h_device(bsb_ctx *ctx, int radius, int width, int rows, const void *dev_rows, void *dev_midT,
                                  void *dev_imgT)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (radius < 1 || width <= 0 || rows < 0 || (rows > 0 && (!dev_rows || !dev_midT)))
        return fail(ctx, BSB_ERR_INVALID, "bsb_bloom_h_device: bad argument");
    return bloom_h_async(ctx, ctx->devs[0], radius, width, rows, static_cast<const float4 *>(dev_rows),
                         static_cast<float4 *>(dev_midT), static_cast<float4 *>(dev_imgT));
}

extern "C" int bsb_bloom_v_device(bsb_ctx *ctx, double strength, int radius, int height, int cols, int nseg,
                                  const void *const *seg_midT, const void *const *seg_imgT, const int *seg_rows,
                                  void *dev_out_rgba, void *dev_out_rgb8)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (radius < 1 || height <= 0 || cols < 0 || !seg_midT || !seg_imgT || !seg_rows || (!dev_out_rgba && !dev_out_rgb8))
        return fail(ctx, BSB_ERR_INVALID, "bsb_bloom_v_device: bad argument");
    return bloom_v_async(ctx, ctx->devs[0], strength, radius, height, cols, nseg, reinterpret_cast<const float4 *const *>(seg_midT),
                         reinterpret_cast<const float4 *const *>(seg_imgT), seg_rows, static_cast<float4 *>(dev_out_rgba),
                         static_cast<uint8_t *>(dev_out_rgb8));
}

// ======================================================================== micro-benchmarks
extern "C" int bsb_measure_fp64_peak(bsb_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return BSB_ERR_INVALID;
    DeviceState &d = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    const int blocks = d.n_sms * 8, iters = 8192;
    BSB_CUDA(ctx, launch_dfma_peak(d.d_misc, blocks, iters, d.stream));  // warm-up at full length: clocks ramp from idle
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        BSB_CUDA(ctx, cudaEventRecord(d.ev[0], d.stream));
        BSB_CUDA(ctx, launch_dfma_peak(d.d_misc, blocks, iters, d.stream));
        BSB_CUDA(ctx, cudaEventRecord(d.ev[1], d.stream));
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0;
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    *tflops = best;
    return BSB_OK;
}

extern "C" int bsb_measure_hbm_copy(bsb_ctx *ctx, size_t bytes, int reps, double *gbs)
{
    if (!ctx || !gbs || bytes == 0 || reps < 1) return BSB_ERR_INVALID;
    DeviceState &d = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    void *a = nullptr, *b = nullptr;
    BSB_CUDA(ctx, cudaMalloc(&a, bytes));
    if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return fail(ctx, BSB_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemsetAsync(a, 1, bytes, d.stream);
    cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, d.stream);
    double best = 0;
    for (int rep = 0; rep < reps; rep++) {
        cudaEventRecord(d.ev[0], d.stream);
        cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, d.stream);
        cudaEventRecord(d.ev[1], d.stream);
        cudaStreamSynchronize(d.stream);
        float ms = 0;
        cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
        const double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    cudaFree(a);
    cudaFree(b);
    *gbs = best;
    return cudaGetLastError() == cudaSuccess ? BSB_OK : fail(ctx, BSB_ERR_CUDA, "hbm copy benchmark failed");
}

// max relative error of the kernel's |pos|^-5 primitive vs pow(q,-2.5), and the largest
// residual |1 - q y0^2| of the MUFU.RSQ64H seed, over n log-spaced q in [q_lo, q_hi]
extern "C" int bsb_selftest_rinv5(bsb_ctx *ctx, double q_lo, double q_hi, int n, double *max_rel_err,
                                  double *max_seed_residual)
{
    if (!ctx || !max_rel_err || !max_seed_residual || !(q_lo > 0) || !(q_hi > q_lo) || n < 1) return BSB_ERR_INVALID;
    DeviceState &d = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    BSB_CUDA(ctx, cudaMemsetAsync(d.d_misc, 0, 2 * sizeof(double), d.stream));
    BSB_CUDA(ctx, launch_rinv5_selftest(q_lo, q_hi, n, d.d_misc, d.stream));
    double out[2];
    BSB_CUDA(ctx, cudaMemcpyAsync(out, d.d_misc, sizeof out, cudaMemcpyDeviceToHost, d.stream));
    BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    *max_rel_err = out[0];
    *max_seed_residual = out[1];
    return BSB_OK;
}
