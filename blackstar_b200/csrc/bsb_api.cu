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
    uint8_t *h_stage = nullptr;                       // kStageSlots pinned chunks for copies into pageable memory
    cudaEvent_t stage_ev[kStageSlots] = {};
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
    int copy_threads = 4;
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

// Device -> caller-owned host memory, asynchronously ordered on d.stream up to the point where the
// bytes have left the device; returns after the last byte is in `dst` only for pageable targets.
//  * page-locked target (cudaHostAlloc / cudaHostRegister, e.g. GHC's pinned ForeignPtr registered by
//    the shim, or torch's pinned tensors): one cudaMemcpyAsync, the DMA engine writes it directly;
//  * pageable target (plain malloc, what mallocForeignPtrBytes hands to the FFI): a cudaMemcpy to
//    pageable memory is staged by the driver through one small buffer and runs at a fraction of the
//    link rate, so the library stages it itself -- the frame is cut into 8 MB chunks that go through
//    a ring of pinned buffers while a few parked host threads move finished chunks into the caller's
//    buffer, so the PCIe transfer and the host copy overlap.
int copy_to_host(bsb_ctx *ctx, DeviceState &d, void *dst, const void *src_dev, size_t bytes)
{
    if (bytes == 0) return BSB_OK;
    BSB_CUDA(ctx, cudaSetDevice(d.dev));
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, dst);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    const bool pinned = pe == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    if (pinned || bytes < kStageChunk / 4 || ctx->copy_threads <= 0) {
        BSB_CUDA(ctx, cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, d.stream));
        return BSB_OK;
    }
    if (!d.h_stage) {
        BSB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&d.h_stage), kStageSlots * kStageChunk, cudaHostAllocDefault));
        for (auto &e : d.stage_ev) BSB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!ctx->pool || ctx->pool->size() != ctx->copy_threads) ctx->pool.reset(new CopyPool(ctx->copy_threads));
    const int nchunks = (int)((bytes + kStageChunk - 1) / kStageChunk);
    const int nt = ctx->pool->size();
    std::vector<std::atomic<int>> copied(nchunks);
    for (auto &c : copied) c.store(0, std::memory_order_relaxed);
    std::atomic<int> enqueued{ 0 };
    std::atomic<int> failed{ 0 };
    uint8_t *out = static_cast<uint8_t *>(dst);
    ctx->pool->begin([&, nt](int k) {
        cudaSetDevice(d.dev);
        for (int i = 0; i < nchunks; i++) {
            while (enqueued.load(std::memory_order_acquire) <= i) {
                if (failed.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            if (cudaEventSynchronize(d.stage_ev[i % kStageSlots]) != cudaSuccess) failed.store(1);
            const size_t off = (size_t)i * kStageChunk;
            const size_t len = std::min(kStageChunk, bytes - off);
            const size_t a = len * k / nt, b = len * (k + 1) / nt;   // my slice of the chunk
            std::memcpy(out + off + a, d.h_stage + (size_t)(i % kStageSlots) * kStageChunk + a, b - a);
            copied[i].fetch_add(1, std::memory_order_release);
        }
    });
    cudaError_t err = cudaSuccess;
    for (int i = 0; i < nchunks && err == cudaSuccess; i++) {
        if (i >= kStageSlots)   // the slot is free once every thread has copied its slice of chunk i - slots
            while (copied[i - kStageSlots].load(std::memory_order_acquire) < nt) std::this_thread::yield();
        const size_t off = (size_t)i * kStageChunk;
        const size_t len = std::min(kStageChunk, bytes - off);
        err = cudaMemcpyAsync(d.h_stage + (size_t)(i % kStageSlots) * kStageChunk, static_cast<const uint8_t *>(src_dev) + off, len,
                              cudaMemcpyDeviceToHost, d.stream);
        if (err == cudaSuccess) err = cudaEventRecord(d.stage_ev[i % kStageSlots], d.stream);
        if (err == cudaSuccess) enqueued.store(i + 1, std::memory_order_release);
    }
    if (err != cudaSuccess) failed.store(1);
    ctx->pool->wait();
    if (err != cudaSuccess) return fail(ctx, BSB_ERR_CUDA, std::string("staged device-to-host copy: ") + cudaGetErrorString(err));
    if (failed.load()) return fail(ctx, BSB_ERR_CUDA, "staged device-to-host copy failed");
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
    const char *v = std::getenv("BSB_TRACE_VARIANT");
    if (v) ctx->trace_variant = std::atoi(v);
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
    if (n_gpus == 0) n_gpus = count;
    std::vector<int> devs(n_gpus);
    for (int k = 0; k < n_gpus; k++) devs[k] = k;
    return bsb_create_on(devs.data(), n_gpus);
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
        if (d.h_stage) cudaFreeHost(d.h_stage);
        for (auto &e : d.stage_ev) if (e) cudaEventDestroy(e);
        if (d.d_aux) cudaFree(d.d_aux);
        if (d.d_u8) cudaFree(d.d_u8);
        if (d.d_vx) cudaFree(d.d_vx);
        if (d.d_vy) cudaFree(d.d_vy);
        if (d.d_misc) cudaFree(d.d_misc);
        for (auto &ev : d.ev) if (ev) cudaEventDestroy(ev);
        if (d.own_stream) cudaStreamDestroy(d.own_stream);
    }
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
    return fail(ctx, BSB_ERR_INVALID, std::string("unknown option ") + key);
}

// ======================================================================== star map
extern "C" int bsb_set_stars(bsb_ctx *ctx, const bsb_star *stars, size_t n)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (n > 0 && !stars) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars: NULL star list");
    if (n > (size_t)1 << 28) return fail(ctx, BSB_ERR_INVALID, "bsb_set_stars: too many stars");
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

// render on all GPUs + gather on GPU 0 + bloom; leaves the frame in devs[0].d_frame.
// Records ev[0..4] on GPU 0: begin, own tile traced, gathered, bloomed.
int render_full_device(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, bsb_stats *st, bool want_float,
                       bool want_rgb8)
{
    if (!cam || !scn) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full: NULL argument");
    if (scn->width <= 0 || scn->height <= 0) return fail(ctx, BSB_ERR_INVALID, "resolution must be positive");
    const int n = (int)ctx->devs.size();
    const int W = scn->width, H = scn->height;
    const size_t npix = (size_t)W * H;
    DeviceState &d0 = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d0.dev));
    int rc = ensure(ctx, d0.d_frame, d0.frame_cap, npix);
    if (rc) return rc;
    // Row tiles.  First frame: equal shares.  Afterwards: proportional to the rate every GPU
    // achieved on its previous tile (rows per millisecond), so that tiles with fewer RK4 steps per
    // ray (the hole, the far sky) get more rows and all GPUs finish together.
    std::vector<int> r0(n), r1(n);
    {
        double total_rate = 0;
        bool known = n > 1;
        for (int k = 0; k < n; k++) { known = known && ctx->devs[k].rows_per_ms > 0; total_rate += ctx->devs[k].rows_per_ms; }
        double acc = 0;
        for (int k = 0; k < n; k++) {
            r0[k] = k == 0 ? 0 : r1[k - 1];
            if (known) {
                acc += ctx->devs[k].rows_per_ms / total_rate;
                r1[k] = k == n - 1 ? H : std::max(r0[k], std::min(H, (int)(acc * H + 0.5)));
            } else {
                r1[k] = (int)((long long)H * (k + 1) / n);
            }
        }
    }
    for (int k = 1; k < n; k++) {
        DeviceState &d = ctx->devs[k];
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        rc = ensure(ctx, d.d_frame, d.frame_cap, (size_t)(r1[k] - r0[k]) * W + 1);
        if (rc) return rc;
    }
    // row tiles: GPU k renders rows [H k/n, H (k+1)/n) of the final image
    for (int k = 0; k < n; k++) {
        DeviceState &d = ctx->devs[k];
        float4 *dst = k == 0 ? d0.d_frame : d.d_frame;
        rc = trace_async(ctx, d, cam, scn, r0[k], r1[k], dst, d.ev[0], d.ev[1]);
        if (rc) return rc;
        d.last_rows = r1[k] - r0[k];
    }
    int launches = 0;  // ray tables + trace on every GPU that has rows
    for (int k = 0; k < n; k++) launches += r1[k] > r0[k] ? 2 : 0;
    // the single collective of the path: gather the tiles on GPU 0 (grouped send/recv)
    if (n > 1) {
        int nrc = ctx->nccl.GroupStart();
        for (int k = 1; k < n && nrc == 0; k++) {
            const size_t cnt = (size_t)(r1[k] - r0[k]) * W * 4;
            if (cnt == 0) continue;
            nrc = ctx->nccl.Recv(d0.d_frame + (size_t)r0[k] * W, cnt, kNcclFloat, k, ctx->comms[0], d0.stream);
            if (nrc == 0) nrc = ctx->nccl.Send(ctx->devs[k].d_frame, cnt, kNcclFloat, 0, ctx->comms[k], ctx->devs[k].stream);
        }
        const int erc = ctx->nccl.GroupEnd();
        if (nrc == 0) nrc = erc;
        if (nrc != 0) return fail(ctx, BSB_ERR_NCCL, std::string("NCCL gather: ") + ctx->nccl.GetErrorString(nrc));
        BSB_CUDA(ctx, cudaSetDevice(d0.dev));
    }
    BSB_CUDA(ctx, cudaSetDevice(d0.dev));
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[2], d0.stream));
    // app/Main.hs:113: bloom only if bloomStrength /= 0
    if (want_rgb8) {
        rc = ensure(ctx, d0.d_u8, d0.u8_cap, npix * 3 + 16);
        if (rc) return rc;
    }
    if (scn->bloom_strength != 0) {
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
    if (st) { st->launches = launches; st->n_gpus = n; st->rays = rays_of(scn, H); }
    return BSB_OK;
}

int collect_full_stats(bsb_ctx *ctx, bsb_stats *st)
{
    const int n = (int)ctx->devs.size();
    double trace_ms = 0;
    for (int k = 0; k < n; k++) {
        DeviceState &d = ctx->devs[k];
        BSB_CUDA(ctx, cudaSetDevice(d.dev));
        BSB_CUDA(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0;
        BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        if (ms > trace_ms) trace_ms = ms;
        d.rows_per_ms = (d.last_rows >= 8 && ms > 1e-3f) ? d.last_rows / (double)ms : 0.0;
        fill_counter_stats(d, st);
    }
    DeviceState &d0 = ctx->devs[0];
    BSB_CUDA(ctx, cudaSetDevice(d0.dev));
    float ms = 0;
    st->trace_ms = trace_ms;
    BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[1], d0.ev[2]));
    st->gather_ms = n > 1 ? ms : 0.0;
    BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[2], d0.ev[3]));
    st->bloom_ms = ms;
    BSB_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[3], d0.ev[4]));
    st->d2h_ms = ms;
    return BSB_OK;
}

}  // namespace

extern "C" int bsb_render_full(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, float *out_rgba,
                               bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!out_rgba) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full: NULL output buffer");
    const auto t0 = std::chrono::steady_clock::now();
    bsb_stats st;
    std::memset(&st, 0, sizeof st);
    int rc = render_full_device(ctx, cam, scn, &st, true, false);
    if (rc) return rc;
    DeviceState &d0 = ctx->devs[0];
    const size_t npix = (size_t)scn->width * scn->height;
    rc = copy_to_host(ctx, d0, out_rgba, d0.d_frame, npix * sizeof(float4));
    if (rc) return rc;
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[4], d0.stream));
    rc = collect_full_stats(ctx, &st);
    if (rc) return rc;
    st.total_ms = ms_since(t0);
    if (stats) *stats = st;
    if (st.capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    return BSB_OK;
}

extern "C" int bsb_render_full_srgb8(bsb_ctx *ctx, const bsb_camera *cam, const bsb_scene *scn, uint8_t *out_rgb8,
                                     bsb_stats *stats)
{
    if (!ctx) return BSB_ERR_INVALID;
    if (!out_rgb8) return fail(ctx, BSB_ERR_INVALID, "bsb_render_full_srgb8: NULL output buffer");
    const auto t0 = std::chrono::steady_clock::now();
    bsb_stats st;
    std::memset(&st, 0, sizeof st);
    int rc = render_full_device(ctx, cam, scn, &st, false, true);
    if (rc) return rc;
    DeviceState &d0 = ctx->devs[0];
    const size_t npix = (size_t)scn->width * scn->height;
    rc = copy_to_host(ctx, d0, out_rgb8, d0.d_u8, npix * 3);  // bloom_ms includes the fused sRGB map
    if (rc) return rc;
    BSB_CUDA(ctx, cudaEventRecord(d0.ev[4], d0.stream));
    rc = collect_full_stats(ctx, &st);
    if (rc) return rc;
    st.total_ms = ms_since(t0);
    if (stats) *stats = st;
    if (st.capped) return fail(ctx, BSB_ERR_STEPCAP, "a ray reached the step cap (the reference would not terminate)");
    return BSB_OK;
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
