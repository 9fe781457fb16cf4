// libplonkish_cuda.so — the C ABI declared in include/plonkish_cuda.h.
//
// Host side of the drop-in for variable_base_msm
// (/root/reference/plonkish_backend/src/util/arithmetic/msm.rs:84-115): per-device
// contexts (stream, scratch arena, resident bases), host<->device staging, point
// chunking, the multi-GPU point sharding of msm.rs:101-114, and two measurement
// helpers.  No CPU fallback anywhere: every compute entry point needs a CUDA device.
#include "../../include/plonkish_cuda.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pk {
static std::atomic<unsigned long long> g_launches{0};
}
#define PK_COUNT_LAUNCH() (pk::g_launches.fetch_add(1, std::memory_order_relaxed))

#include "msm_kernels.cuh"
#include "poly_kernels.cuh"
#include "sumcheck_kernels.cuh"
#include "lookup_kernels.cuh"
#include "dpfq.cuh"

using namespace pk;

// ------------------------------------------------------------------ error state
static thread_local std::string t_last_error;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t err__ = (expr);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return fail(PLONKISH_CUDA_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

// -------------------------------------------------------------------- contexts
static const size_t MAX_POINTS_PER_LAUNCH = (size_t)1 << 26;  // scatter entries keep 26 index bits

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

#define PK_LANES 7  // side streams for small concurrent MSMs (many / open)
struct Ctx {
    int dev = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of scalar chunks, overlapped with the compute stream
    cudaEvent_t chunk_ready[16] = {};
    cudaEvent_t last_done = nullptr;  // end of the last enqueued MSM: orders arena reuse across streams
    bool has_last = false;
    DeviceBuffer arena, scalars, scalars2, bases_tmp, partials, batch_out, open_buf;
    // per-proof temporaries that would otherwise come out of the pool in multi-GB pieces on every call (pool growth
    // showed up as 40-80 ms swings between repetitions of one proof): `tmp` serves the eq table expansion, the division
    // scratch and the grand-product scratch (one at a time: they run under mu, in stream order); `sc_block` backs ONE
    // live sum-check state, a second concurrent state falls back to the pool
    DeviceBuffer tmp, sc_block;
    bool sc_block_busy = false;
    cudaEvent_t buf_free[2] = {};         // batch path: scalar buffer b may be overwritten again
    cudaEvent_t many_start = nullptr;     // many path: side lanes start after this point of the main stream
    struct Lane {                         // extra streams with their own scratch: small independent MSMs run side by side
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DeviceBuffer arena, scalars;
        void *d_res = nullptr;            // 128-byte projective result slot
    } lanes[PK_LANES + 1];                // lanes[PK_LANES]: the "big lane" — every second LARGE MSM of a batch / many call
                                          // runs there with its own full-size workspace, so that its decompose + sort (HBM and
                                          // latency bound) fill the GPU while the previous MSM's accumulate drains and its
                                          // single-wave reduce, item levels and finalize run (blocks of the older kernel are
                                          // dispatched first, so the two accumulates do not fight)
    void *d_out = nullptr;  // [0,64) affine out, [192,256) synth step point, [256,384) running projective sum, [384,448) fixed base
    void *h_out = nullptr;  // pinned mirror of the affine result
    std::mutex mu;
};

// Device memory behind a handle.  Handles are reference counted: the map holds one reference, every call in
// flight (through its BasesView / ScalarsEntry copy) and every sum-check state another, so a *_release racing
// with an enqueued MSM — Rust's Drop on one rayon worker while another still commits — frees the memory only
// after the last user is gone.
struct DevBlock {
    int dev = 0;
    void *ptr = nullptr;
    bool owns = true;        // false: borrows the caller's device memory
    Ctx *pool_ctx = nullptr; // != nullptr: carved from the stream-ordered pool of that context
    size_t bytes = 0;
    ~DevBlock() {
        if (!ptr || !owns) return;
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(dev);
        if (pool_ctx) {
            if (cudaFreeAsync(ptr, pool_ctx->stream) != cudaSuccess) cudaGetLastError();  // ordered after everything enqueued on the context's stream
        } else {
            cudaDeviceSynchronize();
            cudaFree(ptr);
        }
        cudaSetDevice(cur);
    }
};
typedef std::shared_ptr<DevBlock> BlockRef;
static BlockRef make_block(int dev, void *ptr, bool owns, Ctx *pool_ctx, size_t bytes) {
    BlockRef b = std::make_shared<DevBlock>();
    b->dev = dev; b->ptr = ptr; b->owns = owns; b->pool_ctx = pool_ctx; b->bytes = bytes;
    return b;
}

struct BasesEntry {
    int n_shards = 1;                 // 1: whole slice on `dev`; G: shard g on device g
    int dev = 0;
    size_t n = 0;
    std::vector<void *> d_ptr;        // per shard: plain bases, or the table of window multiples (row 0 = the bases)
    std::vector<BlockRef> keep;       // per shard: owner of d_ptr (null for an empty shard)
    std::vector<size_t> shard_n;
    std::vector<uint32_t> table_c;    // per shard: window bits of the table, 0 = plain bases (no table)
    void add_shard(int device, void *d, bool owns, size_t cnt, uint32_t tc) {
        const size_t bytes = cnt * PLONKISH_CUDA_AFFINE_BYTES * (tc ? pk_windows_for(tc) : 1);
        d_ptr.push_back(d); shard_n.push_back(cnt); table_c.push_back(tc);
        keep.push_back(d ? make_block(device, d, owns, nullptr, bytes) : BlockRef());
    }
    size_t device_bytes() const {
        size_t t = 0;
        for (const BlockRef &b : keep) if (b && b->owns) t += b->bytes;
        return t;
    }
};

// What an MSM launch sequence reads its bases from.
struct BasesView {
    const void *ptr = nullptr;
    uint32_t table_c = 0;   // 0: plain affine array; else table T[w*stride + i]
    size_t stride = 0;
    BlockRef keep;          // keeps the memory alive for the duration of the call
};

static std::mutex g_mu;
static std::vector<Ctx *> g_ctx;
static std::map<uint64_t, BasesEntry> g_bases;
static uint64_t g_next_handle = 1;

// Borrowed-slice cache (plonkish_cuda_bases_cached): host address -> registered handle, validated by content.
struct CachedBases {
    uint64_t handle = 0;
    size_t n = 0, bytes = 0;
    int dev = 0, n_gpus = 1;  // n_gpus > 1: sharded over devices 0..n_gpus-1 (dev unused)
    unsigned long long tick = 0;
    std::vector<std::pair<size_t, uint64_t>> samples;  // (position, FNV-1a of the 64-byte point)
};
static std::mutex g_cache_mu;
static std::map<uintptr_t, CachedBases> g_bases_cache;
static unsigned long long g_cache_tick = 0;
static size_t g_cache_limit = 0;  // 0: half of the device's memory

static int grow(DeviceBuffer &b, size_t bytes) {
    if (bytes <= b.bytes) return 0;
    if (b.ptr) {
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaFree(b.ptr));
        b.ptr = nullptr;
        b.bytes = 0;
    }
    CUDA_TRY(cudaMalloc(&b.ptr, bytes));
    b.bytes = bytes;
    return 0;
}

// Resident polynomials come and go once per proof: they are carved from the device's stream-ordered
// pool (kept warm, see create_ctx_locked) so that neither side synchronises the device the way
// cudaMalloc / cudaFree do.  Every entry point ends with its streams joined and idle, so a pointer
// obtained here is usable on all of the context's streams.
static int pool_alloc(Ctx *c, void **p, size_t bytes) {
    CUDA_TRY(cudaMallocAsync(p, bytes, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
static void pool_free(Ctx *c, void *p) {
    if (p && cudaFreeAsync(p, c->stream) != cudaSuccess) cudaGetLastError();
}
// Returns a pooled buffer on every exit path unless release() hands it on.
struct PoolGuard {
    Ctx *c;
    void *p;
    ~PoolGuard() { pool_free(c, p); }
    void *release() { void *q = p; p = nullptr; return q; }
};

// Contexts are created on first use of a device, so a one-process-per-GPU rank only
// ever touches its own GPU.  Caller must not hold g_mu.
static int create_ctx_fill(Ctx *c, int device);
static int create_ctx_locked(int device) {
    Ctx *c = new Ctx();
    const int rc = create_ctx_fill(c, device);
    if (rc != PLONKISH_CUDA_OK) { delete c; return rc; }  // (streams / events created before the failing call die with the process' context)
    g_ctx[device] = c;
    return PLONKISH_CUDA_OK;
}
static int create_ctx_fill(Ctx *c, int device) {
    c->dev = device;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(PLONKISH_CUDA_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    c->sm_count = prop.multiProcessorCount;
    {
        cudaMemPool_t pool;
        unsigned long long keep_all = ~0ull;  // freed resident polynomials stay in the pool for the next proof
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c->chunk_ready[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c->buf_free[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->many_start, cudaEventDisableTiming));
    for (auto &ln : c->lanes) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&ln.d_res, 256));
    }
    CUDA_TRY(cudaMalloc(&c->d_out, 512));  // see the slot map at Ctx::d_out
    CUDA_TRY(cudaMallocHost(&c->h_out, 256));
    return PLONKISH_CUDA_OK;
}

static Ctx *ctx_for(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (device < 0 || (size_t)device >= g_ctx.size()) return nullptr;
    if (!g_ctx[device] && create_ctx_locked(device) != PLONKISH_CUDA_OK) return nullptr;
    return g_ctx[device];
}


// ------------------------------------------------------------- pageable uploads
// The reference hands variable_base_msm borrowed slices (`poly.evals()`, kzg.rs:255): a Rust Vec<Fr> is
// pageable memory.  cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread, one
// thread, a few GB/s, and nothing overlaps it.  Uploads from unregistered host memory therefore go through the
// library's own pinned ring: a few copier threads move 4 MiB pieces into pinned slots and enqueue one
// cudaMemcpyAsync per piece on the destination stream, so the host-side copy of the next pieces runs while the
// DMA engine moves the current one (about the bandwidth of the pinned path once 4+ threads copy).  Pinned /
// registered sources (cudaPointerGetAttributes) keep the direct single cudaMemcpyAsync.
extern "C" void plonkish_cuda_host_copy(void *dst, const void *src, size_t len, int stream);  // host_copy.cpp

class Stager {
public:
    static const size_t SLOT = (size_t)4 << 20;
    // One ring per device in use.  A ring runs one upload at a time; the single-process multi-GPU entry uploads to all
    // its devices at once (one host thread each), so it asks for `shards` rings that split the copier threads between
    // them (set_shards) — with one ring for all, device B's small first chunk waited behind device A's large one.
    static Stager &get(int dev = 0) {
        static std::mutex mu;
        static Stager *rings[PK_MAX_STAGERS] = {nullptr};  // leaked on purpose: worker threads outlive static destruction
        const int shards = shards_().load(std::memory_order_relaxed);
        const int k = shards > 1 ? (dev < 0 ? 0 : dev) % (shards < PK_MAX_STAGERS ? shards : PK_MAX_STAGERS) : 0;
        std::lock_guard<std::mutex> lk(mu);
        if (!rings[k]) rings[k] = new Stager(shards > 1 ? shards : 1);
        return *rings[k];
    }
    static void set_shards(int n) {
        if (n > shards_().load(std::memory_order_relaxed)) shards_().store(n, std::memory_order_relaxed);
    }
    // Copies [src, src + bytes) to device memory dst on `stream`; returns when every piece is enqueued
    // (the source may be reused), not when the DMA is done.
    cudaError_t run(int dev, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
        std::lock_guard<std::mutex> job_lock(job_mu_);
        if (!ensure_ready(dev)) return cudaErrorMemoryAllocation;
        copy_ns_.store(0, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            src_ = (const char *)src; dst_ = (char *)dst; bytes_ = bytes; stream_ = stream; dev_ = dev;
            // pieces shrink with the upload so that every copier gets a few and the first DMA starts early (a 4 MiB
            // upload in 4 MiB pieces would be one thread's memcpy followed by one DMA)
            size_t piece = (bytes / (4 * (size_t)nthreads_) + 0xffff) & ~(size_t)0xffff;
            piece_ = piece < ((size_t)256 << 10) ? ((size_t)256 << 10) : (piece > SLOT ? SLOT : piece);
            npieces_ = (bytes + piece_ - 1) / piece_; next_ = 0; done_ = 0; err_ = cudaSuccess;
            ++job_id_;
        }
        cv_work_.notify_all();
        work(dev);  // the caller copies too
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return done_ == npieces_; });
        base_ += npieces_;
        npieces_ = 0;
        // what the ring sustains on this host right now (several ranks share its memory bandwidth): the host-scalar MSM
        // cuts its points into more, smaller chunks when uploads are slow (enqueue_host_msm)
        // The rate is that of the host-side copies alone (bytes x copiers / time spent copying): it is the host's memory
        // bandwidth that ranks share.  Wall time would also count the waits for ring slots whose DMA sits behind a stream
        // dependency (the batch entry's second buffer waits for the MSM before last) and read as a slow host.
        const double copy_sec = (double)copy_ns_.load(std::memory_order_relaxed) * 1e-9;
        if (copy_sec > 0 && bytes >= ((size_t)32 << 20)) {
            const size_t np = (bytes + piece_ - 1) / piece_;
            const double par = (double)(np < (size_t)nthreads_ ? np : (size_t)nthreads_);
            double gbps = (double)bytes * par / copy_sec * 1e-9;
            if (gbps > 50.0) gbps = 50.0;  // the DMA engine's share of PCIe is the limit above that
            const double old = rate_gbps_.load(std::memory_order_relaxed);
            rate_gbps_.store(old > 0 ? 0.5 * old + 0.5 * gbps : gbps, std::memory_order_relaxed);
        }
        return err_;
    }
    int threads() const { return nthreads_; }
    double rate_gbps() const { return rate_gbps_.load(std::memory_order_relaxed); }  // 0 until the first large upload

private:
    static const int PK_MAX_STAGERS = 16;
    static std::atomic<int> &shards_() { static std::atomic<int> v{1}; return v; }
    explicit Stager(int shards) {
        unsigned hw = std::thread::hardware_concurrency();
        int t = hw >= 32 ? 8 : hw >= 16 ? 6 : hw >= 8 ? 4 : 2;
        if (shards > 1) {  // the rings of one process share the host's cores like the ranks of a torchrun job do
            const int per = (int)(hw / (unsigned)shards);
            t = per < 2 ? 2 : (per < t ? per : t);
        }
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU (torchrun): the ranks share the host's cores
            const long w = atol(e);
            if (w > 1) {
                const int per = (int)(hw / (unsigned long)w);
                t = per < 2 ? 2 : (per < t ? per : t);
            }
        }
        if (const char *e = getenv("PLONKISH_CUDA_COPY_THREADS")) {
            const long v = atol(e);
            if (v >= 1 && v <= 64) t = (int)v;
        }
        nthreads_ = t;
        nslots_ = 2 * t + 2;
        // host_copy.cpp: streaming stores into the ring (measured end to end from pageable memory, 2^24-point MSM: 38.5 ->
        // 38.0 ms with one rank, 45.2 -> 43.3 ms with two ranks sharing the host); PLONKISH_CUDA_STAGE_NT=0 = plain memcpy
        if (const char *e = getenv("PLONKISH_CUDA_STAGE_NT")) stream_stores_ = e[0] != '0';
    }
    bool ensure_ready(int dev) {
        if (slots_.empty()) {
            slots_.resize(nslots_, nullptr);
            for (int i = 0; i < nslots_; ++i) {
                if (cudaHostAlloc(&slots_[i], SLOT, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); slots_.clear(); return false; }
            }
            slot_done_.assign(nslots_, 0);
            slot_ev_.assign(nslots_, std::vector<cudaEvent_t>());
            slot_dev_.assign(nslots_, -1);
            for (int i = 0; i < nthreads_ - 1; ++i) std::thread([this] { worker(); }).detach();
        }
        for (int i = 0; i < nslots_; ++i) {
            if ((int)slot_ev_[i].size() <= dev) slot_ev_[i].resize(dev + 1, nullptr);
            if (!slot_ev_[i][dev]) {
                cudaSetDevice(dev);
                if (cudaEventCreateWithFlags(&slot_ev_[i][dev], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
            }
        }
        return true;
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            int dev;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return job_id_ != seen && next_ < npieces_; });
                seen = job_id_;
                dev = dev_;
            }
            work(dev);
        }
    }
    void work(int dev) {
        cudaSetDevice(dev);
        for (;;) {
            size_t i;
            unsigned long long g;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (next_ >= npieces_) return;
                i = next_++;
                g = base_ + i;
            }
            const int s = (int)(g % (unsigned long long)nslots_);
            {   // the slot's previous piece (global number g - nslots) must have recorded its event, then finished its DMA
                std::unique_lock<std::mutex> lk(mu_);
                cv_slot_.wait(lk, [&] { return g < (unsigned long long)nslots_ || slot_done_[s] == g - nslots_ + 1; });
            }
            cudaError_t e = cudaSuccess;
            if (slot_dev_[s] >= 0) e = cudaEventSynchronize(slot_ev_[s][slot_dev_[s]]);
            const size_t off = i * piece_;
            const size_t len = bytes_ - off < piece_ ? bytes_ - off : piece_;
            const auto c0 = std::chrono::steady_clock::now();
            plonkish_cuda_host_copy(slots_[s], src_ + off, len, stream_stores_ ? 1 : 0);
            copy_ns_.fetch_add((unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - c0).count(),
                               std::memory_order_relaxed);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dst_ + off, slots_[s], len, cudaMemcpyHostToDevice, stream_);
            if (e == cudaSuccess) e = cudaEventRecord(slot_ev_[s][dev], stream_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                slot_dev_[s] = dev;
                slot_done_[s] = g + 1;
                if (e != cudaSuccess && err_ == cudaSuccess) err_ = e;
                ++done_;
            }
            cv_slot_.notify_all();
            cv_done_.notify_all();
        }
    }
    int nthreads_ = 2, nslots_ = 6;
    bool stream_stores_ = true;
    std::atomic<double> rate_gbps_{0.0};
    std::atomic<unsigned long long> copy_ns_{0};
    std::vector<void *> slots_;
    std::vector<std::vector<cudaEvent_t>> slot_ev_;  // [slot][device]
    std::vector<int> slot_dev_;                      // device whose event the slot's last piece recorded
    std::vector<unsigned long long> slot_done_;      // global number + 1 of the last piece recorded on the slot
    std::mutex job_mu_, mu_;
    std::condition_variable cv_work_, cv_done_, cv_slot_;
    const char *src_ = nullptr;
    char *dst_ = nullptr;
    size_t bytes_ = 0, piece_ = SLOT, npieces_ = 0, next_ = 0, done_ = 0;
    unsigned long long base_ = 0, job_id_ = 0;
    cudaStream_t stream_ = nullptr;
    int dev_ = 0;
    cudaError_t err_ = cudaSuccess;
};

static std::atomic<unsigned long long> g_staged_bytes{0};
// Host -> device copy of caller memory on `stream` (current device = c->dev).  Returns once the source may be reused.
// tuning override: PLONKISH_CUDA_STAGE_MIN_MB (uploads of at least this many MiB from pageable memory go through the ring)
static const size_t STAGE_MIN_BYTES = [] {
    const char *e = getenv("PLONKISH_CUDA_STAGE_MIN_MB");
    const long v = e ? atol(e) : 0;
    return (size_t)(v >= 1 && v <= 4096 ? v : 4) << 20;
}();
// Whether an upload of `bytes` from src goes through the staging ring: unregistered (pageable) host memory only.
static bool is_staged_source(const void *src, size_t bytes) {
    static const bool staging = [] {
        const char *e = getenv("PLONKISH_CUDA_STAGING");
        return !(e && e[0] == '0');
    }();
    // From 4 MiB on the ring beats the driver's own pageable path (measured end to end with streaming stores, MSM of
    // 2^17 / 2^18 / 2^19 / 2^20 / 2^21 points = 4 .. 64 MiB: staged 1.06 / 1.48 / 2.38 / 3.74 / 6.12 ms, with the ring only
    // from 32 MiB 1.13 / 1.66 / 2.90 / 4.59 / 6.71 ms; before the streaming stores the driver's path won below 32 MiB).
    if (!staging || bytes < STAGE_MIN_BYTES) return false;
    cudaPointerAttributes attr;
    const cudaError_t q = cudaPointerGetAttributes(&attr, src);
    if (q != cudaSuccess) cudaGetLastError();
    return q == cudaSuccess && attr.type == cudaMemoryTypeUnregistered;
}
static cudaError_t upload(Ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (is_staged_source(src, bytes)) {
        g_staged_bytes.fetch_add(bytes, std::memory_order_relaxed);
        return Stager::get(c->dev).run(c->dev, dst, src, bytes, stream);
    }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
}
extern "C" uint64_t plonkish_cuda_staged_bytes(void) { return g_staged_bytes.load(); }
extern "C" double plonkish_cuda_staging_rate_gbps(void) { return Stager::get().rate_gbps(); }

extern "C" int plonkish_cuda_init(int n_devices) {
    std::lock_guard<std::mutex> lk(g_mu);
    int visible = 0;
    cudaError_t err = cudaGetDeviceCount(&visible);
    if (err != cudaSuccess || visible == 0)
        return fail(PLONKISH_CUDA_E_NO_DEVICE, "no CUDA device: %s", err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    if (n_devices <= 0 || n_devices > visible) n_devices = visible;
    if ((int)g_ctx.size() < n_devices) g_ctx.resize(n_devices, nullptr);
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_device_count(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    return (int)g_ctx.size();
}

static void release_all_scalars();
static void nccl_shutdown();
extern "C" void plonkish_cuda_shutdown(void) {
    nccl_shutdown();
    std::lock_guard<std::mutex> lk(g_mu);
    g_bases.clear();  // the blocks free themselves (DevBlock)
    g_bases_cache.clear();
    release_all_scalars();
    for (Ctx *c : g_ctx) {
        if (!c) continue;
        cudaSetDevice(c->dev);
        cudaDeviceSynchronize();
        cudaFree(c->arena.ptr); cudaFree(c->scalars.ptr); cudaFree(c->scalars2.ptr); cudaFree(c->bases_tmp.ptr); cudaFree(c->partials.ptr);
        cudaFree(c->batch_out.ptr); cudaFree(c->open_buf.ptr); cudaFree(c->tmp.ptr); cudaFree(c->sc_block.ptr);
        for (int i = 0; i < 2; ++i) cudaEventDestroy(c->buf_free[i]);
        cudaEventDestroy(c->many_start);
        for (auto &ln : c->lanes) {
            cudaFree(ln.arena.ptr); cudaFree(ln.scalars.ptr); cudaFree(ln.d_res);
            cudaEventDestroy(ln.done); cudaStreamDestroy(ln.stream);
        }
        cudaFree(c->d_out); cudaFreeHost(c->h_out);
        cudaEventDestroy(c->last_done);
        for (int i = 0; i < 16; ++i) cudaEventDestroy(c->chunk_ready[i]);
        cudaStreamDestroy(c->copy_stream);
        cudaStreamDestroy(c->stream);
        delete c;
    }
    g_ctx.clear();
}

extern "C" const char *plonkish_cuda_last_error(void) { return t_last_error.c_str(); }
extern "C" uint64_t plonkish_cuda_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------ base cache
static bool precompute_enabled() {
    const char *e = getenv("PLONKISH_CUDA_PRECOMPUTE");
    return !(e && e[0] == '0');
}

// Makes `n` bases resident on c's device.  src is a host pointer (src_is_device = false) or
// a device pointer on the same device.  mode: 0 = policy (table of window multiples when it
// fits in free memory), 1 = plain bases, 2 = table or fail.  Caller holds c->mu.
static int make_resident(Ctx *c, const void *src, bool src_is_device, size_t n, int mode, void **out_ptr, uint32_t *out_c, bool *owns) {
    CUDA_TRY(cudaSetDevice(c->dev));
    // One-time, heavyweight: order after whatever stream produced a device-side source.
    if (src_is_device) CUDA_TRY(cudaDeviceSynchronize());
    bool want_table = (mode == 2) || (mode == 0 && precompute_enabled());
    uint32_t tc = pk_table_window_bits((u32)(n > 0xffffffffull ? 0xffffffffull : n));
    if (const char *e = getenv("PLONKISH_CUDA_TABLE_C")) {  // tuning override: window bits of the table
        const long v = atol(e);
        if (v >= 8 && v <= 22) tc = (uint32_t)v;
    }
    const uint32_t tw = pk_windows_for(tc);
    const size_t table_bytes = (size_t)tw * n * PLONKISH_CUDA_AFFINE_BYTES, cur_bytes = n * sizeof(xyzz);
    if (want_table) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const bool fits = (n <= ((size_t)1 << 27)) && (table_bytes + cur_bytes + ((size_t)8 << 30) < free_b);
        if (!fits) {
            if (mode == 2) return fail(PLONKISH_CUDA_E_INVALID, "bases_register: a table of %u x %zu points does not fit in %zu free bytes", tw, n, free_b);
            want_table = false;
        }
    }
    if (!want_table) {
        *out_c = 0;
        if (src_is_device) { *out_ptr = const_cast<void *>(src); *owns = false; return PLONKISH_CUDA_OK; }
        void *d = nullptr;
        CUDA_TRY(cudaMalloc(&d, n * PLONKISH_CUDA_AFFINE_BYTES));
        // stream-ordered with everything that will read it (a plain cudaMemcpy from pageable memory may
        // return while the DMA is still in flight on the legacy stream, which c->stream does not wait for)
        {
            const cudaError_t ce = upload(c, d, src, n * PLONKISH_CUDA_AFFINE_BYTES, c->stream);
            const cudaError_t se = ce == cudaSuccess ? cudaStreamSynchronize(c->stream) : ce;
            if (se != cudaSuccess) { cudaFree(d); return fail(PLONKISH_CUDA_E_CUDA, "bases_register: upload failed: %s", cudaGetErrorString(se)); }
        }
        *out_ptr = d; *owns = true;
        return PLONKISH_CUDA_OK;
    }
    void *table = nullptr, *cur = nullptr, *staged = nullptr;
    cudaError_t ce = cudaMalloc(&table, table_bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&cur, cur_bytes);
    const void *d_src = src;
    if (ce == cudaSuccess && !src_is_device) {
        ce = cudaMalloc(&staged, n * PLONKISH_CUDA_AFFINE_BYTES);
        if (ce == cudaSuccess) ce = upload(c, staged, src, n * PLONKISH_CUDA_AFFINE_BYTES, c->stream);  // ordered before the table kernels
        d_src = staged;
    }
    if (ce == cudaSuccess) {
        pk_enqueue_table_build(d_src, (u32)n, tc, tw, (xyzz *)cur, (affine *)table, c->stream);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    cudaFree(cur);
    cudaFree(staged);
    if (ce != cudaSuccess) {
        cudaFree(table);
        return fail(PLONKISH_CUDA_E_CUDA, "bases_register: building the table of %u x %zu points failed: %s", tw, n, cudaGetErrorString(ce));
    }
    *out_ptr = table; *out_c = tc; *owns = true;
    return PLONKISH_CUDA_OK;
}

static uint64_t publish(const BasesEntry &e) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uint64_t h = g_next_handle++;
    g_bases[h] = e;
    return h;
}

static int register_one(int device, const void *src, bool src_is_device, size_t n, int mode, uint64_t *handle) {
    if (!src || !handle || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "bases_register: null argument or n == 0");
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_register: device %d not initialised (call plonkish_cuda_init)", device);
    BasesEntry e;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        void *d = nullptr;
        uint32_t tc = 0;
        bool owns = true;
        int rc = make_resident(c, src, src_is_device, n, mode, &d, &tc, &owns);
        if (rc) return rc;
        e.n_shards = 1; e.dev = device; e.n = n;
        e.add_shard(device, d, owns, n, tc);
    }
    *handle = publish(e);
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_bases_register(int device, const void *bases, size_t n, uint64_t *handle) {
    return register_one(device, bases, false, n, 0, handle);
}

extern "C" int plonkish_cuda_bases_register_device(int device, const void *d_bases, size_t n, int mode, uint64_t *handle) {
    if (mode < 0 || mode > 2) return fail(PLONKISH_CUDA_E_INVALID, "bases_register_device: mode must be 0, 1 or 2");
    return register_one(device, d_bases, true, n, mode, handle);
}

extern "C" int plonkish_cuda_bases_register_sharded(int n_gpus, const void *bases, size_t n, uint64_t *handle) {
    if (!bases || !handle || n == 0 || n_gpus < 1) return fail(PLONKISH_CUDA_E_INVALID, "bases_register_sharded: bad argument");
    if (n_gpus > plonkish_cuda_device_count()) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_register_sharded: %d GPUs requested, %d initialised", n_gpus, plonkish_cuda_device_count());
    const size_t per = (n + n_gpus - 1) / n_gpus;  // msm.rs:101 chunk_size
    BasesEntry e;
    e.n_shards = n_gpus; e.dev = 0; e.n = n;
    for (int g = 0; g < n_gpus; ++g) {
        const size_t beg = (size_t)g * per;
        const size_t cnt = beg >= n ? 0 : (beg + per <= n ? per : n - beg);
        void *d = nullptr;
        uint32_t tc = 0;
        if (cnt) {
            Ctx *c = ctx_for(g);
            if (!c) return PLONKISH_CUDA_E_NO_DEVICE;
            std::lock_guard<std::mutex> lk(c->mu);
            bool owns = true;
            int rc = make_resident(c, (const char *)bases + beg * PLONKISH_CUDA_AFFINE_BYTES, false, cnt, 0, &d, &tc, &owns);
            if (rc) return rc;
        }
        e.add_shard(g, d, true, cnt, tc);  // an error below frees the earlier shards with `e`
    }
    *handle = publish(e);
    return PLONKISH_CUDA_OK;
}

// The same from device memory: d_bases[g] points at shard g (ceil(n/G) points, the last shards short or empty)
// on device g.  mode as in bases_register_device.
extern "C" int plonkish_cuda_bases_register_sharded_device(int n_gpus, const void *const *d_bases, size_t n, int mode, uint64_t *handle) {
    if (!d_bases || !handle || n == 0 || n_gpus < 1 || mode < 0 || mode > 2) return fail(PLONKISH_CUDA_E_INVALID, "bases_register_sharded_device: bad argument");
    if (n_gpus > plonkish_cuda_device_count()) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_register_sharded_device: %d GPUs requested, %d initialised", n_gpus, plonkish_cuda_device_count());
    const size_t per = (n + n_gpus - 1) / n_gpus;  // msm.rs:101 chunk_size
    BasesEntry e;
    e.n_shards = n_gpus; e.dev = 0; e.n = n;
    for (int g = 0; g < n_gpus; ++g) {
        const size_t beg = (size_t)g * per;
        const size_t cnt = beg >= n ? 0 : (beg + per <= n ? per : n - beg);
        void *d = nullptr;
        uint32_t tc = 0;
        bool owns = true;
        if (cnt) {
            if (!d_bases[g]) return fail(PLONKISH_CUDA_E_INVALID, "bases_register_sharded_device: null shard %d", g);
            Ctx *c = ctx_for(g);
            if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_register_sharded_device: device %d not initialised", g);
            std::lock_guard<std::mutex> lk(c->mu);
            int rc = make_resident(c, d_bases[g], true, cnt, mode, &d, &tc, &owns);
            if (rc) return rc;
        }
        e.add_shard(g, d, owns, cnt, tc);
    }
    *handle = publish(e);
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_bases_release(uint64_t handle) {
    BasesEntry e;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_bases.find(handle);
        if (it == g_bases.end()) return fail(PLONKISH_CUDA_E_INVALID, "bases_release: unknown handle %llu", (unsigned long long)handle);
        e = it->second;
        g_bases.erase(it);
    }
    // `e` holds the map's reference now: the memory is freed here, or by the last call still using the slice
    return PLONKISH_CUDA_OK;
}

// ---- borrowed slices: the cache behind the Rust shim's variable_base_msm --------------------------------------
// Inside variable_base_msm the shim sees a borrowed &[G1Affine] and nothing else (msm.rs:84-87): it cannot know
// whether the slice is a ProverParam's SRS (static: worth a resident table) or a temporary that will be freed and
// whose address the allocator will hand to the next Vec (IPA's folded generators, Hyrax rows, a second setup).  So
// the cache is keyed by address but trusted only after its content fingerprint — the first points, the last one
// and two dozen pseudo-random positions, hashed at registration — matches what is at that address now; a mismatch
// evicts the entry and registers the slice afresh.  A longer slice at a cached address replaces the entry, a
// shorter one uses its prefix (&powers_of_s_g1[..len], univariate/kzg.rs:28).  Entries are evicted least recently
// used first once the resident tables exceed the byte limit (default: half of the device's memory).
static uint64_t fnv1a64(const unsigned char *p, size_t len) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < len; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
static void cache_sample_positions(size_t n, std::vector<size_t> &pos) {
    pos.clear();
    for (size_t i = 0; i < n && i < 8; ++i) pos.push_back(i);
    if (n > 8) pos.push_back(n - 1);
    for (uint64_t k = 1; n > 9 && k <= 24; ++k) pos.push_back((size_t)((k * 0x9e3779b97f4a7c15ull) % n));
}
static void cache_drop_locked(std::map<uintptr_t, CachedBases>::iterator it) {  // caller holds g_cache_mu
    const uint64_t h = it->second.handle;
    g_bases_cache.erase(it);
    plonkish_cuda_bases_release(h);
}

static int bases_cached_impl(int device, int n_gpus, const void *bases_affine64, size_t n, uint64_t *handle) {
    if (!bases_affine64 || !handle || n == 0 || n_gpus < 1) return fail(PLONKISH_CUDA_E_INVALID, "bases_cached: null argument or n == 0");
    const unsigned char *src = (const unsigned char *)bases_affine64;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_bases_cache.find((uintptr_t)src);
    if (it != g_bases_cache.end()) {
        CachedBases &e = it->second;
        // a sharded entry serves exactly its own length (the shard boundaries depend on it)
        bool ok = e.n_gpus == n_gpus && (n_gpus > 1 ? e.n == n : (e.dev == device && e.n >= n));
        for (size_t k = 0; ok && k < e.samples.size(); ++k) {
            if (e.samples[k].first >= n) continue;  // beyond the caller's slice: not ours to read
            ok = fnv1a64(src + e.samples[k].first * PLONKISH_CUDA_AFFINE_BYTES, PLONKISH_CUDA_AFFINE_BYTES) == e.samples[k].second;
        }
        if (ok && n < e.n) ok = fnv1a64(src + (n - 1) * PLONKISH_CUDA_AFFINE_BYTES, PLONKISH_CUDA_AFFINE_BYTES) ==
                                 [&] { unsigned char b[PLONKISH_CUDA_AFFINE_BYTES]; return plonkish_cuda_bases_read(e.handle, n - 1, 1, b) == 0 ? fnv1a64(b, sizeof(b)) : 0; }();
        if (ok) {
            e.tick = ++g_cache_tick;
            *handle = e.handle;
            return PLONKISH_CUDA_OK;
        }
        cache_drop_locked(it);  // stale address, another device, or a longer slice than the one registered
    }
    if (n_gpus > 1) device = 0;
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_cached: device %d not initialised (call plonkish_cuda_init)", device);
    size_t limit = g_cache_limit;
    if (!limit) {
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(device);
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) limit = total_b / 2;
    }
    // (a sharded slice holds 1/G of this on every device; the limit is per device)
    const size_t per_dev = (n + n_gpus - 1) / n_gpus;
    const size_t estimate = per_dev * PLONKISH_CUDA_AFFINE_BYTES * pk_windows_for(pk_table_window_bits((u32)(per_dev > 0xffffffffull ? 0xffffffffull : per_dev)));
    for (;;) {  // make room: least recently used first
        size_t total = 0;
        auto lru = g_bases_cache.end();
        for (auto j = g_bases_cache.begin(); j != g_bases_cache.end(); ++j) {
            total += j->second.bytes;
            if (lru == g_bases_cache.end() || j->second.tick < lru->second.tick) lru = j;
        }
        if (lru == g_bases_cache.end() || total + estimate <= limit) break;
        cache_drop_locked(lru);
    }
    CachedBases e;
    int rc = n_gpus > 1 ? plonkish_cuda_bases_register_sharded(n_gpus, bases_affine64, n, &e.handle)
                        : plonkish_cuda_bases_register(device, bases_affine64, n, &e.handle);
    if (rc) return rc;
    e.n = n; e.dev = device; e.n_gpus = n_gpus; e.tick = ++g_cache_tick;
    {
        std::lock_guard<std::mutex> lk2(g_mu);
        e.bytes = g_bases[e.handle].device_bytes() / (size_t)n_gpus;
    }
    std::vector<size_t> pos;
    cache_sample_positions(n, pos);
    for (size_t q : pos) e.samples.push_back({q, fnv1a64(src + q * PLONKISH_CUDA_AFFINE_BYTES, PLONKISH_CUDA_AFFINE_BYTES)});
    g_bases_cache[(uintptr_t)src] = e;
    *handle = e.handle;
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_bases_cached(int device, const void *bases_affine64, size_t n, uint64_t *handle) {
    return bases_cached_impl(device, 1, bases_affine64, n, handle);
}
// The same for the multi-GPU entry: the slice is sharded over devices 0..n_gpus-1 (bases_register_sharded).
extern "C" int plonkish_cuda_bases_cached_sharded(int n_gpus, const void *bases_affine64, size_t n, uint64_t *handle) {
    if (n_gpus > plonkish_cuda_device_count()) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_cached_sharded: %d GPUs requested, %d initialised", n_gpus, plonkish_cuda_device_count());
    return bases_cached_impl(0, n_gpus, bases_affine64, n, handle);
}

// Drop hook for the owner of a cached slice (ProverParam's Drop in the shim): forget and free it now.
extern "C" int plonkish_cuda_bases_cache_evict(const void *bases_affine64) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_bases_cache.find((uintptr_t)bases_affine64);
    if (it == g_bases_cache.end()) return PLONKISH_CUDA_OK;
    cache_drop_locked(it);
    return PLONKISH_CUDA_OK;
}

// max_bytes of resident tables the cache may hold (0 = default, half of the device's memory); evicts down to it.
extern "C" int plonkish_cuda_bases_cache_limit(size_t max_bytes) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_limit = max_bytes;
    for (;;) {
        size_t total = 0;
        auto lru = g_bases_cache.end();
        for (auto j = g_bases_cache.begin(); j != g_bases_cache.end(); ++j) {
            total += j->second.bytes;
            if (lru == g_bases_cache.end() || j->second.tick < lru->second.tick) lru = j;
        }
        if (!max_bytes || lru == g_bases_cache.end() || total <= max_bytes) break;
        cache_drop_locked(lru);
    }
    return PLONKISH_CUDA_OK;
}

// out[0] = entries, out[1] = bytes of device memory they hold.
extern "C" int plonkish_cuda_bases_cache_stats(size_t out[2]) {
    if (!out) return fail(PLONKISH_CUDA_E_INVALID, "bases_cache_stats: null output");
    std::lock_guard<std::mutex> lk(g_cache_mu);
    out[0] = g_bases_cache.size();
    out[1] = 0;
    for (auto &kv : g_bases_cache) out[1] += kv.second.bytes;
    return PLONKISH_CUDA_OK;
}

static bool lookup_bases(uint64_t handle, BasesEntry &out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_bases.find(handle);
    if (it == g_bases.end()) return false;
    out = it->second;
    return true;
}

// ------------------------------------------------------------- enqueue helpers
// Caller holds c->mu and has made c->dev current.  Enqueues the MSM over n device-
// resident points on `stream`; leaves the projective sum in *d_xyzz_out (device).
static MsmPlan plan_for(const Ctx *c, const BasesView &b, size_t n, uint32_t window_bits) {
    if (b.table_c) {
        MsmPlan p = pk_make_plan_b((u32)n, b.table_c, (u32)b.stride, (u32)c->sm_count);
        // Level 1 fused with the decomposition pays up to 2^21 points (measured, scatter stage: 2^18 0.089 -> 0.037 ms, 2^20
        // 0.30 -> 0.115, 2^21 0.33 -> 0.22, 2^22 0.41 -> 0.41, 2^24 1.03 -> 1.40: one scalar conversion per thread and stage makes
        // the big launches issue bound, while small ones save the digit round trip and a launch's worth of latency).
        static const int fuse = [] { const char *e = getenv("PLONKISH_CUDA_FUSE_L1"); return e ? atoi(e) : -1; }();  // A/B switch: 0 / 1 force
        p.fuse_l1 = fuse >= 0 ? (u32)(fuse != 0) : (n <= ((size_t)1 << 21) ? 1u : 0u);
        // tuning overrides (read per call so that one process can sweep them): the run length of a K3 thread, directly or as
        // the number of waves of resident threads the launch should make
        if (const char *e = getenv("PLONKISH_CUDA_ACC_WAVES")) {
            const double waves = atof(e);
            if (waves > 0) {
                double L = (double)n * p.W / ((double)c->sm_count * 512.0 * waves);
                pk_set_run_length(p, (u32)(L < 16 ? 16 : L > 1024 ? 1024 : L));
            }
        }
        if (const char *e = getenv("PLONKISH_CUDA_ACC_L")) {  // equal runs of L entries instead of the tiers
            const long L = atol(e);
            if (L >= 1 && L <= 4096) pk_set_run_length(p, (u32)L);
        }
        if (const char *e = getenv("PLONKISH_CUDA_ACC_TIERS")) {  // number of tiers (1 = equal runs of rem / (2 R))
            const long J = atol(e);
            if (J >= 1 && J <= 16 && p.acc_tiers) {
                p.acc_tiers = (u32)J;
                p.nthreads1 = pk_acc_threads((unsigned long long)p.n * p.W, p.acc_tiers, p.acc_resident);
            }
        }
        return p;
    }
    return pk_make_plan((u32)n, window_bits, (u32)c->sm_count);
}

static int enqueue_device_msm(Ctx *c, const void *d_scalars, const BasesView &bases, size_t n, uint32_t window_bits,
                              cudaStream_t stream, xyzz **d_result, const StageMarks *marks = nullptr) {
    const size_t first_chunk = n < MAX_POINTS_PER_LAUNCH ? n : MAX_POINTS_PER_LAUNCH;
    const size_t tail_chunk = n % MAX_POINTS_PER_LAUNCH;
    MsmPlan plan0 = plan_for(c, bases, first_chunk, window_bits);
    size_t need = pk_workspace_bytes(plan0);
    if (n > MAX_POINTS_PER_LAUNCH && tail_chunk) {
        const size_t t = pk_workspace_bytes(plan_for(c, bases, tail_chunk, window_bits));
        if (t > need) need = t;
    }
    int rc = grow(c->arena, need);
    if (rc) return rc;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(stream, c->last_done, 0));
    // Chunks of <= 2^26 points run back to back on the stream and chain their
    // projective sums through one slot outside the arena (msm.rs:112-114).
    xyzz *running = (xyzz *)((char *)c->d_out + 256);
    xyzz *prev = nullptr;
    for (size_t done = 0; done < n;) {
        const size_t cnt = (n - done < MAX_POINTS_PER_LAUNCH) ? n - done : MAX_POINTS_PER_LAUNCH;
        MsmPlan plan = (cnt == first_chunk) ? plan0 : plan_for(c, bases, cnt, window_bits);
        MsmWorkspace w = pk_carve_workspace(plan, c->arena.ptr);
        w.result = running;
        pk_enqueue_msm(plan, (const char *)d_scalars + done * PLONKISH_CUDA_SCALAR_BYTES,
                       (const char *)bases.ptr + done * PLONKISH_CUDA_AFFINE_BYTES, w, prev, stream, marks);
        prev = running;
        done += cnt;
    }
    CUDA_TRY(cudaGetLastError());
    *d_result = running;
    return PLONKISH_CUDA_OK;
}

static int mark_done(Ctx *c, cudaStream_t stream) {
    CUDA_TRY(cudaEventRecord(c->last_done, stream));
    c->has_last = true;
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------------------- timer lines
// The reference wraps every MSM in start_timer(|| format!("variable_base_msm-{}", n)) (msm.rs:92, util/timer.rs:19-24:
// ark_std's perf_trace) and benchmark/src/bin/plotter.rs:337-373 rebuilds its cost breakdown from those Start/End
// lines.  A single call keeps that line on the Rust side.  The batch / many / open entry points replace several
// reference calls, so the library prints one Start/End pair per MSM itself, in perf_trace's format, each with that
// MSM's own share of the call (the shares add up to the call's duration: the plotter subtracts them from the parent).
// mode: 0 off, 1 stderr, 2 stdout (PLONKISH_CUDA_TIMER=1 / stdout); depth = nesting depth of the caller's timers.
static std::atomic<int> g_timer_mode{-1}, g_timer_depth{-1};
static int timer_mode() {
    int m = g_timer_mode.load(std::memory_order_relaxed);
    if (m >= 0) return m;
    const char *e = getenv("PLONKISH_CUDA_TIMER");
    m = !e ? 0 : (e[0] == '1' ? 1 : (e[0] == '2' || e[0] == 's') ? 2 : 0);
    g_timer_mode.store(m);
    return m;
}
static int timer_depth() {
    int d = g_timer_depth.load(std::memory_order_relaxed);
    if (d >= 0) return d;
    const char *e = getenv("PLONKISH_CUDA_TIMER_DEPTH");
    d = e ? (int)atol(e) : 0;
    if (d < 0 || d > 32) d = 0;
    g_timer_depth.store(d);
    return d;
}
extern "C" int plonkish_cuda_timer_config(int mode, int depth) {
    if (mode < 0 || mode > 2 || depth < 0 || depth > 32) return fail(PLONKISH_CUDA_E_INVALID, "timer_config: mode 0..2, depth 0..32");
    g_timer_mode.store(mode);
    g_timer_depth.store(depth);
    return PLONKISH_CUDA_OK;
}
static void timer_emit(size_t n, double ms) {
    const int mode = timer_mode();
    if (!mode) return;
    FILE *f = mode == 2 ? stdout : stderr;
    const int indent = 2 * timer_depth();
    std::string pad;
    for (int i = 0; i < indent; ++i) pad += "\xc2\xb7";  // PAD_CHAR of perf_trace (U+00B7)
    char msg[96], dur[48];
    snprintf(msg, sizeof(msg), "variable_base_msm-%zu ", n);
    const double ns = ms * 1e6;
    if (ns >= 1e9) snprintf(dur, sizeof(dur), "%.3fs", ns / 1e9);
    else if (ns >= 1e6) snprintf(dur, sizeof(dur), "%.3fms", ns / 1e6);
    else if (ns >= 1e3) snprintf(dur, sizeof(dur), "%.3f\xc2\xb5s", ns / 1e3);
    else snprintf(dur, sizeof(dur), "%.0fns", ns);
    std::string line = pad + "Start:   " + (std::string(msg, strlen(msg) - 1)) + "\n" + pad + "End:     " + msg;
    for (int w = (int)strlen(msg); w < 75 - indent; ++w) line += '.';
    line += dur;
    line += "\n";
    fputs(line.c_str(), f);
    fflush(f);
}
// One Start/End pair for an MSM the caller timed itself (a host mirror composing several calls).
extern "C" void plonkish_cuda_timer_emit(size_t n, double ms) { timer_emit(n, ms); }
static double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
static void timer_report(size_t n, std::chrono::steady_clock::time_point t0) {
    if (timer_mode()) timer_emit(n, ms_since(t0));
}
// Several MSMs that ran concurrently inside one call: the call's duration is split by the library's own load model
// (0.6 ms of launch latency + 3 ns per point, see enqueue_many), so the lines add up to the call.
static void timer_report_shared(const std::vector<size_t> &ns, std::chrono::steady_clock::time_point t0) {
    if (!timer_mode()) return;
    const double total = ms_since(t0);
    double wsum = 0;
    for (size_t n : ns) wsum += (double)n + 200000.0;
    for (size_t n : ns) timer_emit(n, total * ((double)n + 200000.0) / wsum);
}

// --------------------------------------------------------------- device entry
static int msm_device_common(int device, const void *d_scalars, const BasesView &bases, size_t n, uint32_t window_bits,
                             void *d_out_affine64, void *d_out_xyzz128, void *cuda_stream) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_device: device %d not initialised (call plonkish_cuda_init)", device);
    if (!d_out_affine64 && !d_out_xyzz128) return fail(PLONKISH_CUDA_E_INVALID, "msm_device: no output pointer");
    if (n && (!d_scalars || !bases.ptr)) return fail(PLONKISH_CUDA_E_INVALID, "msm_device: null input");
    if (window_bits && (window_bits < 8 || window_bits > 16)) return fail(PLONKISH_CUDA_E_INVALID, "msm_device: window_bits must be 0 or 8..16");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    cudaStream_t stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, as in the CUDA runtime
    if (n == 0) {
        if (d_out_affine64) CUDA_TRY(cudaMemsetAsync(d_out_affine64, 0, PLONKISH_CUDA_AFFINE_BYTES, stream));
        if (d_out_xyzz128) CUDA_TRY(cudaMemsetAsync(d_out_xyzz128, 0, PLONKISH_CUDA_XYZZ_BYTES, stream));
        return PLONKISH_CUDA_OK;
    }
    xyzz *res = nullptr;
    int rc = enqueue_device_msm(c, d_scalars, bases, n, window_bits, stream, &res);
    if (rc) return rc;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, stream, res, 1u, (affine *)d_out_affine64, (xyzz *)d_out_xyzz128);
    CUDA_TRY(cudaGetLastError());
    return mark_done(c, stream);
}

extern "C" int plonkish_cuda_msm_bn254_g1_device(int device, const void *d_scalars, const void *d_bases, size_t n,
                                                 uint32_t window_bits, void *d_out_affine64, void *d_out_xyzz128, void *cuda_stream) {
    BasesView v;
    v.ptr = d_bases;
    return msm_device_common(device, d_scalars, v, n, window_bits, d_out_affine64, d_out_xyzz128, cuda_stream);
}

static int view_of(uint64_t handle, size_t n, int shard, BasesView &v, int *device, const char *who) {
    BasesEntry e;
    if (!lookup_bases(handle, e)) return fail(PLONKISH_CUDA_E_INVALID, "%s: unknown bases handle %llu", who, (unsigned long long)handle);
    if (shard < 0) {
        if (e.n_shards != 1) return fail(PLONKISH_CUDA_E_INVALID, "%s: handle is sharded; use plonkish_cuda_msm_bn254_g1_multi", who);
        shard = 0;
    }
    if (n > e.shard_n[shard]) return fail(PLONKISH_CUDA_E_INVALID, "%s: n = %zu exceeds the %zu registered bases", who, n, e.shard_n[shard]);
    v.ptr = e.d_ptr[shard];
    v.table_c = e.table_c[shard];
    v.stride = e.shard_n[shard];
    v.keep = e.keep[shard];
    // A short prefix of a long slice (&powers_of_s_g1[..len], univariate/kzg.rs:28) would pay for the table's full
    // bucket set (2^(c-1) buckets sized for the whole slice): below 1/16 of the slice the plain layout on row 0 of
    // the table — the bases themselves — is cheaper.
    // PLONKISH_CUDA_PREFIX_SLACK: how many window bits beyond its own a prefix may inherit (default 3).
    static const u32 slack = [] { const char *e = getenv("PLONKISH_CUDA_PREFIX_SLACK"); const int s = e ? atoi(e) : 3; return (u32)(s < 0 ? 0 : s > 16 ? 16 : s); }();
    if (v.table_c && n && v.table_c > pk_table_window_bits((u32)n) + slack) v.table_c = 0;
    if (device) *device = e.n_shards == 1 ? e.dev : shard;
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_msm_bn254_g1_device_resident(const void *d_scalars, uint64_t bases_handle, size_t n, void *d_out_affine64,
                                                          void *d_out_xyzz128, void *cuda_stream) {
    BasesView v;
    int device = 0;
    int rc = view_of(bases_handle, n, -1, v, &device, "msm_device_resident");
    if (rc) return rc;
    return msm_device_common(device, d_scalars, v, n, 0, d_out_affine64, d_out_xyzz128, cuda_stream);
}

extern "C" int plonkish_cuda_g1_sum_partials_device(int device, const void *d_partials, size_t count, void *d_out_affine64, void *cuda_stream) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sum_partials: device %d not initialised", device);
    if (!d_partials || !d_out_affine64 || count == 0) return fail(PLONKISH_CUDA_E_INVALID, "sum_partials: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    cudaStream_t stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, as in the CUDA runtime
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, stream, (const xyzz *)d_partials, (u32)count, (affine *)d_out_affine64, (xyzz *)nullptr);
    CUDA_TRY(cudaGetLastError());
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------- host scalars, pipelined
// The step's scalars arrive from host memory (32 B per point, 512 MB at 2^24: ~10 ms of
// PCIe).  The points are cut into chunks; chunk k+1 is copied on the copy stream while the
// compute stream decomposes, sorts and accumulates chunk k into its own bucket array
// (MsmPlan::chunk); one bucket reduce at the end adds the arrays.
static int enqueue_host_msm(Ctx *c, const void *h_scalars, const BasesView &bases, size_t n, xyzz **d_result, void *d_dst = nullptr) {
    if (!d_dst) d_dst = c->scalars.ptr;  // where the uploaded scalars live (a caller keeping them resident passes its own buffer)
    // Chunk boundaries.  Only the first chunk's copy is exposed, so it is small; later chunks
    // grow (1/16, 1/4, 11/16 at n >= 2^23; 1/4, 3/4 from 2^20) and each copies while its predecessor
    // computes.  Measured (tools/e2e_chunks.py): 2^24 48.4 ms unchunked, 42.8 ms chunked (38.5 ms with
    // the scalars already resident); 2^22 14.0 -> 13.1 ms; equal 4- and 8-way splits lose to per-chunk costs.
    std::vector<size_t> cuts;  // chunk end offsets
    const char *env = getenv("PLONKISH_CUDA_HOST_CHUNKS");  // tuning override: N equal chunks
    const long forced = env ? atol(env) : 0;
    const char *env_cuts = getenv("PLONKISH_CUDA_HOST_CUTS");  // tuning override: chunk ends as fractions, e.g. "0.25,1"
    if (env_cuts && n >= ((size_t)1 << 20) && n <= MAX_POINTS_PER_LAUNCH) {
        for (const char *q = env_cuts; *q;) {
            char *e = nullptr;
            const double fr = strtod(q, &e);
            if (e == q) break;
            if (fr > 0 && fr <= 1.0) cuts.push_back((size_t)((double)n * fr));
            q = (*e == ',') ? e + 1 : e;
        }
        if (cuts.empty() || cuts.back() != n) cuts.push_back(n);
    } else if (forced >= 1 && forced <= 16) {
        for (long k = 1; k <= forced; ++k) cuts.push_back(n * (size_t)k / (size_t)forced);
    } else if (n >= ((size_t)1 << 23) && n <= MAX_POINTS_PER_LAUNCH) {
        // The three-chunk geometry (1/16, 1/4, 11/16) needs an upload rate of ~44 GB/s: 11/16 of the scalars must arrive
        // while a quarter of the points computes.  The staging ring sustains that for one rank; ranks sharing the host's
        // memory bandwidth do not (measured per rank: 28 GB/s with two ranks, 17 with four, 8.7 with eight), and the
        // MSM itself consumes its scalars at ~16.5 GB/s (32 B x 516 Mpoints/s).  So the chunks grow by g = upload rate /
        // consumption rate, from 1/16 of the points, in as few chunks as that takes and at most eight (every extra chunk
        // costs ~1 ms: one more bucket array in the reduce, smaller launches); with g near one that is eight equal
        // chunks.  Measured, 2^24 points per rank from pageable memory: two ranks 43.0 -> 38.2 ms (five chunks), four
        // ranks 50.8 ms with five chunks (see DESIGN.md 6 for the final figures).  Pinned sources keep three chunks
        // (five lose there: 35.8 -> 38.7 ms).
        const double rate = Stager::get(c->dev).rate_gbps();
        const double g_raw = rate / 16.5;
        if (rate > 0 && g_raw < 2.6 && is_staged_source(h_scalars, n * 32)) {
            const double g = g_raw < 1.0 ? 1.0 : g_raw;
            const int max_chunks = 8;
            double f[max_chunks], sum = 0;
            int m = 0;
            for (double v = 1.0 / 16.0; m < max_chunks && sum < 1.0; v *= g) { f[m++] = v; sum += v; }
            double acc = 0;
            for (int k = 0; k < m; ++k) {
                acc += f[k] / sum;
                size_t cut = k + 1 == m ? n : (size_t)((double)n * acc);
                cut &= ~(size_t)7;  // chunk starts stay 256-byte aligned in the scalar buffer
                if (k + 1 == m) cut = n;
                cuts.push_back(cut);
            }
        } else {
            cuts = {n / 16, 5 * (n / 16), n};
        }
    } else if (n >= ((size_t)1 << 20) && n <= MAX_POINTS_PER_LAUNCH) {
        cuts = {n / 4, n};
    } else {
        const size_t pieces = (n + MAX_POINTS_PER_LAUNCH - 1) / MAX_POINTS_PER_LAUNCH;
        for (size_t k = 1; k <= pieces; ++k) cuts.push_back(n * k / pieces);
    }
    {  // drop empty chunks (tiny n with a forced chunk count)
        std::vector<size_t> kept;
        for (size_t k = 0, prev = 0; k < cuts.size(); ++k) {
            if (cuts[k] > prev) { kept.push_back(cuts[k]); prev = cuts[k]; }
        }
        cuts.swap(kept);
    }
    const size_t nchunks = cuts.size();
    if (nchunks > 16) return fail(PLONKISH_CUDA_E_INVALID, "msm: n = %zu is beyond 16 x 2^26 points", n);
    size_t largest = 0;
    for (size_t k = 0, prev = 0; k < nchunks; prev = cuts[k], ++k) largest = (cuts[k] - prev > largest) ? cuts[k] - prev : largest;
    if (largest > MAX_POINTS_PER_LAUNCH) return fail(PLONKISH_CUDA_E_INVALID, "msm: chunk of %zu points exceeds 2^26", largest);
    // every chunk uses the window width the whole MSM would use, so the bucket layout is shared
    const uint32_t c_all = bases.table_c ? 0 : plan_for(c, bases, n < MAX_POINTS_PER_LAUNCH ? n : MAX_POINTS_PER_LAUNCH, 0).c;
    MsmPlan plan0 = plan_for(c, bases, largest, c_all);
    plan0.nchunks = (u32)nchunks;
    int rc = grow(c->arena, pk_workspace_bytes(plan0));
    if (rc) return rc;
    if (c->has_last) {
        // the scalar buffer and the arena are free once the previous MSM is done; the uploads
        // below run on the copy stream, so it waits as well
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->last_done, 0));
    }
    MsmWorkspace ws0 = pk_carve_workspace(plan0, c->arena.ptr);
    MsmPlan last = plan0;
    for (size_t k = 0, done = 0; k < nchunks; done = cuts[k], ++k) {
        const size_t cnt = cuts[k] - done;
        char *d_chunk = (char *)d_dst + done * PLONKISH_CUDA_SCALAR_BYTES;
        cudaStream_t cs = (nchunks > 1) ? c->copy_stream : c->stream;
        CUDA_TRY(upload(c, d_chunk, (const char *)h_scalars + done * PLONKISH_CUDA_SCALAR_BYTES, cnt * PLONKISH_CUDA_SCALAR_BYTES, cs));
        if (nchunks > 1) {
            CUDA_TRY(cudaEventRecord(c->chunk_ready[k], cs));
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->chunk_ready[k], 0));
        }
        MsmPlan plan = plan_for(c, bases, cnt, c_all);
        plan.chunk = (u32)k;
        plan.nchunks = (u32)nchunks;
        MsmWorkspace w = pk_carve_workspace(plan, c->arena.ptr);
        pk_enqueue_buckets(plan, d_chunk, (const char *)bases.ptr + done * PLONKISH_CUDA_AFFINE_BYTES, w, c->stream);
        last = plan;
    }
    xyzz *running = (xyzz *)((char *)c->d_out + 256);
    ws0.result = running;
    pk_enqueue_reduce(last, ws0, nullptr, c->stream);
    CUDA_TRY(cudaGetLastError());
    *d_result = running;
    return PLONKISH_CUDA_OK;
}

// ----------------------------------------------------------------- host entry
extern "C" int plonkish_cuda_msm_bn254_g1(const void *scalars, const void *bases, uint64_t bases_handle, size_t n, void *out_affine64) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!out_affine64) return fail(PLONKISH_CUDA_E_INVALID, "msm: null output");
    if (n && !scalars) return fail(PLONKISH_CUDA_E_INVALID, "msm: null scalars");
    if (n && !bases && !bases_handle) return fail(PLONKISH_CUDA_E_INVALID, "msm: neither bases nor a handle given");
    int device = 0;
    BasesView view;
    if (bases_handle) {
        int rc0 = view_of(bases_handle, n, -1, view, &device, "msm");
        if (rc0) return rc0;
    }
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm: device %d not initialised (call plonkish_cuda_init; there is no CPU fallback)", device);
    if (n == 0) {
        memset(out_affine64, 0, PLONKISH_CUDA_AFFINE_BYTES);
        return PLONKISH_CUDA_OK;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc = grow(c->scalars, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    if (!bases_handle) {
        rc = grow(c->bases_tmp, n * PLONKISH_CUDA_AFFINE_BYTES);
        if (rc) return rc;
        CUDA_TRY(upload(c, c->bases_tmp.ptr, bases, n * PLONKISH_CUDA_AFFINE_BYTES, c->stream));
        view.ptr = c->bases_tmp.ptr;
    }
    xyzz *res = nullptr;
    rc = enqueue_host_msm(c, scalars, view, n, &res);
    if (rc) return rc;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, c->stream, res, 1u, (affine *)c->d_out, (xyzz *)nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_out, c->d_out, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    rc = mark_done(c, c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_affine64, c->h_out, PLONKISH_CUDA_AFFINE_BYTES);
    timer_report(n, t0);
    return PLONKISH_CUDA_OK;
}

// ---------------------------------------------------------------- batch entry
// `count` MSMs of n points each against one resident base slice — the shape of
// MultilinearKzg::batch_commit (pcs/multilinear/kzg.rs:259-274: one variable_base_msm per
// polynomial, all against pp.eq(num_vars)).  The reference runs them one after another; here
// the copy stream uploads the scalars of MSM j+1 (two device buffers) while the compute stream
// works on MSM j, and the host waits once at the end.
static uint64_t publish_scalars(int dev, void *d_ptr, size_t n);

// PLONKISH_CUDA_BIG_LANE: 0 = off, 1 (default) = the batch entry alternates its MSMs between the main stream and the big
// lane, 2 = the many / open entry does so too for its large MSMs.  Measured on B200: batch_commit of three 2^24-point
// polynomials 101.9 -> 100.9 ms, of three 2^20-point ones 10.4 -> 9.8 ms (only the decompose and the small kernels find
// room beside the previous MSM's reduce: the 1 024-thread sort blocks need a whole SM); open on 2^24 evaluations
// 38.6 -> 39.2 ms, on 2^20 5.29 -> 5.08 ms, hence off there by default.
static int big_lane_mode() {
    static const int mode = [] { const char *e = getenv("PLONKISH_CUDA_BIG_LANE"); return e ? atoi(e) : 1; }();
    return mode;
}
static bool big_lane_on() { return big_lane_mode() >= 1; }


// keep != nullptr: every polynomial's scalars stay resident under keep[j] (its own allocation)
// instead of passing through the two staging buffers.
static int batch_impl(const void *const *scalars_list, size_t count, uint64_t bases_handle, size_t n, void *out_affine64_list, uint64_t *keep) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!out_affine64_list || (count && !scalars_list)) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch: null argument");
    if (!bases_handle) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch: a registered bases handle is required");
    if (count == 0) return PLONKISH_CUDA_OK;
    if (n == 0) { memset(out_affine64_list, 0, count * PLONKISH_CUDA_AFFINE_BYTES); return PLONKISH_CUDA_OK; }
    if (n > MAX_POINTS_PER_LAUNCH) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch: n = %zu exceeds 2^26 points per MSM", n);
    BasesView view;
    int device = 0;
    int rc = view_of(bases_handle, n, -1, view, &device, "msm_batch");
    if (rc) return rc;
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_batch: device %d not initialised", device);
    for (size_t j = 0; j < count; ++j)
        if (!scalars_list[j]) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch: null scalars for MSM %zu", j);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t bytes = n * PLONKISH_CUDA_SCALAR_BYTES;
    std::vector<void *> kept(count, nullptr);
    struct KeptGuard {  // frees the kept buffers unless the call succeeds (publish takes them over)
        Ctx *c; std::vector<void *> &v; bool armed = true;
        ~KeptGuard() { if (armed) for (void *p : v) pool_free(c, p); }
    } kept_guard{c, kept};
    if (keep) {
        for (size_t j = 0; j < count; ++j) {
            if (pool_alloc(c, &kept[j], bytes) != 0) {
                return fail(PLONKISH_CUDA_E_CUDA, "msm_batch: cannot keep %zu x %zu bytes of scalars resident", count, bytes);
            }
        }
    } else if ((rc = grow(c->scalars, bytes)) || (rc = grow(c->scalars2, bytes))) {
        return rc;
    }
    if ((rc = grow(c->batch_out, count * PLONKISH_CUDA_AFFINE_BYTES))) return rc;
    MsmPlan plan = plan_for(c, view, n, 0);
    if ((rc = grow(c->arena, pk_workspace_bytes(plan)))) return rc;  // before anything is in flight: growing synchronises
    void *bufs[2] = {c->scalars.ptr, c->scalars2.ptr};
    const bool timed = timer_mode() != 0;
    // Every second MSM runs on the big lane (own stream, own workspace): its decompose and sort overlap the previous
    // MSM's draining accumulate, reduce, item levels and finalize.  Not when the timer lines are on: they measure each
    // MSM's own span on the one compute stream.  And only when every polynomial has its own resident buffer (keep): with
    // the two recycled upload buffers an MSM on the big lane can get ahead of the first MSM's last chunk, whose end then
    // frees buffer 0 late and stalls the third upload (measured: three 2^24-point MSMs 101.2 -> 109.4 ms).
    Ctx::Lane &big = c->lanes[PK_LANES];
    const bool alternate = big_lane_on() && !timed && count >= 2 && keep != nullptr;
    if (alternate && (rc = grow(big.arena, pk_workspace_bytes(plan)))) return rc;
    std::vector<cudaEvent_t> tev;  // timed: the end of every MSM on the compute stream, for one timer line each
    struct TevGuard { std::vector<cudaEvent_t> &v; ~TevGuard() { for (cudaEvent_t e : v) cudaEventDestroy(e); } } tev_guard{tev};
    if (timed) {
        tev.resize(count + 1, nullptr);
        for (auto &e : tev) CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(tev[0], c->stream));
    }
    // MSM 0: nothing to hide its upload behind, so it goes up in growing chunks, each copying
    // while its predecessor computes (enqueue_host_msm); sizes the arena for the chunked layout.
    {
        xyzz *res0 = nullptr;
        if ((rc = enqueue_host_msm(c, scalars_list[0], view, n, &res0, keep ? kept[0] : bufs[0]))) return rc;
        PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, c->stream, res0, 1u, (affine *)c->batch_out.ptr, (xyzz *)nullptr);
        CUDA_TRY(cudaEventRecord(c->buf_free[0], c->stream));
        if (timed) CUDA_TRY(cudaEventRecord(tev[1], c->stream));
    }
    MsmWorkspace ws = pk_carve_workspace(plan, c->arena.ptr);
    ws.result = (xyzz *)((char *)c->d_out + 256);
    MsmWorkspace ws_big = ws;
    if (alternate) {
        ws_big = pk_carve_workspace(plan, big.arena.ptr);
        ws_big.result = (xyzz *)big.d_res;
        // the big lane starts after everything enqueued so far on the main stream's predecessors (the previous call)
        if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(big.stream, c->last_done, 0));
    }
    bool big_used = false;
    // MSM j >= 1: its upload runs on the copy stream while MSM j-1 computes
    for (size_t j = 1; j < count; ++j) {
        const int b = (int)(j & 1);
        const bool on_big = alternate && (j & 1);
        cudaStream_t st = on_big ? big.stream : c->stream;
        const MsmWorkspace &w = on_big ? ws_big : ws;
        void *dst = keep ? kept[j] : bufs[b];
        if (j >= 2 && !keep) CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->buf_free[b], 0));  // MSM j-2 has consumed this buffer
        CUDA_TRY(upload(c, dst, scalars_list[j], bytes, c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->chunk_ready[b], c->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, c->chunk_ready[b], 0));
        pk_enqueue_msm(plan, dst, view.ptr, w, nullptr, st);
        CUDA_TRY(cudaEventRecord(c->buf_free[b], st));  // scalars are dead after the decompose; recorded after the whole MSM for simplicity
        PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, st, w.result, 1u,
                  (affine *)((char *)c->batch_out.ptr + j * PLONKISH_CUDA_AFFINE_BYTES), (xyzz *)nullptr);
        big_used = big_used || on_big;
        if (timed) CUDA_TRY(cudaEventRecord(tev[j + 1], c->stream));
    }
    if (big_used) {  // join the big lane before the results are read
        CUDA_TRY(cudaEventRecord(big.done, big.stream));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, big.done, 0));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_affine64_list, c->batch_out.ptr, count * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    kept_guard.armed = false;
    if (keep)
        for (size_t j = 0; j < count; ++j) keep[j] = publish_scalars(c->dev, kept[j], n);
    if (timed) {
        // MSM j's line: from the end of MSM j-1 to its own end on the compute stream (uploads overlap the
        // predecessor); the host time before the first event and after the last goes to the first and last line
        const double wall = ms_since(t0);
        std::vector<double> ms(count, 0.0);
        double dev_total = 0;
        for (size_t j = 0; j < count; ++j) {
            float f = 0;
            CUDA_TRY(cudaEventElapsedTime(&f, tev[j], tev[j + 1]));
            ms[j] = f;
            dev_total += f;
        }
        const double rest = wall > dev_total ? wall - dev_total : 0.0;
        ms[0] += rest * 0.5;
        ms[count - 1] += rest * 0.5;
        for (size_t j = 0; j < count; ++j) timer_emit(n, ms[j]);
    }
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_msm_bn254_g1_batch(const void *const *scalars_list, size_t count, uint64_t bases_handle, size_t n,
                                                void *out_affine64_list) {
    return batch_impl(scalars_list, count, bases_handle, n, out_affine64_list, nullptr);
}

// batch_commit that leaves each polynomial's evaluations resident for the later open
// (backend/hyperplonk.rs:201,251 commit the witness / permutation polynomials that :287 opens).
extern "C" int plonkish_cuda_msm_bn254_g1_batch_keep(const void *const *scalars_list, size_t count, uint64_t bases_handle, size_t n,
                                                     void *out_affine64_list, uint64_t *scalars_handles_out) {
    if (!scalars_handles_out) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch_keep: null handle output");
    if (n == 0) return fail(PLONKISH_CUDA_E_INVALID, "msm_batch_keep: n == 0");
    return batch_impl(scalars_list, count, bases_handle, n, out_affine64_list, scalars_handles_out);
}

// ------------------------------------------------------------------ many entry
// `count` independent MSMs, each with its own size and resident base slice — the k quotient
// commitments of MultilinearKzg::open (pcs/multilinear/kzg.rs:291-293 through
// pcs/multilinear.rs:72-107: sizes 2^(k-1), ..., 2, 1 against eqs[k-1..0]).  The quotients do not
// depend on the commitments, so the caller computes them all and hands them over in one call:
// large MSMs run one after another on the main stream (chunk-pipelined uploads), small ones are
// spread over three more streams with their own scratch and overlap with everything else —
// a small MSM is bound by the latency of its dozen short kernels, not by throughput.
struct ManyJob {
    const void *scalars = nullptr;  // host pointer, or device pointer when on_device
    bool on_device = false;
    BasesView view;
    size_t n = 0;
};

// Enqueues every job (results to d_out_list[j], zero for n == 0) and joins the side lanes
// back into c->stream.  Caller holds c->mu, has made c->dev current, and synchronises.
static int enqueue_many(Ctx *c, const std::vector<ManyJob> &jobs, void *d_out_list) {
    const size_t count = jobs.size();
    int rc;
    const int NL = PK_LANES;
    // The two or three largest MSMs (more than a quarter of the largest) run in turn on the main stream;
    // every smaller one goes to the side lane with the least work so far (assigned largest first), where
    // its latency-bound stages fill the gaps of the large ones.
    std::vector<size_t> order;
    size_t max_n = 0;
    for (size_t j = 0; j < count; ++j) {
        if (jobs[j].n == 0) continue;
        order.push_back(j);
        if (jobs[j].n > max_n) max_n = jobs[j].n;
    }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return jobs[a].n > jobs[b].n; });
    size_t SMALL = max_n / 4;
    if (const char *e = getenv("PLONKISH_CUDA_MANY_SMALL_LOG2")) {  // tuning override: MSMs up to 2^v points go to the side lanes
        const long v = atol(e);
        if (v >= 0 && v <= 24) SMALL = (size_t)1 << v;
    }
    std::vector<int> lane_of(count, -1);
    // size every lane's scratch once, before anything is enqueued (growing synchronises the device)
    size_t lane_arena[NL] = {}, lane_scalars[NL] = {}, lane_load[NL] = {}, big_scalars = 0, big_arena = 0;
    unsigned big_seen = 0;
    for (size_t j : order) {
        const ManyJob &jb = jobs[j];
        if (jb.n <= SMALL) {
            int l = 0;
            for (int t = 1; t < NL; ++t)
                if (lane_load[t] < lane_load[l]) l = t;
            lane_of[j] = l;
            lane_load[l] += jb.n + 200000;  // an MSM costs ~0.6 ms of launch latency + 3 ns per point (size sweep), in points
            const size_t a = pk_workspace_bytes(plan_for(c, jb.view, jb.n, 0));
            lane_arena[l] = a > lane_arena[l] ? a : lane_arena[l];
            const size_t sb = jb.on_device ? 0 : jb.n * PLONKISH_CUDA_SCALAR_BYTES;
            lane_scalars[l] = sb > lane_scalars[l] ? sb : lane_scalars[l];
        } else if (jb.on_device) {
            // large resident jobs alternate between the main stream and the big lane (opt-in: PLONKISH_CUDA_BIG_LANE=2)
            const bool alt = big_lane_mode() >= 2 && jb.n <= MAX_POINTS_PER_LAUNCH && (big_seen++ & 1u);
            const size_t a = jb.n <= MAX_POINTS_PER_LAUNCH ? pk_workspace_bytes(plan_for(c, jb.view, jb.n, 0)) : 0;
            if (alt) {
                lane_of[j] = NL;
                big_arena = a > big_arena ? a : big_arena;
            } else if (a && (rc = grow(c->arena, a))) {
                return rc;
            }
        } else {
            big_scalars = jb.n > big_scalars ? jb.n : big_scalars;
        }
    }
    for (int l = 0; l < NL; ++l) {
        if ((rc = grow(c->lanes[l].arena, lane_arena[l])) || (rc = grow(c->lanes[l].scalars, lane_scalars[l]))) return rc;
    }
    if (big_arena && (rc = grow(c->lanes[NL].arena, big_arena))) return rc;
    if (big_scalars && (rc = grow(c->scalars, big_scalars * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_out_list, 0, count * PLONKISH_CUDA_AFFINE_BYTES, c->stream));  // n == 0 entries
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    CUDA_TRY(cudaEventRecord(c->many_start, c->stream));  // lanes start after the memset / the producers of device scalars
    bool lane_used[NL + 1] = {};
    // Enqueue smallest first: the latency-bound small MSMs start while the host is still launching the rest
    // and are done before the large ones need the whole GPU (largest first measured 2 ms slower at k = 20).
    std::reverse(order.begin(), order.end());
    for (size_t j : order) {
        const ManyJob &jb = jobs[j];
        affine *d_out = (affine *)((char *)d_out_list + j * PLONKISH_CUDA_AFFINE_BYTES);
        if (lane_of[j] >= 0) {
            Ctx::Lane &ln = c->lanes[lane_of[j]];
            if (!lane_used[lane_of[j]]) CUDA_TRY(cudaStreamWaitEvent(ln.stream, c->many_start, 0));
            lane_used[lane_of[j]] = true;
            const void *d_sc = jb.scalars;
            if (!jb.on_device) {
                CUDA_TRY(upload(c, ln.scalars.ptr, jb.scalars, jb.n * PLONKISH_CUDA_SCALAR_BYTES, ln.stream));
                d_sc = ln.scalars.ptr;
            }
            MsmPlan plan = plan_for(c, jb.view, jb.n, 0);
            MsmWorkspace w = pk_carve_workspace(plan, ln.arena.ptr);
            w.result = (xyzz *)ln.d_res;
            pk_enqueue_msm(plan, d_sc, jb.view.ptr, w, nullptr, ln.stream);
            PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, ln.stream, (const xyzz *)ln.d_res, 1u, d_out, (xyzz *)nullptr);
        } else {
            xyzz *res = nullptr;
            if (jb.on_device) rc = enqueue_device_msm(c, jb.scalars, jb.view, jb.n, 0, c->stream, &res);
            else rc = enqueue_host_msm(c, jb.scalars, jb.view, jb.n, &res);
            if (rc) return rc;
            PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, c->stream, res, 1u, d_out, (xyzz *)nullptr);
            CUDA_TRY(cudaEventRecord(c->last_done, c->stream));  // the next large MSM reuses the scalar buffer and the arena in stream order
            c->has_last = true;
        }
    }
    CUDA_TRY(cudaGetLastError());
    for (int l = 0; l <= NL; ++l) {
        if (!lane_used[l]) continue;
        CUDA_TRY(cudaEventRecord(c->lanes[l].done, c->lanes[l].stream));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->lanes[l].done, 0));
    }
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_msm_bn254_g1_many(const void *const *scalars_list, const uint64_t *bases_handles, const size_t *ns,
                                               size_t count, void *out_affine64_list) {
    const auto t0 = std::chrono::steady_clock::now();
    if (count == 0) return PLONKISH_CUDA_OK;
    if (!scalars_list || !bases_handles || !ns || !out_affine64_list) return fail(PLONKISH_CUDA_E_INVALID, "msm_many: null argument");
    std::vector<ManyJob> jobs(count);
    int device = -1;
    for (size_t j = 0; j < count; ++j) {
        if (ns[j] == 0) continue;
        if (!scalars_list[j] || !bases_handles[j]) return fail(PLONKISH_CUDA_E_INVALID, "msm_many: MSM %zu lacks scalars or a bases handle", j);
        int dev_j = 0;
        int rc = view_of(bases_handles[j], ns[j], -1, jobs[j].view, &dev_j, "msm_many");
        if (rc) return rc;
        if (device < 0) device = dev_j;
        if (dev_j != device) return fail(PLONKISH_CUDA_E_INVALID, "msm_many: all base slices must live on one device");
        jobs[j].scalars = scalars_list[j];
        jobs[j].n = ns[j];
    }
    if (device < 0) { memset(out_affine64_list, 0, count * PLONKISH_CUDA_AFFINE_BYTES); return PLONKISH_CUDA_OK; }
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_many: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc = grow(c->batch_out, count * PLONKISH_CUDA_AFFINE_BYTES);
    if (rc) return rc;
    if ((rc = enqueue_many(c, jobs, c->batch_out.ptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_affine64_list, c->batch_out.ptr, count * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    timer_report_shared(std::vector<size_t>(ns, ns + count), t0);
    return PLONKISH_CUDA_OK;
}

// Host scalars in, this rank's projective partial out (device memory): the per-rank half of the
// point-sharded MSM (msm.rs:101-111) for the one-process-per-GPU deployment; the ranks then
// all-gather the 128-byte partials and fold them (plonkish_cuda_g1_sum_partials_device).
extern "C" int plonkish_cuda_msm_bn254_g1_host_partial(const void *scalars, uint64_t bases_handle, size_t n, void *d_out_xyzz128) {
    if (!d_out_xyzz128) return fail(PLONKISH_CUDA_E_INVALID, "msm_host_partial: null output");
    if (!bases_handle) return fail(PLONKISH_CUDA_E_INVALID, "msm_host_partial: a registered bases handle is required");
    if (n && !scalars) return fail(PLONKISH_CUDA_E_INVALID, "msm_host_partial: null scalars");
    BasesView view;
    int device = 0;
    int rc = view_of(bases_handle, n, -1, view, &device, "msm_host_partial");
    if (rc) return rc;
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_host_partial: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out_xyzz128, 0, PLONKISH_CUDA_XYZZ_BYTES, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return PLONKISH_CUDA_OK;
    }
    if ((rc = grow(c->scalars, n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    xyzz *res = nullptr;
    if ((rc = enqueue_host_msm(c, scalars, view, n, &res))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_out_xyzz128, res, PLONKISH_CUDA_XYZZ_BYTES, cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_msm_bn254_g1_gather(const void *const *scalar_ptrs, const void *const *base_ptrs, size_t n, void *out_affine64) {
    if (!out_affine64) return fail(PLONKISH_CUDA_E_INVALID, "msm_gather: null output");
    if (n && (!scalar_ptrs || !base_ptrs)) return fail(PLONKISH_CUDA_E_INVALID, "msm_gather: null pointer table");
    std::vector<unsigned char> s(n * PLONKISH_CUDA_SCALAR_BYTES), b(n * PLONKISH_CUDA_AFFINE_BYTES);
    for (size_t i = 0; i < n; ++i) {
        if (!scalar_ptrs[i] || !base_ptrs[i]) return fail(PLONKISH_CUDA_E_INVALID, "msm_gather: null element %zu", i);
        memcpy(&s[i * PLONKISH_CUDA_SCALAR_BYTES], scalar_ptrs[i], PLONKISH_CUDA_SCALAR_BYTES);
        memcpy(&b[i * PLONKISH_CUDA_AFFINE_BYTES], base_ptrs[i], PLONKISH_CUDA_AFFINE_BYTES);
    }
    return plonkish_cuda_msm_bn254_g1(s.data(), b.data(), 0, n, out_affine64);
}

// ------------------------------------------------------------ multi-GPU entry
// One process, G devices (msm.rs:101-114 lifted from rayon threads to GPUs): one host thread per device runs the
// single-GPU host path on its shard — chunk-pipelined upload on the device's copy stream, decompose / sort /
// accumulate / reduce on its compute stream — and leaves one 128-byte projective partial; the partials are gathered
// with ncclAllGather over NVLink on a ncclCommInitAll communicator and device 0 adds them and normalises.  (The
// one-process-per-GPU deployment does the same gather through torch.distributed — plonkish_b200/distributed.py.)
// NCCL is dlopen'ed (libnccl.so.2): the library has no link-time dependency on it and fails loudly without it.
typedef struct ncclComm *pk_ncclComm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(pk_ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(pk_ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, pk_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::vector<pk_ncclComm_t> comms;  // communicator of device g among devices 0..comms.size()-1
};
static std::mutex g_nccl_mu;
static NcclApi g_nccl;

static int nccl_ensure(int n_gpus) {  // caller holds g_nccl_mu
    if (!g_nccl.lib) {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return fail(PLONKISH_CUDA_E_COMM, "msm_multi: cannot load libnccl.so.2: %s", dlerror());
        g_nccl.CommInitAll = (int (*)(pk_ncclComm_t *, int, const int *))dlsym(lib, "ncclCommInitAll");
        g_nccl.CommDestroy = (int (*)(pk_ncclComm_t))dlsym(lib, "ncclCommDestroy");
        g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, pk_ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
        g_nccl.GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
        g_nccl.GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
        g_nccl.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.GetErrorString) {
            dlclose(lib);
            return fail(PLONKISH_CUDA_E_COMM, "msm_multi: libnccl.so.2 lacks an expected symbol");
        }
        g_nccl.lib = lib;
    }
    if ((int)g_nccl.comms.size() == n_gpus) return PLONKISH_CUDA_OK;
    for (pk_ncclComm_t cm : g_nccl.comms) g_nccl.CommDestroy(cm);
    g_nccl.comms.assign(n_gpus, nullptr);
    std::vector<int> devs(n_gpus);
    for (int g = 0; g < n_gpus; ++g) devs[g] = g;
    const int r = g_nccl.CommInitAll(g_nccl.comms.data(), n_gpus, devs.data());
    if (r != 0) {
        g_nccl.comms.clear();
        return fail(PLONKISH_CUDA_E_COMM, "msm_multi: ncclCommInitAll over %d devices failed: %s", n_gpus, g_nccl.GetErrorString(r));
    }
    return PLONKISH_CUDA_OK;
}
static void nccl_shutdown() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    for (pk_ncclComm_t cm : g_nccl.comms) g_nccl.CommDestroy(cm);
    g_nccl.comms.clear();
}

extern "C" int plonkish_cuda_msm_bn254_g1_multi(int n_gpus, const void *scalars, const void *bases, uint64_t bases_handle, size_t n, void *out_affine64) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!out_affine64) return fail(PLONKISH_CUDA_E_INVALID, "msm_multi: null output");
    if (n_gpus < 1 || n_gpus > plonkish_cuda_device_count())
        return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_multi: %d GPUs requested, %d initialised", n_gpus, plonkish_cuda_device_count());
    if (n == 0) { memset(out_affine64, 0, PLONKISH_CUDA_AFFINE_BYTES); return PLONKISH_CUDA_OK; }
    if (!scalars || (!bases && !bases_handle)) return fail(PLONKISH_CUDA_E_INVALID, "msm_multi: null input");
    BasesEntry entry;
    if (bases_handle) {
        if (!lookup_bases(bases_handle, entry)) return fail(PLONKISH_CUDA_E_INVALID, "msm_multi: unknown bases handle");
        if (entry.n_shards != n_gpus || entry.n != n) return fail(PLONKISH_CUDA_E_INVALID, "msm_multi: handle was sharded for %d GPUs / %zu points", entry.n_shards, entry.n);
    }
    const size_t per = (n + n_gpus - 1) / n_gpus;  // msm.rs:101
    std::vector<Ctx *> cs(n_gpus);
    for (int g = 0; g < n_gpus; ++g) {
        cs[g] = ctx_for(g);
        if (!cs[g]) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_multi: device %d not initialised", g);
    }
    std::lock_guard<std::mutex> nccl_lock(g_nccl_mu);  // one multi-GPU call at a time (the communicators are shared)
    if (n_gpus > 1) Stager::set_shards(n_gpus);
    int rc = n_gpus > 1 ? nccl_ensure(n_gpus) : PLONKISH_CUDA_OK;
    if (rc) return rc;
    struct LockAll {  // every device context for the whole call, in device order
        std::vector<Ctx *> &v;
        explicit LockAll(std::vector<Ctx *> &cs_) : v(cs_) { for (Ctx *c : v) c->mu.lock(); }
        ~LockAll() { for (size_t g = v.size(); g-- > 0;) v[g]->mu.unlock(); }
    } lock_all(cs);
    // per device: [0, G) gathered partials, [G] this device's own
    std::vector<int> rcs(n_gpus, 0);
    std::vector<std::string> errs(n_gpus);
    auto shard = [&](int g) {
        Ctx *c = cs[g];
        auto run = [&]() -> int {
            CUDA_TRY(cudaSetDevice(g));
            int r = grow(c->partials, (size_t)(n_gpus + 1) * PLONKISH_CUDA_XYZZ_BYTES);
            if (r) return r;
            char *send = (char *)c->partials.ptr + (size_t)n_gpus * PLONKISH_CUDA_XYZZ_BYTES;
            const size_t beg = (size_t)g * per;
            const size_t cnt = beg >= n ? 0 : (beg + per <= n ? per : n - beg);
            if (cnt == 0) {  // trailing shard of a short MSM: the identity
                if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
                CUDA_TRY(cudaMemsetAsync(send, 0, PLONKISH_CUDA_XYZZ_BYTES, c->stream));
                return PLONKISH_CUDA_OK;
            }
            if ((r = grow(c->scalars, cnt * PLONKISH_CUDA_SCALAR_BYTES))) return r;
            BasesView view;
            if (bases_handle) {
                view.ptr = entry.d_ptr[g];
                view.table_c = entry.table_c[g];
                view.stride = entry.shard_n[g];
                view.keep = entry.keep[g];
            } else {
                if ((r = grow(c->bases_tmp, cnt * PLONKISH_CUDA_AFFINE_BYTES))) return r;
                if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
                CUDA_TRY(upload(c, c->bases_tmp.ptr, (const char *)bases + beg * PLONKISH_CUDA_AFFINE_BYTES, cnt * PLONKISH_CUDA_AFFINE_BYTES, c->stream));
                view.ptr = c->bases_tmp.ptr;
            }
            xyzz *res = nullptr;
            if ((r = enqueue_host_msm(c, (const char *)scalars + beg * PLONKISH_CUDA_SCALAR_BYTES, view, cnt, &res))) return r;
            CUDA_TRY(cudaMemcpyAsync(send, res, PLONKISH_CUDA_XYZZ_BYTES, cudaMemcpyDeviceToDevice, c->stream));
            return PLONKISH_CUDA_OK;
        };
        rcs[g] = run();
        if (rcs[g]) errs[g] = t_last_error;  // the message lives in the worker's thread-local slot
    };
    {
        std::vector<std::thread> workers;
        for (int g = 1; g < n_gpus; ++g) workers.emplace_back(shard, g);
        shard(0);
        for (auto &w : workers) w.join();
    }
    for (int g = 0; g < n_gpus; ++g) {
        if (rcs[g]) { t_last_error = errs[g]; return rcs[g]; }
    }
    if (n_gpus > 1) {
        int r = g_nccl.GroupStart();
        for (int g = 0; r == 0 && g < n_gpus; ++g) {
            char *recv = (char *)cs[g]->partials.ptr;
            r = g_nccl.AllGather(recv + (size_t)n_gpus * PLONKISH_CUDA_XYZZ_BYTES, recv, PLONKISH_CUDA_XYZZ_BYTES, /* ncclChar */ 0, g_nccl.comms[g], cs[g]->stream);
        }
        const int r2 = g_nccl.GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return fail(PLONKISH_CUDA_E_COMM, "msm_multi: ncclAllGather of the partials failed: %s", g_nccl.GetErrorString(r));
    } else {
        CUDA_TRY(cudaSetDevice(0));
        CUDA_TRY(cudaMemcpyAsync(cs[0]->partials.ptr, (char *)cs[0]->partials.ptr + PLONKISH_CUDA_XYZZ_BYTES, PLONKISH_CUDA_XYZZ_BYTES, cudaMemcpyDeviceToDevice, cs[0]->stream));
    }
    CUDA_TRY(cudaSetDevice(0));
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, cs[0]->stream, (const xyzz *)cs[0]->partials.ptr, (u32)n_gpus, (affine *)cs[0]->d_out, (xyzz *)nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cs[0]->h_out, cs[0]->d_out, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, cs[0]->stream));
    for (int g = n_gpus - 1; g >= 0; --g) {  // every device is idle again when the call returns
        CUDA_TRY(cudaSetDevice(g));
        if ((rc = mark_done(cs[g], cs[g]->stream))) return rc;
        CUDA_TRY(cudaStreamSynchronize(cs[g]->stream));
    }
    memcpy(out_affine64, cs[0]->h_out, PLONKISH_CUDA_AFFINE_BYTES);
    timer_report(n, t0);
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------------ per-stage timing
extern "C" int plonkish_cuda_msm_profile_device(int device, const void *d_scalars, const void *d_bases, uint64_t bases_handle, size_t n,
                                                uint32_t window_bits, void *d_out_affine64, double stage_ms[9]) {
    if (!d_scalars || (!d_bases && !bases_handle) || !stage_ms || n == 0 || n > MAX_POINTS_PER_LAUNCH) return fail(PLONKISH_CUDA_E_INVALID, "msm_profile: bad argument");
    if (window_bits && (window_bits < 8 || window_bits > 16)) return fail(PLONKISH_CUDA_E_INVALID, "msm_profile: window_bits must be 0 or 8..16");
    BasesView view;
    view.ptr = d_bases;
    if (bases_handle) {
        int rc0 = view_of(bases_handle, n, -1, view, &device, "msm_profile");
        if (rc0) return rc0;
    }
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_profile: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    CUDA_TRY(cudaDeviceSynchronize());  // inputs may come from another stream; this call is synchronous anyway
    StageMarks marks;
    cudaEvent_t ev[10];
    for (int i = 0; i < 10; ++i) CUDA_TRY(cudaEventCreate(&ev[i]));
    for (int i = 0; i < 9; ++i) marks.ev[i] = ev[i];
    xyzz *res = nullptr;
    int rc = enqueue_device_msm(c, d_scalars, view, n, window_bits, c->stream, &res, &marks);
    if (rc) return rc;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, c->stream, res, 1u, (affine *)(d_out_affine64 ? d_out_affine64 : c->d_out), (xyzz *)nullptr);
    CUDA_TRY(cudaEventRecord(ev[9], c->stream));
    CUDA_TRY(cudaGetLastError());
    rc = mark_done(c, c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 9; ++i) {
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
        stage_ms[i] = ms;
    }
    for (int i = 0; i < 10; ++i) cudaEventDestroy(ev[i]);
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------------------ plan probe
extern "C" int plonkish_cuda_msm_plan(int device, size_t n, uint32_t window_bits, uint64_t bases_handle, uint32_t out[8]) {
    if (!out || n == 0 || n > MAX_POINTS_PER_LAUNCH) return fail(PLONKISH_CUDA_E_INVALID, "msm_plan: bad argument");
    BasesView view;
    if (bases_handle) {
        int rc0 = view_of(bases_handle, n, -1, view, &device, "msm_plan");
        if (rc0) return rc0;
    }
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_plan: device %d not initialised", device);
    MsmPlan p = plan_for(c, view, n, window_bits);
    out[0] = p.c; out[1] = p.W; out[2] = p.hi_bits; out[3] = p.lo_bits; out[4] = p.idx_bits;
    out[5] = p.tile; out[6] = p.L; out[7] = p.nthreads1;
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------- synthetic known-dlog bases
static const int SYNTH_RUN = 16;

// k * G by double-and-add, k a 64-bit integer.
__device__ xyzz mul_generator_u64(unsigned long long k) {
    affine g;
    g.x = fq_one();
    g.y = fq_dbl(fq_one());
    xyzz r = xyzz_identity();
    for (int b = 63; b >= 0; --b) {
        r = xyzz_double(r);
        if ((k >> b) & 1ull) xyzz_madd(r, g.x, g.y);
    }
    return r;
}

__global__ void k_synth_step(affine *d_step, unsigned long long step) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_step = xyzz_to_affine(mul_generator_u64(step));
}

// Thread t writes out[t*16 .. t*16+16): walks P += step*G and normalises its 16
// points with one shared inversion (Montgomery's trick).
__global__ void __launch_bounds__(128) k_synth_bases(affine *__restrict__ out, unsigned long long first, unsigned long long n,
                                                     unsigned long long a, unsigned long long step, const affine *__restrict__ d_step) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long i0 = t * SYNTH_RUN;
    if (i0 >= n) return;
    const int cnt = (n - i0 < (unsigned long long)SYNTH_RUN) ? (int)(n - i0) : SYNTH_RUN;
    const affine d = *d_step;
    xyzz p = mul_generator_u64(a + (first + i0) * step);
    xyzz pts[SYNTH_RUN];
    fe prefix[SYNTH_RUN];
    fe acc = fq_one();
    for (int j = 0; j < cnt; ++j) {
        pts[j] = p;
        prefix[j] = acc;
        if (!xyzz_is_identity(p)) acc = fq_mul(acc, fq_mul(p.zz, p.zzz));
        xyzz_madd(p, d.x, d.y);
    }
    fe inv = fq_inv_fast(acc);
    for (int j = cnt - 1; j >= 0; --j) {
        affine r;
        if (xyzz_is_identity(pts[j])) {
            r.x = fe_zero(); r.y = fe_zero();
        } else {
            const fe i = fq_mul(inv, prefix[j]);  // 1 / (zz*zzz)
            inv = fq_mul(inv, fq_mul(pts[j].zz, pts[j].zzz));
            r.x = fq_mul(pts[j].x, fq_mul(i, pts[j].zzz));
            r.y = fq_mul(pts[j].y, fq_mul(i, pts[j].zz));
        }
        uint4 *q = reinterpret_cast<uint4 *>(out + i0 + j);
        store_fe(q, r.x);
        store_fe(q + 2, r.y);
    }
}

extern "C" int plonkish_cuda_synth_bases_device(int device, void *d_out, size_t first, size_t n, uint64_t a, uint64_t step, void *cuda_stream) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "synth_bases: device %d not initialised", device);
    if (!d_out || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "synth_bases: bad argument");
    if (a >> 31 || step >> 31 || (first + n) >> 32) return fail(PLONKISH_CUDA_E_INVALID, "synth_bases: a, step < 2^31 and first + n < 2^32 required");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    cudaStream_t stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, as in the CUDA runtime
    affine *d_step = (affine *)((char *)c->d_out + 192);
    PK_LAUNCH(k_synth_step, dim3(1), dim3(32), 0, stream, d_step, (unsigned long long)step);
    const unsigned long long threads = (n + SYNTH_RUN - 1) / SYNTH_RUN;
    PK_LAUNCH(k_synth_bases, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, stream, (affine *)d_out,
              (unsigned long long)first, (unsigned long long)n, (unsigned long long)a, (unsigned long long)step, d_step);
    CUDA_TRY(cudaGetLastError());
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------ integer-pipe microbenchmarks
// 16 independent 64-bit accumulators per thread, each fed by mad.wide.u32: no
// dependent chain shorter than 16 instructions, so the multiplier pipe is the limit.
__global__ void __launch_bounds__(256) k_bench_imad_wide(unsigned long long *out, u32 iters, u32 seed) {
    unsigned long long acc[16];
    u32 a = seed + threadIdx.x * 2654435761u, b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (unsigned long long)(a + k) << 7;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + (u32)k), "r"(b));
        b += 0x9e3779b9u;
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 independent 32-bit accumulators fed by mad.lo.u32 (plain IMAD).
__global__ void __launch_bounds__(256) k_bench_imad32(u32 *out, u32 iters, u32 seed) {
    u32 acc[16];
    u32 a = seed + threadIdx.x * 2654435761u, b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = a + k;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(b), "r"(a));
    }
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Four independent 8-word accumulators fed by the library's own carry-chain block
// (cmad8: mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X), the exact instruction
// K3's Montgomery products are made of.  4 wide multiply-adds per cmad8.
__global__ void __launch_bounds__(256) k_bench_chain(u32 *out, u32 iters, u32 seed) {
    u32 acc[4][8];
    u32 a = seed + threadIdx.x * 2654435761u, b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j][k] = a + 8 * j + k;
    u32 carries = 0;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) carries += cmad8(acc[j], a + j, a ^ 0x55u, b + j, b ^ 0xaau, b);
        b += 0x9e3779b9u;
    }
    u32 s = carries;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) s ^= acc[j][k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Narrow carry chains: 8 x mad.lo.cc then 8 x mad.hi.cc per accumulator row (the unfused form of
// a row of partial products: 16 32-bit IMADs with carry instead of 8 IMAD.WIDE).
__global__ void __launch_bounds__(256) k_bench_narrow_chain(u32 *out, u32 iters, u32 seed) {
    u32 acc[2][10];
    u32 a[8];
    u32 b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 2654435761u + k * 97u;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[j][k] = a[k & 7] + j;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            u32 *c = acc[j];
            asm volatile("mad.lo.cc.u32  %0, %10, %18, %0;\n\t"
                "madc.lo.cc.u32 %1, %11, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %12, %18, %2;\n\t"
                "madc.lo.cc.u32 %3, %13, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %14, %18, %4;\n\t"
                "madc.lo.cc.u32 %5, %15, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %16, %18, %6;\n\t"
                "madc.lo.cc.u32 %7, %17, %18, %7;\n\t"
                "addc.u32       %8, %8, 0;\n\t"
                "mad.hi.cc.u32  %1, %10, %18, %1;\n\t"
                "madc.hi.cc.u32 %2, %11, %18, %2;\n\t"
                "madc.hi.cc.u32 %3, %12, %18, %3;\n\t"
                "madc.hi.cc.u32 %4, %13, %18, %4;\n\t"
                "madc.hi.cc.u32 %5, %14, %18, %5;\n\t"
                "madc.hi.cc.u32 %6, %15, %18, %6;\n\t"
                "madc.hi.cc.u32 %7, %16, %18, %7;\n\t"
                "madc.hi.cc.u32 %8, %17, %18, %8;\n\t"
                "addc.u32       %9, %9, 0;"
                : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8]), "+r"(c[9])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b + (u32)j));
        }
        b += 0x9e3779b9u;
    }
    u32 sres = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 10; ++k) sres ^= acc[j][k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sres;
}
// Same row as 8 fused pairs (the even/odd form the library uses: two cmad8 per row).
__global__ void __launch_bounds__(256) k_bench_wide_row(u32 *out, u32 iters, u32 seed) {
    u32 ev[2][8], od[2][8];
    u32 a[8];
    u32 b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 2654435761u + k * 97u;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) { ev[j][k] = a[k] + j; od[j][k] = a[k] ^ j; }
    u32 carries = 0;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            carries += cmad8(ev[j], a[0], a[2], a[4], a[6], b + (u32)j);
            carries += cmad8(od[j], a[1], a[3], a[5], a[7], b + (u32)j);
        }
        b += 0x9e3779b9u;
    }
    u32 sres = carries;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) sres ^= ev[j][k] ^ od[j][k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sres;
}
// out[0] rows per second in the narrow form (16 IMAD + 2 add per row), out[1] in the fused form
// (8 IMAD.WIDE per row): which spelling of a row of partial products the multiplier pipe prefers.
extern "C" int plonkish_cuda_bench_row_forms(int device, double out[2]) {
    Ctx *c = ctx_for(device);
    if (!c || !out) return fail(PLONKISH_CUDA_E_INVALID, "bench_row_forms: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 8, threads = 256;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const u32 iters = 2048;
    for (int v = 0; v < 2; ++v) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CUDA_TRY(cudaEventRecord(e0, c->stream));
            if (v == 0) PK_LAUNCH(k_bench_narrow_chain, dim3(blocks), dim3(threads), 0, c->stream, (u32 *)scratch, iters, 11u + rep);
            if (v == 1) PK_LAUNCH(k_bench_wide_row, dim3(blocks), dim3(threads), 0, c->stream, (u32 *)scratch, iters, 11u + rep);
            CUDA_TRY(cudaEventRecord(e1, c->stream));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        }
        out[v] = (double)blocks * threads * 2.0 * iters / (ms * 1e-3);
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// Two independent Montgomery products per thread per iteration (the K3 instruction mix).
__global__ void __launch_bounds__(256) k_bench_fq_mul(uint4 *out, u32 iters, u32 seed) {
    // every chain value depends on the thread index: a warp-uniform chain would be moved to the
    // uniform datapath (UIMAD) by ptxas and would not load the vector multiplier pipe at all
    fe x = fq_one(), y = fq_one(), z = fq_one();
    x.l[0] ^= seed + threadIdx.x; y.l[1] ^= blockIdx.x + 977u * threadIdx.x; z.l[2] ^= seed + 31u * threadIdx.x;
    x.l[7] &= 0x0fffffffu; y.l[7] &= 0x0fffffffu; z.l[7] &= 0x0fffffffu;
    for (u32 it = 0; it < iters; ++it) {
        x = fq_mul(x, z);
        y = fq_mul(y, z);
    }
    fe r = fq_add(x, y);
    store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), r);
}

// Occupancy experiment: the library's fq_mul stream with `warps_per_sm` resident warps
// (128-thread blocks, dynamic shared memory as the limiter).  Returns products per second.
__global__ void __launch_bounds__(128) k_bench_fq_mul_occ(uint4 *out, u32 iters, u32 seed) {
    extern __shared__ unsigned char pad_smem[];
    fe x = fq_one(), y = fq_one(), z = fq_one();
    x.l[0] ^= seed + threadIdx.x; y.l[1] ^= blockIdx.x + 977u * threadIdx.x; z.l[2] ^= seed + 31u * threadIdx.x;
    x.l[7] &= 0x0fffffffu; y.l[7] &= 0x0fffffffu; z.l[7] &= 0x0fffffffu;
    if (seed == 0xffffffffu) pad_smem[threadIdx.x] = 1;  // keep the allocation alive
    for (u32 it = 0; it < iters; ++it) {
        x = fq_mul(x, z);
        y = fq_mul(y, z);
    }
    fe r = fq_add(x, y);
    store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), r);
}
extern "C" int plonkish_cuda_bench_fq_mul_occupancy(int device, int warps_per_sm, double *out_per_s) {
    Ctx *c = ctx_for(device);
    if (!c || !out_per_s || warps_per_sm < 4 || warps_per_sm > 64 || warps_per_sm % 4) return fail(PLONKISH_CUDA_E_INVALID, "bench_fq_mul_occupancy: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const int blocks_per_sm = warps_per_sm / 4;
    const size_t smem = blocks_per_sm >= 16 ? 0 : (size_t)(220 * 1024 / blocks_per_sm) - 2048;
    CUDA_TRY(cudaFuncSetAttribute(k_bench_fq_mul_occ, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    const unsigned blocks = (unsigned)c->sm_count * blocks_per_sm * 4;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * 128 * 32));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float ms = 0;
    const u32 iters = 256;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_fq_mul_occ, dim3(blocks), dim3(128), smem, c->stream, (uint4 *)scratch, iters, 7u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    CUDA_TRY(cudaGetLastError());
    *out_per_s = (double)blocks * 128 * 2.0 * iters / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

__global__ void __launch_bounds__(128) k_bench_inv(uint4 *out, u32 iters, u32 seed, int fast) {
    fe x = fq_one();
    x.l[0] ^= seed + threadIdx.x; x.l[1] ^= blockIdx.x; x.l[7] &= 0x0fffffffu;
    for (u32 it = 0; it < iters; ++it) {
        x = fast ? fq_inv_fast(x) : fq_inv(x);
        x.l[0] ^= it;  x.l[7] &= 0x0fffffffu;
    }
    store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), x);
}
// Register-resident mixed-addition streams (no memory traffic, no bucket logic), 128-thread
// blocks like k_accumulate: variant 0 = one XYZZ accumulator per thread, 1 = two independent
// accumulators per thread, 2 = one dependent fq_mul chain per thread; MINB = blocks per SM the
// register allocation is capped for (4 -> 128 registers, 2 -> uncapped).
template <int VARIANT, int MINB>
__global__ void __launch_bounds__(128, MINB) k_bench_madd(uint4 *out, u32 iters, u32 seed) {
    fe x = fq_one(), y = fq_dbl(fq_one());
    x.l[0] ^= seed + threadIdx.x; y.l[1] ^= blockIdx.x + 977u * threadIdx.x;
    x.l[7] &= 0x0fffffffu; y.l[7] &= 0x0fffffffu;
    xyzz a = xyzz_identity(), b = xyzz_identity();
    if (VARIANT == 2) {
        for (u32 it = 0; it < iters * 10; ++it) x = fq_mul(x, y);
        store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), x);
        return;
    }
    for (u32 it = 0; it < iters; ++it) {
        xyzz_madd<MulInline>(a, x, y);
        if (VARIANT == 1) xyzz_madd<MulInline>(b, y, x);
        x.l[0] += 2; y.l[0] += 6;
    }
    fe r = fq_add(a.x, fq_add(a.zz, fq_add(b.y, b.zzz)));
    store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), r);
}
// out[v] = field products per second inside the stream of variant v (10 per mixed addition).
extern "C" int plonkish_cuda_bench_madd(int device, double out[5]) {
    Ctx *c = ctx_for(device);
    if (!c || !out) return fail(PLONKISH_CUDA_E_INVALID, "bench_madd: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 4 * 4, threads = 128;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 32));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const u32 iters = 64;
    for (int v = 0; v < 5; ++v) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CUDA_TRY(cudaEventRecord(e0, c->stream));
            if (v == 0) PK_LAUNCH((k_bench_madd<0, 4>), dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 3u + rep);
            if (v == 1) PK_LAUNCH((k_bench_madd<1, 4>), dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 3u + rep);
            if (v == 2) PK_LAUNCH((k_bench_madd<2, 4>), dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 3u + rep);
            if (v == 3) PK_LAUNCH((k_bench_madd<0, 2>), dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 3u + rep);
            if (v == 4) PK_LAUNCH((k_bench_madd<1, 2>), dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 3u + rep);
            CUDA_TRY(cudaEventRecord(e1, c->stream));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double madds = (double)blocks * threads * iters * ((v == 1 || v == 4) ? 2.0 : 1.0);
        out[v] = madds * 10.0 / (ms * 1e-3);
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// Inversions per second: out[0] safegcd (fq_inv_fast), out[1] Fermat ladder (fq_inv).
extern "C" int plonkish_cuda_bench_inversion(int device, double out[2]) {
    Ctx *c = ctx_for(device);
    if (!c || !out) return fail(PLONKISH_CUDA_E_INVALID, "bench_inversion: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 8, threads = 128;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 32));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    for (int fast = 1; fast >= 0; --fast) {
        float ms = 0;
        const u32 iters = 8;
        for (int rep = 0; rep < 2; ++rep) {
            CUDA_TRY(cudaEventRecord(e0, c->stream));
            PK_LAUNCH(k_bench_inv, dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, iters, 5u + rep, fast);
            CUDA_TRY(cudaEventRecord(e1, c->stream));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        }
        out[fast ? 0 : 1] = (double)blocks * threads * iters / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// ---- FP64 pipe probes: could 52-bit double-precision FMAs carry part of the field arithmetic next to the integer pipe?
// 16 independent accumulators per thread, fma.rz.f64 each (DFMA).
__global__ void __launch_bounds__(256) k_bench_dfma(double *out, u32 iters, u32 seed) {
    double acc[16];
    const double a = 1.0 + (double)(seed + threadIdx.x) * 1e-9, b = 1.0 - (double)(blockIdx.x + 1) * 1e-9;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (double)k + a;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[k]) : "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// The same DFMA stream interleaved one to one with independent mad.wide.u32: do the two pipes issue side by side?
__global__ void __launch_bounds__(256) k_bench_dfma_imad(double *out, u32 iters, u32 seed) {
    double acc[8];
    unsigned long long iacc[8];
    const double a = 1.0 + (double)(seed + threadIdx.x) * 1e-9, b = 1.0 - (double)(blockIdx.x + 1) * 1e-9;
    u32 x = seed + threadIdx.x * 2654435761u, y = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] = (double)k + a; iacc[k] = (unsigned long long)(x + k) << 7; }
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[k]) : "d"(b), "d"(a));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(iacc[k]) : "r"(x + (u32)k), "r"(y));
        }
        y += 0x9e3779b9u;
    }
    double s = 0;
    unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += acc[k]; t ^= iacc[k]; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(t & 0xffff);
}
// out[0] = DFMA per second alone, out[1] = DFMA per second and out[2] = mad.wide.u32 per second when interleaved 1:1.
extern "C" int plonkish_cuda_bench_fp64_pipe(int device, double out[3]) {
    Ctx *c = ctx_for(device);
    if (!c || !out) return fail(PLONKISH_CUDA_E_INVALID, "bench_fp64_pipe: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 8, threads = 256;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const u32 iters = 2048;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_dfma, dim3(blocks), dim3(threads), 0, c->stream, (double *)scratch, iters, 41u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[0] = (double)blocks * threads * 16.0 * iters / (ms * 1e-3);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_dfma_imad, dim3(blocks), dim3(threads), 0, c->stream, (double *)scratch, iters, 43u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[1] = out[2] = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// ---- is the accumulate loop bound by the multiplier pipe or by instruction issue?  8 independent mad.wide.u32 per
// iteration, interleaved with 8 * R independent 32-bit adds (IADD3 on the ALU pipe) on other registers.
template <int R>
__global__ void __launch_bounds__(256) k_bench_imad_alu_mix(unsigned long long *out, u32 iters, u32 seed) {
    unsigned long long acc[8];
    u32 extra[8 * (R > 0 ? R : 1)];
    u32 a = seed + threadIdx.x * 2654435761u, b = seed ^ (blockIdx.x * 40503u + 12345u + threadIdx.x * 7919u);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (unsigned long long)(a + k) << 7;
#pragma unroll
    for (int k = 0; k < 8 * (R > 0 ? R : 1); ++k) extra[k] = a ^ (k * 977u);
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + (u32)k), "r"(b));
#pragma unroll
            for (int j = 0; j < R; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(extra[k * R + j]) : "r"(b));
        }
        b += 0x9e3779b9u;
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= acc[k];
#pragma unroll
    for (int k = 0; k < 8 * (R > 0 ? R : 1); ++k) s += extra[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// out[R] = mad.wide.u32 per second with R adds issued next to every multiply, R = 0..3.
extern "C" int plonkish_cuda_bench_issue_mix(int device, double out[4]) {
    Ctx *c = ctx_for(device);
    if (!c || !out) return fail(PLONKISH_CUDA_E_INVALID, "bench_issue_mix: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 8, threads = 256;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const u32 iters = 4096;
    for (int r = 0; r < 4; ++r) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CUDA_TRY(cudaEventRecord(e0, c->stream));
            if (r == 0) PK_LAUNCH((k_bench_imad_alu_mix<0>), dim3(blocks), dim3(threads), 0, c->stream, (unsigned long long *)scratch, iters, 51u + rep);
            if (r == 1) PK_LAUNCH((k_bench_imad_alu_mix<1>), dim3(blocks), dim3(threads), 0, c->stream, (unsigned long long *)scratch, iters, 51u + rep);
            if (r == 2) PK_LAUNCH((k_bench_imad_alu_mix<2>), dim3(blocks), dim3(threads), 0, c->stream, (unsigned long long *)scratch, iters, 51u + rep);
            if (r == 3) PK_LAUNCH((k_bench_imad_alu_mix<3>), dim3(blocks), dim3(threads), 0, c->stream, (unsigned long long *)scratch, iters, 51u + rep);
            CUDA_TRY(cudaEventRecord(e1, c->stream));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        }
        out[r] = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// ---- FP64-pipe mixed additions (dpfq.cuh): register-resident stream, alone and next to the integer-pipe stream
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_bench_dp_madd(uint4 *out, u32 iters, u32 seed) {
    fe x = fq_one(), y = fq_dbl(fq_one());
    x.l[0] ^= seed + threadIdx.x; y.l[1] ^= blockIdx.x + 977u * threadIdx.x;
    x.l[7] &= 0x0fffffffu; y.l[7] &= 0x0fffffffu;
    dxyzz a = dxyzz_identity();
    for (u32 it = 0; it < iters; ++it) {
        dxyzz_madd(a, x, y);
        x.l[0] += 2; y.l[0] += 6;
    }
    const xyzz w = dxyzz_to_words(a);
    store_fe(out + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x), fq_add(w.x, fq_add(w.zz, fq_add(w.y, w.zzz))));
}
// out[0] = FP64-pipe mixed additions per second alone (dp_blocks_per_sm blocks of 128 threads per SM), out[1] = integer-pipe
// mixed additions per second alone (int_blocks_per_sm blocks per SM), out[2] / out[3] = the same two when both kernels run
// at the same time on two streams (each sized to leave room for the other), out[4] = wall ms of the concurrent run.
extern "C" int plonkish_cuda_bench_dp_madd(int device, int dp_blocks_per_sm, int int_blocks_per_sm, double out[5]) {
    Ctx *c = ctx_for(device);
    if (!c || !out || dp_blocks_per_sm < 1 || dp_blocks_per_sm > 4 || int_blocks_per_sm < 1 || int_blocks_per_sm > 4) return fail(PLONKISH_CUDA_E_INVALID, "bench_dp_madd: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned dp_blocks = (unsigned)c->sm_count * dp_blocks_per_sm, int_blocks = (unsigned)c->sm_count * int_blocks_per_sm;
    void *s1 = nullptr, *s2 = nullptr;
    CUDA_TRY(cudaMalloc(&s1, (size_t)dp_blocks * 128 * 32));
    CUDA_TRY(cudaMalloc(&s2, (size_t)int_blocks * 128 * 32));
    cudaEvent_t e[6];
    for (auto &ev : e) CUDA_TRY(cudaEventCreate(&ev));
    cudaStream_t sa = c->stream, sb = c->lanes[0].stream;
    const u32 it_dp = 96, it_int = 192;
    auto launch_dp = [&](cudaStream_t st, u32 seed) {
        if (dp_blocks_per_sm >= 2) PK_LAUNCH((k_bench_dp_madd<2>), dim3(dp_blocks), dim3(128), 0, st, (uint4 *)s1, it_dp, seed);
        else PK_LAUNCH((k_bench_dp_madd<1>), dim3(dp_blocks), dim3(128), 0, st, (uint4 *)s1, it_dp, seed);
    };
    auto launch_int = [&](cudaStream_t st, u32 seed) { PK_LAUNCH((k_bench_madd<0, 4>), dim3(int_blocks), dim3(128), 0, st, (uint4 *)s2, it_int, seed); };
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e[0], sa));
        launch_dp(sa, 3u + rep);
        CUDA_TRY(cudaEventRecord(e[1], sa));
        CUDA_TRY(cudaEventSynchronize(e[1]));
        CUDA_TRY(cudaEventElapsedTime(&ms, e[0], e[1]));
    }
    out[0] = (double)dp_blocks * 128 * it_dp / (ms * 1e-3);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e[0], sa));
        launch_int(sa, 5u + rep);
        CUDA_TRY(cudaEventRecord(e[1], sa));
        CUDA_TRY(cudaEventSynchronize(e[1]));
        CUDA_TRY(cudaEventElapsedTime(&ms, e[0], e[1]));
    }
    out[1] = (double)int_blocks * 128 * it_int / (ms * 1e-3);
    float ms_dp = 0, ms_int = 0, ms_all = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaEventRecord(e[0], sa));
        CUDA_TRY(cudaStreamWaitEvent(sb, e[0], 0));
        CUDA_TRY(cudaEventRecord(e[2], sb));
        launch_int(sb, 9u + rep);
        CUDA_TRY(cudaEventRecord(e[3], sb));
        CUDA_TRY(cudaEventRecord(e[4], sa));
        launch_dp(sa, 7u + rep);
        CUDA_TRY(cudaEventRecord(e[1], sa));
        CUDA_TRY(cudaStreamWaitEvent(sa, e[3], 0));
        CUDA_TRY(cudaEventRecord(e[5], sa));
        CUDA_TRY(cudaEventSynchronize(e[5]));
        CUDA_TRY(cudaEventElapsedTime(&ms_dp, e[4], e[1]));
        CUDA_TRY(cudaEventElapsedTime(&ms_int, e[2], e[3]));
        CUDA_TRY(cudaEventElapsedTime(&ms_all, e[0], e[5]));
    }
    out[2] = (double)dp_blocks * 128 * it_dp / (ms_dp * 1e-3);
    out[3] = (double)int_blocks * 128 * it_int / (ms_int * 1e-3);
    out[4] = ms_all;
    CUDA_TRY(cudaGetLastError());
    for (auto &ev : e) cudaEventDestroy(ev);
    CUDA_TRY(cudaFree(s1));
    CUDA_TRY(cudaFree(s2));
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_bench_integer_pipe(int device, double out[6]) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bench_integer_pipe: device %d not initialised", device);
    if (!out) return fail(PLONKISH_CUDA_E_INVALID, "bench_integer_pipe: null output");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const unsigned blocks = (unsigned)c->sm_count * 8, threads = 256;
    void *scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, (size_t)blocks * threads * 32));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float ms = 0;
    const u32 it_wide = 4096, it_mul = 256;
    for (int rep = 0; rep < 3; ++rep) {  // last repetition is the one reported
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_imad_wide, dim3(blocks), dim3(threads), 0, c->stream, (unsigned long long *)scratch, it_wide, 17u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[0] = (double)blocks * threads * 16.0 * it_wide / (ms * 1e-3);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_fq_mul, dim3(blocks), dim3(threads), 0, c->stream, (uint4 *)scratch, it_mul, 29u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[1] = (double)blocks * threads * 2.0 * it_mul / (ms * 1e-3);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_imad32, dim3(blocks), dim3(threads), 0, c->stream, (u32 *)scratch, it_wide, 31u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[4] = (double)blocks * threads * 16.0 * it_wide / (ms * 1e-3);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        PK_LAUNCH(k_bench_chain, dim3(blocks), dim3(threads), 0, c->stream, (u32 *)scratch, it_wide, 37u + rep);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    out[5] = (double)blocks * threads * 16.0 * it_wide / (ms * 1e-3);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, c->dev);
    out[2] = clock_khz / 1000.0;  // the device's maximum SM clock; bench.py samples the live one
    out[3] = c->sm_count;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(scratch));
    return PLONKISH_CUDA_OK;
}

// ------------------------------------------------------------------ test hooks
// Element-wise probes of the device arithmetic, so the parity tests can pin the PTX
// carry chains and the group law element by element without going through an MSM.
__global__ void k_debug_field(int op, const fe *a, const fe *b, fe *out, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe r;
    switch (op) {
        case 0: r = fq_mul(a[i], b[i]); break;
        case 1: r = fq_add(a[i], b[i]); break;
        case 2: r = fq_sub(a[i], b[i]); break;
        case 3: r = fr_to_canonical(a[i]); break;
        case 4: r = fq_inv(a[i]); break;
        case 5: r = mont_mul<FrMod>(a[i], b[i]); break;
        case 7: r = fq_inv_fast(a[i]); break;
        case 8: r = fq_mul_sum(a[i], b[i], b[i], a[(i + 1) % n]); break;        // a*b + b*a'  (a' = next element of a)
        case 9: r = mont_mul_sum<FrMod>(a[i], a[i], b[i], b[(i + 1) % n]); break;  // Fr: a^2 + b*b'
        case 10: r = fq_sqr(a[i]); break;                                          // symmetric squaring
        case 11: r = mont_sqr<FrMod>(a[i]); break;
        case 12: r = dp_to_mont256(dp_mul(dp_from_mont256(a[i]), dp_from_mont256(b[i]))); break;  // FP64-pipe product, memory form in and out
        case 13: r = dp_to_mont256(dp_sqr(dp_from_mont256(a[i]))); break;
        case 14: r = dp_to_words(dp_from_words(a[i])); break;                                   // limb conversions only
        default: r = fq_neg(a[i]); break;
    }
    out[i] = r;
}
// op 0: out = xyzz(a) + affine(b); 1: out = xyzz(a) + xyzz(b); 2: out = 2*xyzz(a);
// op 3: out[0..64) = to_affine(xyzz(a)).  a, b, out are arrays of 128-byte slots.
__global__ void k_debug_point(int op, const xyzz *a, const xyzz *b, xyzz *out, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz r = a[i];
    switch (op) {
        case 0: xyzz_madd(r, b[i].x, b[i].y); break;
        case 1: r = xyzz_add(a[i], b[i]); break;
        case 2: r = xyzz_double(a[i]); break;
        case 4: {  // mixed addition through the FP64-pipe formulas, converted in and out
            dxyzz d = dxyzz_from_words(a[i]);
            dxyzz_madd(d, b[i].x, b[i].y);
            r = dxyzz_to_words(d);
            break;
        }
        default: {
            const affine q = xyzz_to_affine(a[i]);
            r.x = q.x; r.y = q.y; r.zz = fe_zero(); r.zzz = fe_zero();
        }
    }
    out[i] = r;
}

static int debug_run(int device, int op, bool point, const void *a, const void *b, void *out, size_t n) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "debug: device %d not initialised", device);
    if (!a || !out || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "debug: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t elem = point ? sizeof(xyzz) : sizeof(fe);
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc(&da, n * elem));
    CUDA_TRY(cudaMalloc(&db, n * elem));
    CUDA_TRY(cudaMalloc(&dout, n * elem));
    CUDA_TRY(cudaMemcpyAsync(da, a, n * elem, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(db, b ? b : a, n * elem, cudaMemcpyHostToDevice, c->stream));
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (point) {
        PK_LAUNCH(k_debug_point, dim3(blocks), dim3(128), 0, c->stream, op, (const xyzz *)da, (const xyzz *)db, (xyzz *)dout, (u32)n);
    } else {
        PK_LAUNCH(k_debug_field, dim3(blocks), dim3(128), 0, c->stream, op, (const fe *)da, (const fe *)db, (fe *)dout, (u32)n);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyAsync(out, dout, n * elem, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_debug_field_op(int device, int op, const void *a32, const void *b32, void *out32, size_t n) {
    return debug_run(device, op, false, a32, b32, out32, n);
}
extern "C" int plonkish_cuda_debug_point_op(int device, int op, const void *a128, const void *b128, void *out128, size_t n) {
    return debug_run(device, op, true, a128, b128, out128, n);
}

// ============================================================ resident scalars
// Polynomial evaluations kept in HBM between commit and open: the
// reference re-reads poly.evals() from host memory in every caller (kzg.rs:255, 291).
struct ScalarsEntry {
    int dev = 0;
    size_t n = 0;
    void *d_ptr = nullptr;
    BlockRef keep;  // owner of d_ptr (pooled): copies of the entry keep it alive while a call or a sum-check state uses it
};
static std::map<uint64_t, ScalarsEntry> g_scalars;

static void release_all_sumcheck();
static void release_all_scalars() {  // caller holds g_mu
    release_all_sumcheck();
    g_scalars.clear();
}
static uint64_t publish_scalars(int dev, void *d_ptr, size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uint64_t h = g_next_handle++;
    ScalarsEntry e;
    e.dev = dev; e.n = n; e.d_ptr = d_ptr;
    e.keep = make_block(dev, d_ptr, true, g_ctx[dev], n * PLONKISH_CUDA_SCALAR_BYTES);
    g_scalars[h] = e;
    return h;
}
static bool lookup_scalars(uint64_t handle, ScalarsEntry &out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_scalars.find(handle);
    if (it == g_scalars.end()) return false;
    out = it->second;
    return true;
}

extern "C" int plonkish_cuda_scalars_register(int device, const void *scalars, size_t n, uint64_t *handle) {
    if (!scalars || !handle || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "scalars_register: null argument or n == 0");
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "scalars_register: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    CUDA_TRY(upload(c, d, scalars, n * PLONKISH_CUDA_SCALAR_BYTES, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *handle = publish_scalars(device, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_scalars_release(uint64_t handle) {
    ScalarsEntry e;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_scalars.find(handle);
        if (it == g_scalars.end()) return fail(PLONKISH_CUDA_E_INVALID, "scalars_release: unknown handle %llu", (unsigned long long)handle);
        e = it->second;
        g_scalars.erase(it);
    }
    // `e` holds the map's reference: the buffer returns to the pool here (stream-ordered after everything enqueued on
    // the context's stream), or when the last call / sum-check state still reading it is done
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_scalars_read(uint64_t handle, size_t offset, size_t n, void *out) {
    ScalarsEntry e;
    if (!lookup_scalars(handle, e)) return fail(PLONKISH_CUDA_E_INVALID, "scalars_read: unknown handle %llu", (unsigned long long)handle);
    if (!out || offset > e.n || n > e.n - offset) return fail(PLONKISH_CUDA_E_INVALID, "scalars_read: range [%zu, +%zu) outside %zu scalars", offset, n, e.n);
    Ctx *c = ctx_for(e.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "scalars_read: device %d not initialised", e.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    CUDA_TRY(cudaMemcpyAsync(out, (const char *)e.d_ptr + offset * PLONKISH_CUDA_SCALAR_BYTES, n * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_bases_read(uint64_t handle, size_t offset, size_t n, void *out) {
    BasesEntry e;
    if (!lookup_bases(handle, e)) return fail(PLONKISH_CUDA_E_INVALID, "bases_read: unknown handle %llu", (unsigned long long)handle);
    if (e.n_shards != 1) return fail(PLONKISH_CUDA_E_INVALID, "bases_read: handle is sharded");
    if (!out || offset > e.n || n > e.n - offset) return fail(PLONKISH_CUDA_E_INVALID, "bases_read: range [%zu, +%zu) outside %zu bases", offset, n, e.n);
    Ctx *c = ctx_for(e.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "bases_read: device %d not initialised", e.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    // row 0 of a table is the bases themselves
    CUDA_TRY(cudaMemcpyAsync(out, (const char *)e.d_ptr[0] + offset * PLONKISH_CUDA_AFFINE_BYTES, n * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

// MultilinearPolynomial::eq_xy(y) (poly/multilinear.rs; used by the sum-check prover state, classic.rs:57-61) as a
// resident table: evaluations of eq(x, y) over the boolean hypercube, lowest variable first — the same recurrence
// as the eq tables of the SRS (kzg.rs:178-192), kept as scalars.
extern "C" int plonkish_cuda_eq_table(int device, const void *y, size_t num_vars, uint64_t *handle) {
    if (!handle || (num_vars && !y)) return fail(PLONKISH_CUDA_E_INVALID, "eq_table: null argument");
    if (num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "eq_table: num_vars = %zu exceeds 28", num_vars);
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "eq_table: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t n = (size_t)1 << num_vars, total = 2 * n - 1;
    void *all = nullptr, *out = nullptr;
    int rc = grow(c->tmp, (total + num_vars + 1) * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    all = c->tmp.ptr;
    if ((rc = pool_alloc(c, &out, n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    PoolGuard out_guard{c, out};
    void *d_y = (char *)all + total * PLONKISH_CUDA_SCALAR_BYTES;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    if (num_vars) CUDA_TRY(cudaMemcpyAsync(d_y, y, num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    pk_enqueue_eq_scalars(d_y, (u32)num_vars, all, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaMemcpyAsync(out, (char *)all + (n - 1) * PLONKISH_CUDA_SCALAR_BYTES, n * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *handle = publish_scalars(device, out_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// commit from resident evaluations: variable_base_msm(poly.evals(), pp.eq(k)).into() (kzg.rs:255)
extern "C" int plonkish_cuda_msm_bn254_g1_resident(uint64_t scalars_handle, uint64_t bases_handle, size_t n, void *out_affine64) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!out_affine64) return fail(PLONKISH_CUDA_E_INVALID, "msm_resident: null output");
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "msm_resident: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    if (n > se.n) return fail(PLONKISH_CUDA_E_INVALID, "msm_resident: n = %zu exceeds the %zu resident scalars", n, se.n);
    if (n == 0) { memset(out_affine64, 0, PLONKISH_CUDA_AFFINE_BYTES); return PLONKISH_CUDA_OK; }
    BasesView view;
    int device = 0;
    int rc = view_of(bases_handle, n, -1, view, &device, "msm_resident");
    if (rc) return rc;
    if (device != se.dev) return fail(PLONKISH_CUDA_E_INVALID, "msm_resident: scalars on device %d, bases on device %d", se.dev, device);
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_resident: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    xyzz *res = nullptr;
    if ((rc = enqueue_device_msm(c, se.d_ptr, view, n, 0, c->stream, &res))) return rc;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, c->stream, res, 1u, (affine *)c->d_out, (xyzz *)nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_out, c->d_out, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_affine64, c->h_out, PLONKISH_CUDA_AFFINE_BYTES);
    timer_report(n, t0);
    return PLONKISH_CUDA_OK;
}

// g_prime = sum_i coeffs[i] * poly_i over resident polynomials (pcs/multilinear.rs:203-213).
static int lincomb_impl(const uint64_t *scalars_handles, const void *coeffs, size_t count, size_t n, uint64_t *out_handle, bool padded) {
    if (!scalars_handles || !coeffs || !out_handle || count == 0 || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "fr_linear_combination: bad argument");
    std::vector<ScalarsEntry> es(count);
    for (size_t i = 0; i < count; ++i) {
        if (!lookup_scalars(scalars_handles[i], es[i])) return fail(PLONKISH_CUDA_E_INVALID, "fr_linear_combination: unknown handle %llu", (unsigned long long)scalars_handles[i]);
        if (!padded && es[i].n < n) return fail(PLONKISH_CUDA_E_INVALID, "fr_linear_combination: polynomial %zu holds %zu < %zu scalars", i, es[i].n, n);
        if (es[i].dev != es[0].dev) return fail(PLONKISH_CUDA_E_INVALID, "fr_linear_combination: polynomials live on different devices");
    }
    Ctx *c = ctx_for(es[0].dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_linear_combination: device %d not initialised", es[0].dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)c->sm_count * 8) blocks = (size_t)c->sm_count * 8;
    for (size_t done = 0; done < count; done += PK_LINCOMB_MAX) {
        LincombArgs a;
        a.count = (u32)(count - done < PK_LINCOMB_MAX ? count - done : PK_LINCOMB_MAX);
        a.accumulate = done ? 1u : 0u;
        for (u32 i = 0; i < a.count; ++i) {
            a.poly[i] = (const uint4 *)es[done + i].d_ptr;
            a.len[i] = padded && es[done + i].n < n ? es[done + i].n : n;
            memcpy(a.coeff[i].l, (const char *)coeffs + (done + i) * PLONKISH_CUDA_SCALAR_BYTES, PLONKISH_CUDA_SCALAR_BYTES);
        }
        PK_LAUNCH(k_fr_lincomb, dim3((unsigned)blocks), dim3(256), 0, c->stream, a, n, (uint4 *)d);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}
extern "C" int plonkish_cuda_fr_linear_combination(const uint64_t *scalars_handles, const void *coeffs, size_t count, size_t n, uint64_t *out_handle) {
    return lincomb_impl(scalars_handles, coeffs, count, n, out_handle, false);
}
// Sums of univariate polynomials of different lengths (`f += (scalar, q)`, poly/univariate.rs; the combined quotient and f of
// UnivariateKzg::batch_open over Gemini's folds, pcs/univariate/kzg.rs:330,339-343): a polynomial shorter than n counts as
// zero past its last coefficient.
extern "C" int plonkish_cuda_fr_linear_combination_padded(const uint64_t *scalars_handles, const void *coeffs, size_t count, size_t n, uint64_t *out_handle) {
    return lincomb_impl(scalars_handles, coeffs, count, n, out_handle, true);
}

// MultilinearKzg::open on a resident polynomial (kzg.rs:276-302): the quotients of
// pcs/multilinear.rs:72-107 are produced in HBM and committed from there, so no scalar
// crosses PCIe.  eq_handles[i] = pp.eq(i) for i < num_vars; point = num_vars Montgomery Fr;
// out_comms = num_vars affine points (the order written to the transcript, kzg.rs:299),
// out_eval = f(point) (the `remainder` of kzg.rs:295).
extern "C" int plonkish_cuda_kzg_open_bn254(uint64_t scalars_handle, const uint64_t *eq_handles, const void *point, size_t num_vars,
                                            void *out_comms_affine64, void *out_eval_mont32) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!out_eval_mont32 || (num_vars && (!eq_handles || !point || !out_comms_affine64))) return fail(PLONKISH_CUDA_E_INVALID, "kzg_open: null argument");
    if (num_vars > 26) return fail(PLONKISH_CUDA_E_INVALID, "kzg_open: num_vars = %zu exceeds 26", num_vars);
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "kzg_open: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    const size_t n = (size_t)1 << num_vars;
    if (se.n != n) return fail(PLONKISH_CUDA_E_INVALID, "kzg_open: polynomial holds %zu evaluations, point has %zu variables", se.n, num_vars);  // multilinear.rs:77
    std::vector<ManyJob> jobs(num_vars);
    for (size_t i = 0; i < num_vars; ++i) {
        int dev_i = 0;
        int rc = view_of(eq_handles[i], (size_t)1 << i, -1, jobs[i].view, &dev_i, "kzg_open");
        if (rc) return rc;
        if (dev_i != se.dev) return fail(PLONKISH_CUDA_E_INVALID, "kzg_open: eqs[%zu] lives on device %d, the polynomial on %d", i, dev_i, se.dev);
    }
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "kzg_open: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc;
    // q buffer (2^k), two remainder buffers (2^(k-1), 2^(k-2)), point (k), eval (1)
    const size_t q_elems = n, ra = n / 2 + 1, rb = n / 4 + 1;
    const size_t total = (q_elems + ra + rb + num_vars + 2) * PLONKISH_CUDA_SCALAR_BYTES;
    if ((rc = grow(c->open_buf, total))) return rc;
    if ((rc = grow(c->batch_out, (num_vars + 1) * PLONKISH_CUDA_AFFINE_BYTES))) return rc;
    char *base = (char *)c->open_buf.ptr;
    void *q = base, *rem_a = base + q_elems * 32, *rem_b = base + (q_elems + ra) * 32;
    void *d_point = base + (q_elems + ra + rb) * 32, *d_eval = base + (q_elems + ra + rb + num_vars + 1) * 32;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    if (num_vars) CUDA_TRY(cudaMemcpyAsync(d_point, point, num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    pk_enqueue_quotients(se.d_ptr, (u32)num_vars, d_point, q, rem_a, rem_b, d_eval, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    // largest first on the main stream; the small ones overlap on the side lanes
    for (size_t i = 0; i < num_vars; ++i) {
        jobs[i].scalars = (const char *)q + ((size_t)1 << i) * PLONKISH_CUDA_SCALAR_BYTES;
        jobs[i].on_device = true;
        jobs[i].n = (size_t)1 << i;
    }
    if (num_vars && (rc = enqueue_many(c, jobs, c->batch_out.ptr))) return rc;
    if (num_vars) CUDA_TRY(cudaMemcpyAsync(out_comms_affine64, c->batch_out.ptr, num_vars * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_out, d_eval, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_eval_mont32, c->h_out, PLONKISH_CUDA_SCALAR_BYTES);
    {
        std::vector<size_t> sizes;
        for (size_t i = 0; i < num_vars; ++i) sizes.push_back((size_t)1 << i);
        timer_report_shared(sizes, t0);
    }
    return PLONKISH_CUDA_OK;
}

// UnivariatePolynomial::div_rem by (X - z) on resident coefficients (poly/univariate.rs:144-168): the quotient of
// UnivariateKzg::open (pcs/univariate/kzg.rs:281-282) and, point by point, of batch_open's vanishing polynomials (:327).
// The quotient comes back as a new resident vector of the same length n (coefficient n-1 is zero), the remainder —
// the polynomial's value at z — as one Montgomery Fr.
extern "C" int plonkish_cuda_fr_div_linear(uint64_t scalars_handle, const void *z_mont32, uint64_t *out_quotient_handle, void *out_rem_mont32) {
    if (!z_mont32 || !out_quotient_handle || !out_rem_mont32) return fail(PLONKISH_CUDA_E_INVALID, "fr_div_linear: null argument");
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "fr_div_linear: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_div_linear: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t n = se.n;
    void *q = nullptr, *scratch = nullptr;
    int rc = pool_alloc(c, &q, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard q_guard{c, q};
    // PLONKISH_CUDA_HORNER_LOG_CHUNK: coefficients per thread of the division's two passes, as a power of two (4..8; default 4)
    static const u32 log_chunk = [] {
        const char *e = getenv("PLONKISH_CUDA_HORNER_LOG_CHUNK");
        const int v = e ? atoi(e) : PK_HORNER_LOG_CHUNK;
        return (u32)(v < 4 ? 4 : v > 8 ? 8 : v);
    }();
    const size_t scratch_elems = pk_horner_scratch_elems(n, log_chunk) + 16;
    if ((rc = grow(c->tmp, scratch_elems * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    scratch = c->tmp.ptr;
    char *zs = (char *)scratch, *rem = zs + 8 * PLONKISH_CUDA_SCALAR_BYTES, *work = zs + 16 * PLONKISH_CUDA_SCALAR_BYTES;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    CUDA_TRY(cudaMemcpyAsync(zs, z_mont32, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    pk_enqueue_div_linear(se.d_ptr, n, zs, work, q, rem, c->stream, log_chunk);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_out, rem, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_rem_mont32, c->h_out, PLONKISH_CUDA_SCALAR_BYTES);
    *out_quotient_handle = publish_scalars(c->dev, q_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// `quotients` (pcs/multilinear.rs:72-107) on a resident polynomial, kept: the packed quotient buffer (q_i, 2^i values, at
// element offset 2^i; element 0 is zero) comes back as a resident vector of 2^num_vars scalars and f(point) as one
// Montgomery Fr.  What Zeromorph::open (pcs/multilinear/zeromorph.rs:149-150) starts from: it commits the quotients
// (many_resident below) and reads them again for q_hat and f.
extern "C" int plonkish_cuda_fr_quotients(uint64_t scalars_handle, const void *point, size_t num_vars, uint64_t *out_q_handle, void *out_eval_mont32) {
    if (!out_q_handle || !out_eval_mont32 || (num_vars && !point)) return fail(PLONKISH_CUDA_E_INVALID, "fr_quotients: null argument");
    if (num_vars > PK_ZM_MAX_VARS) return fail(PLONKISH_CUDA_E_INVALID, "fr_quotients: num_vars = %zu exceeds %d", num_vars, PK_ZM_MAX_VARS);
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "fr_quotients: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    const size_t n = (size_t)1 << num_vars;
    if (se.n != n) return fail(PLONKISH_CUDA_E_INVALID, "fr_quotients: polynomial holds %zu evaluations, point has %zu variables", se.n, num_vars);  // multilinear.rs:77
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_quotients: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *q = nullptr;
    int rc = pool_alloc(c, &q, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard q_guard{c, q};
    // two remainder buffers (2^(k-1), 2^(k-2)), point (k), eval (1)
    const size_t ra = n / 2 + 1, rb = n / 4 + 1;
    if ((rc = grow(c->open_buf, (ra + rb + num_vars + 2) * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    char *base = (char *)c->open_buf.ptr;
    void *rem_a = base, *rem_b = base + ra * 32, *d_point = base + (ra + rb) * 32, *d_eval = base + (ra + rb + num_vars + 1) * 32;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    if (num_vars) CUDA_TRY(cudaMemcpyAsync(d_point, point, num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(q, 0, PLONKISH_CUDA_SCALAR_BYTES, c->stream));
    pk_enqueue_quotients(se.d_ptr, (u32)num_vars, d_point, q, rem_a, rem_b, d_eval, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_out, d_eval, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_eval_mont32, c->h_out, PLONKISH_CUDA_SCALAR_BYTES);
    *out_q_handle = publish_scalars(c->dev, q_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// A handle on the sub-range [offset, offset + n) of a resident vector: shares the memory (no copy) and keeps it alive
// until the slice itself is released.  One quotient of fr_quotients or one fold of fr_gemini_folds as a polynomial of
// its own for fr_div_linear / fr_linear_combination(_padded) / msm_bn254_g1_resident.
extern "C" int plonkish_cuda_scalars_slice(uint64_t handle, size_t offset, size_t n, uint64_t *out_handle) {
    if (!out_handle || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "scalars_slice: null argument or n == 0");
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_scalars.find(handle);
    if (it == g_scalars.end()) return fail(PLONKISH_CUDA_E_INVALID, "scalars_slice: unknown scalars handle %llu", (unsigned long long)handle);
    const ScalarsEntry &src = it->second;
    if (offset > src.n || n > src.n - offset) return fail(PLONKISH_CUDA_E_INVALID, "scalars_slice: [%zu, %zu) of %zu resident scalars", offset, offset + n, src.n);
    ScalarsEntry e;
    e.dev = src.dev; e.n = n; e.d_ptr = (char *)src.d_ptr + offset * PLONKISH_CUDA_SCALAR_BYTES;
    e.keep = src.keep;
    const uint64_t h = g_next_handle++;
    g_scalars[h] = e;
    *out_handle = h;
    return PLONKISH_CUDA_OK;
}

// The folds of Gemini::open (pcs/multilinear/gemini.rs:98-108): f_0 = the polynomial's evaluations read as coefficients,
// f_i = merge_into(f_(i-1), point[i-1], 1, 0) (poly/multilinear.rs:599-618) for i = 1..num_vars-1, kept in HBM packed
// like the quotients: f_i (2^(num_vars-i) values) at element offset 2^(num_vars-i) of a resident vector of 2^num_vars
// scalars (elements 0 and 1 are zero).  point: num_vars Montgomery Fr (the last one is not used, as in the reference).
extern "C" int plonkish_cuda_fr_gemini_folds(uint64_t scalars_handle, const void *point, size_t num_vars, uint64_t *out_handle) {
    if (!out_handle || !point) return fail(PLONKISH_CUDA_E_INVALID, "fr_gemini_folds: null argument");
    if (num_vars == 0 || num_vars > PK_ZM_MAX_VARS) return fail(PLONKISH_CUDA_E_INVALID, "fr_gemini_folds: num_vars = %zu (1..%d supported)", num_vars, PK_ZM_MAX_VARS);
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "fr_gemini_folds: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    const size_t n = (size_t)1 << num_vars;
    if (se.n != n) return fail(PLONKISH_CUDA_E_INVALID, "fr_gemini_folds: polynomial holds %zu evaluations, point has %zu variables", se.n, num_vars);
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_gemini_folds: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    if ((rc = grow(c->tmp, (num_vars + 1) * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    CUDA_TRY(cudaMemcpyAsync(c->tmp.ptr, point, num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(d, 0, (n < 2 ? n : 2) * PLONKISH_CUDA_SCALAR_BYTES, c->stream));
    pk_enqueue_gemini_folds(se.d_ptr, (u32)num_vars, c->tmp.ptr, d, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// `count` independent MSMs whose scalars are the sub-ranges [offsets[j], offsets[j] + ns[j]) of ONE resident vector, MSM j
// against the first ns[j] bases of bases_handles[j]: UnivariateKzg::batch_commit_and_write over Zeromorph's quotients
// (zeromorph.rs:150: q_i against powers_of_s_g1[..2^i], every handle the same registered slice) without a scalar crossing
// PCIe; scheduling as in msm_bn254_g1_many (large ones in turn, small ones on the side lanes, one host wait).
extern "C" int plonkish_cuda_msm_bn254_g1_many_resident(uint64_t scalars_handle, const size_t *offsets, const uint64_t *bases_handles, const size_t *ns,
                                                        size_t count, void *out_affine64_list) {
    const auto t0 = std::chrono::steady_clock::now();
    if (count == 0) return PLONKISH_CUDA_OK;
    if (!offsets || !bases_handles || !ns || !out_affine64_list) return fail(PLONKISH_CUDA_E_INVALID, "msm_many_resident: null argument");
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "msm_many_resident: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    std::vector<ManyJob> jobs(count);
    bool any = false;
    for (size_t j = 0; j < count; ++j) {
        if (ns[j] == 0) continue;
        if (offsets[j] > se.n || ns[j] > se.n - offsets[j])
            return fail(PLONKISH_CUDA_E_INVALID, "msm_many_resident: MSM %zu reads [%zu, %zu) of %zu resident scalars", j, offsets[j], offsets[j] + ns[j], se.n);
        if (!bases_handles[j]) return fail(PLONKISH_CUDA_E_INVALID, "msm_many_resident: MSM %zu lacks a bases handle", j);
        int dev_j = 0;
        int rc = view_of(bases_handles[j], ns[j], -1, jobs[j].view, &dev_j, "msm_many_resident");
        if (rc) return rc;
        if (dev_j != se.dev) return fail(PLONKISH_CUDA_E_INVALID, "msm_many_resident: bases of MSM %zu live on device %d, the scalars on %d", j, dev_j, se.dev);
        jobs[j].scalars = (const char *)se.d_ptr + offsets[j] * PLONKISH_CUDA_SCALAR_BYTES;
        jobs[j].on_device = true;
        jobs[j].n = ns[j];
        any = true;
    }
    if (!any) { memset(out_affine64_list, 0, count * PLONKISH_CUDA_AFFINE_BYTES); return PLONKISH_CUDA_OK; }
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "msm_many_resident: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc = grow(c->batch_out, count * PLONKISH_CUDA_AFFINE_BYTES);
    if (rc) return rc;
    if ((rc = enqueue_many(c, jobs, c->batch_out.ptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_affine64_list, c->batch_out.ptr, count * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = mark_done(c, c->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    timer_report_shared(std::vector<size_t>(ns, ns + count), t0);
    return PLONKISH_CUDA_OK;
}

// Zeromorph's q_hat (zeromorph.rs:157-168) from the packed quotients of fr_quotients: q_hat[2^n - 2^i + j] += weights[i] *
// q_i[j] (weights[i] = y^i, num_vars Montgomery Fr); a new resident vector of 2^num_vars coefficients.
extern "C" int plonkish_cuda_zeromorph_q_hat_bn254(uint64_t q_handle, const void *weights_mont32, size_t num_vars, uint64_t *out_handle) {
    if (!out_handle || (num_vars && !weights_mont32)) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_q_hat: null argument");
    if (num_vars > PK_ZM_MAX_VARS) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_q_hat: num_vars = %zu exceeds %d", num_vars, PK_ZM_MAX_VARS);
    ScalarsEntry qe;
    if (!lookup_scalars(q_handle, qe)) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_q_hat: unknown scalars handle %llu", (unsigned long long)q_handle);
    const size_t n = (size_t)1 << num_vars;
    if (qe.n != n) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_q_hat: the quotient buffer holds %zu scalars, expected 2^%zu", qe.n, num_vars);
    Ctx *c = ctx_for(qe.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "zeromorph_q_hat: device %d not initialised", qe.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    ZmWeights w;
    memset(&w, 0, sizeof(w));
    if (num_vars) memcpy(w.w, weights_mont32, num_vars * PLONKISH_CUDA_SCALAR_BYTES);
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_zm_q_hat(qe.d_ptr, w, (u32)num_vars, d, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// Zeromorph's f (zeromorph.rs:175-180): f = z * poly + q_hat, f[0] += c0 (= eval_scalar * eval), f[j] += q_scalars[i] *
// q_i[j] for j < 2^i; poly = the 2^num_vars evaluations read as coefficients (:175).  A new resident vector, which
// UnivariateKzg::open (fr_div_linear + msm_resident against open_pp) then opens at x.
extern "C" int plonkish_cuda_zeromorph_f_bn254(uint64_t poly_handle, uint64_t q_hat_handle, uint64_t q_handle, const void *z_mont32, const void *c0_mont32,
                                               const void *q_scalars_mont32, size_t num_vars, uint64_t *out_handle) {
    if (!out_handle || !z_mont32 || !c0_mont32 || (num_vars && !q_scalars_mont32)) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_f: null argument");
    if (num_vars > PK_ZM_MAX_VARS) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_f: num_vars = %zu exceeds %d", num_vars, PK_ZM_MAX_VARS);
    ScalarsEntry pe, he, qe;
    if (!lookup_scalars(poly_handle, pe) || !lookup_scalars(q_hat_handle, he) || !lookup_scalars(q_handle, qe))
        return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_f: unknown scalars handle");
    const size_t n = (size_t)1 << num_vars;
    if (pe.n != n || he.n != n || qe.n != n) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_f: poly / q_hat / quotients hold %zu / %zu / %zu scalars, expected 2^%zu", pe.n, he.n, qe.n, num_vars);
    if (he.dev != pe.dev || qe.dev != pe.dev) return fail(PLONKISH_CUDA_E_INVALID, "zeromorph_f: polynomials live on different devices");
    Ctx *c = ctx_for(pe.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "zeromorph_f: device %d not initialised", pe.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    ZmWeights w;
    memset(&w, 0, sizeof(w));
    if (num_vars) memcpy(w.w, q_scalars_mont32, num_vars * PLONKISH_CUDA_SCALAR_BYTES);
    fe z, c0;
    memcpy(z.l, z_mont32, PLONKISH_CUDA_SCALAR_BYTES);
    memcpy(c0.l, c0_mont32, PLONKISH_CUDA_SCALAR_BYTES);
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_zm_f(pe.d_ptr, he.d_ptr, qe.d_ptr, w, z, c0, (u32)num_vars, d, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// permutation_z_polys (backend/hyperplonk/prover.rs:252-345): the grand-product polynomials HyperPlonk commits at
// backend/hyperplonk.rs:251-252, from resident witness columns and permutation polynomials.  value_handles[i] /
// sigma_handles[i]: the column polys[*poly] and its permutation polynomial for i < count (in the order of
// pp.permutation_polys); num_chunks = pp.num_permutation_z_polys.  out_handles receives num_chunks resident polynomials.
extern "C" int plonkish_cuda_permutation_z_polys_bn254(const uint64_t *value_handles, const uint64_t *sigma_handles, size_t count, size_t num_chunks,
                                                       size_t num_vars, const void *beta_mont32, const void *gamma_mont32, uint64_t *out_handles) {
    if (!value_handles || !sigma_handles || !beta_mont32 || !gamma_mont32 || !out_handles) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: null argument");
    if (count == 0 || num_chunks == 0 || num_chunks > count) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: %zu polynomials in %zu chunks", count, num_chunks);
    if (num_vars == 0 || num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: num_vars = %zu (1..28 supported)", num_vars);
    const size_t chunk_size = (count + num_chunks - 1) / num_chunks;  // prover.rs:263
    if (chunk_size > PK_PERM_MAX || num_chunks > PK_PERM_MAX) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: at most %d polynomials per chunk and %d chunks", PK_PERM_MAX, PK_PERM_MAX);
    if ((count + chunk_size - 1) / chunk_size != num_chunks) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: %zu polynomials do not fill %zu chunks of %zu", count, num_chunks, chunk_size);
    const size_t n = (size_t)1 << num_vars;
    std::vector<ScalarsEntry> vals(count), sigs(count);
    for (size_t i = 0; i < count; ++i) {
        if (!lookup_scalars(value_handles[i], vals[i]) || !lookup_scalars(sigma_handles[i], sigs[i])) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: unknown scalars handle at %zu", i);
        if (vals[i].n != n || sigs[i].n != n) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: polynomial %zu does not hold 2^%zu evaluations", i, num_vars);
        if (vals[i].dev != vals[0].dev || sigs[i].dev != vals[0].dev) return fail(PLONKISH_CUDA_E_INVALID, "permutation_z_polys: polynomials live on different devices");
    }
    Ctx *c = ctx_for(vals[0].dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "permutation_z_polys: device %d not initialised", vals[0].dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t scratch_elems = pk_perm_z_scratch_elems(num_chunks, n);
    void *work = nullptr;
    int rc = grow(c->tmp, (scratch_elems + 2 + 8) * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    work = c->tmp.ptr;
    std::vector<void *> outs(num_chunks, nullptr);
    struct OutGuard { Ctx *c; std::vector<void *> &v; bool armed = true; ~OutGuard() { if (armed) for (void *p : v) pool_free(c, p); } } out_guard{c, outs};
    for (size_t k = 0; k < num_chunks; ++k)
        if ((rc = pool_alloc(c, &outs[k], n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    char *d_bg = (char *)work + scratch_elems * PLONKISH_CUDA_SCALAR_BYTES, *d_out_table = d_bg + 2 * PLONKISH_CUDA_SCALAR_BYTES;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    unsigned char bg[64];
    memcpy(bg, beta_mont32, 32);
    memcpy(bg + 32, gamma_mont32, 32);
    CUDA_TRY(cudaMemcpyAsync(d_bg, bg, 64, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_out_table, outs.data(), num_chunks * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
    std::vector<PermArgs> chunks(num_chunks);
    for (size_t k = 0; k < num_chunks; ++k) {
        PermArgs &a = chunks[k];
        memset(&a, 0, sizeof(a));
        const size_t first = k * chunk_size;
        a.count = (u32)(count - first < chunk_size ? count - first : chunk_size);
        for (u32 i = 0; i < a.count; ++i) {
            a.value[i] = (const uint4 *)vals[first + i].d_ptr;
            a.sigma[i] = (const uint4 *)sigs[first + i].d_ptr;
            a.id_offset[i] = (unsigned long long)(first + i) << num_vars;   // idx << num_vars, prover.rs:286
        }
    }
    pk_enqueue_perm_z(chunks.data(), (u32)num_chunks, (u32)num_vars, d_bg, work, d_out_table, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    out_guard.armed = false;
    for (size_t k = 0; k < num_chunks; ++k) out_handles[k] = publish_scalars(c->dev, outs[k], n);
    return PLONKISH_CUDA_OK;
}

// One table of the compiled sum-check expression (plonkish_b200/expression.py; k_fr_affine in poly_kernels.cuh):
//   out[b] = constant + identity_coeff * b + sum_i coeffs[i] * poly_i[rotate(b, rotations[i])],  then out[rows[j]] += values[j].
// Covers what the reference's evaluator keeps implicit (piop/sum_check/classic.rs:40-75, 104-126): identity and Lagrange
// polynomials, constants inside a linear factor such as w + beta * id + gamma (backend/hyperplonk/preprocessor.rs:153-165),
// rotated queries (rotation_map, util/arithmetic/bh.rs:104-121, 135-137) and instance polynomials
// (backend/hyperplonk/prover.rs:32-48).  Null pointers switch the constant / the identity term off.
extern "C" int plonkish_cuda_fr_affine_table(int device, size_t num_vars, const uint64_t *poly_handles, const int32_t *rotations, const void *coeffs_mont32,
                                             size_t count, const void *constant_mont32, const void *identity_coeff_mont32, const uint64_t *sparse_rows,
                                             const void *sparse_values_mont32, size_t sparse_count, uint64_t *out_handle) {
    if (!out_handle || (count && (!poly_handles || !coeffs_mont32)) || (sparse_count && (!sparse_rows || !sparse_values_mont32)))
        return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: null argument");
    if (num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: num_vars = %zu exceeds 28", num_vars);
    const size_t n = (size_t)1 << num_vars;
    static const unsigned char FR_ONE_BYTES[32] = {0xfb, 0xff, 0xff, 0x4f, 0x1c, 0x34, 0x96, 0xac, 0x29, 0xcd, 0x60, 0x9f, 0x95, 0x76, 0xfc, 0x36,
                                                   0x2e, 0x46, 0x79, 0x78, 0x6f, 0xa3, 0x6e, 0x66, 0x2f, 0xdf, 0x07, 0x9a, 0xc1, 0x77, 0x0a, 0x0e};
    std::vector<ScalarsEntry> es(count);
    for (size_t i = 0; i < count; ++i) {
        if (!lookup_scalars(poly_handles[i], es[i])) return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: unknown handle %llu", (unsigned long long)poly_handles[i]);
        if (es[i].n != n) return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: polynomial %zu holds %zu evaluations, not 2^%zu", i, es[i].n, num_vars);
        if (es[i].dev != device) return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: polynomial %zu lives on device %d, not %d", i, es[i].dev, device);
        if (rotations && (size_t)(rotations[i] < 0 ? -rotations[i] : rotations[i]) > num_vars)
            return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: rotation %d exceeds num_vars = %zu", rotations[i], num_vars);  // classic.rs:42
    }
    for (size_t j = 0; j < sparse_count; ++j)
        if (sparse_rows[j] >= n) return fail(PLONKISH_CUDA_E_INVALID, "fr_affine_table: row %llu outside 2^%zu", (unsigned long long)sparse_rows[j], num_vars);
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_affine_table: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    int rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES);
    if (rc) return rc;
    PoolGuard d_guard{c, d};
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)c->sm_count * 8) blocks = (size_t)c->sm_count * 8;
    // the first launch writes, further launches (more than PK_AFFINE_MAX sources) add their sources to it through a
    // unit-coefficient read of the output itself
    size_t done = 0;
    bool first = true;
    do {
        AffineArgs a;
        memset(&a, 0, sizeof(a));
        a.num_vars = (u32)num_vars; a.primitive = PK_BH_PRIMITIVES[num_vars]; a.x_inv = PK_BH_X_INVS[num_vars];
        u32 slot = 0;
        if (first) {
            if (constant_mont32) { a.has_constant = 1; memcpy(a.constant.l, constant_mont32, 32); }
            if (identity_coeff_mont32) { a.has_id = 1; memcpy(a.id_coeff.l, identity_coeff_mont32, 32); }
        } else {
            a.poly[0] = (const uint4 *)d; a.rotation[0] = 0;
            memcpy(a.coeff[0].l, FR_ONE_BYTES, 32);
            slot = 1;
        }
        for (; slot < PK_AFFINE_MAX && done < count; ++slot, ++done) {
            a.poly[slot] = (const uint4 *)es[done].d_ptr;
            a.rotation[slot] = rotations ? rotations[done] : 0;
            memcpy(a.coeff[slot].l, (const char *)coeffs_mont32 + done * PLONKISH_CUDA_SCALAR_BYTES, 32);
            if (memcmp(a.coeff[slot].l, FR_ONE_BYTES, 32) != 0) a.has_coeff_mask |= 1u << slot;
        }
        a.count = slot;
        PK_LAUNCH(k_fr_affine, dim3((unsigned)blocks), dim3(256), 0, c->stream, a, n, (uint4 *)d);
        first = false;
    } while (done < count);
    if (sparse_count) {
        if ((rc = grow(c->tmp, sparse_count * (PLONKISH_CUDA_SCALAR_BYTES + sizeof(unsigned long long)) + 64))) return rc;
        char *d_vals = (char *)c->tmp.ptr, *d_rows = d_vals + sparse_count * PLONKISH_CUDA_SCALAR_BYTES;
        CUDA_TRY(cudaMemcpyAsync(d_vals, sparse_values_mont32, sparse_count * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(d_rows, sparse_rows, sparse_count * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        PK_LAUNCH(k_fr_sparse_add, dim3(1), dim3(32), 0, c->stream, (uint4 *)d, (const unsigned long long *)d_rows, (const uint4 *)d_vals, (u32)sparse_count);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// MultilinearPolynomial::evaluate (poly/multilinear.rs:137-156) on a resident polynomial at `count` points of num_vars
// Montgomery Fr each — what prove_sum_check asks for through evaluate_for_rotation (prover.rs:392-400,
// poly/multilinear.rs:191-263: the polynomial at the 2^distance points of rotation_eval_points).  The variables are fixed
// from the highest down like the remainder of `quotients` (multilinear.rs:85-97); the value is the same field element in
// any order.  out_evals: count Montgomery Fr.
extern "C" int plonkish_cuda_fr_evaluate(uint64_t scalars_handle, const void *points_mont32, size_t num_vars, size_t count, void *out_evals_mont32) {
    if (count == 0) return PLONKISH_CUDA_OK;
    if (!out_evals_mont32 || (num_vars && !points_mont32)) return fail(PLONKISH_CUDA_E_INVALID, "fr_evaluate: null argument");
    if (num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "fr_evaluate: num_vars = %zu exceeds 28", num_vars);
    ScalarsEntry se;
    if (!lookup_scalars(scalars_handle, se)) return fail(PLONKISH_CUDA_E_INVALID, "fr_evaluate: unknown scalars handle %llu", (unsigned long long)scalars_handle);
    const size_t n = (size_t)1 << num_vars;
    if (se.n != n) return fail(PLONKISH_CUDA_E_INVALID, "fr_evaluate: polynomial holds %zu evaluations, points have %zu variables", se.n, num_vars);  // multilinear.rs:138
    Ctx *c = ctx_for(se.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_evaluate: device %d not initialised", se.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc;
    const size_t q_elems = n, ra = n / 2 + 1, rb = n / 4 + 1;
    if ((rc = grow(c->open_buf, (q_elems + ra + rb + count * (num_vars + 1) + 2) * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    char *base = (char *)c->open_buf.ptr;
    void *q = base, *rem_a = base + q_elems * 32, *rem_b = base + (q_elems + ra) * 32;
    char *d_points = base + (q_elems + ra + rb) * 32, *d_evals = d_points + count * num_vars * 32;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    if (num_vars) CUDA_TRY(cudaMemcpyAsync(d_points, points_mont32, count * num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    for (size_t p = 0; p < count; ++p)
        pk_enqueue_quotients(se.d_ptr, (u32)num_vars, d_points + p * num_vars * 32, q, rem_a, rem_b, d_evals + p * 32, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned char> host(count * PLONKISH_CUDA_SCALAR_BYTES);
    CUDA_TRY(cudaMemcpyAsync(host.data(), d_evals, host.size(), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_evals_mont32, host.data(), host.size());
    return PLONKISH_CUDA_OK;
}

static int fill_expr(const char *what, size_t num_polys, const void *term_coeffs, const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms,
                     int common_poly, SumcheckExpr &ex);
// A compiled expression on every row: out[b] = (sum_t coeff_t * prod_j poly[fac_t,j][b]) (* poly[common][b]) — the
// compressed input / table polynomial of a lookup (lookup_compressed_poly, backend/hyperplonk/prover.rs:79-137: the
// reference evaluates the expression tree per row; here the tree arrives compiled into product terms over resident
// tables, plonkish_b200/expression.py).  Same term encoding as plonkish_cuda_sumcheck_new.
extern "C" int plonkish_cuda_fr_expression_table(const uint64_t *poly_handles, size_t num_polys, size_t num_vars, const void *term_coeffs,
                                                 const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms, int common_poly,
                                                 uint64_t *out_handle) {
    if (!poly_handles || !out_handle) return fail(PLONKISH_CUDA_E_INVALID, "fr_expression_table: null argument");
    if (num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "fr_expression_table: num_vars = %zu exceeds 28", num_vars);
    SumcheckExpr ex;
    int rc = fill_expr("fr_expression_table", num_polys, term_coeffs, term_offsets, term_polys, num_terms, common_poly, ex);
    if (rc) return rc;
    const size_t n = (size_t)1 << num_vars;
    SumcheckPolys polys;
    memset(&polys, 0, sizeof(polys));
    std::vector<ScalarsEntry> es(num_polys);
    for (size_t p = 0; p < num_polys; ++p) {
        if (!lookup_scalars(poly_handles[p], es[p])) return fail(PLONKISH_CUDA_E_INVALID, "fr_expression_table: unknown scalars handle %llu", (unsigned long long)poly_handles[p]);
        if (es[p].n != n) return fail(PLONKISH_CUDA_E_INVALID, "fr_expression_table: polynomial %zu holds %zu evaluations, expected 2^%zu", p, es[p].n, num_vars);
        if (es[p].dev != es[0].dev) return fail(PLONKISH_CUDA_E_INVALID, "fr_expression_table: polynomials live on different devices");
        polys.p[p] = (const uint4 *)es[p].d_ptr;
    }
    Ctx *c = ctx_for(es[0].dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fr_expression_table: device %d not initialised", es[0].dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    void *d = nullptr;
    if ((rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    PoolGuard d_guard{c, d};
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_expr_rows(polys, ex, n, d, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// lookup_m_poly (backend/hyperplonk/prover.rs:145-192): m[row] = how many input values equal table[row], a value that
// occurs in several table rows counted at the last of them (the HashMap of :151); PLONKISH_CUDA_E_INVALID with
// "Invalid lookup input" when an input value is not in the table (:169-177).
extern "C" int plonkish_cuda_lookup_m_poly_bn254(uint64_t input_handle, uint64_t table_handle, uint64_t *out_handle) {
    if (!out_handle) return fail(PLONKISH_CUDA_E_INVALID, "lookup_m_poly: null argument");
    ScalarsEntry in, tb;
    if (!lookup_scalars(input_handle, in) || !lookup_scalars(table_handle, tb)) return fail(PLONKISH_CUDA_E_INVALID, "lookup_m_poly: unknown scalars handle");
    if (in.n != tb.n || in.dev != tb.dev) return fail(PLONKISH_CUDA_E_INVALID, "lookup_m_poly: input holds %zu values on device %d, table %zu on device %d", in.n, in.dev, tb.n, tb.dev);
    if (in.n >= ((size_t)1 << 30)) return fail(PLONKISH_CUDA_E_INVALID, "lookup_m_poly: %zu rows exceed 2^30", in.n);
    const size_t n = in.n;
    Ctx *c = ctx_for(in.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "lookup_m_poly: device %d not initialised", in.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc;
    const size_t slots = pk_lookup_slots(n);
    if ((rc = grow(c->tmp, (slots + n + 4) * sizeof(u32)))) return rc;
    u32 *d_slots = (u32 *)c->tmp.ptr, *d_counts = d_slots + slots, *d_missing = d_counts + n;
    void *d = nullptr;
    if ((rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    PoolGuard d_guard{c, d};
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_lookup_m(in.d_ptr, tb.d_ptr, (u32)n, d_slots, d_counts, d_missing, d, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    u32 missing = 0;
    CUDA_TRY(cudaMemcpyAsync(&missing, d_missing, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (missing) return fail(PLONKISH_CUDA_E_INVALID, "lookup_m_poly: Invalid lookup input");  // Error::InvalidSnark, prover.rs:176-178
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// lookup_h_poly (backend/hyperplonk/prover.rs:206-250): h = 1 / (gamma + input) - m / (gamma + table), one of the
// polynomials committed at backend/hyperplonk.rs:251-252.
extern "C" int plonkish_cuda_lookup_h_poly_bn254(uint64_t input_handle, uint64_t table_handle, uint64_t m_handle, const void *gamma_mont32,
                                                 uint64_t *out_handle) {
    if (!out_handle || !gamma_mont32) return fail(PLONKISH_CUDA_E_INVALID, "lookup_h_poly: null argument");
    ScalarsEntry in, tb, mm;
    if (!lookup_scalars(input_handle, in) || !lookup_scalars(table_handle, tb) || !lookup_scalars(m_handle, mm)) return fail(PLONKISH_CUDA_E_INVALID, "lookup_h_poly: unknown scalars handle");
    if (in.n != tb.n || in.n != mm.n || in.dev != tb.dev || in.dev != mm.dev) return fail(PLONKISH_CUDA_E_INVALID, "lookup_h_poly: the three polynomials differ in size or device");
    const size_t n = in.n;
    Ctx *c = ctx_for(in.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "lookup_h_poly: device %d not initialised", in.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc;
    if ((rc = grow(c->tmp, 2 * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    void *d = nullptr;
    if ((rc = pool_alloc(c, &d, n * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
    PoolGuard d_guard{c, d};
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    CUDA_TRY(cudaMemcpyAsync(c->tmp.ptr, gamma_mont32, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
    pk_enqueue_lookup_h(in.d_ptr, tb.d_ptr, mm.d_ptr, c->tmp.ptr, n, d, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_handle = publish_scalars(c->dev, d_guard.release(), n);
    return PLONKISH_CUDA_OK;
}

// ============================================================== fixed-base MSM
// fixed_base_msm (msm.rs:67-81) over a window table of one base (msm.rs:16-31) followed by
// batch_normalize (kzg.rs:204-207, univariate/kzg.rs:196-199): out[i] = scalars[i] * base, affine.
static const size_t FIXED_BATCH = (size_t)1 << 22;  // scalars per launch pair; bounds the projective scratch (512 MiB)

struct FixedTable {
    void *table = nullptr, *offsets = nullptr, *tmp = nullptr;  // tmp: max(table build, FIXED_BATCH) xyzz
    void release() { cudaFree(table); cudaFree(offsets); cudaFree(tmp); table = offsets = tmp = nullptr; }
    ~FixedTable() { release(); }  // every early return gives the table back
};
// Caller holds c->mu.  d_base: one affine point in device memory.
static int fixed_table_build(Ctx *c, const void *d_base, size_t max_batch, FixedTable &ft) {
    const size_t entries = (size_t)PK_FIXED_W * PK_FIXED_ROW;
    const size_t tmp_pts = entries > max_batch ? entries : max_batch;
    CUDA_TRY(cudaMalloc(&ft.table, entries * sizeof(affine)));
    CUDA_TRY(cudaMalloc(&ft.offsets, PK_FIXED_W * sizeof(affine)));
    CUDA_TRY(cudaMalloc(&ft.tmp, tmp_pts * sizeof(xyzz)));
    pk_enqueue_fixed_table(d_base, (affine *)ft.offsets, (xyzz *)ft.tmp, (affine *)ft.table, c->stream);
    CUDA_TRY(cudaGetLastError());
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_fixed_base_msm_bn254_g1(int device, const void *base_affine64, const void *scalars, size_t n, void *out_affine64_list) {
    if (!base_affine64 || (n && (!scalars || !out_affine64_list))) return fail(PLONKISH_CUDA_E_INVALID, "fixed_base_msm: null argument");
    if (n == 0) return PLONKISH_CUDA_OK;
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "fixed_base_msm: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t batch = n < FIXED_BATCH ? n : FIXED_BATCH;
    int rc;
    if ((rc = grow(c->scalars, batch * PLONKISH_CUDA_SCALAR_BYTES)) || (rc = grow(c->bases_tmp, batch * PLONKISH_CUDA_AFFINE_BYTES))) return rc;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    CUDA_TRY(cudaMemcpyAsync((char *)c->d_out + 384, base_affine64, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream));
    FixedTable ft;
    if ((rc = fixed_table_build(c, (char *)c->d_out + 384, batch, ft))) { ft.release(); return rc; }
    for (size_t done = 0; done < n; done += batch) {
        const size_t cnt = n - done < batch ? n - done : batch;
        CUDA_TRY(upload(c, c->scalars.ptr, (const char *)scalars + done * PLONKISH_CUDA_SCALAR_BYTES, cnt * PLONKISH_CUDA_SCALAR_BYTES, c->stream));
        pk_enqueue_fixed_base(c->scalars.ptr, (u32)cnt, (const affine *)ft.table, (xyzz *)ft.tmp, (affine *)c->bases_tmp.ptr, c->stream);
        CUDA_TRY(cudaMemcpyAsync((char *)out_affine64_list + done * PLONKISH_CUDA_AFFINE_BYTES, c->bases_tmp.ptr, cnt * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaGetLastError());
    rc = mark_done(c, c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    ft.release();
    return rc;
}

// MultilinearKzg::setup's prover half (kzg.rs:167-212) on the device: eq tables from
// ss (num_vars Montgomery Fr), times g1 by fixed-base MSM, normalised, and every slice
// eqs[k] (2^k bases, k = 0..num_vars) registered as resident bases — the SRS never exists
// in host memory.  handles_out receives num_vars + 1 handles.
extern "C" int plonkish_cuda_kzg_setup_eqs_bn254(int device, const void *g1_affine64, const void *ss, size_t num_vars, uint64_t *handles_out) {
    if (!g1_affine64 || !handles_out || (num_vars && !ss)) return fail(PLONKISH_CUDA_E_INVALID, "kzg_setup_eqs: null argument");
    if (num_vars > 26) return fail(PLONKISH_CUDA_E_INVALID, "kzg_setup_eqs: num_vars = %zu exceeds 26", num_vars);
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "kzg_setup_eqs: device %d not initialised", device);
    std::vector<BasesEntry> entries(num_vars + 1);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CUDA_TRY(cudaSetDevice(c->dev));
        const size_t total = ((size_t)2 << num_vars) - 1;
        const size_t batch = total < FIXED_BATCH ? total : FIXED_BATCH;
        void *d_eq = nullptr, *d_pts = nullptr, *d_ss = nullptr;
        FixedTable ft;
        auto cleanup = [&] { cudaFree(d_eq); cudaFree(d_pts); cudaFree(d_ss); ft.release(); };
        if (cudaMalloc(&d_eq, total * PLONKISH_CUDA_SCALAR_BYTES) != cudaSuccess || cudaMalloc(&d_pts, total * PLONKISH_CUDA_AFFINE_BYTES) != cudaSuccess ||
            cudaMalloc(&d_ss, (num_vars + 1) * PLONKISH_CUDA_SCALAR_BYTES) != cudaSuccess) {
            cleanup();
            return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: out of device memory for %zu points", total);
        }
        {
            cudaError_t e1 = c->has_last ? cudaStreamWaitEvent(c->stream, c->last_done, 0) : cudaSuccess;
            if (e1 == cudaSuccess && num_vars) e1 = cudaMemcpyAsync(d_ss, ss, num_vars * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream);
            if (e1 == cudaSuccess) e1 = cudaMemcpyAsync((char *)c->d_out + 384, g1_affine64, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream);
            if (e1 != cudaSuccess) { cleanup(); return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: %s", cudaGetErrorString(e1)); }
        }
        pk_enqueue_eq_scalars(d_ss, (u32)num_vars, d_eq, (u32)c->sm_count, c->stream);
        int rc = fixed_table_build(c, (char *)c->d_out + 384, batch, ft);
        if (rc) { cleanup(); return rc; }
        for (size_t done = 0; done < total; done += batch) {
            const size_t cnt = total - done < batch ? total - done : batch;
            pk_enqueue_fixed_base((const char *)d_eq + done * PLONKISH_CUDA_SCALAR_BYTES, (u32)cnt, (const affine *)ft.table, (xyzz *)ft.tmp,
                                  (affine *)((char *)d_pts + done * PLONKISH_CUDA_AFFINE_BYTES), c->stream);
        }
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            cleanup();
            return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: kernel sequence failed");
        }
        cudaFree(d_eq); d_eq = nullptr;
        ft.release();
        for (size_t k = 0; k <= num_vars; ++k) {
            const size_t nk = (size_t)1 << k;
            const void *src = (const char *)d_pts + (nk - 1) * PLONKISH_CUDA_AFFINE_BYTES;
            void *d = nullptr;
            uint32_t tc = 0;
            bool owns = true;
            rc = make_resident(c, src, true, nk, 0, &d, &tc, &owns);
            if (!rc && !owns) {  // plain slice: give it its own storage, d_pts goes away
                if (cudaMalloc(&d, nk * PLONKISH_CUDA_AFFINE_BYTES) != cudaSuccess) rc = fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: out of device memory");
                else if (cudaMemcpyAsync(d, src, nk * PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) rc = fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: copy failed");
                owns = true;
            }
            if (rc) {
                cleanup();  // the slices registered so far are freed with `entries`
                return rc;
            }
            BasesEntry &e = entries[k];
            e.n_shards = 1; e.dev = device; e.n = nk;
            e.add_shard(device, d, true, nk, tc);
        }
        const cudaError_t sync_err = cudaStreamSynchronize(c->stream);
        cleanup();
        if (sync_err != cudaSuccess) return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_eqs: %s", cudaGetErrorString(sync_err));
    }
    for (size_t k = 0; k <= num_vars; ++k) handles_out[k] = publish(entries[k]);
    return PLONKISH_CUDA_OK;
}

// UnivariateKzg::setup's G1 half (pcs/univariate/kzg.rs:175-195) on the device: powers(s).take(n), times g1 by
// fixed-base MSM, normalised, registered as one resident slice (powers_of_s_g1, the bases of commit_coeffs,
// univariate/kzg.rs:24-30).  The G2 half (two points for the verifier) stays with the caller.
extern "C" int plonkish_cuda_kzg_setup_powers_bn254(int device, const void *g1_affine64, const void *s, size_t n, uint64_t *handle) {
    if (!g1_affine64 || !s || !handle || n == 0) return fail(PLONKISH_CUDA_E_INVALID, "kzg_setup_powers: null argument or n == 0");
    if (n > ((size_t)1 << 27)) return fail(PLONKISH_CUDA_E_INVALID, "kzg_setup_powers: n = %zu exceeds 2^27", n);
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "kzg_setup_powers: device %d not initialised", device);
    BasesEntry e;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CUDA_TRY(cudaSetDevice(c->dev));
        const size_t batch = n < FIXED_BATCH ? n : FIXED_BATCH;
        void *d_pow = nullptr, *d_pts = nullptr, *d_s = nullptr;
        FixedTable ft;
        auto cleanup = [&] { cudaFree(d_pow); cudaFree(d_pts); cudaFree(d_s); ft.release(); };
        if (cudaMalloc(&d_pow, n * PLONKISH_CUDA_SCALAR_BYTES) != cudaSuccess || cudaMalloc(&d_pts, n * PLONKISH_CUDA_AFFINE_BYTES) != cudaSuccess ||
            cudaMalloc(&d_s, PLONKISH_CUDA_SCALAR_BYTES) != cudaSuccess) {
            cleanup();
            return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_powers: out of device memory for %zu points", n);
        }
        {
            cudaError_t e1 = c->has_last ? cudaStreamWaitEvent(c->stream, c->last_done, 0) : cudaSuccess;
            if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(d_s, s, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream);
            if (e1 == cudaSuccess) e1 = cudaMemcpyAsync((char *)c->d_out + 384, g1_affine64, PLONKISH_CUDA_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream);
            if (e1 != cudaSuccess) { cleanup(); return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_powers: %s", cudaGetErrorString(e1)); }
        }
        pk_enqueue_fr_powers(d_s, n, d_pow, c->stream);
        int rc = fixed_table_build(c, (char *)c->d_out + 384, batch, ft);
        if (rc) { cleanup(); return rc; }
        for (size_t done = 0; done < n; done += batch) {
            const size_t cnt = n - done < batch ? n - done : batch;
            pk_enqueue_fixed_base((const char *)d_pow + done * PLONKISH_CUDA_SCALAR_BYTES, (u32)cnt, (const affine *)ft.table, (xyzz *)ft.tmp,
                                  (affine *)((char *)d_pts + done * PLONKISH_CUDA_AFFINE_BYTES), c->stream);
        }
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            cleanup();
            return fail(PLONKISH_CUDA_E_CUDA, "kzg_setup_powers: kernel sequence failed");
        }
        cudaFree(d_pow); d_pow = nullptr;
        ft.release();
        void *d = nullptr;
        uint32_t tc = 0;
        bool owns = true;
        rc = make_resident(c, d_pts, true, n, 0, &d, &tc, &owns);
        if (rc) { cleanup(); return rc; }
        if (!owns) { d_pts = nullptr; owns = true; }  // plain slice: the entry adopts the points as they are
        e.n_shards = 1; e.dev = device; e.n = n;
        e.add_shard(device, d, true, n, tc);
        cleanup();
    }
    *handle = publish(e);
    return PLONKISH_CUDA_OK;
}

// ================================================================== sum check
// ClassicSumCheck<EvaluationsProver>::prove (piop/sum_check/classic.rs:208-240) as a round-by-round
// state on the device: the caller owns the transcript (it hashes each round message to draw the
// challenge, classic.rs:226-229), the device owns the tables.
struct SumcheckState {
    int dev = 0;
    u32 num_vars = 0, round = 0, num_polys = 0;
    SumcheckExpr ex;
    std::vector<const void *> cur;          // current table of every polynomial
    std::vector<BlockRef> keep;             // the callers' resident polynomials, alive as long as the state reads them
    std::vector<void *> buf_a, buf_b;       // state-owned halves: 2^(k-1) and 2^(k-2) evaluations per polynomial
    void *block = nullptr;                  // one pooled allocation behind buf_a / buf_b / scratch
    void *partials = nullptr, *d_out = nullptr, *d_chal = nullptr, *d_final = nullptr;
};
static std::map<uint64_t, SumcheckState> g_sumcheck;

static void release_all_sumcheck() {  // caller holds g_mu
    for (auto &kv : g_sumcheck) {
        Ctx *c = (size_t)kv.second.dev < g_ctx.size() ? g_ctx[kv.second.dev] : nullptr;
        if (c && kv.second.block == c->sc_block.ptr) { c->sc_block_busy = false; continue; }  // freed with the context
        cudaSetDevice(kv.second.dev);
        cudaFree(kv.second.block);
    }
    g_sumcheck.clear();
}

// The flattened expression sum_t coeff_t * prod_j poly[fac_t,j] (* poly[common]) as the kernels take it; shared by the
// sum-check state and plonkish_cuda_fr_expression_table.  Returns 0 or a failure already recorded with `what`.
static int fill_expr(const char *what, size_t num_polys, const void *term_coeffs, const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms,
                     int common_poly, SumcheckExpr &ex) {
    if (!term_offsets || (num_terms && !term_coeffs)) return fail(PLONKISH_CUDA_E_INVALID, "%s: null argument", what);
    if (num_polys == 0 || num_polys > PK_SC_MAX_POLYS) return fail(PLONKISH_CUDA_E_INVALID, "%s: %zu polynomials (1..%d supported)", what, num_polys, PK_SC_MAX_POLYS);
    if (num_terms == 0 || num_terms > PK_SC_MAX_TERMS) return fail(PLONKISH_CUDA_E_INVALID, "%s: %zu terms (1..%d supported)", what, num_terms, PK_SC_MAX_TERMS);
    if (common_poly >= (int)num_polys) return fail(PLONKISH_CUDA_E_INVALID, "%s: common factor %d out of range", what, common_poly);
    memset(&ex, 0, sizeof(ex));
    ex.num_terms = (u32)num_terms; ex.num_polys = (u32)num_polys; ex.common = common_poly < 0 ? -1 : common_poly;
    u32 degree = 0;
    for (size_t t = 0; t < num_terms; ++t) {
        const uint32_t beg = term_offsets[t], end = term_offsets[t + 1];
        if (end < beg || end - beg > PK_SC_MAX_FACTORS) return fail(PLONKISH_CUDA_E_INVALID, "%s: term %zu has %u factors (0..%d supported)", what, t, end - beg, PK_SC_MAX_FACTORS);
        ex.nfac[t] = (unsigned char)(end - beg);
        for (uint32_t j = beg; j < end; ++j) {
            if (!term_polys || term_polys[j] >= num_polys) return fail(PLONKISH_CUDA_E_INVALID, "%s: term %zu names polynomial %u of %zu", what, t, term_polys ? term_polys[j] : 0u, num_polys);
            ex.fac[t][j - beg] = (unsigned char)term_polys[j];
        }
        memcpy(ex.coeff[t].l, (const char *)term_coeffs + t * PLONKISH_CUDA_SCALAR_BYTES, PLONKISH_CUDA_SCALAR_BYTES);
        static const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};  // R mod r
        ex.has_coeff[t] = memcmp(ex.coeff[t].l, FR_ONE, 32) != 0;
        if (end - beg > degree) degree = end - beg;
    }
    if (ex.common >= 0) degree += 1;
    if (degree < 1) degree = 1;
    ex.degree = degree;
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_sumcheck_new(const uint64_t *poly_handles, size_t num_polys, size_t num_vars, const void *term_coeffs,
                                          const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms, int common_poly,
                                          uint64_t *state_handle) {
    if (!poly_handles || !state_handle) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: null argument");
    if (num_vars == 0 || num_vars > 28) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: num_vars = %zu (1..28 supported)", num_vars);  // classic.rs:41
    SumcheckState st;
    st.num_vars = (u32)num_vars; st.num_polys = (u32)num_polys;
    int rc_expr = fill_expr("sumcheck_new", num_polys, term_coeffs, term_offsets, term_polys, num_terms, common_poly, st.ex);
    if (rc_expr) return rc_expr;
    if (st.ex.degree > PK_SC_MAX_DEGREE) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: degree %u exceeds %d", st.ex.degree, PK_SC_MAX_DEGREE);
    const size_t n = (size_t)1 << num_vars;
    for (size_t p = 0; p < num_polys; ++p) {
        ScalarsEntry se;
        if (!lookup_scalars(poly_handles[p], se)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: unknown scalars handle %llu", (unsigned long long)poly_handles[p]);
        if (se.n != n) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: polynomial %zu holds %zu evaluations, expected 2^%zu", p, se.n, num_vars);
        if (p == 0) st.dev = se.dev;
        if (se.dev != st.dev) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_new: polynomials live on different devices");
        st.cur.push_back(se.d_ptr);
        st.keep.push_back(se.keep);
    }
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_new: device %d not initialised", st.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t half = n / 2, quarter = n / 4 ? n / 4 : 1;
    const size_t max_blocks = (size_t)c->sm_count * 16;
    const size_t elems = num_polys * (half + quarter) + max_blocks * PK_SC_MAX_DEGREE + PK_SC_MAX_DEGREE + 1 + num_polys;
    int rc;
    if (!c->sc_block_busy) {
        if ((rc = grow(c->sc_block, elems * PLONKISH_CUDA_SCALAR_BYTES))) return rc;
        st.block = c->sc_block.ptr;
        c->sc_block_busy = true;
    } else if ((rc = pool_alloc(c, &st.block, elems * PLONKISH_CUDA_SCALAR_BYTES))) {
        return rc;
    }
    char *q = (char *)st.block;
    for (size_t p = 0; p < num_polys; ++p) { st.buf_a.push_back(q); q += half * 32; }
    for (size_t p = 0; p < num_polys; ++p) { st.buf_b.push_back(q); q += quarter * 32; }
    st.partials = q; q += max_blocks * PK_SC_MAX_DEGREE * 32;
    st.d_out = q; q += PK_SC_MAX_DEGREE * 32;
    st.d_chal = q; q += 32;
    st.d_final = q;
    std::lock_guard<std::mutex> lk2(g_mu);
    const uint64_t h = g_next_handle++;
    g_sumcheck[h] = st;
    *state_handle = h;
    return PLONKISH_CUDA_OK;
}

static bool lookup_sumcheck(uint64_t handle, SumcheckState &out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_sumcheck.find(handle);
    if (it == g_sumcheck.end()) return false;
    out = it->second;
    return true;
}

extern "C" int plonkish_cuda_sumcheck_degree(uint64_t state_handle) {
    SumcheckState st;
    if (!lookup_sumcheck(state_handle, st)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_degree: unknown state %llu", (unsigned long long)state_handle);
    return (int)st.ex.degree;
}

// The round message without its first entry: out[x-1] = sum_b expr(.., X = x, b) for x = 1..degree
// (eval.rs:101-131; the caller sets evals[0] = sum - evals[1], eval.rs:128).
extern "C" int plonkish_cuda_sumcheck_round(uint64_t state_handle, void *out_evals) {
    SumcheckState st;
    if (!out_evals) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round: null output");
    if (!lookup_sumcheck(state_handle, st)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round: unknown state %llu", (unsigned long long)state_handle);
    if (st.round >= st.num_vars) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round: all %u rounds are done", st.num_vars);
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_round: device %d not initialised", st.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    SumcheckPolys polys;
    memset(&polys, 0, sizeof(polys));
    for (u32 p = 0; p < st.num_polys; ++p) polys.p[p] = (const uint4 *)st.cur[p];
    const u32 size = 1u << (st.num_vars - st.round - 1);  // ProverState::size, classic.rs:86-88
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_sumcheck_round(polys, st.ex, size, st.partials, st.d_out, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_evals, st.d_out, (size_t)st.ex.degree * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

// The factored round of a zero check (SumcheckExpr::common_sum): out[x-1] = G(x) = sum_b S_b * expr(.., X = x, b) for
// x = 1..degree-1, S_b = the common factor's pair sum.  Valid when the common factor is eq(x, y) times a constant; the
// round polynomial is then h(X) = (1 - y_r + X (2 y_r - 1)) G(X), and the caller rebuilds the reference's message
// (eval.rs:101-131) from G and the running sum — one evaluation point fewer per pair than plonkish_cuda_sumcheck_round.
extern "C" int plonkish_cuda_sumcheck_round_factored(uint64_t state_handle, void *out_evals) {
    SumcheckState st;
    if (!out_evals) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round_factored: null output");
    if (!lookup_sumcheck(state_handle, st)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round_factored: unknown state %llu", (unsigned long long)state_handle);
    if (st.round >= st.num_vars) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round_factored: all %u rounds are done", st.num_vars);
    if (st.ex.common < 0 || st.ex.degree < 2) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_round_factored: the expression has no common factor");
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_round_factored: device %d not initialised", st.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    SumcheckPolys polys;
    memset(&polys, 0, sizeof(polys));
    for (u32 p = 0; p < st.num_polys; ++p) polys.p[p] = (const uint4 *)st.cur[p];
    const u32 size = 1u << (st.num_vars - st.round - 1);
    SumcheckExpr ex = st.ex;
    ex.degree = st.ex.degree - 1;
    ex.common_sum = 1;
    if (c->has_last) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->last_done, 0));
    pk_enqueue_sumcheck_round(polys, ex, size, st.partials, st.d_out, (u32)c->sm_count, c->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_evals, st.d_out, (size_t)ex.degree * PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

// ProverState::next_round (classic.rs:90-141): every table is fixed at the challenge.
extern "C" int plonkish_cuda_sumcheck_fix_var(uint64_t state_handle, const void *challenge) {
    SumcheckState st;
    if (!challenge) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_fix_var: null challenge");
    if (!lookup_sumcheck(state_handle, st)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_fix_var: unknown state %llu", (unsigned long long)state_handle);
    if (st.round >= st.num_vars) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_fix_var: all %u variables are fixed", st.num_vars);
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_fix_var: device %d not initialised", st.dev);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CUDA_TRY(cudaSetDevice(c->dev));
        SumcheckFoldArgs a;
        memset(&a, 0, sizeof(a));
        std::vector<void *> &dst = (st.round % 2 == 0) ? st.buf_a : st.buf_b;
        for (u32 p = 0; p < st.num_polys; ++p) { a.in[p] = (const uint4 *)st.cur[p]; a.out[p] = (uint4 *)dst[p]; }
        const u32 size = 1u << (st.num_vars - st.round - 1);
        CUDA_TRY(cudaMemcpyAsync(st.d_chal, challenge, PLONKISH_CUDA_SCALAR_BYTES, cudaMemcpyHostToDevice, c->stream));
        pk_enqueue_sumcheck_fold(a, st.num_polys, st.d_chal, size, (u32)c->sm_count, c->stream);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // the caller's challenge buffer is free again
        for (u32 p = 0; p < st.num_polys; ++p) st.cur[p] = dst[p];
        st.keep.clear();  // the folded tables are state-owned: the callers' polynomials may go
        st.round += 1;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_sumcheck[state_handle] = st;
    return PLONKISH_CUDA_OK;
}

// ProverState::into_evals (classic.rs:143-149): every polynomial at the point of all challenges.
extern "C" int plonkish_cuda_sumcheck_final_evals(uint64_t state_handle, void *out_evals) {
    SumcheckState st;
    if (!out_evals) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_final_evals: null output");
    if (!lookup_sumcheck(state_handle, st)) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_final_evals: unknown state %llu", (unsigned long long)state_handle);
    if (st.round != st.num_vars) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_final_evals: %u of %u variables fixed", st.round, st.num_vars);  // classic.rs:144
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_final_evals: device %d not initialised", st.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    for (u32 p = 0; p < st.num_polys; ++p)
        CUDA_TRY(cudaMemcpyAsync((char *)st.d_final + (size_t)p * 32, st.cur[p], 32, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(out_evals, st.d_final, (size_t)st.num_polys * 32, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PLONKISH_CUDA_OK;
}

extern "C" int plonkish_cuda_sumcheck_free(uint64_t state_handle) {
    SumcheckState st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_sumcheck.find(state_handle);
        if (it == g_sumcheck.end()) return fail(PLONKISH_CUDA_E_INVALID, "sumcheck_free: unknown state %llu", (unsigned long long)state_handle);
        st = it->second;
        g_sumcheck.erase(it);
    }
    Ctx *c = ctx_for(st.dev);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "sumcheck_free: device %d not initialised", st.dev);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    if (st.block == c->sc_block.ptr) c->sc_block_busy = false;  // the next state reuses it in stream order
    else pool_free(c, st.block);
    return PLONKISH_CUDA_OK;
}

// ================================================================= transcript
// Keccak-f[1600], host code: the permutation behind the reference's Keccak256Transcript (util/transcript.rs:100-131,
// util/hash.rs:5-8 — the sha3 crate on the Rust side).  The transcript itself (absorb / squeeze order, byte orders)
// lives in the host mirrors (plonkish_b200/transcript.py); only the 24 rounds are here, because they dominate a pure
// Python transcript.  Lane (x, y) is state[x + 5 y].
extern "C" void plonkish_cuda_keccak_f1600(uint64_t state[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};  // [x][y]
    auto rol = [](uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; };
    uint64_t *a = state, b[25];
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5], d[5];
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], ROT[x][y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[round];
    }
}


#ifdef PK_STAGE_PROF
// Experiment builds only: cycles per phase of the staged partition kernels, block 7's view (tools/stage_phase_probe.py).
extern "C" int plonkish_cuda_debug_stage_prof(int device, unsigned long long out[16], int reset) {
    Ctx *c = ctx_for(device);
    if (!c) return fail(PLONKISH_CUDA_E_NO_DEVICE, "debug_stage_prof: device %d not initialised", device);
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->dev));
    CUDA_TRY(cudaDeviceSynchronize());
    if (out) CUDA_TRY(cudaMemcpyFromSymbol(out, pk::g_stage_prof, 16 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[16] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(pk::g_stage_prof, z, sizeof(z)));
    }
    return PLONKISH_CUDA_OK;
}
#endif
