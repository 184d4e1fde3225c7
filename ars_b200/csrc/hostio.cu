// Host <-> device copies for callers that hold PAGEABLE memory (the numpy arrays of the drop-in module).
//
// cudaMemcpyAsync from / to pageable memory goes through the driver's own bounce buffer on one CPU thread (~10 GB/s):
// a 350 MB clip takes 35 ms to upload where the bus needs 7.  Here large pageable transfers go through a small ring of
// pinned slots instead: a few worker threads copy pageable -> slot (or slot -> pageable) in parallel while the DMA
// engine moves the previous slot, so the transfer runs at the slower of the bus and the host's memcpy bandwidth.
// Pinned (or registered) memory is detected and copied directly.  Also here: a pool of pinned blocks the Python layer
// wraps as numpy arrays for its results, so that a render's output lands in page-locked memory at bus speed and the
// caller still receives an ordinary ndarray.
#include "ars_common.cuh"

#include <algorithm>
#include <condition_variable>
#include <thread>

namespace ars {

namespace {

class CopyPool {
  public:
    void run(void* dst, const void* src, size_t bytes) {
        start();
        const int T = (int)workers_.size();
        const size_t part = ((bytes / (size_t)T) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = static_cast<char*>(dst);
            src_ = static_cast<const char*>(src);
            bytes_ = bytes;
            part_ = std::max<size_t>(part, 4096);
            pending_ = T;
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
        quit_ = false;
    }
    ~CopyPool() { if (!workers_.empty()) stop(); }

  private:
    void start() {
        if (!workers_.empty()) return;
        unsigned hw = std::thread::hardware_concurrency();
        int T = (int)std::max(2u, std::min(8u, hw ? hw / 2 : 4u));
        if (const char* e = getenv("ARS_COPY_THREADS")) T = std::max(1, std::min(32, atoi(e)));
        for (int i = 0; i < T; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    void loop(int id) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (quit_) return;
            char* d = dst_;
            const char* s = src_;
            const size_t lo = std::min(bytes_, part_ * (size_t)id), hi = std::min(bytes_, part_ * (size_t)(id + 1));
            lk.unlock();
            if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, part_ = 0;
    int pending_ = 0;
    unsigned long long gen_ = 0;
    bool quit_ = false;
};

constexpr int SLOTS = 3;
constexpr size_t SLOT_BYTES = (size_t)32 << 20;
constexpr size_t STAGE_MIN_BYTES = (size_t)4 << 20;     // smaller pageable copies go straight through the driver

struct Ring {
    void* slot[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t fence = nullptr;
    void init() {
        if (stream) return;
        ARS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ARS_CUDA(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
        for (int i = 0; i < SLOTS; ++i) {
            ARS_CUDA(cudaMallocHost(&slot[i], SLOT_BYTES));
            ARS_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
    }
    void release() {
        if (!stream) return;
        cudaStreamSynchronize(stream);
        for (int i = 0; i < SLOTS; ++i) { cudaFreeHost(slot[i]); cudaEventDestroy(ev[i]); slot[i] = nullptr; ev[i] = nullptr; }
        cudaEventDestroy(fence);
        cudaStreamDestroy(stream);
        stream = nullptr;
        fence = nullptr;
    }
};

CopyPool g_pool;
Ring g_ring;
std::mutex g_io_mu;          // one staged transfer at a time (the ring is shared)

struct PinnedBlock { void* p; size_t bytes; };
std::vector<PinnedBlock> g_free_blocks;          // cached result buffers, returned by the numpy finalizers
std::map<void*, size_t> g_live_blocks;
size_t g_free_bytes = 0;
std::mutex g_block_mu;
constexpr size_t BLOCK_CACHE_BYTES = (size_t)4 << 30;

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

static int g_staging = 1;
void host_staging_enable(int on) { g_staging = on ? 1 : 0; }

// host -> device; ordered: after everything already on `after`, and `after` continues once the data has arrived
void host_upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t after) {
    if (!bytes) return;
    if (!g_staging || bytes < STAGE_MIN_BYTES || !is_pageable(h_src)) {
        ARS_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, after));
        return;
    }
    std::lock_guard<std::mutex> lk(g_io_mu);
    g_ring.init();
    ARS_CUDA(cudaEventRecord(g_ring.fence, after));                  // the destination may still be in use by earlier work
    ARS_CUDA(cudaStreamWaitEvent(g_ring.stream, g_ring.fence, 0));
    const char* src = static_cast<const char*>(h_src);
    char* dst = static_cast<char*>(d_dst);
    int k = 0;
    for (size_t off = 0; off < bytes; off += SLOT_BYTES, ++k) {
        const int s = k % SLOTS;
        const size_t len = std::min(SLOT_BYTES, bytes - off);
        // the slot's last transfer -- of this call or of an earlier one -- must have left it (an event that was never
        // recorded counts as complete)
        ARS_CUDA(cudaEventSynchronize(g_ring.ev[s]));
        g_pool.run(g_ring.slot[s], src + off, len);
        ARS_CUDA(cudaMemcpyAsync(dst + off, g_ring.slot[s], len, cudaMemcpyHostToDevice, g_ring.stream));
        ARS_CUDA(cudaEventRecord(g_ring.ev[s], g_ring.stream));
    }
    ARS_CUDA(cudaEventRecord(g_ring.fence, g_ring.stream));
    ARS_CUDA(cudaStreamWaitEvent(after, g_ring.fence, 0));
    // the slots are reused by the next transfer only after their events; nothing else to wait for here
}

// device -> host; `after`: the stream that produced the data.  Returns when the bytes are in h_dst (a pageable
// destination cannot be filled asynchronously); a pinned destination is copied asynchronously on `after`.
void host_download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t after) {
    if (!bytes) return;
    if (!g_staging || bytes < STAGE_MIN_BYTES || !is_pageable(h_dst)) {
        ARS_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, after));
        return;
    }
    std::lock_guard<std::mutex> lk(g_io_mu);
    g_ring.init();
    ARS_CUDA(cudaEventRecord(g_ring.fence, after));
    ARS_CUDA(cudaStreamWaitEvent(g_ring.stream, g_ring.fence, 0));
    for (int s = 0; s < SLOTS; ++s) ARS_CUDA(cudaEventSynchronize(g_ring.ev[s]));      // (slots may hold an upload in flight)
    const char* src = static_cast<const char*>(d_src);
    char* dst = static_cast<char*>(h_dst);
    const size_t nchunk = (bytes + SLOT_BYTES - 1) / SLOT_BYTES;
    auto issue = [&](size_t k) {
        const size_t off = k * SLOT_BYTES, len = std::min(SLOT_BYTES, bytes - off);
        const int s = (int)(k % SLOTS);
        ARS_CUDA(cudaMemcpyAsync(g_ring.slot[s], src + off, len, cudaMemcpyDeviceToHost, g_ring.stream));
        ARS_CUDA(cudaEventRecord(g_ring.ev[s], g_ring.stream));
    };
    for (size_t k = 0; k < std::min<size_t>(SLOTS, nchunk); ++k) issue(k);
    for (size_t k = 0; k < nchunk; ++k) {
        const size_t off = k * SLOT_BYTES, len = std::min(SLOT_BYTES, bytes - off);
        const int s = (int)(k % SLOTS);
        ARS_CUDA(cudaEventSynchronize(g_ring.ev[s]));
        g_pool.run(dst + off, g_ring.slot[s], len);
        if (k + SLOTS < nchunk) issue(k + SLOTS);
    }
}

// ---- pinned result blocks (ars_host_alloc / ars_host_free) ----
void* host_block_alloc(size_t bytes) {
    if (!bytes) bytes = 1;
    const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    {
        std::lock_guard<std::mutex> lk(g_block_mu);
        int best = -1;
        for (size_t i = 0; i < g_free_blocks.size(); ++i)
            if (g_free_blocks[i].bytes >= want && g_free_blocks[i].bytes <= want + want / 4 + ((size_t)1 << 20) &&
                (best < 0 || g_free_blocks[i].bytes < g_free_blocks[(size_t)best].bytes)) best = (int)i;
        if (best >= 0) {
            PinnedBlock b = g_free_blocks[(size_t)best];
            g_free_blocks.erase(g_free_blocks.begin() + best);
            g_free_bytes -= b.bytes;
            g_live_blocks[b.p] = b.bytes;
            return b.p;
        }
    }
    void* p = nullptr;
    ARS_CUDA(cudaMallocHost(&p, want));
    std::lock_guard<std::mutex> lk(g_block_mu);
    g_live_blocks[p] = want;
    return p;
}

void host_block_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_block_mu);
    auto it = g_live_blocks.find(p);
    if (it == g_live_blocks.end()) return;
    const size_t bytes = it->second;
    g_live_blocks.erase(it);
    if (ctx_ready() && g_free_bytes + bytes <= BLOCK_CACHE_BYTES) {
        g_free_blocks.push_back({p, bytes});
        g_free_bytes += bytes;
    } else {
        cudaFreeHost(p);
    }
}

void hostio_release() {
    {
        std::lock_guard<std::mutex> lk(g_io_mu);
        g_ring.release();
        g_pool.stop();
    }
    std::lock_guard<std::mutex> lk(g_block_mu);
    for (auto& b : g_free_blocks) cudaFreeHost(b.p);
    g_free_blocks.clear();
    g_free_bytes = 0;
    // (live blocks belong to numpy arrays that still exist: they are freed when their finalizers run)
}

}  // namespace ars
