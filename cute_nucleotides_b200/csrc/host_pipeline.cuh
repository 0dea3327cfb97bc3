// host_pipeline.cuh -- the host-slice path behind cn_n_to_bits_host / cn_bits_to_n_host (and the base-5 pair) on ONE
// device: a process-wide pool of staging copier threads and a per-thread ring of {stream, event, pinned + device
// staging} slots.  The calling thread only ISSUES work -- staging copies are posted to the pool and waited for when
// their result is needed -- so for every chunk the copy into pinned staging, the H2D DMA, the kernel, the D2H DMA and
// the copy out of staging all overlap with the neighbouring chunks'.  (multi_device.cuh fans one call out over
// several of these pipelines, one per GPU.)
//
// What bounds this path was measured with tools/host_ceiling (profiles/host_ceiling_r02_*.jsonl): the PCIe link moves
// 55 GB/s one way (45 + 45 both ways); cudaHostRegister of the caller's pages in place is NOT an alternative to staging
// (4-8 GB/s on 4 KiB pages, does not scale over threads, and stalls a concurrent DMA stream to 1-4 GB/s); the staging
// copies compete with the DMA engines for host memory bandwidth, which is why the copier count is bounded.
#pragma once
#include "cn_launch.cuh"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <cerrno>
#include <sys/mman.h>
#include <immintrin.h>

namespace {

// ------------------------------------------------------------------------------------------------
// Copy flavours of the staging copier: glibc memcpy, or AVX2 non-temporal (streaming) stores that write the destination
// around the cache.  Measured inside the pipeline on a 16-vCPU B200 box (profiles/host_r02_pageable_sweep.jsonl, 4 GiB,
// 10 copiers, 4 MiB chunks): copying INTO pinned staging is as fast or faster with memcpy (48 vs 44 Gnt/s encode: the
// chunk is still in the last-level cache when the DMA engine reads it), copying OUT of staging into the caller's
// buffer -- which nobody reads again soon -- is faster with streaming stores (39 vs 34 Gnt/s decode).  Those are the
// defaults; CN_HOST_COPY_IN / CN_HOST_COPY_OUT = memcpy | nt override them for A/B runs.
// ------------------------------------------------------------------------------------------------
__attribute__((target("avx2"))) inline void copy_nt_avx2(uint8_t *dst, const uint8_t *src, size_t bytes)
{
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > bytes) head = bytes;
    memcpy(dst, src, head);
    dst += head; src += head; bytes -= head;
    const size_t body = bytes & ~(size_t)127;
    for (size_t i = 0; i < body; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    _mm_sfence();
    memcpy(dst + body, src + body, bytes - body);
}
inline bool copy_flavour_nt(const char *name, bool dflt)
{
    if (!__builtin_cpu_supports("avx2")) return false;
    const char *env = std::getenv(name);
    if (env && !strcmp(env, "nt")) return true;
    if (env && !strcmp(env, "memcpy")) return false;
    return dflt;
}
const bool g_copy_in_nt = copy_flavour_nt("CN_HOST_COPY_IN", false);
const bool g_copy_out_nt = copy_flavour_nt("CN_HOST_COPY_OUT", true);
inline void staging_copy(uint8_t *dst, const uint8_t *src, size_t bytes, bool nt)
{
    if (nt) copy_nt_avx2(dst, src, bytes);
    else memcpy(dst, src, bytes);
}

// ------------------------------------------------------------------------------------------------
// staging copier.  Pageable caller memory cannot be DMA'd, so it is copied through pinned staging by a small
// process-wide pool (CN_HOST_THREADS / cn_set_host_threads, default min(12, 5/8 of the cores) including the caller: the
// measured optimum on a 16-vCPU box is 10 -- fewer cannot feed the PCIe link, more take host memory bandwidth away
// from the DMA engines).
// post() queues the slices of one copy and returns at once; wait() blocks until they are done and lets the
// waiting thread execute queued slices itself.  The pool is created on first use and intentionally never
// destroyed (its threads only ever wait on a condition variable).
// ------------------------------------------------------------------------------------------------
std::atomic<int> g_host_threads{0};          // 0: default

class CopyPool {
public:
    struct Job { int pending = 0; };         // guarded by the pool mutex

    static CopyPool &get()
    {
        static CopyPool *pool = new CopyPool();      // leaked on purpose: no join at process exit
        return *pool;
    }

    // queue dst <- src as slices of at most `slice` bytes
    void post(Job &job, void *dst, const void *src, size_t bytes, bool nt, size_t slice = kSlice)
    {
        if (bytes == 0) return;
        ensure_threads();
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t off = 0; off < bytes; off += slice) {
                size_t n = bytes - off < slice ? bytes - off : slice;
                queue_.push_back(Task{static_cast<uint8_t *>(dst) + off, static_cast<const uint8_t *>(src) + off, n, &job, nullptr, nullptr, nt ? (size_t)1 : 0, 0});
                job.pending++;
            }
        }
        cv_.notify_all();
    }

    // queue fn(ctx, cuts[i], cuts[i+1]) for every i: arbitrary pieces of staging work (the gather / scatter of a batch of
    // small sequences), posted under one lock with one wake-up
    void post_fn(Job &job, void (*fn)(void *, size_t, size_t), void *ctx, const std::vector<size_t> &cuts)
    {
        if (cuts.size() < 2) return;
        ensure_threads();
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i + 1 < cuts.size(); i++) {
                queue_.push_back(Task{nullptr, nullptr, 0, &job, fn, ctx, cuts[i], cuts[i + 1]});
                job.pending++;
            }
        }
        cv_.notify_all();
    }

    // returns when every slice posted on `job` has been copied; the caller copies slices too while it waits
    void wait(Job &job)
    {
        std::unique_lock<std::mutex> lk(mu_);
        while (job.pending > 0) {
            if (!queue_.empty()) {
                Task t = queue_.front();
                queue_.pop_front();
                lk.unlock();
                t.run();
                lk.lock();
                if (--t.job->pending == 0) done_cv_.notify_all();
            } else {
                done_cv_.wait(lk);
            }
        }
    }

    bool done(Job &job)
    {
        std::lock_guard<std::mutex> lk(mu_);
        return job.pending == 0;
    }

    // small enough that a mid-sized call (1 MiB = two 512 KiB chunks) still spreads over several copiers, large enough that
    // queue traffic stays negligible (32 tasks per 4 MiB chunk)
    static constexpr size_t kSlice = (size_t)128 << 10;

private:
    struct Task {
        uint8_t *dst; const uint8_t *src; size_t bytes; Job *job;
        void (*fn)(void *, size_t, size_t); void *ctx; size_t a, b;
        void run() const { if (fn) fn(ctx, a, b); else staging_copy(dst, src, bytes, a != 0); }     // a: streaming-store flag of a plain copy
    };

    CopyPool() = default;

    static int wanted_threads()
    {
        int n = g_host_threads.load(std::memory_order_relaxed);
        if (n <= 0) {
            if (const char *env = std::getenv("CN_HOST_THREADS")) n = std::atoi(env);
        }
        if (n <= 0) {
            unsigned hc = std::thread::hardware_concurrency();
            n = (int)(hc * 5 / 8 > 12 ? 12 : hc * 5 / 8);
        }
        if (n < 1) n = 1;
        if (n > 64) n = 64;
        return n;
    }
    // the pool only ever grows: `n` copiers in total means n - 1 pool threads plus the waiting caller
    void ensure_threads()
    {
        const int want = wanted_threads() - 1;
        if (started_.load(std::memory_order_acquire) >= want) return;
        std::lock_guard<std::mutex> lk(mu_);
        while (started_.load(std::memory_order_relaxed) < want) {
            std::thread([this] { run(); }).detach();
            started_.fetch_add(1, std::memory_order_release);
        }
    }
    void run()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return !queue_.empty(); });
            Task t = queue_.front();
            queue_.pop_front();
            lk.unlock();
            t.run();
            lk.lock();
            if (--t.job->pending == 0) done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Task> queue_;
    std::atomic<int> started_{0};
};

// A large pageable destination is usually a freshly allocated Vec whose pages have never been touched: left alone, every
// 4 KiB page faults separately inside the copy-out loop (the C++ mirror measured 1 GiB decodes at 14-17 GiB/s into a fresh
// vector vs 35+ into a touched one).  MADV_POPULATE_WRITE (Linux >= 5.14) faults a whole range in with one call; the pool
// runs it for chunk k's destination while chunk k is still on the GPU.  Content is never modified.
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
std::atomic<bool> g_populate_ok{true};
void populate_range(void *, size_t begin, size_t end)
{
    if (!g_populate_ok.load(std::memory_order_relaxed)) return;
    const size_t lo = begin & ~(size_t)4095, hi = (end + 4095) & ~(size_t)4095;
    if (madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_POPULATE_WRITE) != 0 && errno == EINVAL)
        g_populate_ok.store(false, std::memory_order_relaxed);           // kernel too old: stop asking
}

// ------------------------------------------------------------------------------------------------
// host-slice pipeline.  Per calling thread and device: kSlots slots, each with its own stream, a pinned and a
// device staging buffer per direction.  Chunk c uses slot c % kSlots.
// ------------------------------------------------------------------------------------------------
constexpr int kSlots = 8;
std::atomic<size_t> g_host_chunk_pageable{(size_t)4 << 20};   // ASCII bytes per chunk when a side of the call is pageable

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *pin_big = nullptr, *pin_small = nullptr;     // ASCII-sized / packed-sized pinned staging
    uint8_t *dev_big = nullptr, *dev_small = nullptr;
    size_t pin_cap = 0, dev_cap = 0;                      // ASCII bytes the staging pairs were sized for
    CopyPool::Job in_job, out_job;                        // staging copies into / out of this slot that are still running
};

struct HostPipe {
    int device = -1;
    Slot slot[kSlots];
    unsigned long long *d_counter = nullptr;              // kSlots invalid-byte counters (one per slot stream)
    unsigned long long *h_counter = nullptr;              // pinned landing area for them
    uint8_t *tiny_in = nullptr, *tiny_out = nullptr;      // pinned staging of the single-launch small-input path
    size_t tiny_bytes = 0;
    bool ready = false;

    static void free_staging(Slot &sl)
    {
        if (sl.pin_big) cudaFreeHost(sl.pin_big);
        if (sl.pin_small) cudaFreeHost(sl.pin_small);
        if (sl.dev_big) cudaFree(sl.dev_big);
        if (sl.dev_small) cudaFree(sl.dev_small);
        sl.pin_big = sl.pin_small = sl.dev_big = sl.dev_small = nullptr;
        sl.pin_cap = sl.dev_cap = 0;
    }
    void destroy()
    {
        for (auto &sl : slot) {
            free_staging(sl);
            if (sl.done) cudaEventDestroy(sl.done);
            if (sl.stream) cudaStreamDestroy(sl.stream);
            sl.done = nullptr;
            sl.stream = nullptr;
        }
        if (d_counter) cudaFree(d_counter);
        if (h_counter) cudaFreeHost(h_counter);
        if (tiny_in) cudaFreeHost(tiny_in);
        if (tiny_out) cudaFreeHost(tiny_out);
        d_counter = h_counter = nullptr;
        tiny_in = tiny_out = nullptr;
        ready = false;
    }
    // A thread that used the host-slice calls releases its staging when it exits (thread_local destructors of a
    // live process run while the CUDA runtime is still loaded; failures during teardown are ignored).
    ~HostPipe() { if (ready) { destroy(); cudaGetLastError(); } }
};

thread_local HostPipe t_pipe;
thread_local int t_device = -1;      // -1: whatever device is current

int pipe_prepare(HostPipe &p)
{
    int dev = 0;
    if (t_device >= 0) CN_CUDA(cudaSetDevice(t_device));
    CN_CUDA(cudaGetDevice(&dev));
    if (p.ready && p.device == dev) return CN_OK;
    p.destroy();
    p.device = dev;
    // Only the cheap parts are created up front; the chunk-sized staging of a slot is allocated the first time a
    // call needs it (slot_ensure), so a thread that only ever makes small calls pins 2 x 256 KiB, not hundreds of MiB.
    for (auto &sl : p.slot) {
        CN_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CN_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    p.tiny_bytes = g_host_small + 64;
    if (cudaMalloc(&p.d_counter, kSlots * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc(&p.h_counter, kSlots * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&p.tiny_in, p.tiny_bytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&p.tiny_out, p.tiny_bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        p.destroy();
        return fail(CN_ERR_NOMEM, "host pipeline: small staging allocation failed");
    }
    p.ready = true;
    return CN_OK;
}

// Staging of one slot for chunks of up to `chunk` ASCII bytes, on first use or when a larger chunk size is asked for:
// pinned buffers when a side of the call is pageable, device buffers when the copy engines are used (strategy 0).
// Failure-atomic per pair: either both buffers of a pair exist with the recorded capacity, or neither does.
int slot_ensure(Slot &sl, size_t chunk, bool need_pinned, bool need_device)
{
    const size_t big = chunk, small = chunk / 4 + 64;
    if (need_pinned && sl.pin_cap < chunk) {
        if (sl.pin_big) cudaFreeHost(sl.pin_big);
        if (sl.pin_small) cudaFreeHost(sl.pin_small);
        sl.pin_big = sl.pin_small = nullptr;
        sl.pin_cap = 0;
        if (cudaHostAlloc(&sl.pin_big, big, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&sl.pin_small, small, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            if (sl.pin_big) cudaFreeHost(sl.pin_big);
            sl.pin_big = sl.pin_small = nullptr;
            return fail(CN_ERR_NOMEM, "host pipeline: pinned staging allocation of %zu bytes failed", big + small);
        }
        sl.pin_cap = chunk;
    }
    if (need_device && sl.dev_cap < chunk) {
        if (sl.dev_big) cudaFree(sl.dev_big);
        if (sl.dev_small) cudaFree(sl.dev_small);
        sl.dev_big = sl.dev_small = nullptr;
        sl.dev_cap = 0;
        if (cudaMalloc(&sl.dev_big, big) != cudaSuccess || cudaMalloc(&sl.dev_small, small) != cudaSuccess) {
            cudaGetLastError();
            if (sl.dev_big) cudaFree(sl.dev_big);
            sl.dev_big = sl.dev_small = nullptr;
            return fail(CN_ERR_NOMEM, "host pipeline: device staging allocation of %zu bytes failed", big + small);
        }
        sl.dev_cap = chunk;
    }
    return CN_OK;
}

// true when [p, p+bytes) is page-locked memory the device can DMA / dereference directly
bool is_pinned(const void *p, size_t bytes)
{
    if (bytes == 0) return true;
    cudaPointerAttributes a0{}, a1{};
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, static_cast<const uint8_t *>(p) + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
}

// what a host-slice call asks of the encode kernel (decode ignores it)
struct HostMode {
    int enc_mode = cn::kEncPlain;          // kEncPlain / kEncCount / kEncLutExact (2-bit codec), or base-5 checked
    uint64_t *invalid_out = nullptr;       // where the number of bytes outside the alphabet goes (may be null)
};

// One implementation for both directions and both codecs on ONE device: `big` is the ASCII side, `small` the packed side.
//   encode: src = ASCII (len bytes)          dst = packed (8*words bytes)
//   decode: src = packed (8*nwords bytes)    dst = ASCII (len bytes)
int host_codec_one(const Codec &cd, bool encode, const uint8_t *src, uint8_t *dst, size_t len, const HostMode &mode)
{
    HostPipe &p = t_pipe;
    int rc = pipe_prepare(p);
    if (rc != CN_OK) return rc;
    CopyPool &pool = CopyPool::get();

    const size_t nwords = cd.words(len);
    const size_t src_bytes = encode ? len : nwords * 8;
    const size_t dst_bytes = encode ? nwords * 8 : len;
    // every chunk's kernel adds into the counter of its slot; they are summed at the end
    const bool counted = encode && mode.enc_mode != cn::kEncPlain;
    auto run_encode = [&](const void *in, size_t nt, void *out, int slot, cudaStream_t s) {
        return cd.enc(in, nt, out, mode.enc_mode, counted ? p.d_counter + slot : nullptr, s);
    };
    auto finish_counted = [&](int slots_used) -> int {
        if (!counted) return CN_OK;
        cudaStream_t s = p.slot[0].stream;               // every slot stream has been drained by now
        CN_CUDA(cudaMemcpyAsync(p.h_counter, p.d_counter, slots_used * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CN_CUDA(cudaStreamSynchronize(s));
        unsigned long long v = 0;
        for (int k = 0; k < slots_used; k++) v += p.h_counter[k];
        if (mode.invalid_out) *mode.invalid_out = v;
        return CN_OK;
    };

    // Small inputs (the reference's own bench is 40 000 nt) are latency-bound: skip the pointer queries and
    // the copy engines, stage through the small pinned buffers and let ONE kernel read and write them in place
    // over PCIe -- a single launch and a single synchronisation.
    if (len <= g_host_small && src_bytes <= p.tiny_bytes && dst_bytes <= p.tiny_bytes) {
        Slot &sl = p.slot[0];
        if (counted) CN_CUDA(cudaMemsetAsync(p.d_counter, 0, sizeof(unsigned long long), sl.stream));
        memcpy(p.tiny_in, src, src_bytes);
        rc = encode ? run_encode(p.tiny_in, len, p.tiny_out, 0, sl.stream) : cd.dec(p.tiny_in, nwords, len, p.tiny_out, sl.stream);
        if (rc != CN_OK) { cudaStreamSynchronize(sl.stream); cudaGetLastError(); return rc; }
        CN_CUDA(cudaStreamSynchronize(sl.stream));
        memcpy(dst, p.tiny_out, dst_bytes);
        return finish_counted(1);
    }

    const bool src_pinned = is_pinned(src, src_bytes);
    const bool dst_pinned = is_pinned(dst, dst_bytes);
    const bool zero_copy = g_host_strategy.load(std::memory_order_relaxed) == 1;

    // zero-copy with both sides page-locked: a single kernel streams over PCIe in both directions
    if (zero_copy && src_pinned && dst_pinned) {
        cudaStream_t s = p.slot[0].stream;
        if (counted) CN_CUDA(cudaMemsetAsync(p.d_counter, 0, sizeof(unsigned long long), s));
        rc = encode ? run_encode(src, len, dst, 0, s) : cd.dec(src, nwords, len, dst, s);
        if (rc != CN_OK) { cudaStreamSynchronize(s); cudaGetLastError(); return rc; }
        CN_CUDA(cudaStreamSynchronize(s));
        return finish_counted(1);
    }

    // Staging holds `stage` ASCII bytes and stage/4 (+64) packed bytes per slot; the base-5 codec packs 8 bytes per
    // 27 nucleotides (> 1/4), so its chunks are 3/4 of the staging size.  Pinned callers are DMA'd in large chunks
    // (fewer, longer transfers); pageable callers in small ones, so that a chunk is still in the CPU's last-level
    // cache when the DMA engine reads it and many chunks are in flight at once.
    const bool staged = !src_pinned || !dst_pinned;
    const size_t stage = staged ? g_host_chunk_pageable.load(std::memory_order_relaxed) : g_host_chunk.load(std::memory_order_relaxed);
    const size_t max_nt = cd.group == 32 ? stage : stage / 4 * 3;
    // Nucleotides per chunk: a multiple of `unit` (whole words, and whole warp tiles / 16-byte vectors where the
    // staging size allows).  Mid-sized inputs are cut into ~2*kSlots chunks so that the stages overlap.
    size_t unit = cd.group == 32 ? 4096 : (size_t)cn::kB5WarpBytes;
    if (max_nt < unit) unit = cd.group;
    size_t chunk = max_nt / unit * unit;
    if (len / (2 * kSlots) < chunk) {
        size_t c8 = (len / (2 * kSlots) + unit - 1) / unit * unit;
        size_t floor_nt = ((size_t)512 << 10) / unit * unit;
        if (floor_nt == 0 || floor_nt > chunk) floor_nt = chunk;
        chunk = c8 < floor_nt ? floor_nt : (c8 < chunk ? c8 : chunk);
    }
    // A large pageable destination is usually a freshly allocated Vec: ask for transparent huge pages so the
    // first-touch faults taken while copying the result out are per 2 MiB, not per 4 KiB (hint only).
    if (!dst_pinned && dst_bytes >= ((size_t)8 << 20)) {
        const uintptr_t lo = (addr(dst) + 0x1FFFFF) & ~(uintptr_t)0x1FFFFF, hi = (addr(dst) + dst_bytes) & ~(uintptr_t)0x1FFFFF;
        if (hi > lo) (void)madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
    }

    const bool populate = !dst_pinned && dst_bytes >= ((size_t)8 << 20) && g_populate_ok.load(std::memory_order_relaxed);

    const size_t nchunks = (len + chunk - 1) / chunk;
    const int slots_used = nchunks < (size_t)kSlots ? (int)nchunks : kSlots;
    for (int k = 0; k < slots_used; k++) {
        Slot &sl = p.slot[k];
        rc = slot_ensure(sl, stage, staged, !zero_copy);
        if (rc != CN_OK) return rc;
        if (counted) CN_CUDA(cudaMemsetAsync(p.d_counter + k, 0, sizeof(unsigned long long), sl.stream));
    }

    // geometry of chunk c
    struct Geo { size_t nt, words, in_off, in_bytes, out_off, out_bytes; };
    auto geo = [&](size_t c) {
        Geo g;
        const size_t first = c * chunk;
        g.nt = len - first < chunk ? len - first : chunk;
        g.words = cd.words(g.nt);
        const size_t word_off = first / cd.group * 8;              // `first` is a multiple of the group size
        g.in_off = encode ? first : word_off;
        g.in_bytes = encode ? g.nt : g.words * 8;
        g.out_off = encode ? word_off : first;
        g.out_bytes = encode ? g.words * 8 : g.nt;
        return g;
    };
    auto pin_in = [&](Slot &sl) { return encode ? sl.pin_big : sl.pin_small; };
    auto pin_out = [&](Slot &sl) { return encode ? sl.pin_small : sl.pin_big; };

    // The schedule, with S = kSlots (chunk k lives in slot k % S):
    //   copy-in(k)  may start once the H2D copy of chunk k-S has finished        (chunk k-S retired)
    //   issue(k)    = H2D + kernel + D2H + event, once copy-in(k) and copy-out(k-S) are done
    //   copy-out(k) starts when chunk k is retired (its event has completed)
    // in_next = chunks whose copy-in has been posted, issued = chunks submitted to the GPU, retired = chunks whose event
    // has been waited for and whose copy-out has been posted; retired <= issued <= in_next <= retired + S.
    size_t in_next = 0, issued = 0, retired = 0;
    int first_error = CN_OK;
    auto post_in = [&](size_t k) {
        Slot &sl = p.slot[k % kSlots];
        if (!src_pinned) { Geo g = geo(k); pool.post(sl.in_job, pin_in(sl), src + g.in_off, g.in_bytes, g_copy_in_nt); }
    };
    auto retire = [&](size_t k) -> int {                              // wait for chunk k's DMA, start copying its result out
        Slot &sl = p.slot[k % kSlots];
        cudaError_t e = cudaEventSynchronize(sl.done);
        if (e != cudaSuccess) return fail(CN_ERR_CUDA, "host pipeline: chunk %zu failed: %s", k, cudaGetErrorString(e));
        if (!dst_pinned) { Geo g = geo(k); pool.post(sl.out_job, dst + g.out_off, pin_out(sl), g.out_bytes, g_copy_out_nt); }
        return CN_OK;
    };
    // retire, in order, every ISSUED chunk whose event has completed (block_first: wait for the oldest one)
    auto retire_ready = [&](bool block_first) -> int {
        while (retired < issued) {
            if (!block_first && cudaEventQuery(p.slot[retired % kSlots].done) != cudaSuccess) { cudaGetLastError(); break; }
            block_first = false;
            int r = retire(retired);
            retired++;
            if (r != CN_OK) return r;
        }
        return CN_OK;
    };
    auto feed = [&] { while (in_next < nchunks && in_next < retired + kSlots) post_in(in_next++); };

    for (size_t k = 0; k < nchunks; k++) {
        Slot &sl = p.slot[k % kSlots];
        if ((rc = retire_ready(false)) != CN_OK) { first_error = rc; break; }
        feed();
        while (in_next <= k) {                                        // chunk k's slot still belongs to chunk k-S: wait for that one
            if ((rc = retire_ready(true)) != CN_OK) { first_error = rc; break; }
            feed();
        }
        if (first_error != CN_OK) break;
        pool.wait(sl.in_job);                                         // chunk k is in pinned staging
        pool.wait(sl.out_job);                                        // chunk k-S has left the slot's output staging
        const Geo g = geo(k);
        const uint8_t *h_in = src_pinned ? src + g.in_off : pin_in(sl);
        uint8_t *h_out = dst_pinned ? dst + g.out_off : pin_out(sl);
        const int slot_idx = (int)(k % kSlots);
        cudaError_t e = cudaSuccess;
        if (zero_copy) {
            // kernel dereferences the pinned staging (or the caller's pinned side) directly
            rc = encode ? run_encode(h_in, g.nt, h_out, slot_idx, sl.stream) : cd.dec(h_in, g.words, g.nt, h_out, sl.stream);
        } else {
            uint8_t *dev_in = encode ? sl.dev_big : sl.dev_small, *dev_out = encode ? sl.dev_small : sl.dev_big;
            e = cudaMemcpyAsync(dev_in, h_in, g.in_bytes, cudaMemcpyHostToDevice, sl.stream);
            rc = e != cudaSuccess ? CN_OK
                 : (encode ? run_encode(dev_in, g.nt, dev_out, slot_idx, sl.stream) : cd.dec(dev_in, g.words, g.nt, dev_out, sl.stream));
            if (e == cudaSuccess && rc == CN_OK) e = cudaMemcpyAsync(h_out, dev_out, g.out_bytes, cudaMemcpyDeviceToHost, sl.stream);
        }
        if (e == cudaSuccess && rc == CN_OK) e = cudaEventRecord(sl.done, sl.stream);
        if (e != cudaSuccess) rc = fail(CN_ERR_CUDA, "host pipeline: submitting chunk %zu failed: %s", k, cudaGetErrorString(e));
        if (rc != CN_OK) { first_error = rc; break; }
        issued = k + 1;
        if (populate) pool.post_fn(sl.out_job, populate_range, nullptr, {addr(dst + g.out_off), addr(dst + g.out_off) + g.out_bytes});
    }

    // drain: wait for what is in flight, finish the copies (also on the error path: nothing may still be running into
    // the caller's buffers, or out of staging a later call will reuse, when this call returns)
    if (first_error == CN_OK) {
        while (retired < nchunks) {
            rc = retire(retired);
            retired++;
            if (rc != CN_OK) { first_error = rc; break; }
        }
    }
    for (int k = 0; k < slots_used; k++) {
        Slot &sl = p.slot[k];
        if (first_error != CN_OK) cudaStreamSynchronize(sl.stream);
        pool.wait(sl.in_job);
        pool.wait(sl.out_job);
    }
    if (first_error != CN_OK) { cudaGetLastError(); return first_error; }
    return finish_counted(slots_used);
}

// ------------------------------------------------------------------------------------------------
// Many independent small sequences in ONE call (cn_n_to_bits_host_batch / cn_bits_to_n_host_batch).  One call per
// read is launch-latency-bound on a GPU (16 us for the reference's own 40 000-nt bench shape, SURVEY 7 hard part 3);
// here the sequences are laid end to end in pinned staging, each padded to a whole word with zero bytes (code 0, so
// every sequence's last word gets the zero high bits the contract asks for), and ONE ordinary encode / decode launch
// covers thousands of them.  Gather and scatter are spread over the copier pool; chunks flow through the same ring.
// ------------------------------------------------------------------------------------------------
struct BatchCtx {
    unsigned group;                       // nucleotides per word: 32 (2-bit) or 27 (base-5)
    bool encode;
    const void *const *in;
    void *const *out;
    const size_t *lens;
    const size_t *woff;                   // word offset of every sequence inside its chunk
    uint8_t *pin_big, *pin_small;
};
void batch_gather(void *c, size_t a, size_t b)
{
    const BatchCtx &x = *static_cast<const BatchCtx *>(c);
    for (size_t i = a; i < b; i++) {
        const size_t len = x.lens[i];
        if (!len) continue;
        if (x.encode) {
            const size_t padded = (len + x.group - 1) / x.group * x.group;
            uint8_t *dst = x.pin_big + x.woff[i] * x.group;
            memcpy(dst, x.in[i], len);
            if (padded > len) memset(dst + len, 0, padded - len);          // byte 0 -> code 0 / digit 0 in either codec
        } else {
            memcpy(x.pin_small + x.woff[i] * 8, x.in[i], (len + x.group - 1) / x.group * 8);
        }
    }
}
void batch_scatter(void *c, size_t a, size_t b)
{
    const BatchCtx &x = *static_cast<const BatchCtx *>(c);
    for (size_t i = a; i < b; i++) {
        const size_t len = x.lens[i];
        if (!len) continue;
        if (x.encode) memcpy(x.out[i], x.pin_small + x.woff[i] * 8, (len + x.group - 1) / x.group * 8);
        else memcpy(x.out[i], x.pin_big + x.woff[i] * x.group, len);
    }
}

int host_batch_one(const Codec &cd, bool encode, const void *const *in, const size_t *lens, size_t count, void *const *out)
{
    HostPipe &p = t_pipe;
    int rc = pipe_prepare(p);
    if (rc != CN_OK) return rc;
    CopyPool &pool = CopyPool::get();
    const size_t stage = g_host_chunk_pageable.load(std::memory_order_relaxed);
    const size_t stage_words = stage / 32;        // packed staging holds stage / 4 bytes; ASCII staging then needs <= stage bytes in either codec

    // plan: consecutive sequences are grouped into chunks of at most `stage` padded bytes; a sequence that does not fit a
    // chunk on its own goes through the ordinary streaming path afterwards
    struct Chunk { size_t i0, i1, words; };
    std::vector<Chunk> chunks;
    std::vector<size_t> woff(count), big;
    Chunk cur{0, 0, 0};
    for (size_t i = 0; i < count; i++) {
        const size_t w = cd.words(lens[i]);
        if (lens[i] && (!in[i] || !out[i])) return fail(CN_ERR_ARG, "batch: null pointer for sequence %zu", i);
        if (w > stage_words) {
            big.push_back(i);
            woff[i] = 0;
            // a big sequence still ends the running chunk so that chunks stay ranges of consecutive indices
            if (cur.i1 > cur.i0) chunks.push_back(cur);
            cur = Chunk{i + 1, i + 1, 0};
            continue;
        }
        if (cur.words + w > stage_words) {
            chunks.push_back(cur);
            cur = Chunk{i, i, 0};
        }
        woff[i] = cur.words;
        cur.words += w;
        cur.i1 = i + 1;
    }
    if (cur.i1 > cur.i0) chunks.push_back(cur);

    const size_t nchunks = chunks.size();
    const int slots_used = nchunks < (size_t)kSlots ? (int)nchunks : kSlots;
    for (int k = 0; k < slots_used; k++)
        if ((rc = slot_ensure(p.slot[k], stage, true, true)) != CN_OK) return rc;
    std::vector<BatchCtx> ctx(slots_used ? slots_used : 1);

    // gather / scatter of one chunk, cut into pool tasks of >= 128 KiB of payload
    std::vector<size_t> cuts;
    auto post_ranges = [&](CopyPool::Job &job, void (*fn)(void *, size_t, size_t), BatchCtx *c, const Chunk &ch) {
        cuts.clear();
        cuts.push_back(ch.i0);
        size_t acc = 0;
        for (size_t i = ch.i0; i < ch.i1; i++) {
            acc += lens[i];
            if (acc >= ((size_t)128 << 10) || i + 1 == ch.i1) { cuts.push_back(i + 1); acc = 0; }
        }
        pool.post_fn(job, fn, c, cuts);
    };
    int first_error = CN_OK;
    auto scatter_chunk = [&](size_t k) -> int {                       // wait for chunk k's DMA, start scattering its result
        Slot &sl = p.slot[k % kSlots];
        cudaError_t e = cudaEventSynchronize(sl.done);
        if (e != cudaSuccess) return fail(CN_ERR_CUDA, "batch pipeline: chunk %zu failed: %s", k, cudaGetErrorString(e));
        post_ranges(sl.out_job, batch_scatter, &ctx[k % kSlots], chunks[k]);
        return CN_OK;
    };
    // gather(k+1) is posted before chunk k is issued, so the copier pool never waits for the issuing thread
    auto start_gather = [&](size_t k) -> int {
        Slot &sl = p.slot[k % kSlots];
        if (k >= (size_t)kSlots) {
            int r = scatter_chunk(k - kSlots);                        // frees the slot: its previous chunk leaves staging
            if (r != CN_OK) return r;
            pool.wait(sl.out_job);                                    // the slot's context is about to be rewritten
        }
        ctx[k % kSlots] = BatchCtx{cd.group, encode, in, out, lens, woff.data(), sl.pin_big, sl.pin_small};
        post_ranges(sl.in_job, batch_gather, &ctx[k % kSlots], chunks[k]);
        return CN_OK;
    };
    if (nchunks && (rc = start_gather(0)) != CN_OK) first_error = rc;
    for (size_t k = 0; k < nchunks && first_error == CN_OK; k++) {
        Slot &sl = p.slot[k % kSlots];
        if (k + 1 < nchunks && (rc = start_gather(k + 1)) != CN_OK) { first_error = rc; break; }
        pool.wait(sl.in_job);
        const size_t words = chunks[k].words, nt = words * cd.group;
        uint8_t *h_in = encode ? sl.pin_big : sl.pin_small, *h_out = encode ? sl.pin_small : sl.pin_big;
        uint8_t *d_in = encode ? sl.dev_big : sl.dev_small, *d_out = encode ? sl.dev_small : sl.dev_big;
        const size_t in_bytes = encode ? nt : words * 8, out_bytes = encode ? words * 8 : nt;
        cudaError_t e = cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, sl.stream);
        rc = e != cudaSuccess ? CN_OK
             : (encode ? cd.enc(d_in, nt, d_out, cn::kEncPlain, nullptr, sl.stream) : cd.dec(d_in, words, nt, d_out, sl.stream));
        if (e == cudaSuccess && rc == CN_OK) e = cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, sl.stream);
        if (e == cudaSuccess && rc == CN_OK) e = cudaEventRecord(sl.done, sl.stream);
        if (e != cudaSuccess) rc = fail(CN_ERR_CUDA, "batch pipeline: submitting chunk %zu failed: %s", k, cudaGetErrorString(e));
        if (rc != CN_OK) { first_error = rc; break; }
    }
    if (first_error == CN_OK) {
        for (size_t k = nchunks > (size_t)kSlots ? nchunks - kSlots : 0; k < nchunks; k++)
            if ((rc = scatter_chunk(k)) != CN_OK) { first_error = rc; break; }
    }
    for (int k = 0; k < slots_used; k++) {
        if (first_error != CN_OK) cudaStreamSynchronize(p.slot[k].stream);
        pool.wait(p.slot[k].in_job);
        pool.wait(p.slot[k].out_job);
    }
    if (first_error != CN_OK) { cudaGetLastError(); return first_error; }
    for (size_t i : big) {
        const size_t len = lens[i];
        rc = host_codec_one(cd, encode, static_cast<const uint8_t *>(in[i]), static_cast<uint8_t *>(out[i]), len, HostMode{});
        if (rc != CN_OK) return rc;
    }
    return CN_OK;
}

}  // namespace
