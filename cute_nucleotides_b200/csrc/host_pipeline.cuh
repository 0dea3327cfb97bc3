// host_pipeline.cuh -- the host-slice path behind cn_n_to_bits_host / cn_bits_to_n_host (and the base-5 pair):
// a process-wide pool of staging copier threads and a per-thread ring of {stream, event, pinned + device staging}
// slots that keeps H2D copies, kernels and D2H copies of consecutive chunks in flight at once.
#pragma once
#include "cn_launch.cuh"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/mman.h>

namespace {

// ------------------------------------------------------------------------------------------------
// staging copier.  Pageable caller memory cannot be DMA'd, so it is copied through pinned staging;
// one thread moves ~10 GB/s (less into never-touched pages), PCIe Gen5 moves > 50 GB/s, so large
// copies are cut into slices executed by a small process-wide pool (CN_HOST_THREADS, default
// min(8, cores/2)) with the calling thread taking a slice too.  The pool is created on first use and
// intentionally never destroyed (its threads only ever wait on a condition variable).
// ------------------------------------------------------------------------------------------------
class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool *pool = new CopyPool();      // leaked on purpose: no join at process exit
        return *pool;
    }

    void copy(void *dst, const void *src, size_t bytes)
    {
        constexpr size_t kMinSlice = (size_t)256 << 10;
        size_t parts = bytes / kMinSlice;
        if (parts > threads_.size() + 1) parts = threads_.size() + 1;
        if (parts <= 1) { memcpy(dst, src, bytes); return; }
        const size_t slice = ((bytes / parts) + 4095) & ~(size_t)4095;
        Job job;
        size_t off = slice;                            // slice 0 is the caller's
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (off < bytes) {
                size_t n = bytes - off < slice ? bytes - off : slice;
                queue_.push_back(Task{static_cast<uint8_t *>(dst) + off, static_cast<const uint8_t *>(src) + off, n, &job});
                job.pending++;
                off += n;
            }
        }
        cv_.notify_all();
        memcpy(dst, src, slice < bytes ? slice : bytes);
        std::unique_lock<std::mutex> lk(mu_);
        job.done_cv.wait(lk, [&] { return job.pending == 0; });
    }

private:
    struct Job { int pending = 0; std::condition_variable done_cv; };
    struct Task { uint8_t *dst; const uint8_t *src; size_t bytes; Job *job; };

    CopyPool()
    {
        int n = 0;
        if (const char *env = std::getenv("CN_HOST_THREADS")) n = std::atoi(env) - 1;
        else {
            unsigned hc = std::thread::hardware_concurrency();
            n = (int)(hc / 2 > 8 ? 8 : hc / 2) - 1;
        }
        if (n < 0) n = 0;
        if (n > 63) n = 63;
        for (int i = 0; i < n; i++) {
            threads_.emplace_back([this] { run(); });
            threads_.back().detach();
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
            memcpy(t.dst, t.src, t.bytes);
            lk.lock();
            if (--t.job->pending == 0) t.job->done_cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Task> queue_;
    std::vector<std::thread> threads_;
};

inline void staged_copy(void *dst, const void *src, size_t bytes)
{
    if (bytes < ((size_t)512 << 10)) memcpy(dst, src, bytes);
    else CopyPool::get().copy(dst, src, bytes);
}

// ------------------------------------------------------------------------------------------------
// host-slice pipeline.  Per calling thread: kSlots slots, each with its own stream, a pinned and a
// device staging buffer per direction.  Chunk c uses slot c % kSlots, so while chunk c's H2D copy
// runs, chunk c-1's kernel and chunk c-2's D2H copy run on other streams/copy engines.
// ------------------------------------------------------------------------------------------------
constexpr int kSlots = 4;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *pin_big = nullptr, *pin_small = nullptr;     // ASCII-sized / packed-sized pinned staging
    uint8_t *dev_big = nullptr, *dev_small = nullptr;
    bool busy = false;
    size_t dst_off = 0, dst_bytes = 0;                    // where the staged result goes once `done`
};

struct HostPipe {
    int device = -1;
    size_t chunk = 0;
    Slot slot[kSlots];
    unsigned long long *d_counter = nullptr;              // invalid-byte counter of the checked encode
    uint8_t *tiny_in = nullptr, *tiny_out = nullptr;      // pinned staging of the single-launch small-input path
    size_t tiny_bytes = 0;
    bool ready = false;

    void destroy()
    {
        for (auto &sl : slot) {
            if (sl.pin_big) cudaFreeHost(sl.pin_big);
            if (sl.pin_small) cudaFreeHost(sl.pin_small);
            if (sl.dev_big) cudaFree(sl.dev_big);
            if (sl.dev_small) cudaFree(sl.dev_small);
            if (sl.done) cudaEventDestroy(sl.done);
            if (sl.stream) cudaStreamDestroy(sl.stream);
            sl = Slot{};
        }
        if (d_counter) cudaFree(d_counter);
        if (tiny_in) cudaFreeHost(tiny_in);
        if (tiny_out) cudaFreeHost(tiny_out);
        d_counter = nullptr;
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
    if (p.ready && p.device == dev && p.chunk == g_host_chunk) {
        // A previous call that failed half-way returned early and may have left slots marked busy with ITS destination
        // offsets; never let them be retired into this call's buffers.  (Normally nothing is busy here.)
        for (auto &sl : p.slot)
            if (sl.busy) { cudaStreamSynchronize(sl.stream); cudaGetLastError(); sl.busy = false; sl.dst_bytes = 0; }
        return CN_OK;
    }
    p.destroy();
    p.device = dev;
    p.chunk = g_host_chunk;
    // Only the cheap parts are created up front; the chunk-sized staging of a slot is allocated the first time a
    // call needs it (slot_ensure), so a thread that only ever makes small calls pins 2 x 256 KiB, not 80 MiB.
    for (auto &sl : p.slot) {
        CN_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CN_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    p.tiny_bytes = g_host_small + 64;
    if (cudaMalloc(&p.d_counter, sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc(&p.tiny_in, p.tiny_bytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&p.tiny_out, p.tiny_bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        p.destroy();
        return fail(CN_ERR_NOMEM, "host pipeline: small staging allocation failed");
    }
    p.ready = true;
    return CN_OK;
}

// chunk-sized staging of one slot, on first use: pinned buffers when a side of the call is pageable, device buffers
// when the copy engines are used (strategy 0)
int slot_ensure(HostPipe &p, Slot &sl, bool need_pinned, bool need_device)
{
    const size_t big = p.chunk, small = p.chunk / 4 + 64;
    if (need_pinned && !sl.pin_big) {
        if (cudaHostAlloc(&sl.pin_big, big, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&sl.pin_small, small, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(CN_ERR_NOMEM, "host pipeline: pinned staging allocation of %zu bytes failed", big + small);
        }
    }
    if (need_device && !sl.dev_big) {
        if (cudaMalloc(&sl.dev_big, big) != cudaSuccess || cudaMalloc(&sl.dev_small, small) != cudaSuccess) {
            cudaGetLastError();
            return fail(CN_ERR_NOMEM, "host pipeline: device staging allocation of %zu bytes failed", big + small);
        }
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

// One implementation for both directions and both codecs: `big` is the ASCII side, `small` the packed side.
//   encode: src = ASCII (len bytes)          dst = packed (8*words bytes)
//   decode: src = packed (8*nwords bytes)    dst = ASCII (len bytes)
int host_codec(const Codec &cd, bool encode, const uint8_t *src, uint8_t *dst, size_t len, uint64_t *invalid_out = nullptr)
{
    HostPipe &p = t_pipe;
    int rc = pipe_prepare(p);
    if (rc != CN_OK) return rc;

    const size_t nwords = cd.words(len);
    const size_t src_bytes = encode ? len : nwords * 8;
    const size_t dst_bytes = encode ? nwords * 8 : len;
    // Staging holds p.chunk ASCII bytes and p.chunk/4 (+64) packed bytes per slot; the base-5 codec packs
    // 8 bytes per 27 nucleotides (> 1/4), so its chunks are 3/4 of the staging size.
    const size_t max_nt = cd.group == 32 ? p.chunk : p.chunk / 4 * 3;
    // checked encode (2-bit codec only): every chunk's kernel adds into one device counter, read back at the end
    const bool checked = invalid_out != nullptr;
    auto run_encode = [&](const void *in, size_t nt, void *out, cudaStream_t s) {
        return checked ? encode_checked_device(in, nt, out, p.d_counter, s) : cd.enc(in, nt, out, s);
    };
    auto finish_checked = [&]() -> int {
        if (!checked) return CN_OK;
        unsigned long long v = 0;
        CN_CUDA(cudaMemcpy(&v, p.d_counter, sizeof v, cudaMemcpyDeviceToHost));
        *invalid_out = v;
        return CN_OK;
    };
    if (checked) CN_CUDA(cudaMemset(p.d_counter, 0, sizeof(unsigned long long)));

    // Small inputs (the reference's own bench is 40 000 nt) are latency-bound: skip the pointer queries and
    // the copy engines, stage through slot 0's pinned buffers and let ONE kernel read and write them in place
    // over PCIe -- a single launch and a single synchronisation.
    if (len <= g_host_small && len <= max_nt && src_bytes <= p.tiny_bytes && dst_bytes <= p.tiny_bytes) {
        Slot &sl = p.slot[0];
        uint8_t *pin_in = p.tiny_in, *pin_out = p.tiny_out;
        memcpy(pin_in, src, src_bytes);
        rc = encode ? run_encode(pin_in, len, pin_out, sl.stream) : cd.dec(pin_in, nwords, len, pin_out, sl.stream);
        if (rc != CN_OK) return rc;
        CN_CUDA(cudaStreamSynchronize(sl.stream));
        memcpy(dst, pin_out, dst_bytes);
        return finish_checked();
    }

    const bool src_pinned = is_pinned(src, src_bytes);
    const bool dst_pinned = is_pinned(dst, dst_bytes);
    const bool zero_copy = g_host_strategy == 1;

    // zero-copy with both sides page-locked: a single kernel streams over PCIe in both directions
    if (zero_copy && src_pinned && dst_pinned) {
        cudaStream_t s = p.slot[0].stream;
        rc = encode ? run_encode(src, len, dst, s) : cd.dec(src, nwords, len, dst, s);
        if (rc != CN_OK) return rc;
        CN_CUDA(cudaStreamSynchronize(s));
        return finish_checked();
    }

    // Nucleotides per chunk: a multiple of `unit` (whole words, and whole warp tiles / 16-byte vectors where the
    // staging size allows).  Mid-sized inputs are cut into ~8 chunks so that staging copies, both DMA directions
    // and the kernel overlap; large inputs use the full staging size.
    size_t unit = cd.group == 32 ? 4096 : (size_t)cn::kB5WarpBytes;
    if (max_nt < unit) unit = cd.group;
    size_t chunk = max_nt / unit * unit;
    if (len / 8 < chunk) {
        size_t c8 = (len / 8 + unit - 1) / unit * unit;
        size_t floor_nt = ((size_t)1 << 20) / unit * unit;
        if (floor_nt == 0 || floor_nt > chunk) floor_nt = chunk;
        chunk = c8 < floor_nt ? floor_nt : (c8 < chunk ? c8 : chunk);
    }
    // A large pageable destination is usually a freshly allocated Vec: ask for transparent huge pages so the
    // first-touch faults taken while copying the result out are per 2 MiB, not per 4 KiB (hint only).
    if (!dst_pinned && dst_bytes >= ((size_t)8 << 20)) {
        const uintptr_t lo = (addr(dst) + 0x1FFFFF) & ~(uintptr_t)0x1FFFFF, hi = (addr(dst) + dst_bytes) & ~(uintptr_t)0x1FFFFF;
        if (hi > lo) (void)madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
    }
    size_t done_nt = 0;
    int c = 0;
    int first_error = CN_OK;
    while (done_nt < len) {
        Slot &sl = p.slot[c % kSlots];
        // retire whatever this slot was doing kSlots chunks ago
        if (sl.busy) {
            CN_CUDA(cudaEventSynchronize(sl.done));
            if (sl.dst_bytes) staged_copy(dst + sl.dst_off, encode ? sl.pin_small : sl.pin_big, sl.dst_bytes);
            sl.busy = false;
        }
        rc = slot_ensure(p, sl, !src_pinned || !dst_pinned, !zero_copy);
        if (rc != CN_OK) { first_error = rc; break; }
        const size_t nt = (len - done_nt < chunk) ? len - done_nt : chunk;
        const size_t words = cd.words(nt);
        const size_t word_off = done_nt / cd.group * 8;             // done_nt is a multiple of the group size
        const size_t in_off = encode ? done_nt : word_off;
        const size_t in_bytes = encode ? nt : words * 8;
        const size_t out_off = encode ? word_off : done_nt;
        const size_t out_bytes = encode ? words * 8 : nt;
        uint8_t *pin_in = encode ? sl.pin_big : sl.pin_small;
        uint8_t *pin_out = encode ? sl.pin_small : sl.pin_big;
        uint8_t *dev_in = encode ? sl.dev_big : sl.dev_small;
        uint8_t *dev_out = encode ? sl.dev_small : sl.dev_big;

        const uint8_t *h_in = src + in_off;
        if (!src_pinned) { staged_copy(pin_in, h_in, in_bytes); h_in = pin_in; }
        uint8_t *h_out = dst_pinned ? dst + out_off : pin_out;

        if (zero_copy) {
            // kernel dereferences the pinned staging (or the caller's pinned side) directly
            rc = encode ? run_encode(h_in, nt, h_out, sl.stream) : cd.dec(h_in, words, nt, h_out, sl.stream);
        } else {
            CN_CUDA(cudaMemcpyAsync(dev_in, h_in, in_bytes, cudaMemcpyHostToDevice, sl.stream));
            rc = encode ? run_encode(dev_in, nt, dev_out, sl.stream) : cd.dec(dev_in, words, nt, dev_out, sl.stream);
            if (rc == CN_OK) CN_CUDA(cudaMemcpyAsync(h_out, dev_out, out_bytes, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (rc != CN_OK) { first_error = rc; break; }
        CN_CUDA(cudaEventRecord(sl.done, sl.stream));
        sl.busy = true;
        sl.dst_off = out_off;
        sl.dst_bytes = dst_pinned ? 0 : out_bytes;
        done_nt += nt;
        c++;
    }
    // drain in submission order
    for (int k = 0; k < kSlots; k++) {
        Slot &sl = p.slot[(c + k) % kSlots];
        if (!sl.busy) continue;
        cudaError_t e = cudaEventSynchronize(sl.done);
        if (e != cudaSuccess && first_error == CN_OK)
            first_error = fail(CN_ERR_CUDA, "host pipeline drain failed: %s", cudaGetErrorString(e));
        if (e == cudaSuccess && sl.dst_bytes) staged_copy(dst + sl.dst_off, encode ? sl.pin_small : sl.pin_big, sl.dst_bytes);
        sl.busy = false;
    }
    if (first_error == CN_OK) first_error = finish_checked();
    return first_error;
}

}  // namespace
