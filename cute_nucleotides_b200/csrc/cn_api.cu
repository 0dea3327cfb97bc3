// cn_api.cu -- the extern "C" boundary declared in include/cute_nucleotides_cuda.h.
//
// Two families of entry points:
//   * device-resident (cn_encode_device / cn_decode_device): one kernel launch on the caller's stream;
//     this is the roofline path (BASELINE configs 2-5).
//   * host-slice (cn_n_to_bits_host / cn_bits_to_n_host): the literal drop-in for the reference's
//     fn(&[u8]) -> Vec<u64> / fn(&[u64], usize) -> Vec<u8> (src/n_to_bits.rs:34, :51); PCIe-bound, so
//     it is a chunked pipeline that keeps both DMA directions and the kernel busy at once.
// There is no CPU fallback anywhere in this file: without a CUDA device every call fails with
// CN_ERR_CUDA.
#include "../../include/cute_nucleotides_cuda.h"
#include "codec_kernels.cuh"
#include "codec5_kernels.cuh"

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/mman.h>

namespace {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
thread_local char t_err[512] = "";
const char kPanicText[] = "The length is greater than the number of nucleotides!";   // n_to_bits.rs:53

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CN_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(CN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

std::atomic<uint64_t> g_launches{0};

// ------------------------------------------------------------------------------------------------
// tuning state (process-wide, set once by harnesses; plain loads on the hot path)
// ------------------------------------------------------------------------------------------------
struct Tuning { int vec, unroll, threads; };
// measured best on B200 (profiles/tune_r01.jsonl): 256-bit accesses, one vector per thread, 256 threads
Tuning g_tune[2] = {{32, 1, 256}, {32, 1, 256}};
int g_host_strategy = 0;
size_t g_host_chunk = (size_t)16 << 20;      // ASCII bytes per pipeline chunk
size_t g_host_small = (size_t)256 << 10;     // nucleotides at or below which the single-launch path is used

bool tuning_ok(int vec, int unroll, int threads)
{
    return (vec == 16 || vec == 32) && (unroll == 1 || unroll == 2 || unroll == 4 || unroll == 8) &&
           (threads == 128 || threads == 256 || threads == 512);
}

// ------------------------------------------------------------------------------------------------
// launch plumbing
// ------------------------------------------------------------------------------------------------
struct EncArgs {
    const uint8_t *in; uint32_t *out32; size_t nvec;
    const uint8_t *n0; size_t len, edge_first, edge_total; unsigned shift_bytes;
};
struct DecArgs {
    const uint32_t *bits32; uint8_t *out; size_t nvec; unsigned code_shift;
    uint8_t *out0; size_t head, edge_from, len;
};

constexpr size_t kMaxGrid = 0x7FFFFFFFull;

template <int VEC, int U, int T, bool MIS>
cudaError_t enc_launch(const EncArgs &a, cudaStream_t s)
{
    size_t tile = (size_t)U * T;
    size_t blocks = (a.nvec + tile - 1) / tile;
    if (blocks == 0) blocks = 1;                       // the ragged end still needs its warp
    if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
    cn::encode_kernel<VEC, U, T, MIS><<<(unsigned)blocks, T, 0, s>>>(a.in, a.out32, a.nvec, a.n0, a.len,
                                                                    a.edge_first, a.edge_total, a.shift_bytes, nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <int VEC, int U, int T, bool MIS>
cudaError_t dec_launch(const DecArgs &a, cudaStream_t s)
{
    size_t tile = (size_t)U * T;
    size_t blocks = (a.nvec + tile - 1) / tile;
    if (blocks == 0) blocks = 1;
    if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
    cn::decode_kernel<VEC, U, T, MIS><<<(unsigned)blocks, T, 0, s>>>(a.bits32, a.out, a.nvec, a.code_shift, a.out0,
                                                                    a.head, a.edge_from, a.len);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <bool ENC, int VEC, int U, typename A>
cudaError_t pick_threads(int threads, const A &a, cudaStream_t s)
{
    if constexpr (ENC) {
        switch (threads) {
        case 128: return enc_launch<VEC, U, 128, false>(a, s);
        case 512: return enc_launch<VEC, U, 512, false>(a, s);
        default:  return enc_launch<VEC, U, 256, false>(a, s);
        }
    } else {
        switch (threads) {
        case 128: return dec_launch<VEC, U, 128, false>(a, s);
        case 512: return dec_launch<VEC, U, 512, false>(a, s);
        default:  return dec_launch<VEC, U, 256, false>(a, s);
        }
    }
}
template <bool ENC, int VEC, typename A>
cudaError_t pick_unroll(const Tuning &t, const A &a, cudaStream_t s)
{
    switch (t.unroll) {
    case 1:  return pick_threads<ENC, VEC, 1>(t.threads, a, s);
    case 2:  return pick_threads<ENC, VEC, 2>(t.threads, a, s);
    case 8:  return pick_threads<ENC, VEC, 8>(t.threads, a, s);
    default: return pick_threads<ENC, VEC, 4>(t.threads, a, s);
    }
}

inline uintptr_t addr(const void *p) { return reinterpret_cast<uintptr_t>(p); }

// The tuning asked for 256-bit accesses but the pointer is only 16-byte aligned: use 128-bit accesses
// and twice the unroll so each thread keeps the same number of bytes in flight.
inline Tuning narrow(Tuning t)
{
    if (t.vec == 32) { t.vec = 16; if (t.unroll < 8) t.unroll *= 2; }
    return t;
}

int encode_device(const void *d_n, size_t len, void *d_out, cudaStream_t s)
{
    if (len == 0) return CN_OK;
    if (!d_n || !d_out) return fail(CN_ERR_ARG, "cn_encode_device: null pointer");
    if (addr(d_out) & 7) return fail(CN_ERR_ARG, "cn_encode_device: output must be 8-byte aligned");
    const Tuning t = g_tune[CN_DIR_ENCODE];
    const size_t total32 = cn_words_for_len(len) * 2;          // output u32 units
    EncArgs a{};
    a.n0 = static_cast<const uint8_t *>(d_n);
    a.len = len;
    a.out32 = static_cast<uint32_t *>(d_out);
    a.edge_total = total32;
    cudaError_t e;
    const unsigned mis = (unsigned)(addr(d_n) & 15);
    if (mis == 0) {
        a.in = a.n0;
        if (t.vec == 32 && (addr(d_n) & 31) == 0) {
            a.nvec = len >> 5;
            a.edge_first = a.nvec * 2;
            e = pick_unroll<true, 32>(t, a, s);
        } else {
            a.nvec = len >> 4;
            a.edge_first = a.nvec;
            e = pick_unroll<true, 16>(narrow(t), a, s);
        }
    } else {
        // the body reads aligned vectors i and i+1 for group i; keep both inside [d_n, d_n + len)
        size_t spans = (len + mis) >> 4;
        a.in = a.n0 - mis;
        a.shift_bytes = mis;
        a.nvec = spans > 0 ? spans - 1 : 0;
        a.edge_first = a.nvec;
        e = enc_launch<16, 4, 256, true>(a, s);
    }
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "encode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

// encode + validation in one pass (fixed launch shapes: the measured-best ones)
template <int VEC, int U, int T, bool MIS>
cudaError_t enc_launch_checked(const EncArgs &a, unsigned long long *counter, cudaStream_t s)
{
    size_t tile = (size_t)U * T;
    size_t blocks = (a.nvec + tile - 1) / tile;
    if (blocks == 0) blocks = 1;
    if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
    cn::encode_kernel<VEC, U, T, MIS, true><<<(unsigned)blocks, T, 0, s>>>(a.in, a.out32, a.nvec, a.n0, a.len,
                                                                          a.edge_first, a.edge_total, a.shift_bytes, counter);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

int encode_checked_device(const void *d_n, size_t len, void *d_out, void *d_invalid, cudaStream_t s)
{
    if (!d_invalid || (addr(d_invalid) & 7)) return fail(CN_ERR_ARG, "cn_encode_checked_device: counter must be a non-null 8-byte aligned device pointer");
    if (len == 0) return CN_OK;
    if (!d_n || !d_out) return fail(CN_ERR_ARG, "cn_encode_checked_device: null pointer");
    if (addr(d_out) & 7) return fail(CN_ERR_ARG, "cn_encode_checked_device: output must be 8-byte aligned");
    unsigned long long *counter = static_cast<unsigned long long *>(d_invalid);
    EncArgs a{};
    a.n0 = static_cast<const uint8_t *>(d_n);
    a.len = len;
    a.out32 = static_cast<uint32_t *>(d_out);
    a.edge_total = cn_words_for_len(len) * 2;
    cudaError_t e;
    const unsigned mis = (unsigned)(addr(d_n) & 15);
    if (mis == 0) {
        a.in = a.n0;
        if ((addr(d_n) & 31) == 0) {
            a.nvec = len >> 5;
            a.edge_first = a.nvec * 2;
            e = enc_launch_checked<32, 1, 256, false>(a, counter, s);
        } else {
            a.nvec = len >> 4;
            a.edge_first = a.nvec;
            e = enc_launch_checked<16, 2, 256, false>(a, counter, s);
        }
    } else {
        size_t spans = (len + mis) >> 4;
        a.in = a.n0 - mis;
        a.shift_bytes = mis;
        a.nvec = spans > 0 ? spans - 1 : 0;
        a.edge_first = a.nvec;
        e = enc_launch_checked<16, 4, 256, true>(a, counter, s);
    }
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "checked encode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

int decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, cudaStream_t s)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!d_bits || !d_out) return fail(CN_ERR_ARG, "cn_decode_device: null pointer");
    if (addr(d_bits) & 7) return fail(CN_ERR_ARG, "cn_decode_device: packed input must be 8-byte aligned");
    const Tuning t = g_tune[CN_DIR_DECODE];
    DecArgs a{};
    a.bits32 = static_cast<const uint32_t *>(d_bits);
    a.out0 = static_cast<uint8_t *>(d_out);
    a.len = len;
    cudaError_t e;
    const unsigned mis = (unsigned)(addr(d_out) & 15);
    if (mis == 0) {
        a.out = a.out0;
        if (t.vec == 32 && (addr(d_out) & 31) == 0) {
            a.nvec = len >> 5;
            a.edge_from = a.nvec << 5;
            e = pick_unroll<false, 32>(t, a, s);
        } else {
            a.nvec = len >> 4;
            a.edge_from = a.nvec << 4;
            e = pick_unroll<false, 16>(narrow(t), a, s);
        }
    } else {
        size_t head = 16 - mis;                    // nucleotides until the destination is 16-byte aligned
        if (head > len) head = len;
        a.head = head;
        a.out = a.out0 + head;
        a.code_shift = (unsigned)(2 * head);
        a.nvec = (len - head) >> 4;
        a.edge_from = head + (a.nvec << 4);
        e = dec_launch<16, 4, 256, true>(a, s);
    }
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "decode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

// ---- base-5 codec (src/n_to_bits2.rs) ----------------------------------------------------------------
// Tile staging of the base-5 kernels: bit 0 = cp.async.bulk (TMA) tile load in encode, bit 1 = TMA tile store in
// decode.  Measured on B200 at 10 GiB (profiles/codec5_r01.json): LDG/STS staging 6118 / 6500 GB/s, TMA 6870 / 6712 GB/s,
// so TMA is the default; CN_B5_TMA=0..3 selects the alternatives for A/B runs.
const int g_b5_tma = std::getenv("CN_B5_TMA") ? std::atoi(std::getenv("CN_B5_TMA")) : 3;

inline size_t words2_for_len(size_t len) { return len / 27 + ((len % 27) ? 1 : 0); }

int encode2_device(const void *d_n, size_t len, void *d_out, cudaStream_t s)
{
    if (len == 0) return CN_OK;
    if (!d_n || !d_out) return fail(CN_ERR_ARG, "cn_encode2_device: null pointer");
    if (addr(d_out) & 7) return fail(CN_ERR_ARG, "cn_encode2_device: output must be 8-byte aligned");
    const size_t total = words2_for_len(len);
    const uint8_t *in = static_cast<const uint8_t *>(d_n);
    uint64_t *out = static_cast<uint64_t *>(d_out);
    if ((addr(d_n) & 15) == 0 && (addr(d_out) & 31) == 0) {                  // 128-bit ASCII loads, 256-bit packed stores
        const size_t ntiles = (len / 27) / cn::kB5WarpWords;                 // tiles of complete words
        const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
        if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "cn_encode2_device: input too large for one launch");
        if (g_b5_tma & 1) cn::b5_encode_kernel<true><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(in, out, len, ntiles, total);
        else cn::b5_encode_kernel<false><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(in, out, len, ntiles, total);
    } else {
        size_t blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        cn::b5_encode_scalar_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, len, total);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "base-5 encode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

int decode2_device(const void *d_bits, size_t nwords, size_t len, void *d_out, cudaStream_t s)
{
    if (nwords > (~(size_t)0) / 27 || len > nwords * 27) return fail(CN_ERR_LENGTH, "%s", kPanicText);   // n_to_bits2.rs:79-81
    if (len == 0) return CN_OK;
    if (!d_bits || !d_out) return fail(CN_ERR_ARG, "cn_decode2_device: null pointer");
    if (addr(d_bits) & 7) return fail(CN_ERR_ARG, "cn_decode2_device: packed input must be 8-byte aligned");
    const size_t total = words2_for_len(len);
    const uint64_t *bits = static_cast<const uint64_t *>(d_bits);
    uint8_t *out = static_cast<uint8_t *>(d_out);
    if ((addr(d_out) & 15) == 0 && (addr(d_bits) & 31) == 0) {
        const size_t ntiles = (len / 27) / cn::kB5WarpWords;
        const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
        if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "cn_decode2_device: input too large for one launch");
        if (g_b5_tma & 2) cn::b5_decode_kernel<true><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(bits, out, len, ntiles, total);
        else cn::b5_decode_kernel<false><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(bits, out, len, ntiles, total);
    } else {
        size_t blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        cn::b5_decode_scalar_kernel<<<(unsigned)blocks, 256, 0, s>>>(bits, out, len, total);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "base-5 decode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

// What the host pipeline needs to know about a codec: nucleotides per word and the device entry points.
struct Codec {
    unsigned group;                                                        // 32 (2-bit) or 27 (base-5)
    int (*enc)(const void *, size_t, void *, cudaStream_t);
    int (*dec)(const void *, size_t, size_t, void *, cudaStream_t);
    size_t words(size_t nt) const { return nt / group + ((nt % group) ? 1 : 0); }
};
const Codec kCodec2bit{32, encode_device, decode_device};
const Codec kCodecBase5{27, encode2_device, decode2_device};

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
        d_counter = nullptr;
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
    if (p.ready && p.device == dev && p.chunk == g_host_chunk) return CN_OK;
    p.destroy();
    p.device = dev;
    p.chunk = g_host_chunk;
    for (auto &sl : p.slot) {
        CN_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CN_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        if (cudaHostAlloc(&sl.pin_big, p.chunk, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&sl.pin_small, p.chunk / 4 + 64, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&sl.dev_big, p.chunk) != cudaSuccess ||
            cudaMalloc(&sl.dev_small, p.chunk / 4 + 64) != cudaSuccess) {
            cudaGetLastError();
            p.destroy();
            return fail(CN_ERR_NOMEM, "host pipeline: staging allocation of %zu bytes per slot failed", p.chunk);
        }
    }
    if (cudaMalloc(&p.d_counter, sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        p.destroy();
        return fail(CN_ERR_NOMEM, "host pipeline: counter allocation failed");
    }
    p.ready = true;
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
    if (len <= g_host_small && len <= max_nt) {
        Slot &sl = p.slot[0];
        uint8_t *pin_in = encode ? sl.pin_big : sl.pin_small;
        uint8_t *pin_out = encode ? sl.pin_small : sl.pin_big;
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

// internal stream + events for cn_time_*_device
struct Timer {
    cudaStream_t stream = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};
thread_local Timer t_timer;

int timer_prepare(Timer &t)
{
    if (t.stream) return CN_OK;
    CN_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    CN_CUDA(cudaEventCreate(&t.t0));
    CN_CUDA(cudaEventCreate(&t.t1));
    return CN_OK;
}

}  // namespace

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int cn_abi_version(void) { return CN_ABI_VERSION; }

int cn_init(int device)
{
    int count = 0;
    CN_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(CN_ERR_ARG, "cn_init: device %d out of range (0..%d)", device, count - 1);
    CN_CUDA(cudaSetDevice(device));
    CN_CUDA(cudaFree(nullptr));            // force context creation so later calls do not pay for it
    t_device = device;
    return CN_OK;
}

int cn_device_count(int *count)
{
    if (!count) return fail(CN_ERR_ARG, "cn_device_count: null pointer");
    CN_CUDA(cudaGetDeviceCount(count));
    return CN_OK;
}

const char *cn_last_error(void) { return t_err; }
const char *cn_length_panic_message(void) { return kPanicText; }
size_t cn_words_for_len(size_t len) { return (len >> 5) + ((len & 31) ? 1 : 0); }
uint64_t cn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int cn_n_to_bits_host(const uint8_t *n, size_t len, uint64_t *out)
{
    if (len == 0) return CN_OK;                                 // reference: zero-size alloc; here: nothing to do
    if (!n || !out) return fail(CN_ERR_ARG, "cn_n_to_bits_host: null pointer");
    return host_codec(kCodec2bit, true, n, reinterpret_cast<uint8_t *>(out), len);
}

int cn_bits_to_n_host(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!bits || !out) return fail(CN_ERR_ARG, "cn_bits_to_n_host: null pointer");
    return host_codec(kCodec2bit, false, reinterpret_cast<const uint8_t *>(bits), out, len);
}

int cn_n_to_bits_checked_host(const uint8_t *n, size_t len, uint64_t *out, uint64_t *invalid_count)
{
    if (!invalid_count) return fail(CN_ERR_ARG, "cn_n_to_bits_checked_host: null counter");
    *invalid_count = 0;
    if (len == 0) return CN_OK;
    if (!n || !out) return fail(CN_ERR_ARG, "cn_n_to_bits_checked_host: null pointer");
    return host_codec(kCodec2bit, true, n, reinterpret_cast<uint8_t *>(out), len, invalid_count);
}

int cn_encode_checked_device(const void *d_n, size_t len, void *d_out, void *d_invalid_count, void *stream)
{
    return encode_checked_device(d_n, len, d_out, d_invalid_count, static_cast<cudaStream_t>(stream));
}

int cn_encode_device(const void *d_n, size_t len, void *d_out, void *stream)
{
    return encode_device(d_n, len, d_out, static_cast<cudaStream_t>(stream));
}

int cn_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream)
{
    return decode_device(d_bits, nwords, len, d_out, static_cast<cudaStream_t>(stream));
}

size_t cn_words2_for_len(size_t len) { return words2_for_len(len); }

int cn_n_to_bits2_host(const uint8_t *n, size_t len, uint64_t *out)
{
    if (len == 0) return CN_OK;
    if (!n || !out) return fail(CN_ERR_ARG, "cn_n_to_bits2_host: null pointer");
    return host_codec(kCodecBase5, true, n, reinterpret_cast<uint8_t *>(out), len);
}

int cn_bits_to_n2_host(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (nwords > (~(size_t)0) / 27 || len > nwords * 27) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!bits || !out) return fail(CN_ERR_ARG, "cn_bits_to_n2_host: null pointer");
    return host_codec(kCodecBase5, false, reinterpret_cast<const uint8_t *>(bits), out, len);
}

int cn_encode2_device(const void *d_n, size_t len, void *d_out, void *stream)
{
    return encode2_device(d_n, len, d_out, static_cast<cudaStream_t>(stream));
}

int cn_decode2_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream)
{
    return decode2_device(d_bits, nwords, len, d_out, static_cast<cudaStream_t>(stream));
}

int cn_generate2_device(void *d_out, size_t offset, size_t len, uint64_t seed, int alphabet, void *stream)
{
    if (alphabet != 5 && alphabet != 12) return fail(CN_ERR_ARG, "cn_generate2_device: alphabet must be 5 or 12");
    if (len == 0) return CN_OK;
    if (!d_out) return fail(CN_ERR_ARG, "cn_generate2_device: null pointer");
    size_t blocks = (len + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    cn::b5_generate_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint8_t *>(d_out), offset, len, seed, alphabet);
    CN_CUDA(cudaGetLastError());
    return CN_OK;
}

int cn_generate_device(void *d_out, size_t offset, size_t len, uint64_t seed, int alphabet, void *stream)
{
    if (alphabet != 4 && alphabet != 10) return fail(CN_ERR_ARG, "cn_generate_device: alphabet must be 4 or 10");
    if (offset & 7) return fail(CN_ERR_ARG, "cn_generate_device: offset must be a multiple of 8");
    if (len == 0) return CN_OK;
    if (!d_out) return fail(CN_ERR_ARG, "cn_generate_device: null pointer");
    size_t groups = (len + 7) >> 3;
    size_t blocks = (groups + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    cn::generate_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint8_t *>(d_out), offset, len, seed, alphabet);
    CN_CUDA(cudaGetLastError());
    return CN_OK;
}

int cn_generate_words_device(void *d_out, size_t first_word, size_t nwords, uint64_t seed, void *stream)
{
    if (nwords == 0) return CN_OK;
    if (!d_out || (addr(d_out) & 7)) return fail(CN_ERR_ARG, "cn_generate_words_device: null or misaligned pointer");
    size_t blocks = (nwords + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    cn::generate_words_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint64_t *>(d_out), first_word, nwords, seed);
    CN_CUDA(cudaGetLastError());
    return CN_OK;
}

int cn_device_malloc(void **d_ptr, size_t bytes)
{
    if (!d_ptr) return fail(CN_ERR_ARG, "cn_device_malloc: null pointer");
    *d_ptr = nullptr;
    if (bytes == 0) return CN_OK;
    cudaError_t e = cudaMalloc(d_ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CN_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    return CN_OK;
}
int cn_device_free(void *d_ptr) { if (d_ptr) CN_CUDA(cudaFree(d_ptr)); return CN_OK; }

int cn_host_malloc_pinned(void **h_ptr, size_t bytes)
{
    if (!h_ptr) return fail(CN_ERR_ARG, "cn_host_malloc_pinned: null pointer");
    *h_ptr = nullptr;
    if (bytes == 0) return CN_OK;
    cudaError_t e = cudaHostAlloc(h_ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CN_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    return CN_OK;
}
int cn_host_free_pinned(void *h_ptr) { if (h_ptr) CN_CUDA(cudaFreeHost(h_ptr)); return CN_OK; }

int cn_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    if (bytes) CN_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return CN_OK;
}
int cn_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    if (bytes) CN_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return CN_OK;
}
int cn_stream_synchronize(void *stream) { CN_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); return CN_OK; }

int cn_time_encode_device(const void *d_n, size_t len, void *d_out, int iters, float *ms_total)
{
    if (!ms_total || iters < 1) return fail(CN_ERR_ARG, "cn_time_encode_device: bad arguments");
    Timer &t = t_timer;
    int rc = timer_prepare(t);
    if (rc != CN_OK) return rc;
    CN_CUDA(cudaStreamSynchronize(t.stream));
    CN_CUDA(cudaEventRecord(t.t0, t.stream));
    for (int i = 0; i < iters; i++) {
        rc = encode_device(d_n, len, d_out, t.stream);
        if (rc != CN_OK) return rc;
    }
    CN_CUDA(cudaEventRecord(t.t1, t.stream));
    CN_CUDA(cudaEventSynchronize(t.t1));
    CN_CUDA(cudaEventElapsedTime(ms_total, t.t0, t.t1));
    return CN_OK;
}

int cn_time_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, int iters, float *ms_total)
{
    if (!ms_total || iters < 1) return fail(CN_ERR_ARG, "cn_time_decode_device: bad arguments");
    Timer &t = t_timer;
    int rc = timer_prepare(t);
    if (rc != CN_OK) return rc;
    CN_CUDA(cudaStreamSynchronize(t.stream));
    CN_CUDA(cudaEventRecord(t.t0, t.stream));
    for (int i = 0; i < iters; i++) {
        rc = decode_device(d_bits, nwords, len, d_out, t.stream);
        if (rc != CN_OK) return rc;
    }
    CN_CUDA(cudaEventRecord(t.t1, t.stream));
    CN_CUDA(cudaEventSynchronize(t.t1));
    CN_CUDA(cudaEventElapsedTime(ms_total, t.t0, t.t1));
    return CN_OK;
}

int cn_set_tuning(int direction, int vec, int unroll, int threads)
{
    if ((direction != CN_DIR_ENCODE && direction != CN_DIR_DECODE) || !tuning_ok(vec, unroll, threads))
        return fail(CN_ERR_ARG, "cn_set_tuning: unsupported (direction=%d vec=%d unroll=%d threads=%d)", direction, vec, unroll, threads);
    g_tune[direction] = Tuning{vec, unroll, threads};
    return CN_OK;
}

int cn_get_tuning(int direction, int *vec, int *unroll, int *threads)
{
    if ((direction != CN_DIR_ENCODE && direction != CN_DIR_DECODE) || !vec || !unroll || !threads)
        return fail(CN_ERR_ARG, "cn_get_tuning: bad arguments");
    *vec = g_tune[direction].vec; *unroll = g_tune[direction].unroll; *threads = g_tune[direction].threads;
    return CN_OK;
}

int cn_set_host_strategy(int strategy, size_t chunk_bytes)
{
    if (strategy != 0 && strategy != 1) return fail(CN_ERR_ARG, "cn_set_host_strategy: strategy must be 0 or 1");
    if (chunk_bytes) {
        if (chunk_bytes < 4096 || (chunk_bytes & 4095) || chunk_bytes > ((size_t)1 << 30))
            return fail(CN_ERR_ARG, "cn_set_host_strategy: chunk must be a multiple of 4096 in [4 KiB, 1 GiB]");
        g_host_chunk = chunk_bytes;
    }
    g_host_strategy = strategy;
    return CN_OK;
}

}  // extern "C"
