// cn_launch.cuh -- device-resident entry points: pick a kernel instantiation for the pointers' alignment and the
// current tuning, launch it on the caller's stream.  One launch per call, no synchronisation, no allocation.
#pragma once
#include "cn_common.cuh"
#include "codec_kernels.cuh"
#include "codec5_kernels.cuh"
#include "segmented_kernels.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// tuning state (process-wide harness knobs).  Each knob is ONE atomic word, so a call that races with a
// cn_set_* sees either the old or the new setting, never a torn one; every call snapshots what it needs once.
// ------------------------------------------------------------------------------------------------
struct Tuning { int vec, unroll, threads; };
constexpr uint32_t pack_tuning(int vec, int unroll, int threads) { return (uint32_t)vec | ((uint32_t)unroll << 8) | ((uint32_t)threads << 16); }
// measured best on B200 (profiles/tune_r01.jsonl): 256-bit accesses, one vector per thread, 256 threads
std::atomic<uint32_t> g_tune[2] = {{pack_tuning(32, 1, 256)}, {pack_tuning(32, 1, 256)}};
inline Tuning load_tuning(int direction)
{
    const uint32_t v = g_tune[direction].load(std::memory_order_relaxed);
    return Tuning{(int)(v & 0xFF), (int)((v >> 8) & 0xFF), (int)(v >> 16)};
}
std::atomic<int> g_host_strategy{0};
std::atomic<size_t> g_host_chunk{(size_t)16 << 20};      // ASCII bytes per pipeline chunk
const size_t g_host_small = (size_t)256 << 10;           // nucleotides at or below which the single-launch path is used

bool tuning_ok(int vec, int unroll, int threads)
{
    return (vec == 16 || vec == 32) && (unroll == 1 || unroll == 2 || unroll == 4 || unroll == 8) &&
           (threads == 64 || threads == 128 || threads == 256 || threads == 512);
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
        case 64:  return enc_launch<VEC, U, 64, false>(a, s);
        case 128: return enc_launch<VEC, U, 128, false>(a, s);
        case 512: return enc_launch<VEC, U, 512, false>(a, s);
        default:  return enc_launch<VEC, U, 256, false>(a, s);
        }
    } else {
        switch (threads) {
        case 64:  return dec_launch<VEC, U, 64, false>(a, s);
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

// The tuning asked for 256-bit accesses but the pointer is only 16-byte aligned: use 128-bit accesses
// and twice the unroll so each thread keeps the same number of bytes in flight.
inline Tuning narrow(Tuning t)
{
    if (t.vec == 32) { t.vec = 16; if (t.unroll < 8) t.unroll *= 2; }
    return t;
}

// encode + validation / _lut-exact semantics in one pass (fixed launch shapes: the measured-best ones for the kernel
// that carries the extra ALU work)
template <int VEC, int U, int T, bool MIS, int MODE>
cudaError_t enc_launch_mode(const EncArgs &a, unsigned long long *counter, cudaStream_t s)
{
    size_t tile = (size_t)U * T;
    size_t blocks = (a.nvec + tile - 1) / tile;
    if (blocks == 0) blocks = 1;
    if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
    cn::encode_kernel<VEC, U, T, MIS, MODE><<<(unsigned)blocks, T, 0, s>>>(a.in, a.out32, a.nvec, a.n0, a.len,
                                                                          a.edge_first, a.edge_total, a.shift_bytes, counter);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <int MODE>
cudaError_t enc_dispatch_mode(const EncArgs &a0, unsigned long long *counter, cudaStream_t s)
{
    EncArgs a = a0;
    const unsigned mis = (unsigned)(addr(a.n0) & 15);
    if (mis == 0) {
        a.in = a.n0;
        if ((addr(a.n0) & 31) == 0) {
            a.nvec = a.len >> 5;
            a.edge_first = a.nvec * 2;
            // two vectors in flight per thread hide the extra ALU work: 1.845 ms vs 1.843 ms unchecked at 10 GiB
            // (one vector per thread, the best shape for the plain kernel, costs +12 % here)
            return enc_launch_mode<32, 2, 256, false, MODE>(a, counter, s);
        }
        a.nvec = a.len >> 4;
        a.edge_first = a.nvec;
        return enc_launch_mode<16, 4, 256, false, MODE>(a, counter, s);
    }
    size_t spans = (a.len + mis) >> 4;
    a.in = a.n0 - mis;                 // rounded down to 16 B; the body never dereferences the vector that starts before n0
    a.shift_bytes = mis;
    a.nvec = spans > 0 ? spans - 1 : 0;
    a.edge_first = a.nvec;
    return enc_launch_mode<16, 4, 256, true, MODE>(a, counter, s);
}

// mode: cn::kEncPlain / kEncCount / kEncLutExact.  `counter` (device, 8-byte aligned) is incremented by the number of
// bytes outside the alphabet in the non-plain modes; it may be null for kEncLutExact.
int encode_device(const void *d_n, size_t len, void *d_out, int mode, unsigned long long *counter, cudaStream_t s)
{
    if (mode != cn::kEncPlain && mode != cn::kEncCount && mode != cn::kEncLutExact) return fail(CN_ERR_ARG, "encode: bad mode %d", mode);
    if (mode == cn::kEncCount && !counter) return fail(CN_ERR_ARG, "cn_encode_checked_device: counter must be a non-null 8-byte aligned device pointer");
    if (counter && (addr(counter) & 7)) return fail(CN_ERR_ARG, "encode: the invalid-byte counter must be 8-byte aligned");
    if (len == 0) return CN_OK;
    if (!d_n || !d_out) return fail(CN_ERR_ARG, "cn_encode_device: null pointer");
    if (addr(d_out) & 7) return fail(CN_ERR_ARG, "cn_encode_device: output must be 8-byte aligned");
    const size_t total32 = cn_words_for_len(len) * 2;          // output u32 units
    EncArgs a{};
    a.n0 = static_cast<const uint8_t *>(d_n);
    a.len = len;
    a.out32 = static_cast<uint32_t *>(d_out);
    a.edge_total = total32;
    cudaError_t e;
    if (mode == cn::kEncCount) e = enc_dispatch_mode<cn::kEncCount>(a, counter, s);
    else if (mode == cn::kEncLutExact) e = enc_dispatch_mode<cn::kEncLutExact>(a, counter, s);
    else {
        const Tuning t = load_tuning(CN_DIR_ENCODE);
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
            // the body reads aligned vectors i and i+1 for group i >= 1, all inside [d_n, d_n + len); group 0 (whose first
            // vector starts before d_n) is encoded from the caller's own bytes by the edge warp
            size_t spans = (len + mis) >> 4;
            a.in = a.n0 - mis;
            a.shift_bytes = mis;
            a.nvec = spans > 0 ? spans - 1 : 0;
            a.edge_first = a.nvec;
            e = enc_launch<16, 4, 256, true>(a, s);
        }
    }
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "encode kernel launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

int encode_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, cudaStream_t s)
{
    if (nout < 1 || nout > cn::kMaxPeers || !d_outs) return fail(CN_ERR_ARG, "cn_encode_multi_device: 1..%d destinations", cn::kMaxPeers);
    if (len == 0) return CN_OK;
    if (!d_n || (addr(d_n) & 31)) return fail(CN_ERR_ARG, "cn_encode_multi_device: input must be 32-byte aligned");
    cn::PeerOuts outs{};
    for (int d = 0; d < nout; d++) {
        if (!d_outs[d] || (addr(d_outs[d]) & 7)) return fail(CN_ERR_ARG, "cn_encode_multi_device: destination %d null or not 8-byte aligned", d);
        outs.p[d] = static_cast<uint32_t *>(d_outs[d]);
    }
    const size_t nvec = len >> 5;
    size_t blocks = (nvec + 255) / 256;
    if (blocks == 0) blocks = 1;
    if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "cn_encode_multi_device: input too large for one launch");
    cn::encode_multi_kernel<256><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const uint8_t *>(d_n), outs, nout, nvec, len,
                                                                 nvec * 2, cn_words_for_len(len) * 2);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "multi-destination encode launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

// base-5 flavour of the above: d_n 16-byte aligned, destinations 32-byte aligned
int encode2_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, cudaStream_t s)
{
    if (nout < 1 || nout > cn::kMaxPeers || !d_outs) return fail(CN_ERR_ARG, "cn_encode2_multi_device: 1..%d destinations", cn::kMaxPeers);
    if (len == 0) return CN_OK;
    if (!d_n || (addr(d_n) & 15)) return fail(CN_ERR_ARG, "cn_encode2_multi_device: input must be 16-byte aligned");
    cn::PeerOuts64 more{};
    for (int d = 0; d < nout; d++) {
        if (!d_outs[d] || (addr(d_outs[d]) & 31)) return fail(CN_ERR_ARG, "cn_encode2_multi_device: destination %d null or not 32-byte aligned", d);
        if (d > 0) more.p[d - 1] = static_cast<uint64_t *>(d_outs[d]);
    }
    const size_t total = len / 27 + ((len % 27) ? 1 : 0);
    const size_t ntiles = (len / 27) / cn::kB5WarpWords;
    const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
    if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "cn_encode2_multi_device: input too large for one launch");
    cn::b5_encode_kernel<true, cn::kEncPlain, true><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(
        static_cast<const uint8_t *>(d_n), static_cast<uint64_t *>(d_outs[0]), len, ntiles, total, nullptr, more, nout - 1);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "multi-destination base-5 encode launch failed: %s", cudaGetErrorString(e));
    return CN_OK;
}

// ---- CUDA IPC: let another process (one process per GPU) map a buffer of this one -----------------------
// cudaIpcGetMemHandle describes the whole ALLOCATION that contains a pointer; a tensor handed out by a caching
// allocator may sit at an offset inside it, so the offset is exported alongside (cuMemGetAddressRange, resolved at
// run time so the library has no link-time dependency on libcuda).
int ipc_export(void *d_ptr, void *handle64, size_t *offset)
{
    if (!d_ptr || !handle64 || !offset) return fail(CN_ERR_ARG, "cn_ipc_export: null pointer");
    typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CN_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    if (!fn) return fail(CN_ERR_CUDA, "cn_ipc_export: cuMemGetAddressRange unavailable");
    unsigned long long base = 0;
    size_t size = 0;
    int rc = reinterpret_cast<GetRange>(fn)(&base, &size, (unsigned long long)addr(d_ptr));
    if (rc != 0) return fail(CN_ERR_CUDA, "cn_ipc_export: cuMemGetAddressRange failed (%d)", rc);
    cudaIpcMemHandle_t h;
    CN_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(base)));
    static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, sizeof h);
    *offset = (size_t)(addr(d_ptr) - base);
    return CN_OK;
}

int ipc_open(const void *handle64, size_t offset, void **d_ptr)
{
    if (!handle64 || !d_ptr) return fail(CN_ERR_ARG, "cn_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    void *base = nullptr;
    CN_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = static_cast<uint8_t *>(base) + offset;
    return CN_OK;
}

int ipc_close(void *d_ptr, size_t offset)
{
    if (!d_ptr) return CN_OK;
    CN_CUDA(cudaIpcCloseMemHandle(static_cast<uint8_t *>(d_ptr) - offset));
    return CN_OK;
}

int decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, cudaStream_t s)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!d_bits || !d_out) return fail(CN_ERR_ARG, "cn_decode_device: null pointer");
    if (addr(d_bits) & 7) return fail(CN_ERR_ARG, "cn_decode_device: packed input must be 8-byte aligned");
    const Tuning t = load_tuning(CN_DIR_DECODE);
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

template <int MODE>
cudaError_t encode2_launch(const uint8_t *in, size_t len, uint64_t *out, unsigned long long *counter, cudaStream_t s)
{
    const size_t total = words2_for_len(len);
    if ((addr(in) & 15) == 0 && (addr(out) & 31) == 0) {                      // 128-bit ASCII loads, 256-bit packed stores
        const size_t ntiles = (len / 27) / cn::kB5WarpWords;                 // tiles of complete words
        const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
        if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
        if (g_b5_tma & 1) cn::b5_encode_kernel<true, MODE><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(in, out, len, ntiles, total, counter);
        else cn::b5_encode_kernel<false, MODE><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(in, out, len, ntiles, total, counter);
    } else if ((addr(out) & 31) == 0) {
        // ASCII side misaligned: shifted staging.  Only tiles whose trailing aligned vector lies inside the buffer.
        const unsigned mis = (unsigned)(addr(in) & 15);
        size_t ntiles = (len / 27) / cn::kB5WarpWords;
        const size_t fit = len + mis >= 16 ? (len + mis - 16) / cn::kB5WarpBytes : 0;
        if (fit < ntiles) ntiles = fit;
        const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
        if (blocks > kMaxGrid) return cudaErrorInvalidConfiguration;
        cn::b5_encode_mis_kernel<MODE><<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(in, mis, out, len, ntiles, total, counter);
    } else {
        size_t blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        cn::b5_encode_scalar_kernel<MODE><<<(unsigned)blocks, 256, 0, s>>>(in, out, len, total, counter);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// mode / counter as for encode_device; the alphabet here is {A,C,G,T,U,N} in either case
int encode2_device(const void *d_n, size_t len, void *d_out, int mode, unsigned long long *counter, cudaStream_t s)
{
    if (mode != cn::kEncPlain && mode != cn::kEncCount && mode != cn::kEncLutExact) return fail(CN_ERR_ARG, "encode2: bad mode %d", mode);
    if (mode == cn::kEncCount && !counter) return fail(CN_ERR_ARG, "cn_encode2_checked_device: counter must be a non-null 8-byte aligned device pointer");
    if (counter && (addr(counter) & 7)) return fail(CN_ERR_ARG, "encode2: the invalid-byte counter must be 8-byte aligned");
    if (len == 0) return CN_OK;
    if (!d_n || !d_out) return fail(CN_ERR_ARG, "cn_encode2_device: null pointer");
    if (addr(d_out) & 7) return fail(CN_ERR_ARG, "cn_encode2_device: output must be 8-byte aligned");
    const uint8_t *in = static_cast<const uint8_t *>(d_n);
    uint64_t *out = static_cast<uint64_t *>(d_out);
    cudaError_t e = mode == cn::kEncCount      ? encode2_launch<cn::kEncCount>(in, len, out, counter, s)
                    : mode == cn::kEncLutExact ? encode2_launch<cn::kEncLutExact>(in, len, out, counter, s)
                                               : encode2_launch<cn::kEncPlain>(in, len, out, nullptr, s);
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
    } else if ((addr(d_bits) & 31) == 0) {
        const unsigned mis = (unsigned)(addr(d_out) & 15);      // ASCII side misaligned: shifted tile, byte stores at the edges
        const size_t ntiles = (len / 27) / cn::kB5WarpWords;
        const size_t blocks = (ntiles + 1 + cn::kB5Warps - 1) / cn::kB5Warps;
        if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "cn_decode2_device: input too large for one launch");
        cn::b5_decode_mis_kernel<<<(unsigned)blocks, cn::kB5Warps * 32, 0, s>>>(bits, out, mis, len, ntiles, total);
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

// ---- many sequences resident in HBM, tightly concatenated (segmented_kernels.cuh) ----------------------------------
int segmented_device(bool encode, const void *d_src, const void *d_offsets, const void *d_word_offsets, size_t count,
                     size_t total_len, size_t total_words, void *d_dst, cudaStream_t s)
{
    const char *who = encode ? "cn_encode_segmented_device" : "cn_decode_segmented_device";
    if (count == 0 || total_words == 0) return CN_OK;
    if (!d_src || !d_dst || !d_offsets || !d_word_offsets) return fail(CN_ERR_ARG, "%s: null pointer", who);
    if ((addr(d_offsets) | addr(d_word_offsets)) & 7) return fail(CN_ERR_ARG, "%s: offset arrays must be 8-byte aligned", who);
    if (addr(encode ? d_dst : d_src) & 7) return fail(CN_ERR_ARG, "%s: packed words must be 8-byte aligned", who);
    if (total_len > total_words * 32 || total_words > total_len + count) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    const size_t blocks = (total_words + cn::kSegWords - 1) / cn::kSegWords;
    if (blocks > kMaxGrid) return fail(CN_ERR_ARG, "%s: too many words for one launch", who);
    const uint64_t *off = static_cast<const uint64_t *>(d_offsets), *woff = static_cast<const uint64_t *>(d_word_offsets);
    if (encode)
        cn::encode_segmented_kernel<<<(unsigned)blocks, cn::kSegThreads, 0, s>>>(static_cast<const uint8_t *>(d_src), off, woff, count,
                                                                                total_words, total_len, static_cast<uint64_t *>(d_dst));
    else
        cn::decode_segmented_kernel<<<(unsigned)blocks, cn::kSegThreads, 0, s>>>(static_cast<const uint64_t *>(d_src), off, woff, count,
                                                                                total_words, static_cast<uint8_t *>(d_dst));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CN_ERR_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
    return CN_OK;
}

// What the host pipeline needs to know about a codec: nucleotides per word and the device entry points.
struct Codec {
    unsigned group;                                                        // 32 (2-bit) or 27 (base-5)
    int (*enc)(const void *, size_t, void *, int mode, unsigned long long *counter, cudaStream_t);
    int (*dec)(const void *, size_t, size_t, void *, cudaStream_t);
    size_t words(size_t nt) const { return nt / group + ((nt % group) ? 1 : 0); }
};
const Codec kCodec2bit{32, encode_device, decode_device};
const Codec kCodecBase5{27, encode2_device, decode2_device};

}  // namespace
