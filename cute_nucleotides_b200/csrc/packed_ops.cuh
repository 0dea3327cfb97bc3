// packed_ops.cuh -- operations on the packed 2-bit words themselves: Hamming distance, complement,
// reverse complement.  The reference motivates them (README.md:21-23, 45, 415-418: "since the nucleotides are packed,
// operations like Hamming distance can be done very quickly") but ships no code for them, so PARITY IS UNPINNED: the
// semantics are defined through the reference's own codec instead --
//     hamming(a, b, len)        = #{ i < len : bits_to_n_lut(a)[i] != bits_to_n_lut(b)[i] }
//     complement(a, len)        = n_to_bits_lut(map A<->T, C<->G over bits_to_n_lut(a, len))
//     reverse_complement(a,len) = n_to_bits_lut(reverse of the above)
// and the test-side checker computes them exactly that way, through ASCII.
//
// With codes A=00 C=01 T=10 G=11 (src/n_to_bits.rs:8-21) the complement is code ^ 0b10, i.e. XOR with 0xAAAA...;
// a mismatch is a non-zero 2-bit field of a ^ b, i.e. popc(((x | x >> 1) & 0x5555...)).  All three are pure HBM
// streams: hamming reads 2 x 0.25 B/nt and writes nothing, the other two read 0.25 and write 0.25 B/nt.
#pragma once
#include "host_pipeline.cuh"

namespace cn {

constexpr uint64_t kLowBits = 0x5555555555555555ull;     // low bit of every 2-bit field
constexpr uint64_t kCompMask = 0xAAAAAAAAAAAAAAAAull;    // high bit of every field: A<->T, C<->G

// mask of the fields of word w that hold real nucleotides (all of them except in the last, ragged word)
__device__ __forceinline__ uint64_t valid_mask(size_t w, size_t len)
{
    const size_t first = w << 5;
    if (first + 32 <= len) return ~0ull;
    if (first >= len) return 0ull;
    return (1ull << (2 * (len - first))) - 1ull;
}

struct alignas(16) u64x2 { uint64_t a, b; };
struct alignas(32) u64x4 { uint64_t v[4]; };
__device__ __forceinline__ u64x4 ld_stream_u64x4(const void *p)
{
    u64x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u64x4(void *p, const u64x4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.v[0]), "l"(v.v[1]), "l"(v.v[2]), "l"(v.v[3]) : "memory");
}

__device__ __forceinline__ uint32_t mismatches(uint64_t a, uint64_t b)
{
    const uint64_t x = a ^ b;
    return (uint32_t)__popcll((x | (x >> 1)) & kLowBits);
}

// ---- Hamming distance: grid-stride over groups of 4 words (256-bit loads), one atomicAdd per warp ---------------
// `a`, `b`: 32-byte aligned when nvec > 0 (the host side falls back to nvec = 0 otherwise); words [4*nvec, nwords) and
// the ragged last word are finished by the grid's first warp.
template <int UNROLL>
__global__ void __launch_bounds__(256)
hamming_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, size_t nvec, size_t nwords, size_t len,
               unsigned long long *__restrict__ result)
{
    uint32_t count = 0;                                   // < 2^32: a thread sees at most UNROLL * 128 nt per iteration ...
    unsigned long long total = 0;                         // ... and is flushed into 64 bits every iteration
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride * UNROLL) {
        u64x4 x[UNROLL], y[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t j = i + (size_t)u * stride;
            if (j < nvec) { x[u] = ld_stream_u64x4(a + 4 * j); y[u] = ld_stream_u64x4(b + 4 * j); }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t j = i + (size_t)u * stride;
            if (j < nvec) {
#pragma unroll
                for (int k = 0; k < 4; k++) count += mismatches(x[u].v[k], y[u].v[k]);
            }
        }
        total += count;
        count = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        for (size_t w = 4 * nvec + threadIdx.x; w < nwords; w += 32) {
            const uint64_t m = valid_mask(w, len);
            total += mismatches(a[w] & m, b[w] & m);
        }
    }
    // warp reduce (64-bit through two 32-bit shuffles), one atomic per warp that saw a mismatch
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_down_sync(0xFFFFFFFFu, total, off);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(result, total);
}

// ---- complement: out = (in ^ 0xAAAA...) & valid -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
complement_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t nvec, size_t nwords, size_t len)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nvec) {
        u64x4 x = ld_stream_u64x4(in + 4 * i);
#pragma unroll
        for (int k = 0; k < 4; k++) x.v[k] = (x.v[k] ^ kCompMask) & valid_mask(4 * i + k, len);
        st_stream_u64x4(out + 4 * i, x);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= 256 - 32) {
        for (size_t w = 4 * nvec + (threadIdx.x & 31); w < nwords; w += 32) out[w] = (in[w] ^ kCompMask) & valid_mask(w, len);
    }
}

// reverse the order of the 32 two-bit fields of a word: reverse all 64 bits, then swap the two bits of every field back
__device__ __forceinline__ uint64_t reverse_fields(uint64_t x)
{
    const uint64_t y = __brevll(x);
    return ((y >> 1) & kLowBits) | ((y & kLowBits) << 1);
}

// ---- reverse complement.  With W = ceil(len/32) and s = 32 W - len pad fields in the last input word, the field-reversed
// array FR[k] = reverse_fields(in[W-1-k]) holds the answer shifted up by s fields, so
//     out[k] = (FR[k] >> 2s | FR[k+1] << (64 - 2s)) ^ 0xAAAA..., masked to the valid fields.
// One thread per output word; its second input word is its neighbour's first (L1/L2 hit, no extra HBM traffic). -------
__global__ void __launch_bounds__(256)
reverse_complement_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t nwords, size_t len)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords) return;
    const unsigned shift = (unsigned)(2 * ((nwords << 5) - len));           // 0..62 bits
    uint64_t first = __ldg(in + (nwords - 1 - k));
    if (k == 0) first &= valid_mask(nwords - 1, len);                       // pad fields of the ragged input word do not count
    const uint64_t lo = reverse_fields(first);
    uint64_t v = lo;
    if (shift) {
        const uint64_t hi = k + 1 < nwords ? reverse_fields(__ldg(in + (nwords - 2 - k))) : 0ull;
        v = (lo >> shift) | (hi << (64 - shift));
    }
    out[k] = (v ^ kCompMask) & valid_mask(k, len);
}

// The same with 256-bit accesses: thread t produces output words 4t..4t+3 from the 5 input words W-1-4t .. W-5-4t, which
// always lie in exactly two ALIGNED input vectors (A = the one holding word W-1-4t, B = the one below it); which lanes of
// A and B are needed depends only on (W-1) % 4, so the selection is uniform over the grid.  A warp reads 1 KiB + one
// neighbouring vector (B of lane l is A of lane l+1: an L1 hit) and writes 1 KiB.  Threads that would touch the ragged top
// word, the bottom of the array, or memory past `nwords` fall back to the word-at-a-time formulation.  `in` and `out`
// 32-byte aligned.
__global__ void __launch_bounds__(256)
reverse_complement_vec_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t nwords, size_t len)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t k0 = 4 * t;
    if (k0 >= nwords) return;
    const unsigned shift = (unsigned)(2 * ((nwords << 5) - len));           // 0..62 bits
    const size_t base = nwords - 1 - k0;                                    // highest input word this thread needs
    const size_t vb = base >> 2;
    if (t > 0 && k0 + 4 <= nwords && vb >= 1 && 4 * vb + 3 < nwords) {
        const u64x4 A = ld_stream_u64x4(in + 4 * vb), B = ld_stream_u64x4(in + 4 * (vb - 1));
        uint64_t c[5];
        switch (base & 3) {
        case 3:  c[0] = A.v[3]; c[1] = A.v[2]; c[2] = A.v[1]; c[3] = A.v[0]; c[4] = B.v[3]; break;
        case 2:  c[0] = A.v[2]; c[1] = A.v[1]; c[2] = A.v[0]; c[3] = B.v[3]; c[4] = B.v[2]; break;
        case 1:  c[0] = A.v[1]; c[1] = A.v[0]; c[2] = B.v[3]; c[3] = B.v[2]; c[4] = B.v[1]; break;
        default: c[0] = A.v[0]; c[1] = B.v[3]; c[2] = B.v[2]; c[3] = B.v[1]; c[4] = B.v[0]; break;
        }
        if (base < 4) c[4] = 0;                                             // word W-5-4t does not exist: FR beyond the array is 0
        uint64_t fr[5];
#pragma unroll
        for (int i = 0; i < 5; i++) fr[i] = reverse_fields(c[i]);
        u64x4 o;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = fr[i];
            if (shift) v = (fr[i] >> shift) | (fr[i + 1] << (64 - shift));
            o.v[i] = (v ^ kCompMask) & valid_mask(k0 + i, len);
        }
        st_stream_u64x4(out + k0, o);
        return;
    }
    for (size_t k = k0; k < k0 + 4 && k < nwords; k++) {
        uint64_t first = in[nwords - 1 - k];
        if (k == 0) first &= valid_mask(nwords - 1, len);
        const uint64_t lo = reverse_fields(first);
        uint64_t v = lo;
        if (shift) {
            const uint64_t hi = k + 1 < nwords ? reverse_fields(in[nwords - 2 - k]) : 0ull;
            v = (lo >> shift) | (hi << (64 - shift));
        }
        out[k] = (v ^ kCompMask) & valid_mask(k, len);
    }
}

}  // namespace cn

namespace {

// *d_result (device, 8-byte aligned uint64) is INCREMENTED by the number of mismatching nucleotides among the first len
int hamming_device(const void *d_a, const void *d_b, size_t nwords, size_t len, void *d_result, cudaStream_t s)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (!d_result || (addr(d_result) & 7)) return fail(CN_ERR_ARG, "cn_hamming_device: result must be a non-null 8-byte aligned device pointer");
    if (len == 0) return CN_OK;
    if (!d_a || !d_b || ((addr(d_a) | addr(d_b)) & 7)) return fail(CN_ERR_ARG, "cn_hamming_device: null or misaligned pointer");
    const size_t used = cn_words_for_len(len);                       // words beyond ceil(len/32) are never read
    const size_t nvec = ((addr(d_a) | addr(d_b)) & 31) == 0 ? (len >> 5) / 4 : 0;   // whole words only; ragged word -> tail warp
    static const size_t per_sm = std::getenv("CN_HAMMING_CTAS") ? (size_t)std::atoi(std::getenv("CN_HAMMING_CTAS")) : 4;      // A/B knob
    size_t blocks = (nvec + 255) / 256;
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    if (blocks == 0) blocks = 1;
    static const int unroll = std::getenv("CN_HAMMING_UNROLL") ? std::atoi(std::getenv("CN_HAMMING_UNROLL")) : 4;   // A/B knob
    const uint64_t *a = static_cast<const uint64_t *>(d_a), *b = static_cast<const uint64_t *>(d_b);
    unsigned long long *res = static_cast<unsigned long long *>(d_result);
    // measured at 2 x 2.5 GiB (tools/bench_packed.py): unroll 1 / 2 / 4 / 8 at 8 CTAs per SM -> 7109 / 7142 / 7365 / 7308 GB/s;
    // unroll 4 at 2 / 4 / 8 CTAs per SM -> 7341 / 7385 / 7365 (a compute-free pure read tops out at 7547)
    if (unroll == 1) cn::hamming_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(a, b, nvec, used, len, res);
    else if (unroll == 4) cn::hamming_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(a, b, nvec, used, len, res);
    else if (unroll == 8) cn::hamming_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(a, b, nvec, used, len, res);
    else cn::hamming_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(a, b, nvec, used, len, res);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CN_CUDA(cudaGetLastError());
    return CN_OK;
}

int complement_device(const void *d_in, size_t nwords, size_t len, void *d_out, bool reverse, cudaStream_t s)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!d_in || !d_out || ((addr(d_in) | addr(d_out)) & 7)) return fail(CN_ERR_ARG, "cn_complement_device: null or misaligned pointer");
    const size_t used = cn_words_for_len(len);
    const uint64_t *in = static_cast<const uint64_t *>(d_in);
    uint64_t *out = static_cast<uint64_t *>(d_out);
    if (reverse) {
        if (d_in == d_out) return fail(CN_ERR_ARG, "cn_reverse_complement_device: cannot run in place");
        if (((addr(d_in) | addr(d_out)) & 31) == 0) {
            const size_t blocks = ((used + 3) / 4 + 255) / 256;
            if (blocks > 0x7FFFFFFFull) return fail(CN_ERR_ARG, "cn_reverse_complement_device: input too large for one launch");
            cn::reverse_complement_vec_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, used, len);
        } else {
            const size_t blocks = (used + 255) / 256;
            if (blocks > 0x7FFFFFFFull) return fail(CN_ERR_ARG, "cn_reverse_complement_device: input too large for one launch");
            cn::reverse_complement_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, used, len);
        }
    } else {
        const size_t nvec = ((addr(d_in) | addr(d_out)) & 31) == 0 ? used / 4 : 0;
        size_t blocks = (nvec + 255) / 256;
        if (blocks == 0) blocks = 1;
        if (blocks > 0x7FFFFFFFull) return fail(CN_ERR_ARG, "cn_complement_device: input too large for one launch");
        cn::complement_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, nvec, used, len);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CN_CUDA(cudaGetLastError());
    return CN_OK;
}

// Host-slice flavours: the packed words cross PCIe once each way; one synchronous pass through temporary device buffers
// (these are conveniences for host-resident data -- the ops pay off on words that are already in HBM).
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(CN_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes); }
        return CN_OK;
    }
};

int hamming_host(const uint64_t *a, const uint64_t *b, size_t nwords, size_t len, uint64_t *result)
{
    if (!result) return fail(CN_ERR_ARG, "cn_hamming_host: null result");
    *result = 0;
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!a || !b) return fail(CN_ERR_ARG, "cn_hamming_host: null pointer");
    if (t_device >= 0) CN_CUDA(cudaSetDevice(t_device));
    const size_t bytes = cn_words_for_len(len) * 8;
    DevBuf da, db, dr;
    int rc;
    if ((rc = da.alloc(bytes)) != CN_OK || (rc = db.alloc(bytes)) != CN_OK || (rc = dr.alloc(8)) != CN_OK) return rc;
    CN_CUDA(cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
    CN_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
    CN_CUDA(cudaMemsetAsync(dr.p, 0, 8, cudaStreamPerThread));
    if ((rc = hamming_device(da.p, db.p, bytes / 8, len, dr.p, cudaStreamPerThread)) != CN_OK) return rc;
    CN_CUDA(cudaMemcpyAsync(result, dr.p, 8, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    CN_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return CN_OK;
}

int complement_host(const uint64_t *in, size_t nwords, size_t len, uint64_t *out, bool reverse)
{
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    if (len == 0) return CN_OK;
    if (!in || !out) return fail(CN_ERR_ARG, "cn_complement_host: null pointer");
    if (t_device >= 0) CN_CUDA(cudaSetDevice(t_device));
    const size_t bytes = cn_words_for_len(len) * 8;
    DevBuf di, dout;
    int rc;
    if ((rc = di.alloc(bytes)) != CN_OK || (rc = dout.alloc(bytes)) != CN_OK) return rc;
    CN_CUDA(cudaMemcpyAsync(di.p, in, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
    if ((rc = complement_device(di.p, bytes / 8, len, dout.p, reverse, cudaStreamPerThread)) != CN_OK) return rc;
    CN_CUDA(cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    CN_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return CN_OK;
}

}  // namespace
