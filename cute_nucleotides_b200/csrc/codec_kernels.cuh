// codec_kernels.cuh -- sm_100a kernels for the 2-bit nucleotide codec.
//
// What is computed (reference contract, /root/reference/src/n_to_bits.rs:34-47 and :51-69):
//   encode: nucleotide i -> 2-bit code at bits 2(i&31).. of u64 word i>>5, A/a=0 C/c=1 T/t/U/u=2 G/g=3,
//           unused high bits of the last word zero.  On a little-endian device the packed stream is
//           byte-granular: output byte j = codes of nucleotides 4j..4j+3, so a thread that reads 16
//           consecutive ASCII bytes owns exactly one aligned u32 of output and no cross-thread
//           exchange is needed.
//   decode: code -> 'A','C','T','G' (BITS_LUT, :23-30), always upper case.
//
// How (B200-first, not a translation of the AVX2 code):
//   * the code of an ASCII nucleotide is bits 1..2 of the byte; per 32-bit register holding 4 bytes
//     ((x & 0x06060606) * 0x00820820) >> 24 packs the 4 codes into one byte -- the multiply constant of
//     n_to_bits_mul (:223-231) applied to one GPU register (1 LOP3 + 1 IMAD per 4 nt), and PRMT gathers
//     the four top bytes (the GPU's pshufb).
//   * decode spreads 16 codes into 16 PRMT selector nibbles (PRMT + 2x SHF/LOP3 per 8 nt) and uses
//     PRMT against the register-resident LUT word 'A','C','T','G' (same LUT word as lut_i32, :276).
//   * memory: every byte is touched once, so loads are ld.global.nc.L1::no_allocate (128- or 256-bit,
//     sm_100 has LDG.E.256) and stores st.global.L1::no_allocate; each thread issues all of its loads
//     before the first use so >= UNROLL x 16/32 B per thread are in flight.
//
// The kernels are HBM-bound (1.25 B of traffic per nucleotide in either direction, <1.5 integer
// ops per nucleotide); tensor cores do not apply.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace cn {

// ------------------------------------------------------------------------------------------------
// register-level codec
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kAsciiCodeMask = 0x06060606u;                                   // bits 1..2 of each byte
constexpr uint32_t kPackMul = (1u << 23) | (1u << 17) | (1u << 11) | (1u << 5);    // 0x00820820
constexpr uint32_t kLutWord = 0x47544341u;                                         // 'A','C','T','G'

// PRMT with a run-time selector.  __byte_perm() ANDs a variable selector with 0x7777 (an extra LOP3 per use);
// every selector built in this file already has nibbles <= 7, so the raw instruction is used instead.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// (a & mask) | (b & ~mask) in ONE LOP3 (immLut 0xE4 = c ? a : b); nvcc splits the two-constant form into two.
__device__ __forceinline__ uint32_t bitselect(uint32_t a, uint32_t b, uint32_t mask)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}

// 4 ASCII bytes -> their 4 codes in the TOP byte of the result (lower bytes are garbage)
__device__ __forceinline__ uint32_t pack4_top(uint32_t x) { return (x & kAsciiCodeMask) * kPackMul; }

// 16 ASCII bytes -> 16 codes (one u32 of the packed stream)
__device__ __forceinline__ uint32_t pack16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3)
{
    uint32_t lo = __byte_perm(pack4_top(x0), pack4_top(x1), 0x0073);   // [x0.top, x1.top, -, -]
    uint32_t hi = __byte_perm(pack4_top(x2), pack4_top(x3), 0x0073);
    return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ uint32_t pack16(const uint4 &v) { return pack16(v.x, v.y, v.z, v.w); }

// 8 codes (16 bits, in the low or high half of w) -> two PRMT selectors -> 8 ASCII bytes
template <bool kHigh>
__device__ __forceinline__ void unpack8(uint32_t w, uint32_t &a, uint32_t &b)
{
    uint32_t s = __byte_perm(w, 0u, kHigh ? 0x4342 : 0x4140);   // [b, 0, b', 0]
    s = (s | (s << 4)) & 0x0F0F0F0Fu;
    s = (s | (s << 2)) & 0x33333333u;                           // nibble k = code k
    a = prmt(kLutWord, 0u, s);                                  // PRMT reads the low 16 selector bits
    b = prmt(kLutWord, 0u, s >> 16);
}

// 16 codes (one u32 of the packed stream) -> 16 ASCII bytes
__device__ __forceinline__ uint4 unpack16(uint32_t w)
{
    uint4 r;
    unpack8<false>(w, r.x, r.y);
    unpack8<true>(w, r.z, r.w);
    return r;
}

// ------------------------------------------------------------------------------------------------
// validation fused into the encode pass (optional): which bytes are outside {A,C,G,T,U,a,c,g,t,u}?
// A byte is valid iff b & 0xD8 (everything but the case bit and the low 3 bits) equals what its low 3 bits
// predict: a/c/g (index 1,3,7) -> 0x40, t/u (index 4,5) -> 0x50; indices 0,2,6 can never match.  Per 32-bit register: pack the four 3-bit indices
// into PRMT selector nibbles (the shift runs on the FMA pipe as a multiply-high), one PRMT to pack, one PRMT
// lookup, one LOP3 for the masked byte, one LOP3 to XOR and OR-accumulate: 5 ALU-pipe ops per 4 nucleotides.
// Bit 3 / bit 7 of a byte leak into selector nibbles as PRMT's sign-replicate flag; that can only turn a
// match into a mismatch next to a byte that is itself invalid, so the accumulated word is an exact
// "anything invalid?" flag, and a thread that sees it set recounts its bytes one by one.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t invalid_accumulate(uint32_t acc, uint32_t x)
{
    uint32_t t = bitselect(x, __umulhi(x, 0x10000000u), 0x8F8F8F8Fu);     // byte0 = n0 | i1<<4, byte2 = n2 | i3<<4
    uint32_t expect = prmt(0x40FF40FFu, 0x40FF5050u, __byte_perm(t, 0u, 0x4420));
    return acc | ((x & 0xD8D8D8D8u) ^ expect);                            // non-zero <=> some byte is invalid
}
__device__ __forceinline__ bool invalid_byte(uint32_t b)
{
    uint32_t c = b | 0x20u;
    return !(c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'u');
}
__device__ __forceinline__ uint32_t count_invalid4(uint32_t x)
{
    return (uint32_t)invalid_byte(x & 0xFFu) + (uint32_t)invalid_byte((x >> 8) & 0xFFu) +
           (uint32_t)invalid_byte((x >> 16) & 0xFFu) + (uint32_t)invalid_byte(x >> 24);
}
// one atomic per warp, and only when something was invalid
__device__ __forceinline__ void report_invalid(unsigned long long *counter, uint32_t mine)
{
    uint32_t total = __reduce_add_sync(0xFFFFFFFFu, mine);
    if (total && (threadIdx.x & 31) == 0) atomicAdd(counter, (unsigned long long)total);
}

// Exact per-byte version (slow path of the _lut-exact mode): 0x80 in every byte of x that is outside the alphabet.
// The selector nibbles are masked to 3 bits, so no byte can influence its neighbour's lookup.
__device__ __forceinline__ uint32_t invalid_bytes_mask(uint32_t x)
{
    uint32_t t = (x & 0x07070707u) | ((x >> 4) & 0x70707070u);            // byte0 = i0 | i1<<4, byte2 = i2 | i3<<4
    uint32_t expect = prmt(0x40FF40FFu, 0x40FF5050u, __byte_perm(t, 0u, 0x4420));
    uint32_t diff = (x & 0xD8D8D8D8u) ^ expect;
    return (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;  // SWAR "byte != 0"
}
// n_to_bits_lut semantics (BYTE_LUT, src/n_to_bits.rs:8-21, :42): every byte outside the alphabet encodes as 0.
// Clears bits 1..2 (the code bits) of the offending bytes so the ordinary pack then yields code 0 for them.
__device__ __forceinline__ uint32_t lut_exact_clean(uint32_t x, uint32_t &n_invalid)
{
    const uint32_t bad = invalid_bytes_mask(x);
    n_invalid += __popc(bad);
    return x & ~((bad >> 5) | (bad >> 6));
}

// what the encode kernel does besides packing
enum : int {
    kEncPlain = 0,       // codes = (byte >> 1) & 3 for every byte, like the reference's SIMD encoders
    kEncCount = 1,       // + count the bytes outside {A,C,G,T,U,a,c,g,t,u}
    kEncLutExact = 2,    // bytes outside the alphabet encode as 0, exactly like n_to_bits_lut; counted too when a counter is given
};

// ------------------------------------------------------------------------------------------------
// streaming loads / stores
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream16(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
struct alignas(32) u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ld_stream32(const void *p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]),
                   "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream4(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(void *p, uint32_t v)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream8(void *p, uint32_t a, uint32_t b)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_stream16(void *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream32(void *p, const uint4 &a, const uint4 &b)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// ragged ends.  One warp finishes what the vector body does not cover; lane l owns output u32
// (encode) / output 16-byte group (decode) number first+l.  Byte accesses only, any alignment.
// ------------------------------------------------------------------------------------------------

// encode u32 units [first, total): unit u covers nucleotides 16u..16u+15, zero-padded beyond len
// (the zero padding is the "unused high bits of the last word are zero" rule, n_to_bits.rs:35).
template <bool LUT_EXACT = false>
__device__ __forceinline__ uint32_t encode_edge(const uint8_t *__restrict__ n, size_t len,
                                                uint32_t *__restrict__ out32, size_t first, size_t total, unsigned lane)
{
    uint32_t invalid = 0;
    for (size_t u = first + lane; u < total; u += 32) {
        uint32_t code = 0;
        size_t base = u << 4;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            size_t i = base + k;
            uint32_t byte = (i < len) ? (uint32_t)n[i] : 0u;      // 0 -> code 0
            const bool bad = i < len && invalid_byte(byte);
            if (bad) invalid++;
            if (LUT_EXACT && bad) byte = 0u;                      // BYTE_LUT: everything unlisted -> 0
            code |= ((byte >> 1) & 3u) << (2 * k);
        }
        out32[u] = code;
    }
    return invalid;                                               // bytes outside the alphabet among those this lane read
}

// decode nucleotides [from, len) with byte stores; bits32 is the packed stream viewed as u32
__device__ __forceinline__ void decode_edge(const uint32_t *__restrict__ bits32, uint8_t *__restrict__ out,
                                            size_t from, size_t len, unsigned lane)
{
    for (size_t i = from + lane; i < len; i += 32) {
        uint32_t w = bits32[i >> 4];
        out[i] = (uint8_t)(kLutWord >> (8 * ((w >> (2 * (i & 15))) & 3u)));
    }
}

// ------------------------------------------------------------------------------------------------
// encode body.  VEC = ASCII bytes per thread per load (16 -> LDG.128 + STG.32, 32 -> LDG.256 + STG.64).
// A CTA owns THREADS*UNROLL consecutive vectors; within one unrolled step a warp reads 32*VEC
// contiguous bytes and writes 32*VEC/4 contiguous bytes.
//   in      : ASCII; when MISALIGN the pointer has been rounded DOWN to 16 B and `shift_bytes`
//             (1..15) says where nucleotide 0 sits inside the first vector
//   nvec    : number of full VEC-byte groups the body covers
//   edge_*  : the warp-sized ragged end, run by the last warp of the last CTA
// ------------------------------------------------------------------------------------------------
template <int VEC, int UNROLL, int THREADS, bool MISALIGN, int MODE = kEncPlain>
__global__ void __launch_bounds__(THREADS)
encode_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out32, size_t nvec,
              const uint8_t *__restrict__ n0, size_t len, size_t edge_first, size_t edge_total, unsigned shift_bytes,
              unsigned long long *__restrict__ invalid_counter)
{
    static_assert(VEC == 16 || VEC == 32, "VEC");
    constexpr size_t kTile = (size_t)THREADS * UNROLL;
    const size_t tile0 = (size_t)blockIdx.x * kTile;
    const size_t t = tile0 + threadIdx.x;
    uint32_t n_invalid = 0;                // bytes outside the alphabet seen by this thread (MODE != plain, and the edge warp)
    // kEncCount: validate the registers that were just encoded; the exact count is only taken on a mismatch
    auto check16 = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        if constexpr (MODE == kEncCount) {
            uint32_t bad = invalid_accumulate(invalid_accumulate(invalid_accumulate(invalid_accumulate(0u, a), b), c), d);
            if (bad) n_invalid += count_invalid4(a) + count_invalid4(b) + count_invalid4(c) + count_invalid4(d);
        }
    };
    // 16 ASCII bytes -> one packed u32.  kEncLutExact: the cheap "anything invalid?" test runs on every register; only a
    // thread that sees a hit takes the exact per-byte path that zeroes the offending codes (valid data pays nothing more
    // than kEncCount does).
    auto enc16 = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d) -> uint32_t {
        if constexpr (MODE == kEncLutExact) {
            uint32_t bad = invalid_accumulate(invalid_accumulate(invalid_accumulate(invalid_accumulate(0u, a), b), c), d);
            if (bad) {
                a = lut_exact_clean(a, n_invalid); b = lut_exact_clean(b, n_invalid);
                c = lut_exact_clean(c, n_invalid); d = lut_exact_clean(d, n_invalid);
            }
        }
        return pack16(a, b, c, d);
    };

    if constexpr (!MISALIGN) {
        if (tile0 + kTile <= nvec) {                      // full tile: no bounds checks
            if constexpr (VEC == 16) {
                uint4 v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) v[u] = ld_stream16(in + ((t + (size_t)u * THREADS) << 4));
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    st_stream4(out32 + t + (size_t)u * THREADS, enc16(v[u].x, v[u].y, v[u].z, v[u].w));
                    check16(v[u].x, v[u].y, v[u].z, v[u].w);
                }
            } else {
                u32x8 v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) v[u] = ld_stream32(in + ((t + (size_t)u * THREADS) << 5));
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    st_stream8(out32 + ((t + (size_t)u * THREADS) << 1),
                               enc16(v[u].v[0], v[u].v[1], v[u].v[2], v[u].v[3]),
                               enc16(v[u].v[4], v[u].v[5], v[u].v[6], v[u].v[7]));
                    check16(v[u].v[0], v[u].v[1], v[u].v[2], v[u].v[3]);
                    check16(v[u].v[4], v[u].v[5], v[u].v[6], v[u].v[7]);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                size_t i = t + (size_t)u * THREADS;
                if (i < nvec) {
                    if constexpr (VEC == 16) {
                        uint4 v = ld_stream16(in + (i << 4));
                        st_stream4(out32 + i, enc16(v.x, v.y, v.z, v.w));
                        check16(v.x, v.y, v.z, v.w);
                    } else {
                        u32x8 v = ld_stream32(in + (i << 5));
                        st_stream8(out32 + (i << 1), enc16(v.v[0], v.v[1], v.v[2], v.v[3]),
                                   enc16(v.v[4], v.v[5], v.v[6], v.v[7]));
                        check16(v.v[0], v.v[1], v.v[2], v.v[3]);
                        check16(v.v[4], v.v[5], v.v[6], v.v[7]);
                    }
                }
            }
        }
    } else {
        // input not 16-byte aligned: read the two aligned vectors that straddle each 16-nt group and
        // funnel-shift.  (VEC is 16 here.)  The second load of thread t is the first of thread t+1,
        // so it is served by L1/L2, not HBM: plain ld.global.nc.
        const unsigned word_off = shift_bytes >> 2, bit_off = (shift_bytes & 3u) * 8u;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t i = t + (size_t)u * THREADS;
            // group 0 would need the aligned vector that starts before the caller's buffer: the edge warp encodes it
            // from the caller's own bytes instead, so nothing outside [n0, n0 + len) is ever read
            if (i < nvec && i > 0) {
                const uint4 *p = reinterpret_cast<const uint4 *>(in) + i;
                uint4 a = __ldg(p), b = __ldg(p + 1);
                uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                uint32_t x[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    uint32_t s = w[k];
                    if (word_off == 1) s = w[k + 1];
                    if (word_off == 2) s = w[k + 2];
                    if (word_off == 3) s = w[k + 3];
                    x[k] = s;
                }
                const uint32_t y0 = __funnelshift_r(x[0], x[1], bit_off), y1 = __funnelshift_r(x[1], x[2], bit_off),
                               y2 = __funnelshift_r(x[2], x[3], bit_off), y3 = __funnelshift_r(x[3], x[4], bit_off);
                st_stream4(out32 + i, enc16(y0, y1, y2, y3));
                check16(y0, y1, y2, y3);
            }
        }
    }

    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= THREADS - 32) {
        constexpr bool kExact = MODE == kEncLutExact;
        if (MISALIGN && edge_first > 0) n_invalid += encode_edge<kExact>(n0, len, out32, 0, 1, threadIdx.x & 31);
        n_invalid += encode_edge<kExact>(n0, len, out32, edge_first, edge_total, threadIdx.x & 31);
    }
    if constexpr (MODE != kEncPlain) { if (invalid_counter) report_invalid(invalid_counter, n_invalid); }
}

// ------------------------------------------------------------------------------------------------
// encode + assemble in one kernel (multi-GPU): the shard is read from local HBM ONCE and every packed u64 is
// stored to up to kMaxPeers destinations -- the rank's own assembled buffer and the peer-mapped buffers of
// the other ranks (CUDA IPC mappings; the stores travel over NVLink / NVSwitch).  Replaces "encode, then
// all-gather": no intermediate packed shard, no separate collective, and the link transfer overlaps the HBM
// reads.  `in` must be 32-byte aligned; each destination pointer is already offset to this shard's first word.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct PeerOuts { uint32_t *p[kMaxPeers]; };

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
encode_multi_kernel(const uint8_t *__restrict__ in, PeerOuts outs, int nout, size_t nvec,
                    size_t len, size_t edge_first, size_t edge_total)
{
    const size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (i < nvec) {
        u32x8 v = ld_stream32(in + (i << 5));
        const uint32_t lo = pack16(v.v[0], v.v[1], v.v[2], v.v[3]), hi = pack16(v.v[4], v.v[5], v.v[6], v.v[7]);
#pragma unroll
        for (int d = 0; d < kMaxPeers; d++)
            if (d < nout) st_stream8(outs.p[d] + (i << 1), lo, hi);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= THREADS - 32) {
        for (int d = 0; d < nout; d++) encode_edge(in, len, outs.p[d], edge_first, edge_total, threadIdx.x & 31);
    }
}

// ------------------------------------------------------------------------------------------------
// decode body.  VEC = ASCII bytes produced per thread per store (16 -> LDG.32 + STG.128,
// 32 -> LDG.64 + STG.256).
//   bits32     : packed stream as u32 (16 codes each)
//   out        : ASCII destination of the body, VEC-aligned; it starts `head` nucleotides into the
//                sequence (head = 0 when the caller's buffer is already aligned)
//   code_shift : 2*head, the bit offset of the body inside the packed stream (0..30, MISALIGN only)
// ------------------------------------------------------------------------------------------------
template <int VEC, int UNROLL, int THREADS, bool MISALIGN>
__global__ void __launch_bounds__(THREADS)
decode_kernel(const uint32_t *__restrict__ bits32, uint8_t *__restrict__ out, size_t nvec, unsigned code_shift,
              uint8_t *__restrict__ out0, size_t head, size_t edge_from, size_t len)
{
    static_assert(VEC == 16 || VEC == 32, "VEC");
    constexpr size_t kTile = (size_t)THREADS * UNROLL;
    const size_t tile0 = (size_t)blockIdx.x * kTile;
    const size_t t = tile0 + threadIdx.x;

    if constexpr (!MISALIGN) {
        if (tile0 + kTile <= nvec) {
            if constexpr (VEC == 16) {
                uint32_t w[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) w[u] = ld_stream4(bits32 + t + (size_t)u * THREADS);
#pragma unroll
                for (int u = 0; u < UNROLL; u++) st_stream16(out + ((t + (size_t)u * THREADS) << 4), unpack16(w[u]));
            } else {
                uint2 w[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) w[u] = ld_stream8(bits32 + ((t + (size_t)u * THREADS) << 1));
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
                    st_stream32(out + ((t + (size_t)u * THREADS) << 5), unpack16(w[u].x), unpack16(w[u].y));
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                size_t i = t + (size_t)u * THREADS;
                if (i < nvec) {
                    if constexpr (VEC == 16) {
                        st_stream16(out + (i << 4), unpack16(ld_stream4(bits32 + i)));
                    } else {
                        uint2 w = ld_stream8(bits32 + (i << 1));
                        st_stream32(out + (i << 5), unpack16(w.x), unpack16(w.y));
                    }
                }
            }
        }
    } else {
        // destination not 16-byte aligned: the body starts `head` nucleotides in, so each 16-nt group
        // straddles two packed u32s; the caller guarantees bits32[i+1] exists for every i < nvec.
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t i = t + (size_t)u * THREADS;
            if (i < nvec) {
                uint32_t lo = __ldg(bits32 + i), hi = __ldg(bits32 + i + 1);
                st_stream16(out + (i << 4), unpack16(__funnelshift_r(lo, hi, code_shift)));
            }
        }
    }

    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= THREADS - 32) {
        const unsigned lane = threadIdx.x & 31;
        if (MISALIGN) decode_edge(bits32, out0, 0, head, lane);      // nucleotides before the aligned body
        decode_edge(bits32, out0, edge_from, len, lane);             // nucleotides after it
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic data, bit-identical to the generator the tests use on the host
// (counter-based splitmix64).  One hash feeds 8 nucleotides.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t letter(uint32_t lane_byte, int alphabet)
{
    // "ACGT" = 0x54474341 ; "ACGTUacgtu" looked up from two packed constants
    if (alphabet == 4) return (0x54474341u >> (8 * (lane_byte & 3u))) & 0xFFu;
    uint32_t k = (lane_byte * 10u) >> 8;                                  // 0..9
    const uint64_t lo = 0x6763615554474341ull;                            // A C G T U a c g
    return k < 8 ? (uint32_t)(lo >> (8 * k)) & 0xFFu : (k == 8 ? (uint32_t)'t' : (uint32_t)'u');
}

// out[i] for i in [0, len) = letter of global nucleotide offset+i; offset is a multiple of 8
__global__ void generate_kernel(uint8_t *__restrict__ out, size_t offset, size_t len, uint64_t seed, int alphabet)
{
    size_t groups = (len + 7) >> 3;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        uint64_t h = mix64(seed, (offset >> 3) + g);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lo |= letter((uint32_t)(h >> (8 * k)) & 0xFFu, alphabet) << (8 * k);
            hi |= letter((uint32_t)(h >> (8 * k + 32)) & 0xFFu, alphabet) << (8 * k);
        }
        size_t i = g << 3;
        if (i + 8 <= len && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) {
            *reinterpret_cast<uint2 *>(out + i) = make_uint2(lo, hi);
        } else {
            uint64_t both = ((uint64_t)hi << 32) | lo;
            for (int k = 0; k < 8 && i + k < len; k++) out[i + k] = (uint8_t)(both >> (8 * k));
        }
    }
}

__global__ void generate_words_kernel(uint64_t *__restrict__ out, size_t first_word, size_t nwords, uint64_t seed)
{
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x)
        out[w] = mix64(seed, first_word + w);
}

}  // namespace cn
