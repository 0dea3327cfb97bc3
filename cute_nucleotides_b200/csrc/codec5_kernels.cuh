// codec5_kernels.cuh -- sm_100a kernels for cute-nucleotides' base-5 codec (src/n_to_bits2.rs):
// {A, C, T/U, G, N} -> digits 0..4, three digits -> 7 bits (a + 5b + 25c), nine triplets = 27 nucleotides
// per u64 (bit 63 unused).  Contract = n_to_bits2_lut / bits_to_n2_lut (src/n_to_bits2.rs:37-107).
//
// Shape of the problem on a GPU: 27 ASCII bytes <-> 8 packed bytes (1.296 B of HBM traffic per
// nucleotide), and 27-byte groups are not aligned to anything -- but 32 groups are 864 bytes = 54 x 16 B.
// So a WARP owns a tile of 128 words (3456 ASCII bytes, 16-byte aligned): global memory is touched only
// with coalesced 128/256-bit accesses, the tile is staged in a warp-private slice of shared memory (only
// __syncwarp, no block barrier), and each LANE owns 4 consecutive words = 108 ASCII bytes = exactly 27
// aligned shared-memory words.  Lanes are 27 words apart, and 27 is coprime with the 32 banks, so every
// LDS.32 / STS.32 is conflict-free, and because 108 is a multiple of 4 every byte offset inside a lane's
// block is a compile-time constant (no per-lane funnel shifts).
//
//   encode, per lane: 27 LDS.32; digits come from a
//     PRMT lookup on the low 3 ASCII bits (exactly the reference's pshufb LUT, src/n_to_bits2.rs:127-136:
//     'A'&7=1 'C'&7=3 'T'&7=4 'U'&7=5 'N'&7=6 'G'&7=7, case-insensitive for free); a triplet sitting in
//     bytes 0..2 (or 1..3) of a register is ONE multiply: x * (1<<16 | 5<<8 | 25) leaves a+5b+25c in
//     byte 2 (byte 3) with no carries -- the GPU analogue of maddubs(5,25)+add (:160-163).
//   decode, per lane and word: q = e*205>>10 (= e/5), c = e*41>>10 (= e/25), and the three digits as PRMT
//     selector nibbles are e + 11q + 176c; PRMT against 'A','C','T','G','N' emits the letters (the
//     reciprocal-multiply idea of :209-211,241-247 on 32-bit registers); the four 27-byte strings of a lane
//     are spliced into 27 whole words with constant funnel shifts and written with 27 STS.32.
#pragma once
#include "codec_kernels.cuh"

namespace cn {

constexpr int kB5Nt = 27;                       // nucleotides per word
constexpr int kB5WarpWords = 128;               // words per warp tile (4 per lane)
constexpr int kB5WarpBytes = kB5WarpWords * kB5Nt;        // 3456
constexpr int kB5WarpVecs = kB5WarpBytes / 16;            // 216 uint4
constexpr int kB5SmemPerWarp = kB5WarpBytes;
// warps per CTA.  Measured at 10 GiB (encode / decode GB/s): 2 -> 7273 / 6762, 4 -> 7127 / 6743, 8 -> 6905 / 6772.
// A CTA's shared memory is only released when its last warp retires, so small CTAs recycle tile slots sooner;
// 2 warps still allows 32 CTAs = 64 warps per SM.
constexpr int kB5Warps = 2;
constexpr uint32_t kB5Mul = (1u << 16) | (5u << 8) | 25u; // 0x010519

// ---- cp.async.bulk (TMA 1-D bulk copy) + mbarrier helpers ------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 4 ASCII bytes -> 4 base-5 digits.  The byte's low 3 bits index the 8-entry LUT held in two registers.
__device__ __forceinline__ uint32_t b5_digits4(uint32_t x)
{
    // nibble-pack the four 3-bit indices: byte0 = i0 | i1<<4, byte2 = i2 | i3<<4 (bit 3 / 7 come from bit 7 of
    // an ASCII byte, i.e. 0, so PRMT's sign-replicate selector bit stays clear)
    // Measured on B200 (tools/ab, 10 GiB): this form -- two-constant mask that nvcc splits into two LOP3s, and
    // __byte_perm's implicit selector mask -- runs encode at 6910 GB/s; the leaner single-LOP3 / raw-PRMT form
    // has 20 % fewer ALU ops yet runs at 6680 GB/s.  The kernel is latency-, not ALU-bound; keep what is faster.
    uint32_t t = (x & 0x87878787u) | ((x >> 4) & 0x78787878u);
    uint32_t sel = __byte_perm(t, 0u, 0x4420);
    return __byte_perm(0x01000000u, 0x03040202u, sel);      // idx: 0,1(A),2 -> 0; 3(C) -> 1; 4(T),5(U) -> 2; 6(N) -> 4; 7(G) -> 3
}

// one byte -> its digit (the same low-3-bit LUT, scalar form)
__device__ __forceinline__ uint32_t b5_digit(uint32_t byte) { return (uint32_t)(0x0304020201000000ull >> (8 * (byte & 7))) & 7u; }

// ---- validation fused into the base-5 encode (optional) ----------------------------------------------------
// A byte belongs to {A,C,G,T,U,N,a,c,g,t,u,n} iff b & 0xD8 equals what its low 3 bits predict: a/c/g (1,3,7) -> 0x40,
// t/u (4,5) -> 0x50, n (6) -> 0x48; indices 0 and 2 can never match.  The lookup reuses the PRMT selector b5_digits4
// builds anyway (2 extra ALU ops per 4 nucleotides); the accumulated word is an exact "anything invalid in these
// bytes?" flag (a byte >= 0x80 can also flag its neighbour, never hide itself), so a lane that sees it set recounts /
// re-derives its bytes exactly.
// exact: 0x80 in every byte of x outside the 12-letter alphabet
__device__ __forceinline__ uint32_t b5_invalid_bytes_mask(uint32_t x)
{
    uint32_t t = (x & 0x07070707u) | ((x >> 4) & 0x70707070u);
    uint32_t expect = prmt(0x40FF40FFu, 0x40485050u, __byte_perm(t, 0u, 0x4420));
    uint32_t diff = (x & 0xD8D8D8D8u) ^ expect;
    return (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
}
__device__ __forceinline__ bool b5_invalid_byte(uint32_t b)
{
    uint32_t c = b | 0x20u;
    return !(c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'u' || c == 'n');
}
// digits of 4 bytes under the given mode; in the non-plain modes `flag` accumulates the cheap any-invalid test
template <int MODE>
__device__ __forceinline__ uint32_t b5_digits4_mode(uint32_t x, uint32_t &flag)
{
    if constexpr (MODE == kEncPlain) return b5_digits4(x);
    uint32_t t = (x & 0x87878787u) | ((x >> 4) & 0x78787878u);            // as in b5_digits4
    uint32_t sel = __byte_perm(t, 0u, 0x4420);
    // raw PRMT for both lookups: a selector nibble can only have bit 3 set for a byte >= 0x80, which raises the flag and
    // sends the lane to the exact recompute anyway -- so the & 0x7777 that __byte_perm adds is not needed here
    flag |= (x & 0xD8D8D8D8u) ^ prmt(0x40FF40FFu, 0x40485050u, sel);
    return prmt(0x01000000u, 0x03040202u, sel);
}

// Slow path of a lane whose flag fired: its four words are recomputed byte by byte from the staged tile (exact count; in
// kEncLutExact every byte outside the alphabet becomes digit 0, which is what BYTE_LUT of n_to_bits2_lut yields,
// src/n_to_bits2.rs:8-23).  Deliberately rolled loops: this path is rare, and its register footprint must not set the
// kernel's (the speculative fast path around it runs at the plain kernel's 30 registers).
template <int MODE>
__device__ __noinline__ uint32_t b5_recompute_lane(const uint8_t *__restrict__ bytes, uint2 (&w)[4])
{
    uint32_t invalid = 0;
#pragma unroll 1
    for (int g = 0; g < 4; g++) {
        uint64_t word = 0;
#pragma unroll 1
        for (int t = 0; t < 9; t++) {
            uint32_t e = 0, mul = 1;
#pragma unroll 1
            for (int j = 0; j < 3; j++) {
                const uint32_t byte = bytes[27 * g + 3 * t + j];
                uint32_t dgt = b5_digit(byte);
                if (b5_invalid_byte(byte)) { invalid++; if (MODE == kEncLutExact) dgt = 0; }
                e += dgt * mul;
                mul *= 5;
            }
            word |= (uint64_t)e << (7 * t);
        }
        w[g] = make_uint2((uint32_t)word, (uint32_t)(word >> 32));
    }
    return invalid;
}

// 108 digits (the bytes of d[0..26]) -> four packed words.  Triplet j sits at bytes 3j..3j+2; the pattern repeats
// every 3 registers: offset 0 (product byte 2), 3 and 2 (funnel shift first), 1 (product byte 3).
__device__ __forceinline__ void b5_pack108(const uint32_t (&d)[27], uint2 (&w)[4])
{
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t e[9];
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const int j = 9 * g + t, m = 3 * (j >> 2), r = j & 3;
            uint32_t v = r == 0 ? d[m] : r == 1 ? __funnelshift_r(d[m], d[m + 1], 24)
                       : r == 2 ? __funnelshift_r(d[m + 1], d[m + 2], 16) : d[m + 2];
            e[t] = __byte_perm(v * kB5Mul, 0u, r == 3 ? 0x4443 : 0x4442);       // [a+5b+25c, 0, 0, 0]
        }
        w[g].x = e[0] + (e[1] << 7) + (e[2] << 14) + (e[3] << 21) + (e[4] << 28);
        w[g].y = (e[4] >> 4) + (e[5] << 3) + (e[6] << 10) + (e[7] << 17) + (e[8] << 24);
    }
}

// packed word -> 27 ASCII bytes in a[0..6] (byte 27 is 'A', to be replaced by the caller)
__device__ __forceinline__ void b5_unpack27(uint2 w, uint32_t (&a)[7])
{
    uint32_t e[9];
    e[0] = w.x & 0x7F;          e[1] = (w.x >> 7) & 0x7F;   e[2] = (w.x >> 14) & 0x7F;  e[3] = (w.x >> 21) & 0x7F;
    e[4] = __funnelshift_r(w.x, w.y, 28) & 0x7F;
    e[5] = (w.y >> 3) & 0x7F;   e[6] = (w.y >> 10) & 0x7F;  e[7] = (w.y >> 17) & 0x7F;  e[8] = (w.y >> 24) & 0x7F;
    uint32_t s[9];                                   // 12-bit selector triples: a | b<<4 | c<<8
#pragma unroll
    for (int t = 0; t < 9; t++) {
        // (e * 205) >> 10 = e / 5 and (e * 41) >> 10 = e / 25, exact for e < 128; the shift is folded into a
        // multiply-high so both run on the FMA pipe (the ALU pipe is this kernel's busiest unit)
        uint32_t q = __umulhi(e[t], 205u << 22);
        uint32_t c = __umulhi(e[t], 41u << 22);
        s[t] = e[t] + 11u * q + 176u * c;            // (e - 5q) + 16 (q - 5c) + 256 c
    }
    uint32_t n0 = s[0] + (s[1] << 12) + (s[2] << 24);
    uint32_t n1 = (s[2] >> 8) + (s[3] << 4) + (s[4] << 16) + (s[5] << 28);
    uint32_t n2 = (s[5] >> 4) + (s[6] << 8) + (s[7] << 20);
    uint32_t n3 = s[8];
    const uint32_t lo = 0x47544341u, hi = 0x0000004Eu;      // 'A','C','T','G' | 'N',0,0,0
    a[0] = __byte_perm(lo, hi, n0);  a[1] = __byte_perm(lo, hi, n0 >> 16);
    a[2] = __byte_perm(lo, hi, n1);  a[3] = __byte_perm(lo, hi, n1 >> 16);
    a[4] = __byte_perm(lo, hi, n2);  a[5] = __byte_perm(lo, hi, n2 >> 16);
    a[6] = __byte_perm(lo, hi, n3);
}

// four 27-byte strings (a[g][0..6], byte 27 of each is filler) -> 27 words holding the 108 bytes back to back
__device__ __forceinline__ void b5_splice108(const uint32_t (&a)[4][7], uint32_t (&o)[27])
{
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = a[0][k];
    o[6] = __byte_perm(a[0][6], a[1][0], 0x4210);                 // bytes 24,25,26 | next 0
#pragma unroll
    for (int k = 0; k < 6; k++) o[7 + k] = __funnelshift_r(a[1][k], a[1][k + 1], 8);
    o[13] = __byte_perm(a[1][6], a[2][0], 0x5421);                // bytes 25,26 | next 0,1
#pragma unroll
    for (int k = 0; k < 6; k++) o[14 + k] = __funnelshift_r(a[2][k], a[2][k + 1], 16);
    o[20] = __byte_perm(a[2][6], a[3][0], 0x6542);                // byte 26 | next 0,1,2
#pragma unroll
    for (int k = 0; k < 6; k++) o[21 + k] = __funnelshift_r(a[3][k], a[3][k + 1], 24);
}

// ---- scalar paths for whatever the warp tiles do not cover (ragged end, unaligned buffers) -------------

// words [first, total): bytes past len count as digit 0 (src/n_to_bits2.rs:59-70).  Returns the number of bytes outside
// the alphabet among those this thread read (MODE != kEncPlain).
template <int MODE = kEncPlain>
__device__ __forceinline__ uint32_t b5_encode_scalar(const uint8_t *__restrict__ n, size_t len, uint64_t *__restrict__ out,
                                                     size_t first, size_t total, size_t lane, size_t stride)
{
    uint32_t invalid = 0;
    auto digit = [&](size_t i) -> uint32_t {
        if (i >= len) return 0u;
        const uint32_t byte = n[i];
        if constexpr (MODE != kEncPlain) {
            if (b5_invalid_byte(byte)) { invalid++; if (MODE == kEncLutExact) return 0u; }
        }
        return b5_digit(byte);
    };
    for (size_t w = first + lane; w < total; w += stride) {
        uint64_t word = 0;
        size_t base = w * kB5Nt;
#pragma unroll 1
        for (int t = 0; t < 9; t++) {
            size_t i = base + 3 * t;
            uint32_t a = digit(i), b = digit(i + 1), c = digit(i + 2);
            word |= (uint64_t)(a + 5 * b + 25 * c) << (7 * t);
        }
        out[w] = word;
    }
    return invalid;
}

__device__ __forceinline__ void b5_decode_scalar(const uint64_t *__restrict__ bits, uint8_t *__restrict__ out, size_t len,
                                                 size_t first, size_t total, size_t lane, size_t stride)
{
    for (size_t w = first + lane; w < total; w += stride) {
        uint64_t word = bits[w];
        size_t base = w * kB5Nt;
#pragma unroll 1
        for (int t = 0; t < 9; t++) {
            uint32_t e = (uint32_t)(word >> (7 * t)) & 0x7Fu;
            uint32_t dig[3] = {e % 5u, (e / 5u) % 5u, e / 25u};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                size_t i = base + 3 * t + k;
                if (i < len) out[i] = (uint8_t)(0x0000004E47544341ull >> (8 * dig[k]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// encode: grid of CTAs with kB5Warps warps; global warp g owns tile g while g < ntiles; warp `ntiles`
// finishes words [ntiles*128, total_words) with the scalar path.  `in` must be 16-byte aligned when
// ntiles > 0.
// ------------------------------------------------------------------------------------------------------
// more destinations of the same packed words (multi-GPU: peer-mapped buffers of the other ranks, see encode_multi_kernel)
struct PeerOuts64 { uint64_t *p[kMaxPeers]; };

template <bool TMA, int MODE = kEncPlain, bool MULTI = false>
__global__ void __launch_bounds__(kB5Warps * 32)
b5_encode_kernel(const uint8_t *__restrict__ in, uint64_t *__restrict__ out, size_t len, size_t ntiles, size_t total_words,
                 unsigned long long *__restrict__ invalid_counter, PeerOuts64 more = PeerOuts64{}, int nmore = 0)
{
    uint32_t n_invalid = 0;
    __shared__ __align__(128) uint8_t smem[kB5Warps * kB5SmemPerWarp];
    __shared__ uint64_t bars[kB5Warps];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t g = (size_t)blockIdx.x * kB5Warps + warp;
    if (g < ntiles) {
        uint8_t *tile = smem + warp * kB5SmemPerWarp;
        const uint8_t *src = in + g * kB5WarpBytes;
        if constexpr (TMA) {
            // one lane asks the TMA unit for the whole 3456-byte tile; nobody spends LSU slots or registers on it
            if (lane == 0) {
                mbar_init(&bars[warp], 1);
                mbar_expect_tx(&bars[warp], kB5WarpBytes);
                bulk_load(tile, src, kB5WarpBytes, &bars[warp]);
            }
            __syncwarp();
            mbar_wait(&bars[warp], 0);
        } else {
            uint4 v[7];
#pragma unroll
            for (int k = 0; k < 7; k++) {                      // 216 vectors: 6 full rounds + 24 lanes
                unsigned i = lane + 32 * k;
                if (i < kB5WarpVecs) v[k] = ld_stream16(src + 16 * i);
            }
#pragma unroll
            for (int k = 0; k < 7; k++) {
                unsigned i = lane + 32 * k;
                if (i < kB5WarpVecs) *reinterpret_cast<uint4 *>(tile + 16 * i) = v[k];
            }
            __syncwarp();
        }
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile) + 27 * lane;     // this lane's 108 bytes
        uint32_t d[27], flag = 0;
        if constexpr (MODE == kEncPlain) {
#pragma unroll
            for (int k = 0; k < 27; k++) d[k] = b5_digits4(tw[k]);
        } else {
            // volatile: keeps the 27 tile loads in program order.  Without it ptxas hoists all of them above the validation
            // chain (with TMA staging nothing else orders them) and the kernel lands at 72-80 registers -- a third of the
            // occupancy, 26 % slower; in order, the packer consumes the digits as they arrive and 32 registers suffice.
            const volatile uint32_t *vtw = tw;
#pragma unroll
            for (int k = 0; k < 27; k++) d[k] = b5_digits4_mode<MODE>(vtw[k], flag);
        }
        uint2 w[4];
        b5_pack108(d, w);                                        // speculative: right unless the flag fired
        if constexpr (MODE != kEncPlain) {
            if (flag) n_invalid += b5_recompute_lane<MODE>(tile + 108 * lane, w);   // rare: something in these 108 bytes is off-alphabet
        }
        st_stream32(out + g * kB5WarpWords + 4 * lane, make_uint4(w[0].x, w[0].y, w[1].x, w[1].y),
                    make_uint4(w[2].x, w[2].y, w[3].x, w[3].y));
        if constexpr (MULTI) {
#pragma unroll
            for (int k = 0; k < kMaxPeers; k++)
                if (k < nmore) st_stream32(more.p[k] + g * kB5WarpWords + 4 * lane, make_uint4(w[0].x, w[0].y, w[1].x, w[1].y),
                                           make_uint4(w[2].x, w[2].y, w[3].x, w[3].y));
        }
    } else if (g == ntiles) {
        n_invalid += b5_encode_scalar<MODE>(in, len, out, ntiles * kB5WarpWords, total_words, lane, 32);
        if constexpr (MULTI) {
            for (int k = 0; k < nmore; k++) b5_encode_scalar<kEncPlain>(in, len, more.p[k], ntiles * kB5WarpWords, total_words, lane, 32);
        }
    }
    if constexpr (MODE != kEncPlain) { if (invalid_counter) report_invalid(invalid_counter, n_invalid); }
}

// ------------------------------------------------------------------------------------------------------
// encode, ASCII pointer NOT 16-byte aligned (mis = in & 15, 1..15).  TMA and 128-bit loads need aligned global
// addresses, so the warp stages the 217 ALIGNED vectors that cover its tile; in shared memory the tile then starts
// `mis` bytes in, i.e. each lane's 108 bytes start (mis & 3) bytes into a word: one more LDS and a funnel shift per
// word, the shift amount the same for every lane.  The host only hands out tiles whose trailing vector lies inside
// the buffer (ntiles * 3456 + 16 <= len + mis).
// ------------------------------------------------------------------------------------------------------
constexpr int kB5MisVecs = kB5WarpVecs + 1;                     // 217
constexpr int kB5MisSmemPerWarp = kB5MisVecs * 16;              // 3472

template <int MODE = kEncPlain>
__global__ void __launch_bounds__(kB5Warps * 32)
b5_encode_mis_kernel(const uint8_t *__restrict__ in, unsigned mis, uint64_t *__restrict__ out, size_t len, size_t ntiles,
                     size_t total_words, unsigned long long *__restrict__ invalid_counter)
{
    uint32_t n_invalid = 0;
    __shared__ __align__(16) uint8_t smem[kB5Warps * kB5MisSmemPerWarp];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t g = (size_t)blockIdx.x * kB5Warps + warp;
    if (g < ntiles) {
        uint8_t *tile = smem + warp * kB5MisSmemPerWarp;
        const uint8_t *src = (in - mis) + g * kB5WarpBytes;     // 16-byte aligned
        uint4 v[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            unsigned i = lane + 32 * k;
            if (i < kB5MisVecs) {
                if (g == 0 && i == 0) {
                    // the first aligned vector starts `mis` bytes BEFORE the caller's buffer: take only the caller's bytes
                    uint32_t w4[4] = {0u, 0u, 0u, 0u};
                    for (unsigned b = mis; b < 16; b++) w4[b >> 2] |= (uint32_t)in[b - mis] << (8 * (b & 3));
                    v[k] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                } else {
                    v[k] = ld_stream16(src + 16 * i);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            unsigned i = lane + 32 * k;
            if (i < kB5MisVecs) *reinterpret_cast<uint4 *>(tile + 16 * i) = v[k];
        }
        __syncwarp();
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile) + (mis >> 2) + 27 * lane;
        const unsigned shift = (mis & 3u) * 8u;
        uint32_t d[27], flag = 0;
        uint32_t prev = tw[0];
#pragma unroll
        for (int k = 0; k < 27; k++) {
            uint32_t next = tw[k + 1];
            d[k] = b5_digits4_mode<MODE>(__funnelshift_r(prev, next, shift), flag);
            prev = next;
        }
        uint2 w[4];
        b5_pack108(d, w);
        if constexpr (MODE != kEncPlain) {
            if (flag) n_invalid += b5_recompute_lane<MODE>(tile + mis + 108 * lane, w);
        }
        st_stream32(out + g * kB5WarpWords + 4 * lane, make_uint4(w[0].x, w[0].y, w[1].x, w[1].y),
                    make_uint4(w[2].x, w[2].y, w[3].x, w[3].y));
    } else if (g == ntiles) {
        n_invalid += b5_encode_scalar<MODE>(in, len, out, ntiles * kB5WarpWords, total_words, lane, 32);
    }
    if constexpr (MODE != kEncPlain) { if (invalid_counter) report_invalid(invalid_counter, n_invalid); }
}

// generic encode for buffers the tiled kernel cannot take (ASCII pointer not 16-byte aligned)
template <int MODE = kEncPlain>
__global__ void __launch_bounds__(256)
b5_encode_scalar_kernel(const uint8_t *__restrict__ in, uint64_t *__restrict__ out, size_t len, size_t total_words,
                        unsigned long long *__restrict__ invalid_counter)
{
    uint32_t n_invalid = b5_encode_scalar<MODE>(in, len, out, 0, total_words, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
                                                (size_t)gridDim.x * blockDim.x);
    if constexpr (MODE != kEncPlain) { if (invalid_counter) report_invalid(invalid_counter, n_invalid); }
}

// ------------------------------------------------------------------------------------------------------
// decode: same tiling; `out` must be 16-byte aligned when ntiles > 0.  Tiles cover only words whose 27
// nucleotides are all wanted (27*128*ntiles <= len); the scalar warp writes the rest up to len.
// ------------------------------------------------------------------------------------------------------
template <bool TMA>
__global__ void __launch_bounds__(kB5Warps * 32)
b5_decode_kernel(const uint64_t *__restrict__ bits, uint8_t *__restrict__ out, size_t len, size_t ntiles, size_t total_words)
{
    __shared__ __align__(128) uint8_t smem[kB5Warps * kB5SmemPerWarp];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t g = (size_t)blockIdx.x * kB5Warps + warp;
    if (g < ntiles) {
        uint8_t *tile = smem + warp * kB5SmemPerWarp;
        uint32_t *tw = reinterpret_cast<uint32_t *>(tile);
        const uint64_t *src = bits + g * kB5WarpWords;
        const u32x8 w = ld_stream32(src + 4 * lane);                 // this lane's 4 words
        uint32_t a[4][7];
#pragma unroll
        for (int c = 0; c < 4; c++) b5_unpack27(make_uint2(w.v[2 * c], w.v[2 * c + 1]), a[c]);
        uint32_t o[27];
        b5_splice108(a, o);
#pragma unroll
        for (int k = 0; k < 27; k++) tw[27 * lane + k] = o[k];
        uint8_t *dst = out + g * kB5WarpBytes;
        if constexpr (TMA) {
            fence_proxy_async_smem();                           // generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) { bulk_store(dst, tile, kB5WarpBytes); bulk_store_wait_read(); }
        } else {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 7; k++) {
                unsigned i = lane + 32 * k;
                if (i < kB5WarpVecs) st_stream16(dst + 16 * i, *reinterpret_cast<const uint4 *>(tile + 16 * i));
            }
        }
    } else if (g == ntiles) {
        b5_decode_scalar(bits, out, len, ntiles * kB5WarpWords, total_words, lane, 32);
    }
}

// ------------------------------------------------------------------------------------------------------
// decode, ASCII destination NOT 16-byte aligned (mis = out & 15).  The tile is laid out in shared memory `mis`
// bytes in, so shared vector i corresponds to the ALIGNED global vector (out - mis) + 16 i: vectors 1..215 are
// stored whole, the partial first and last vectors byte by byte (neighbouring tiles share those vectors and each
// writes only its own bytes).  A lane's 108 bytes start (mis & 3) bytes into a word: the words are funnel-shifted
// and the word straddling two lanes is completed with the next lane's first word (one shuffle).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kB5Warps * 32)
b5_decode_mis_kernel(const uint64_t *__restrict__ bits, uint8_t *__restrict__ out, unsigned mis, size_t len, size_t ntiles,
                     size_t total_words)
{
    __shared__ __align__(16) uint8_t smem[kB5Warps * kB5MisSmemPerWarp];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t g = (size_t)blockIdx.x * kB5Warps + warp;
    if (g < ntiles) {
        uint8_t *tile = smem + warp * kB5MisSmemPerWarp;
        uint32_t *tw = reinterpret_cast<uint32_t *>(tile) + (mis >> 2) + 27 * lane;
        const u32x8 w = ld_stream32(bits + g * kB5WarpWords + 4 * lane);
        uint32_t a[4][7];
#pragma unroll
        for (int c = 0; c < 4; c++) b5_unpack27(make_uint2(w.v[2 * c], w.v[2 * c + 1]), a[c]);
        uint32_t o[27];
        b5_splice108(a, o);
        const unsigned s = mis & 3u;
        const uint32_t next0 = __shfl_down_sync(0xFFFFFFFFu, o[0], 1);
        if (s == 0) {
#pragma unroll
            for (int k = 0; k < 27; k++) tw[k] = o[k];
        } else {
            const unsigned sh = 8u * (4u - s);                  // word k+1 = last s bytes of o[k], first 4-s bytes of o[k+1]
            if (lane == 0) tw[0] = __funnelshift_r(0u, o[0], sh);
#pragma unroll
            for (int k = 0; k < 26; k++) tw[k + 1] = __funnelshift_r(o[k], o[k + 1], sh);
            tw[27] = __funnelshift_r(o[26], next0, sh);         // lane 31: the extra bytes are past the tile, never stored
        }
        __syncwarp();
        uint8_t *dst = (out - mis) + g * kB5WarpBytes;          // 16-byte aligned
#pragma unroll
        for (int k = 0; k < 7; k++) {
            unsigned i = lane + 32 * k;
            if (i >= 1 && i < kB5MisVecs - 1) st_stream16(dst + 16 * i, *reinterpret_cast<const uint4 *>(tile + 16 * i));
        }
        if (lane < 16) { if (lane >= mis) dst[lane] = tile[lane]; }                                       // first, partial vector
        else if (lane - 16 < mis) dst[16 * (kB5MisVecs - 1) + lane - 16] = tile[16 * (kB5MisVecs - 1) + lane - 16];   // last
    } else if (g == ntiles) {
        b5_decode_scalar(bits, out, len, ntiles * kB5WarpWords, total_words, lane, 32);
    }
}

__global__ void __launch_bounds__(256)
b5_decode_scalar_kernel(const uint64_t *__restrict__ bits, uint8_t *__restrict__ out, size_t len, size_t total_words)
{
    b5_decode_scalar(bits, out, len, 0, total_words, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// synthetic 5-letter data, bit-identical to the host generator used by the tests:
// alphabet 5 = "ACGTN"[(b*5)>>8], alphabet 12 = "ACGTUNacgtun"[(b*12)>>8]
__global__ void b5_generate_kernel(uint8_t *__restrict__ out, size_t offset, size_t len, uint64_t seed, int alphabet)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        size_t gidx = offset + i;
        uint64_t h = mix64(seed, gidx >> 3);
        uint32_t b = (uint32_t)(h >> (8 * (gidx & 7))) & 0xFFu;
        uint32_t k = alphabet == 5 ? (b * 5u) >> 8 : (b * 12u) >> 8;
        const uint64_t lo = 0x63614E5554474341ull;               // A C G T U N a c
        const uint32_t hi = 0x6E757467u;                         // g t u n
        uint32_t ch;
        if (alphabet == 5) ch = k < 4 ? (uint32_t)(lo >> (8 * k)) & 0xFFu : (uint32_t)'N';
        else ch = k < 8 ? (uint32_t)(lo >> (8 * k)) & 0xFFu : (hi >> (8 * (k - 8))) & 0xFFu;
        out[i] = (uint8_t)ch;
    }
}

}  // namespace cn
