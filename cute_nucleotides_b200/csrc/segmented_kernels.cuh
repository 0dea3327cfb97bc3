// segmented_kernels.cuh -- many independent sequences that ALREADY live in HBM, back to back (a read set produced by a GPU
// parser, a batch copied over once): `d_n` holds the sequences tightly concatenated, sequence i = bytes
// [off[i], off[i+1]); the packed side gives every sequence whole words, sequence i = words [woff[i], woff[i+1]) with
// woff[i+1] - woff[i] = ceil(len_i / 32) -- exactly what n_to_bits_lut / bits_to_n_lut produce per sequence
// (src/n_to_bits.rs:34-69), zero high bits in every sequence's last word included.  The device-resident counterpart of
// cn_n_to_bits_host_batch.
//
// Shape on the GPU: a CTA (64 threads, 4 words each) owns 256 consecutive packed words.  Their ASCII bytes are one contiguous range of at most 8 KiB
// (sequences are tight on the ASCII side; only each sequence's last word covers fewer than 32 bytes), at an arbitrary
// alignment.  All global traffic is aligned 128-bit vectors between HBM and a shared-memory tile; the irregular part --
// which sequence a word belongs to, the byte offset of its 32 nucleotides, the ragged last word -- is handled against the
// tile.  Sequence lookup: interpolation guesses refined by a galloping search (two loads when the guess is right), first for
// the CTA's two end words, then per thread between them.
#pragma once
#include "codec_kernels.cuh"

namespace cn {

constexpr int kSegThreads = 64;                            // threads per CTA
constexpr int kSegWpt = 4;                                 // words per thread: thread t owns words w0 + t + 64 j
constexpr int kSegWords = kSegThreads * kSegWpt;           // 256 words = up to 8 KiB of ASCII per CTA
// Few threads per tile on purpose: a CTA goes through three dependent global-memory phases (locate its sequences, load
// the tile, store), so throughput is bytes in flight per SM over that latency; 64-thread CTAs let 27 tiles (216 KiB) be
// resident per SM where 256-thread CTAs allowed 8 (the first version: 2.2 TB/s).
constexpr int kSegTileBytes = kSegWords * 32 + 64;         // the CTA's ASCII range, alignment slack on both sides, and the
                                                           // 4 bytes a funnel shift reads past a ragged last word
// The tile is kept in shared memory with one pad word after every 8: a warp's threads work on words 32 bytes apart, which
// unpadded puts all 32 lanes on 4 banks (8-way conflicts; the L1/shared pipe was 95 % busy and capped the first version at
// 3.1 TB/s); with the pad, lane t's k-th word lands on bank (9 t + k) mod 32 -- conflict-free.
constexpr int kSegTileWords = kSegTileBytes / 4 + kSegTileBytes / 32 + 1;
__device__ __forceinline__ uint32_t seg_slot(uint32_t word_index) { return word_index + (word_index >> 3); }

// index s of the sequence that owns packed word w: the last s with woff[s] <= w (empty sequences in between are skipped
// because woff[s + 1] <= w for them).  `hint` is a guess; the search gallops away from it in whichever direction is needed
// and finishes with a binary search, so a good guess costs two loads and a bad one O(log distance).
__device__ __forceinline__ size_t seg_find(const uint64_t *__restrict__ woff, size_t count, size_t hint, uint64_t w)
{
    size_t lo = hint < count ? hint : count - 1, hi;
    if (woff[lo] > w) {                                    // guessed too far: gallop back (woff[0] = 0 <= w ends it)
        size_t step = 1;
        do { hi = lo; lo = lo > step ? lo - step : 0; step <<= 1; } while (woff[lo] > w);
    } else {                                               // gallop forward: woff[lo] <= w < woff[min(hi, count)]
        size_t step = 1;
        hi = lo + 1;
        while (hi < count && woff[hi] <= w) { lo = hi; hi += step; step <<= 1; }
        if (hi > count) hi = count;
    }
    while (lo + 1 < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        if (woff[mid] <= w) lo = mid; else hi = mid;
    }
    return lo;
}

// word w of W lies about w/W of the way through the sequences -- exact for equal-length reads
__device__ __forceinline__ size_t seg_guess(uint64_t w, uint64_t total_words, size_t count)
{
    return (size_t)((double)w * (double)count / (double)total_words);        // only a hint: rounding does not matter
}

struct SegGeo { uint64_t p; uint32_t avail; };            // ASCII byte offset of a word's first nucleotide, nucleotides in it

__device__ __forceinline__ SegGeo seg_geo(const uint64_t *__restrict__ off, const uint64_t *__restrict__ woff, size_t s, uint64_t w)
{
    SegGeo g;
    g.p = off[s] + 32 * (w - woff[s]);
    const uint64_t left = off[s + 1] - g.p;
    g.avail = left < 32 ? (uint32_t)left : 32u;
    return g;
}

// ------------------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------------------
// What both kernels do first: locate the sequence of every word of the CTA and agree on the ASCII byte range [lo, hi) the
// words cover.  The lookup runs on CONSECUTIVE words per thread (thread t: words 4t..4t+3), so a thread searches once and
// then only steps to the next sequence when a word crosses a boundary; the per-word result (low 32 bits of the ASCII offset,
// nucleotides in the word) goes through shared memory to the thread that processes the word (strided ownership keeps the
// tile accesses of a warp 32 bytes apart instead of 128).
struct SegShared {
    uint64_t first, last, lo, hi;
    uint32_t p32[kSegWords];              // low 32 bits of every word's ASCII byte offset (differences inside a tile are < 2^14)
    uint8_t avail[kSegWords];             // nucleotides in the word, 0 = past the end
};

__device__ __forceinline__ void seg_locate(const uint64_t *__restrict__ off, const uint64_t *__restrict__ woff, size_t count,
                                           uint64_t total_words, uint64_t w0, SegShared &sh)
{
    const uint32_t nw = (uint32_t)(total_words - w0 < kSegWords ? total_words - w0 : kSegWords);      // words of this CTA
    // the CTA's two end words are located from a global interpolation guess, every other word interpolates between them
    if (threadIdx.x == 0) sh.first = seg_find(woff, count, seg_guess(w0, total_words, count), w0);
    if (threadIdx.x == 32) sh.last = seg_find(woff, count, seg_guess(w0 + nw - 1, total_words, count), w0 + nw - 1);
    __syncthreads();
    const uint32_t i0 = kSegWpt * threadIdx.x;
    if (i0 < nw) {
        // all words of the CTA in ONE sequence (long sequences: most CTAs): nothing to search
        const float ratio = nw > 1 ? (float)(sh.last - sh.first) / (float)(nw - 1) : 0.f;
        size_t s = sh.first == sh.last ? (size_t)sh.first : seg_find(woff, count, sh.first + (size_t)(ratio * (float)i0), w0 + i0);
        uint64_t ws = woff[s], we = woff[s + 1], os = off[s], oe = off[s + 1];
#pragma unroll
        for (int j = 0; j < kSegWpt; j++) {
            const uint32_t i = i0 + j;
            const uint64_t w = w0 + i;
            uint32_t avail = 0, p32 = 0;
            if (i < nw) {
                while (w >= we) { s++; ws = we; we = woff[s + 1]; os = oe; oe = off[s + 1]; }       // next non-empty sequence
                const uint64_t p = os + 32 * (w - ws), left = oe - p;
                avail = left < 32 ? (uint32_t)left : 32u;
                p32 = (uint32_t)p;
                if (i == 0) sh.lo = p;
                if (i == nw - 1) sh.hi = p + avail;
            }
            sh.p32[i] = p32;
            sh.avail[i] = (uint8_t)avail;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kSegWpt; j++) { sh.p32[i0 + j] = 0; sh.avail[i0 + j] = 0; }
    }
    __syncthreads();
}

// Encode needs no tile: a word's (up to) 32 bytes lie inside three ALIGNED 16-byte vectors, which the thread loads
// straight from global memory (a warp's words are consecutive, so the three loads of a warp cover one contiguous ~1 KiB
// span; the third vector of a lane is the first of the next: an L1 hit) and aligns in registers -- a select network picks
// the 9 words that contain the bytes, funnel shifts do the byte alignment.  Only the lookup result goes through shared
// memory.  (The first version staged the ASCII range in a shared-memory tile like the decoder below still does; ncu showed it
// issue-bound -- 230 instructions per word, mostly tile address arithmetic -- at 3.3 Tnt/s.)
__global__ void __launch_bounds__(kSegThreads)
encode_segmented_kernel(const uint8_t *__restrict__ n, const uint64_t *__restrict__ off, const uint64_t *__restrict__ woff,
                        size_t count, uint64_t total_words, uint64_t total_bytes, uint64_t *__restrict__ out)
{
    __shared__ SegShared sh;
    const uint64_t w0 = (uint64_t)blockIdx.x * kSegWords;
    seg_locate(off, woff, count, total_words, w0, sh);
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(n), a_end = a0 + total_bytes;
    const uint64_t hi32 = sh.lo & ~0xFFFFFFFFull;                        // p32 holds the low half; the tile spans < 2^14 bytes
    const uint32_t lo32 = (uint32_t)sh.lo;
#pragma unroll
    for (int j = 0; j < kSegWpt; j++) {
        const uint32_t i = threadIdx.x + kSegThreads * j, avail = sh.avail[i];
        if (!avail) continue;
        const uint32_t p32 = sh.p32[i];
        const uint64_t addr = a0 + ((hi32 | p32) + (p32 < lo32 ? 0x100000000ull : 0ull));       // of the word's first byte
        const uint64_t va = addr & ~(uint64_t)15;
        uint32_t r[12];
        if (va >= a0 && va + 48 <= a_end) {
            const uint4 v0 = ld_stream16(reinterpret_cast<const void *>(va));
            const uint4 v1 = ld_stream16(reinterpret_cast<const void *>(va + 16));
            const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(va + 32));
            r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
            r[8] = v2.x; r[9] = v2.y; r[10] = v2.z; r[11] = v2.w;
        } else {                                                          // a vector sticks out of the buffer: take only its bytes
#pragma unroll
            for (int q = 0; q < 12; q++) {                                // (unrolled: r[] must stay in registers)
                uint32_t v = 0;
#pragma unroll 1
                for (int k = 0; k < 4; k++) {
                    const uint64_t a = va + 4 * q + k;
                    if (a >= a0 && a < a_end) v |= (uint32_t)n[a - a0] << (8 * k);
                }
                r[q] = v;
            }
        }
        // the word's bytes start (addr & 15) bytes into r[]: drop (addr & 15) >> 2 whole registers, then shift by the rest
        const unsigned wsel = (unsigned)(addr & 15) >> 2, shift = (unsigned)(addr & 3) * 8u;
        uint32_t t[10], u[9], x[8];
#pragma unroll
        for (int k = 0; k < 10; k++) t[k] = (wsel & 2u) ? r[k + 2] : r[k];
#pragma unroll
        for (int k = 0; k < 9; k++) u[k] = (wsel & 1u) ? t[k + 1] : t[k];
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = __funnelshift_r(u[k], u[k + 1], shift);
        // bytes at and beyond `avail` belong to the next sequence (or to nobody): their codes are cleared in the packed word,
        // which leaves exactly the zeroed high bits of n_to_bits_lut's last word (src/n_to_bits.rs:35).  Branch-free: full and
        // ragged words of a warp run the same instructions.
        const uint64_t keep = avail >= 32 ? ~0ull : ((1ull << (2 * avail)) - 1ull);
        st_stream8(out + w0 + i, pack16(x[0], x[1], x[2], x[3]) & (uint32_t)keep, pack16(x[4], x[5], x[6], x[7]) & (uint32_t)(keep >> 32));
    }
}

// ------------------------------------------------------------------------------------------------------------
// decode: the mirror image.  Every thread expands its word into the tile (whole shared-memory words where it owns all
// four bytes, single bytes at its two ragged ends), then the CTA writes the tile out with aligned 128-bit stores; the
// partial vectors at the two ends of the CTA's range are shared with the neighbouring CTAs and written byte by byte.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSegThreads)
decode_segmented_kernel(const uint64_t *__restrict__ bits, const uint64_t *__restrict__ off, const uint64_t *__restrict__ woff,
                        size_t count, uint64_t total_words, uint8_t *__restrict__ out)
{
    __shared__ uint32_t tile[kSegTileWords];
    __shared__ SegShared sh;
    const uint64_t w0 = (uint64_t)blockIdx.x * kSegWords;
    uint2 word[kSegWpt];
#pragma unroll
    for (int j = 0; j < kSegWpt; j++) {                                    // the packed words do not depend on the lookup: load them first
        const uint64_t w = w0 + threadIdx.x + kSegThreads * j;
        word[j] = w < total_words ? ld_stream8(bits + w) : make_uint2(0u, 0u);
    }
    seg_locate(off, woff, count, total_words, w0, sh);
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(out);
    const uint64_t lo_addr = (a0 + sh.lo) & ~(uint64_t)15, hi_addr = (a0 + sh.hi + 15) & ~(uint64_t)15;
    const uint32_t tile0 = (uint32_t)(lo_addr - a0);
#pragma unroll
    for (int j = 0; j < kSegWpt; j++) {
        const uint32_t i = threadIdx.x + kSegThreads * j;
        const int avail = (int)sh.avail[i];
        if (!avail) continue;
        const uint4 a = unpack16(word[j].x), b = unpack16(word[j].y);
        const uint32_t y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const uint32_t o = sh.p32[i] - tile0, base = o >> 2;
        const unsigned s = o & 3u;
        // Shifted image: word k of the tile window holds decoded bytes 4k - s .. 4k - s + 3.  Words [k0, k1) are wholly this
        // word's (one store each, predicated, no divergence between full and ragged words of a warp); the head word 0 (from
        // byte s on, when s > 0) and the tail word k1 (its first (avail + s) & 3 bytes) are shared with the neighbours and
        // written byte by byte.
        uint32_t *t0 = tile + seg_slot(base);
        const uint32_t c = 8u - (base & 7u);                              // tile words until the next pad word
        auto slot_ptr = [&](uint32_t k) { return t0 + k + (k >= c ? 1u : 0u); };
        const uint32_t k1 = ((uint32_t)avail + s) >> 2, rem = ((uint32_t)avail + s) & 3u;
        uint32_t tail = 0;
        if (s == 0 && k1 > 0) *slot_ptr(0) = y[0];
#pragma unroll
        for (uint32_t k = 1; k < 9; k++) {
            const uint32_t v = __funnelshift_l(y[k - 1], k < 8 ? y[k] : 0u, s * 8u);
            if (k < 8 && k < k1) *slot_ptr(k) = v;
            if (k == k1) tail = v;
        }
        if (s != 0 || k1 == 0) {                                          // head: bytes s .. 3 of tile word 0 (fewer if avail is tiny)
            const uint32_t v0 = y[0] << (s * 8u), stop = s + (uint32_t)avail < 4u ? s + (uint32_t)avail : 4u;
            for (uint32_t q = s; q < stop; q++) reinterpret_cast<uint8_t *>(slot_ptr(0))[q] = (uint8_t)(v0 >> (8 * q));
        }
        if (k1 > 0)                                                       // tail: the first `rem` bytes of tile word k1
            for (uint32_t q = 0; q < rem; q++) reinterpret_cast<uint8_t *>(slot_ptr(k1))[q] = (uint8_t)(tail >> (8 * q));
    }
    __syncthreads();
    const uint32_t nvec = (uint32_t)((hi_addr - lo_addr) >> 4);
    const uint32_t begin = (uint32_t)(a0 + sh.lo - lo_addr), end = (uint32_t)(a0 + sh.hi - lo_addr);   // tile bytes this CTA owns
    for (uint32_t v = threadIdx.x; v < nvec; v += kSegThreads) {
        const uint32_t *src = tile + 4 * v + (v >> 1);                    // = seg_slot(4 v); the vector's 4 words never straddle a pad
        const uint4 val = make_uint4(src[0], src[1], src[2], src[3]);
        uint8_t *dst = reinterpret_cast<uint8_t *>(lo_addr + 16ull * v);
        if (16 * v >= begin && 16 * v + 16 <= end) {
            st_stream16(dst, val);
        } else {                                                          // the CTA's first / last vector is shared with a neighbour
            const uint32_t q4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll 1
            for (uint32_t k = 0; k < 16; k++)
                if (16 * v + k >= begin && 16 * v + k < end) dst[k] = (uint8_t)(q4[k >> 2] >> (8 * (k & 3)));
        }
    }
}

}  // namespace cn
