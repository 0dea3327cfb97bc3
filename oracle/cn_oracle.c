/*
 * cn_oracle.c -- CPU restatement of cute-nucleotides' 2-bit codec (and, further down, its
 * base-5 "n_to_bits2" codec).
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library built from it.
 * The product path (cute_nucleotides_b200/csrc) never links, loads or calls anything here.
 *
 * Why a restatement: the reference is a Rust crate and neither rustc nor cargo exist in this
 * image, so the reference itself cannot be compiled (`oracle/_ref` is therefore absent and the
 * CPU baseline is kind "port").  Parity is pinned by the reference's own known-answer tests
 * (src/n_to_bits.rs:408-470, src/n_to_bits2.rs:270-299), which tests/test_oracle.py replays
 * against every function below, and by an independent numpy restatement (oracle/cn_oracle_np.py).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -mbmi2 -mpclmul; NOT -march=native, the .so travels
 * to a different host).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <immintrin.h>

#define CN_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Tables.  src/n_to_bits.rs:8-21 (BYTE_LUT, 128 entries, everything unlisted is 0) and        */
/* src/n_to_bits.rs:23-30 (BITS_LUT).  The reference indexes BYTE_LUT unchecked, so a byte     */
/* >= 0x80 is an out-of-bounds read there; the oracle defines it as code 0 and reports it      */
/* through oracle_count_invalid() so tests never rely on it.                                   */
/* ------------------------------------------------------------------------------------------ */
static uint8_t g_byte_lut[256];
static const uint8_t g_bits_lut[4] = { 'A', 'C', 'T', 'G' };   /* index = 2-bit code */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void)
{
    memset(g_byte_lut, 0, sizeof g_byte_lut);
    g_byte_lut['a'] = 0; g_byte_lut['A'] = 0;
    g_byte_lut['c'] = 1; g_byte_lut['C'] = 1;
    g_byte_lut['t'] = 2; g_byte_lut['T'] = 2;
    g_byte_lut['u'] = 2; g_byte_lut['U'] = 2;
    g_byte_lut['g'] = 3; g_byte_lut['G'] = 3;
}

CN_EXPORT size_t oracle_words_for_len(size_t len)
{   /* src/n_to_bits.rs:35  (n.len() >> 5) + (n.len() & 31 != 0) */
    return (len >> 5) + ((len & 31) ? 1 : 0);
}

CN_EXPORT size_t oracle_count_invalid(const uint8_t *n, size_t len)
{   /* bytes outside {ACGTUacgtu}: the reference's variants disagree on these (SURVEY 8a notes) */
    size_t bad = 0;
    for (size_t i = 0; i < len; i++) {
        switch (n[i]) {
        case 'A': case 'C': case 'G': case 'T': case 'U':
        case 'a': case 'c': case 'g': case 't': case 'u': break;
        default: bad++;
        }
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------ */
/* Scalar oracle pair -- THE parity reference.                                                 */
/* ------------------------------------------------------------------------------------------ */

/* src/n_to_bits.rs:34-47 n_to_bits_lut: zeroed words, nucleotide i -> bits 2(i&31).. of word i>>5 */
CN_EXPORT void oracle_n_to_bits_lut(const uint8_t *n, size_t len, uint64_t *out)
{
    pthread_once(&g_once, init_tables);
    size_t words = oracle_words_for_len(len);
    for (size_t w = 0; w < words; w++) out[w] = 0;
    for (size_t i = 0; i < len; i++) {
        size_t word = i >> 5;
        unsigned shift = (unsigned)(i & 31) << 1;
        out[word] |= (uint64_t)g_byte_lut[n[i]] << shift;
    }
}

/* src/n_to_bits.rs:51-69 bits_to_n_lut.  Returns 1 where the reference panics
 * ("The length is greater than the number of nucleotides!", :52-54), else 0. */
CN_EXPORT int oracle_bits_to_n_lut(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > (nwords << 5)) return 1;
    for (size_t i = 0; i < len; i++) {
        uint64_t cur = bits[i >> 5];
        unsigned shift = (unsigned)(i & 31) << 1;
        out[i] = g_bits_lut[(cur >> shift) & 3u];
    }
    return 0;
}

/* Tail rule shared by all four SIMD encoders (src/n_to_bits.rs:109-111, 160-162, 201-203,
 * 253-255): if len & 31 != 0 the last word is n_to_bits_lut of the remaining bytes. */
static void encode_tail(const uint8_t *n, size_t len, uint64_t *out)
{
    size_t body = len >> 5;
    if (len & 31) oracle_n_to_bits_lut(n + (body << 5), len & 31, out + body);
}

/* ------------------------------------------------------------------------------------------ */
/* AVX2 / BMI2 encoders                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* src/n_to_bits.rs:80-115 n_to_bits_pext: four pext(q, 0x0606..06) per 32 bytes */
__attribute__((target("avx2,bmi2")))
CN_EXPORT void oracle_n_to_bits_pext(const uint8_t *n, size_t len, uint64_t *out)
{
    const uint64_t keep = 0x0606060606060606ull;
    size_t body = len >> 5;
    for (size_t w = 0; w < body; w++) {
        uint64_t q[4];
        memcpy(q, n + (w << 5), 32);
        out[w] = _pext_u64(q[0], keep) | (_pext_u64(q[1], keep) << 16)
               | (_pext_u64(q[2], keep) << 32) | (_pext_u64(q[3], keep) << 48);
    }
    encode_tail(n, len, out);
}

/* src/n_to_bits.rs:121-166 n_to_bits_shift: and 0x06, >>1, >>7 (16-bit), >>12 (32-bit), pshufb */
__attribute__((target("avx2")))
CN_EXPORT void oracle_n_to_bits_shift(const uint8_t *n, size_t len, uint64_t *out)
{
    const __m256i keep = _mm256_set1_epi8(0x06);
    const __m256i pick = _mm256_set_epi32(-1, -1, -1, 0x0C080400, -1, -1, -1, 0x0C080400);
    size_t body = len >> 5;
    for (size_t w = 0; w < body; w++) {
        __m256i v = _mm256_and_si256(_mm256_loadu_si256((const __m256i *)(n + (w << 5))), keep);
        __m256i p = _mm256_or_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 7));
        __m256i q = _mm256_or_si256(p, _mm256_srli_epi32(p, 12));
        __m256i r = _mm256_shuffle_epi8(q, pick);
        uint64_t lo = (uint64_t)_mm256_extract_epi64(r, 0);
        uint64_t hi = (uint64_t)_mm256_extract_epi64(r, 2);
        out[w] = lo | (hi << 32);
    }
    encode_tail(n, len, out);
}

/* src/n_to_bits.rs:172-207 n_to_bits_movemask (the reference's fastest encoder):
 * permute4x64(0xD8), <<6 and <<5 on 64-bit lanes, unpackhi/lo_epi8, two movemasks */
__attribute__((target("avx2")))
static inline uint64_t movemask_word(const uint8_t *p)
{
    __m256i v = _mm256_permute4x64_epi64(_mm256_loadu_si256((const __m256i *)p), 0xD8);
    __m256i bit1 = _mm256_slli_epi64(v, 6);      /* ASCII bit 1 -> byte MSB */
    __m256i bit2 = _mm256_slli_epi64(v, 5);      /* ASCII bit 2 -> byte MSB */
    uint64_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_unpackhi_epi8(bit1, bit2));
    uint64_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_unpacklo_epi8(bit1, bit2));
    return (hi << 32) | lo;
}

__attribute__((target("avx2")))
CN_EXPORT void oracle_n_to_bits_movemask(const uint8_t *n, size_t len, uint64_t *out)
{
    size_t body = len >> 5;
    for (size_t w = 0; w < body; w++) out[w] = movemask_word(n + (w << 5));
    encode_tail(n, len, out);
}

/* src/n_to_bits.rs:213-259 n_to_bits_mul: and 0x06, mullo_epi32 by 1<<23|1<<17|1<<11|1<<5
 * (:223-231), pshufb picks the top byte of every 32-bit lane.  This is the formulation the
 * GPU encode kernel uses per 32-bit register. */
__attribute__((target("avx2")))
CN_EXPORT void oracle_n_to_bits_mul(const uint8_t *n, size_t len, uint64_t *out)
{
    const __m256i keep = _mm256_set1_epi8(0x06);
    const uint32_t m = (1u << 23) | (1u << 17) | (1u << 11) | (1u << 5);
    const __m256i mul = _mm256_set1_epi32((int)m);
    const __m256i pick = _mm256_set_epi32(-1, -1, -1, 0x0F0B0703, -1, -1, -1, 0x0F0B0703);
    size_t body = len >> 5;
    for (size_t w = 0; w < body; w++) {
        __m256i v = _mm256_and_si256(_mm256_loadu_si256((const __m256i *)(n + (w << 5))), keep);
        __m256i r = _mm256_shuffle_epi8(_mm256_mullo_epi32(v, mul), pick);
        uint64_t lo = (uint64_t)_mm256_extract_epi64(r, 0);
        uint64_t hi = (uint64_t)_mm256_extract_epi64(r, 2);
        out[w] = lo | (hi << 32);
    }
    encode_tail(n, len, out);
}

/* ------------------------------------------------------------------------------------------ */
/* AVX2 / BMI2 / PCLMUL decoders.  Like the reference they decode WHOLE words: `out` must have */
/* room for 32*nwords bytes (src/n_to_bits.rs:271, 317, 352); the caller keeps the first `len`.*/
/* Return 1 where the reference panics (:266-268, 310-312, 347-349).                           */
/* ------------------------------------------------------------------------------------------ */
#define LUT_WORD ((int)('A' | ('C' << 8) | ('T' << 16) | ((unsigned)'G' << 24)))

/* src/n_to_bits.rs:265-303 bits_to_n_shuffle (the reference's fastest decoder) */
__attribute__((target("avx2")))
static inline __m256i shuffle_word(uint64_t word)
{
    const __m256i dup4 = _mm256_set_epi32(0x07070707, 0x06060606, 0x05050505, 0x04040404,
                                          0x03030303, 0x02020202, 0x01010101, 0x00000000);
    const __m256i keep = _mm256_set1_epi16(0x0C03);
    const __m256i lut = _mm256_set_epi32('G', 'T', 'C', LUT_WORD, 'G', 'T', 'C', LUT_WORD);
    __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi64x((long long)word), dup4);
    __m256i s = _mm256_srli_epi16(v, 4);
    v = _mm256_and_si256(_mm256_blend_epi16(v, s, 0xAA), keep);
    return _mm256_shuffle_epi8(lut, v);
}

__attribute__((target("avx2")))
CN_EXPORT int oracle_bits_to_n_shuffle(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > (nwords << 5)) return 1;
    for (size_t w = 0; w < nwords; w++)
        _mm256_storeu_si256((__m256i *)(out + (w << 5)), shuffle_word(bits[w]));
    return 0;
}

/* src/n_to_bits.rs:309-340 bits_to_n_pdep: four pdep(word >> 16k, 0x0303..03) then pshufb(lut) */
__attribute__((target("avx2,bmi2")))
CN_EXPORT int oracle_bits_to_n_pdep(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > (nwords << 5)) return 1;
    const uint64_t spread = 0x0303030303030303ull;
    const __m256i lut = _mm256_set_epi32(0, 0, 0, LUT_WORD, 0, 0, 0, LUT_WORD);
    for (size_t w = 0; w < nwords; w++) {
        uint64_t x = bits[w];
        __m256i v = _mm256_set_epi64x((long long)_pdep_u64(x >> 48, spread), (long long)_pdep_u64(x >> 32, spread),
                                      (long long)_pdep_u64(x >> 16, spread), (long long)_pdep_u64(x, spread));
        _mm256_storeu_si256((__m256i *)(out + (w << 5)), _mm256_shuffle_epi8(lut, v));
    }
    return 0;
}

/* src/n_to_bits.rs:346-404 bits_to_n_clmul: bytes spread to 32-bit lanes, carry-less multiply
 * by 1 | 1<<6 | 1<<12 | 1<<18 (:357-365), and 0x03, pshufb(lut); two 128-bit stores per word */
__attribute__((target("ssse3,pclmul,sse4.1")))
CN_EXPORT int oracle_bits_to_n_clmul(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > (nwords << 5)) return 1;
    const __m128i lo_sel = _mm_set_epi32((int)0xFFFFFF03u, (int)0xFFFFFF02u, (int)0xFFFFFF01u, (int)0xFFFFFF00u);
    const __m128i hi_sel = _mm_set_epi32((int)0xFFFFFF07u, (int)0xFFFFFF06u, (int)0xFFFFFF05u, (int)0xFFFFFF04u);
    const __m128i mul = _mm_set_epi64x(0, (long long)(1ull | (1ull << 6) | (1ull << 12) | (1ull << 18)));
    const __m128i keep = _mm_set1_epi8(0x03);
    const __m128i lut = _mm_set1_epi32(LUT_WORD);
    for (size_t w = 0; w < nwords; w++) {
        __m128i v = _mm_set1_epi64x((long long)bits[w]);
        __m128i lo = _mm_shuffle_epi8(v, lo_sel);
        __m128i hi = _mm_shuffle_epi8(v, hi_sel);
        __m128i lo0 = _mm_clmulepi64_si128(lo, mul, 0x00), lo1 = _mm_clmulepi64_si128(lo, mul, 0x0F);
        __m128i hi0 = _mm_clmulepi64_si128(hi, mul, 0x00), hi1 = _mm_clmulepi64_si128(hi, mul, 0x0F);
        __m128i a = _mm_and_si128(_mm_unpacklo_epi64(lo0, lo1), keep);
        __m128i b = _mm_and_si128(_mm_unpacklo_epi64(hi0, hi1), keep);
        _mm_storeu_si128((__m128i *)(out + (w << 5)), _mm_shuffle_epi8(lut, a));
        _mm_storeu_si128((__m128i *)(out + (w << 5) + 16), _mm_shuffle_epi8(lut, b));
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Multi-threaded drivers for the timed CPU baseline: the reference is single-threaded; the    */
/* bench shards its fastest variants by sequence offset, one pthread per core.     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int kind;                 /* 0 enc movemask, 1 dec shuffle, 2 enc lut, 3 dec lut */
    const void *in; void *out;
    size_t first_word, n_words, len_total;
} shard_t;

static void *shard_main(void *arg)
{
    shard_t *s = (shard_t *)arg;
    size_t nt0 = s->first_word << 5;
    size_t nt1 = nt0 + (s->n_words << 5);
    if (nt1 > s->len_total) nt1 = s->len_total;
    if (nt1 <= nt0) return NULL;
    switch (s->kind) {
    case 0: oracle_n_to_bits_movemask((const uint8_t *)s->in + nt0, nt1 - nt0, (uint64_t *)s->out + s->first_word); break;
    case 2: oracle_n_to_bits_lut((const uint8_t *)s->in + nt0, nt1 - nt0, (uint64_t *)s->out + s->first_word); break;
    case 1: {   /* whole words while they fit inside len; the ragged last word goes through _lut */
        size_t full = (nt1 - nt0) >> 5;
        oracle_bits_to_n_shuffle((const uint64_t *)s->in + s->first_word, full, full << 5, (uint8_t *)s->out + nt0);
        size_t done = nt0 + (full << 5);
        if (done < nt1)
            oracle_bits_to_n_lut((const uint64_t *)s->in + s->first_word + full, 1, nt1 - done, (uint8_t *)s->out + done);
        break; }
    case 3: oracle_bits_to_n_lut((const uint64_t *)s->in + s->first_word, s->n_words, nt1 - nt0, (uint8_t *)s->out + nt0); break;
    }
    return NULL;
}

static int run_sharded(int kind, const void *in, void *out, size_t len, int threads)
{
    pthread_once(&g_once, init_tables);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    size_t words = oracle_words_for_len(len);
    size_t per = (words + (size_t)threads - 1) / (size_t)threads;
    pthread_t tid[256]; shard_t sh[256];
    int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t first = (size_t)t * per;
        if (first >= words) break;
        size_t cnt = words - first < per ? words - first : per;
        sh[t] = (shard_t){ kind, in, out, first, cnt, len };
        if (threads == 1) { shard_main(&sh[t]); continue; }
        if (pthread_create(&tid[t], NULL, shard_main, &sh[t]) != 0) return 2;
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    return 0;
}

/* kind: 0 = n_to_bits_movemask, 2 = n_to_bits_lut */
CN_EXPORT int oracle_encode_mt(int kind, const uint8_t *n, size_t len, uint64_t *out, int threads)
{
    if (kind != 0 && kind != 2) return 3;
    return run_sharded(kind, n, out, len, threads);
}

/* kind: 1 = bits_to_n_shuffle, 3 = bits_to_n_lut; out needs exactly len bytes */
CN_EXPORT int oracle_decode_mt(int kind, const uint64_t *bits, size_t nwords, size_t len, uint8_t *out, int threads)
{
    if (kind != 1 && kind != 3) return 3;
    if (len > (nwords << 5)) return 1;
    return run_sharded(kind, bits, out, len, threads);
}

/* The reference's own bench shape: one call per sequence, single thread, the output allocated inside the timed
 * region (benches/bench_n_to_bits.rs:6-7, 15-19, 44-47).  Sequence i is buf[offsets[i] .. offsets[i+1]); `reps` passes over
 * all of them; encode with n_to_bits_movemask then (roundtrip != 0) decode with bits_to_n_shuffle.  Returns seconds. */
CN_EXPORT double oracle_time_small_calls(const uint8_t *buf, const uint64_t *offsets, size_t count, int reps, int roundtrip)
{
    struct timespec t0, t1;
    volatile uint64_t sink = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++) {
        for (size_t i = 0; i < count; i++) {
            const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
            const size_t words = oracle_words_for_len(len);
            uint64_t *w = malloc((words ? words : 1) * 8);
            oracle_n_to_bits_movemask(buf + offsets[i], len, w);
            if (roundtrip) {
                uint8_t *o = aligned_alloc(32, (words ? words : 1) * 32);       /* the SIMD decoder writes whole words (:271-272) */
                oracle_bits_to_n_shuffle(w, words, len, o);
                sink += o[0];
                free(o);
            }
            sink += w[0];
            free(w);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    (void)sink;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------ */
/* Deterministic synthetic input, bit-identical to cn_generate_device / cn_generate_words_device*/
/* in the CUDA library (counter-based splitmix64).               */
/* One 64-bit hash feeds 8 nucleotides (one byte lane each).                                   */
/*   alphabet 4  : "ACGT"[lane & 3]                                                            */
/*   alphabet 10 : "ACGTUacgtu"[(lane * 10) >> 8]                                              */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

CN_EXPORT int oracle_generate(uint8_t *out, size_t offset, size_t len, uint64_t seed, int alphabet)
{
    static const char a4[] = "ACGT", a10[] = "ACGTUacgtu";
    if (alphabet != 4 && alphabet != 10) return 3;
    size_t i = 0;
    while (i < len) {
        size_t g = offset + i;
        uint64_t h = mix64(seed, g >> 3);
        for (unsigned lane = (unsigned)(g & 7); lane < 8 && i < len; lane++, i++) {
            unsigned b = (unsigned)(h >> (8 * lane)) & 0xFFu;
            out[i] = (uint8_t)(alphabet == 4 ? a4[b & 3] : a10[(b * 10) >> 8]);
        }
    }
    return 0;
}

CN_EXPORT void oracle_generate_words(uint64_t *out, size_t first_word, size_t nwords, uint64_t seed)
{
    for (size_t w = 0; w < nwords; w++) out[w] = mix64(seed, first_word + w);
}

typedef struct { uint8_t *out; size_t offset, len; uint64_t seed; int alphabet; } gen_t;
static void *gen_main(void *arg) { gen_t *g = (gen_t *)arg; oracle_generate(g->out, g->offset, g->len, g->seed, g->alphabet); return NULL; }

CN_EXPORT int oracle_generate_mt(uint8_t *out, size_t offset, size_t len, uint64_t seed, int alphabet, int threads)
{
    if (alphabet != 4 && alphabet != 10) return 3;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    size_t per = ((len + (size_t)threads - 1) / (size_t)threads + 7) & ~(size_t)7;
    pthread_t tid[256]; gen_t g[256]; int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t first = (size_t)t * per;
        if (first >= len) break;
        size_t cnt = len - first < per ? len - first : per;
        g[t] = (gen_t){ out + first, offset + first, cnt, seed, alphabet };
        if (pthread_create(&tid[t], NULL, gen_main, &g[t]) != 0) return 2;
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    return 0;
}

/* What decode(encode(x)) must equal: upper case with U -> T (SURVEY 8a "Round-trip"). */
CN_EXPORT void oracle_canonical(const uint8_t *n, size_t len, uint8_t *out)
{
    pthread_once(&g_once, init_tables);
    for (size_t i = 0; i < len; i++) out[i] = g_bits_lut[g_byte_lut[n[i]]];
}

/* 1 if the CPU has what the SIMD restatements need */
CN_EXPORT int oracle_cpu_ok(void)
{
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("pclmul");
}

/* ========================================================================================== */
/* Base-5 codec: {A, C, T/U, G, N} -> digits 0..4, three digits -> 7 bits (a + 5b + 25c),      */
/* nine triplets (27 nucleotides) per u64.  Restates src/n_to_bits2.rs.                        */
/* ========================================================================================== */
static uint8_t g_byte_lut2[256];
static const uint8_t g_bits_lut2[8] = { 'A', 'C', 'T', 'G', 'N', 0, 0, 0 };   /* src/n_to_bits2.rs:25-33 (5 entries) */
static pthread_once_t g_once2 = PTHREAD_ONCE_INIT;

static void init_tables2(void)
{   /* src/n_to_bits2.rs:8-23 */
    memset(g_byte_lut2, 0, sizeof g_byte_lut2);
    g_byte_lut2['a'] = 0; g_byte_lut2['A'] = 0;
    g_byte_lut2['c'] = 1; g_byte_lut2['C'] = 1;
    g_byte_lut2['t'] = 2; g_byte_lut2['T'] = 2;
    g_byte_lut2['u'] = 2; g_byte_lut2['U'] = 2;
    g_byte_lut2['g'] = 3; g_byte_lut2['G'] = 3;
    g_byte_lut2['n'] = 4; g_byte_lut2['N'] = 4;
}

/* ------------------------------------------------------------------------------------------ */
/* Operations on packed words.  PARITY UNPINNED: the reference holds no code for   */
/* them (README.md:21-23, 45, 415-418 only motivate them), so they are DEFINED through the        */
/* reference's own codec, via ASCII: decode with bits_to_n_lut, operate on letters, re-encode     */
/* with n_to_bits_lut.  Returns 1 where bits_to_n_lut would panic.                                */
/* ------------------------------------------------------------------------------------------ */
CN_EXPORT int oracle_hamming(const uint64_t *a, const uint64_t *b, size_t nwords, size_t len, uint64_t *result)
{
    uint8_t *x = malloc(len ? len : 1), *y = malloc(len ? len : 1);
    int rc = oracle_bits_to_n_lut(a, nwords, len, x);
    if (rc == 0) rc = oracle_bits_to_n_lut(b, nwords, len, y);
    uint64_t d = 0;
    if (rc == 0) for (size_t i = 0; i < len; i++) d += x[i] != y[i];
    *result = d;
    free(x); free(y);
    return rc;
}

CN_EXPORT int oracle_complement(const uint64_t *bits, size_t nwords, size_t len, uint64_t *out, int reverse)
{
    uint8_t *x = malloc(len ? len : 1), *y = malloc(len ? len : 1);
    int rc = oracle_bits_to_n_lut(bits, nwords, len, x);
    if (rc == 0) {
        for (size_t i = 0; i < len; i++) {
            uint8_t c = x[i];
            c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : 'C';
            y[reverse ? len - 1 - i : i] = c;
        }
        oracle_n_to_bits_lut(y, len, out);
    }
    free(x); free(y);
    return rc;
}

CN_EXPORT size_t oracle_count_invalid2(const uint8_t *n, size_t len)
{   /* bytes outside {ACGTUNacgtun}: n_to_bits2_lut maps them to digit 0 (:8-23), n_to_bits2_pext by their low 3 bits (:127-136) */
    size_t bad = 0;
    for (size_t i = 0; i < len; i++) {
        switch (n[i]) {
        case 'A': case 'C': case 'G': case 'T': case 'U': case 'N':
        case 'a': case 'c': case 'g': case 't': case 'u': case 'n': break;
        default: bad++;
        }
    }
    return bad;
}

CN_EXPORT size_t oracle_words2_for_len(size_t len)
{   /* src/n_to_bits2.rs:38 */
    return len / 27 + ((len % 27) ? 1 : 0);
}

/* src/n_to_bits2.rs:37-74 n_to_bits2_lut */
CN_EXPORT void oracle_n_to_bits2_lut(const uint8_t *n, size_t len, uint64_t *out)
{
    pthread_once(&g_once2, init_tables2);
    size_t words = oracle_words2_for_len(len);
    for (size_t w = 0; w < words; w++) out[w] = 0;
    size_t triplets = len / 3;
    for (size_t t = 0; t < triplets; t++) {
        const uint8_t *p = n + 3 * t;
        uint64_t enc = (uint64_t)g_byte_lut2[p[0]] + 5u * g_byte_lut2[p[1]] + 25u * g_byte_lut2[p[2]];
        out[t / 9] |= enc << ((t % 9) * 7);
    }
    size_t left = len % 3;                          /* :59-70 ragged last triplet */
    if (left) {
        const uint8_t *p = n + 3 * triplets;
        uint64_t enc = g_byte_lut2[p[0]];
        if (left >= 2) enc += 5u * g_byte_lut2[p[1]];
        out[triplets / 9] |= enc << ((triplets % 9) * 7);
    }
}

/* src/n_to_bits2.rs:78-107 bits_to_n2_lut.  Decodes whole triplets, so `out` needs
 * 3*ceil(len/3) bytes (the reference allocates 27*nwords, :87); the caller keeps `len`.
 * Returns 1 where the reference panics (:79-81).  A 7-bit field above 124 cannot come from the
 * encoder; the reference would index BITS_LUT out of bounds there, the oracle yields 0 bytes. */
CN_EXPORT int oracle_bits_to_n2_lut(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > nwords * 27) return 1;
    size_t triplets = len / 3 + ((len % 3) ? 1 : 0);
    for (size_t t = 0; t < triplets; t++) {
        unsigned cur = (unsigned)(bits[t / 9] >> ((t % 9) * 7)) & 0x7Fu;
        out[3 * t] = g_bits_lut2[cur % 5];
        out[3 * t + 1] = g_bits_lut2[(cur / 5) % 5];
        out[3 * t + 2] = g_bits_lut2[(cur / 25) & 7];
    }
    return 0;
}

/* src/n_to_bits2.rs:118-189 n_to_bits2_pext: pshufb LUT on the low 3 ASCII bits (:127-136),
 * permutevar8x32 to make both halves lane-local, two pshufb to split a / (b,c), maddubs by (5,25),
 * add, pext(0x007F..) to squeeze 16-bit lanes to 7 bits.  Vector body covers (len-5)/27 words
 * (:120, the 32-byte load must stay inside the slice), the rest goes through n_to_bits2_lut. */
__attribute__((target("avx2,bmi2")))
CN_EXPORT void oracle_n_to_bits2_pext(const uint8_t *n, size_t len, uint64_t *out)
{
    size_t body = len < 5 ? 0 : (len - 5) / 27;
    size_t words = oracle_words2_for_len(len);
    long long lutq = 0;
    lutq |= 0ll << ((('A') & 7) << 3);
    lutq |= 1ll << ((('C') & 7) << 3);
    lutq |= 2ll << ((('T') & 7) << 3);
    lutq |= 2ll << ((('U') & 7) << 3);
    lutq |= 3ll << ((('G') & 7) << 3);
    lutq |= 4ll << ((('N') & 7) << 3);
    const __m256i lut = _mm256_set1_epi64x(lutq);
    const __m256i bridge = _mm256_set_epi32(6, 5, 4, 3, 3, 2, 1, 0);
    const __m256i first = _mm256_set_epi16(-1, -1, -1, -1, (short)0xFF1C, (short)0xFF19, (short)0xFF16, (short)0xFF13,
                                           -1, -1, -1, (short)0xFF0C, (short)0xFF09, (short)0xFF06, (short)0xFF03, (short)0xFF00);
    const __m256i rest = _mm256_set_epi16(-1, -1, -1, -1, 0x1E1D, 0x1B1A, 0x1817, 0x1514,
                                          -1, -1, -1, 0x0E0D, 0x0B0A, 0x0807, 0x0504, 0x0201);
    const __m256i w5_25 = _mm256_set1_epi16(0x1905);
    const uint64_t squeeze = 0x007F007F007F007Full;
    const uint8_t *p = n;
    for (size_t w = 0; w < body; w++, p += 27) {
        __m256i v = _mm256_shuffle_epi8(lut, _mm256_loadu_si256((const __m256i *)p));
        v = _mm256_permutevar8x32_epi32(v, bridge);
        __m256i a = _mm256_shuffle_epi8(v, first);
        __m256i bc = _mm256_maddubs_epi16(_mm256_shuffle_epi8(v, rest), w5_25);
        __m256i e = _mm256_add_epi16(a, bc);
        uint64_t q0 = (uint64_t)_mm256_extract_epi64(e, 0), q1 = (uint64_t)_mm256_extract_epi64(e, 1),
                 q2 = (uint64_t)_mm256_extract_epi64(e, 2);
        out[w] = _pext_u64(q0, squeeze) | (q1 << 28) | (_pext_u64(q2, squeeze) << 35);
    }
    if (body < words) oracle_n_to_bits2_lut(n + body * 27, len - body * 27, out + body);
}

/* src/n_to_bits2.rs:196-268 bits_to_n2_pdep: pdep(0x7F7F..) pads 7-bit fields to bytes, pshufb widens
 * to 16 bits, division / remainder by 5 and 25 through 16-bit reciprocal multiplies (:209-211,241-247),
 * three pshufb + or to interleave a,b,c, permutevar8x32 to close the lane gap, pshufb(lut).
 * Stores 32 bytes per word at stride 27: `out` needs 27*nwords + 5 bytes (:201). */
__attribute__((target("avx2,bmi2")))
CN_EXPORT int oracle_bits_to_n2_pdep(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out)
{
    if (len > nwords * 27) return 1;
    const uint64_t pad = 0x7F7F7F7F7F7F7F7Full;
    const __m256i widen = _mm256_set_epi16(-1, -1, -1, (short)0xFF04, (short)0xFF03, (short)0xFF02, (short)0xFF01, (short)0xFF00,
                                           -1, -1, -1, -1, (short)0xFF03, (short)0xFF02, (short)0xFF01, (short)0xFF00);
    const __m256i five = _mm256_set1_epi16(5);
    const __m256i inv5 = _mm256_set1_epi16((short)((1u << 16) / 5 + 1));
    const __m256i inv25 = _mm256_set1_epi16((short)((1u << 16) / 25 + 1));
    const __m256i sel_a = _mm256_set_epi64x((long long)0xFFFFFF08FFFF06FFull, (long long)0xFF04FFFF02FFFF00ull,
                                            (long long)0xFFFFFF08FFFF06FFull, (long long)0xFF04FFFF02FFFF00ull);
    const __m256i sel_b = _mm256_set_epi64x((long long)0xFFFF08FFFF06FFFFull, (long long)0x04FFFF02FFFF00FFull,
                                            (long long)0xFFFF08FFFF06FFFFull, (long long)0x04FFFF02FFFF00FFull);
    const __m256i sel_c = _mm256_set_epi64x((long long)0xFF08FFFF06FFFF04ull, (long long)0xFFFF02FFFF00FFFFull,
                                            (long long)0xFF08FFFF06FFFF04ull, (long long)0xFFFF02FFFF00FFFFull);
    const __m256i close_gap = _mm256_set_epi32(7, 7, 6, 5, 4, 2, 1, 0);
    const __m256i lut = _mm256_set1_epi64x((long long)('A' | ('C' << 8) | ('T' << 16) | ((uint64_t)'G' << 24) | ((uint64_t)'N' << 32)));
    uint8_t *p = out;
    for (size_t w = 0; w < nwords; w++, p += 27) {
        int64_t cur = (int64_t)bits[w];
        int64_t lo = (int64_t)_pdep_u64((uint64_t)cur, pad);
        int64_t hi = ((cur >> 56) << 32) | (lo >> 32);
        __m256i v = _mm256_shuffle_epi8(_mm256_set_epi64x(0, hi, 0, lo), widen);
        __m256i f5 = _mm256_mullo_epi16(v, inv5), f25 = _mm256_mullo_epi16(v, inv25);
        __m256i a = _mm256_shuffle_epi8(_mm256_mulhi_epu16(f5, five), sel_a);
        __m256i b = _mm256_shuffle_epi8(_mm256_mulhi_epu16(f25, five), sel_b);
        __m256i c = _mm256_shuffle_epi8(_mm256_mulhi_epu16(v, inv25), sel_c);
        __m256i abc = _mm256_permutevar8x32_epi32(_mm256_or_si256(_mm256_or_si256(a, b), c), close_gap);
        _mm256_storeu_si256((__m256i *)p, _mm256_shuffle_epi8(lut, abc));
    }
    return 0;
}

/* synthetic 5-letter data: "ACGTN"[(lane_byte * 5) >> 8] (alphabet 5), "ACGTUNacgtun" (alphabet 12) */
CN_EXPORT int oracle_generate2(uint8_t *out, size_t offset, size_t len, uint64_t seed, int alphabet)
{
    static const char a5[] = "ACGTN", a12[] = "ACGTUNacgtun";
    if (alphabet != 5 && alphabet != 12) return 3;
    size_t i = 0;
    while (i < len) {
        size_t g = offset + i;
        uint64_t h = mix64(seed, g >> 3);
        for (unsigned lane = (unsigned)(g & 7); lane < 8 && i < len; lane++, i++) {
            unsigned b = (unsigned)(h >> (8 * lane)) & 0xFFu;
            out[i] = (uint8_t)(alphabet == 5 ? a5[(b * 5) >> 8] : a12[(b * 12) >> 8]);
        }
    }
    return 0;
}

/* what bits_to_n2(n_to_bits2(x)) must equal: upper case, U -> T */
CN_EXPORT void oracle_canonical2(const uint8_t *n, size_t len, uint8_t *out)
{
    pthread_once(&g_once2, init_tables2);
    for (size_t i = 0; i < len; i++) out[i] = g_bits_lut2[g_byte_lut2[n[i]]];
}
