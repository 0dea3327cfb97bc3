/*
 * cute_nucleotides_cuda.h -- C ABI of the B200 (sm_100a) implementation of cute-nucleotides' 2-bit
 * nucleotide codec.  This is the drop-in boundary: the entry points are exactly what a Rust `extern "C"`
 * block (`mod ffi` in rust/src/lib.rs) binds to put `n_to_bits_cuda` / `bits_to_n_cuda` next
 * to the reference's `n_to_bits_{lut,pext,shift,movemask,mul}` / `bits_to_n_{lut,shuffle,pdep,clmul}`.
 *
 * Reference interface replaced (paths relative to the reference repo):
 *   pub fn n_to_bits_*(n: &[u8]) -> Vec<u64>                src/n_to_bits.rs:34, 80, 121, 172, 213
 *   pub fn bits_to_n_*(bits: &[u64], len: usize) -> Vec<u8> src/n_to_bits.rs:51, 265, 309, 346
 *   pub fn n_to_bits2_*(n: &[u8]) -> Vec<u64>               src/n_to_bits2.rs:37, 118
 *   pub fn bits_to_n2_*(bits: &[u64], len: usize) -> Vec<u8> src/n_to_bits2.rs:78, 196
 *
 * Data contract (identical to n_to_bits_lut / bits_to_n_lut, src/n_to_bits.rs:34-69):
 *   - nucleotide i occupies bits 2*(i&31) .. 2*(i&31)+1 of 64-bit word i>>5
 *   - codes: A/a = 0, C/c = 1, T/t/U/u = 2, G/g = 3 (case-insensitive)
 *   - ceil(len/32) words are produced; unused high bits of the last word are zero
 *   - decoding always yields upper-case 'A','C','T','G'
 *   - bytes outside {A,C,G,T,U,a,c,g,t,u}: the reference's variants disagree with each other.  The plain entry
 *     points map them by (byte >> 1) & 3 like its SIMD encoders (pext/shift/movemask/mul); mode
 *     CN_ENC_LUT_EXACT of the *_ex_* entry points maps them to 0 exactly like n_to_bits_lut's BYTE_LUT
 *     (src/n_to_bits.rs:8-21, :42), so either behaviour of the reference is reproducible bit for bit.
 *
 * Conventions: every function returns CN_OK (0) or a CN_ERR_* code and never throws or aborts;
 * cn_last_error() returns a thread-local description of the last failure on the calling thread.
 * Outputs are ALWAYS caller-allocated: no memory allocated here is ever handed to the caller to free
 * with a foreign allocator (Rust's Vec owns what Rust allocated).
 * All entry points are thread-safe; the host-slice calls keep per-thread staging state (released when the thread
 * exits).  Input and output buffers of one call must not overlap (the reference's functions cannot alias either: they
 * return a fresh Vec); decoding back over the buffer that was ENCODED earlier is fine (separate calls).
 * The cn_set_* knobs are process-wide; each is a single atomic word, so a call racing with a setter sees the old or
 * the new value, never a mix (a call snapshots its knobs once).
 * No entry point reads memory outside the buffers it is given (unaligned buffers included).
 */
#ifndef CUTE_NUCLEOTIDES_CUDA_H
#define CUTE_NUCLEOTIDES_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CN_OK          0
#define CN_ERR_LENGTH  1   /* len > 32 * nwords: where the reference panics (src/n_to_bits.rs:52-54) */
#define CN_ERR_CUDA    2   /* a CUDA runtime call failed; see cn_last_error() */
#define CN_ERR_ARG     3   /* null / misaligned pointer or bad enum */
#define CN_ERR_NOMEM   4   /* staging allocation failed */

#define CN_ABI_VERSION 2

#if defined(__GNUC__)
#define CN_API __attribute__((visibility("default")))
#else
#define CN_API
#endif

/* ---- library ---------------------------------------------------------------------------------- */
CN_API int cn_abi_version(void);
/* Select the CUDA device used by the host-slice calls of the calling thread (default: current). */
CN_API int cn_init(int device);
CN_API int cn_device_count(int *count);
CN_API const char *cn_last_error(void);
/* The reference's panic text for CN_ERR_LENGTH, so a shim can re-raise it verbatim. */
CN_API const char *cn_length_panic_message(void);
/* ceil(len / 32): words produced by the encoders (src/n_to_bits.rs:35, :83-84). */
CN_API size_t cn_words_for_len(size_t len);
/* Number of codec kernels this library has launched in this process (bench.py "gpu_launches"). */
CN_API uint64_t cn_launch_count(void);

/* ---- host-slice entry points: the drop-in for the Rust signatures ------------------------------ */
/* n_to_bits_*(n) -- src/n_to_bits.rs:34/80/121/172/213.  `n` is `len` ASCII bytes in host memory (any
 * alignment, pageable or pinned); `out` receives cn_words_for_len(len) words.  len == 0 is a no-op. */
CN_API int cn_n_to_bits_host(const uint8_t *n, size_t len, uint64_t *out);
/* bits_to_n_*(bits, len) -- src/n_to_bits.rs:51/265/309/346.  Writes exactly `len` bytes to `out`.
 * Returns CN_ERR_LENGTH when len > 32 * nwords. */
CN_API int cn_bits_to_n_host(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out);

/* Many independent sequences in ONE call.  The reference's own bench shape (40 000 nt, benches/bench_n_to_bits.rs:10,
 * 38-39) is launch-latency-bound on a GPU one call at a time; here sequence i -- seqs[i][0 .. lens[i]) -- is encoded into
 * outs[i][0 .. cn_words_for_len(lens[i])) exactly as cn_n_to_bits_host would, but thousands of sequences share one kernel
 * launch and one pair of PCIe transfers.  The inverse decodes lens[i] nucleotides from bits[i] (which must hold
 * cn_words_for_len(lens[i]) words) into outs[i].  Zero-length sequences are skipped (their pointers may be NULL).  With a
 * device set (cn_set_devices) the batch is cut into runs of consecutive sequences, one run per device. */
CN_API int cn_n_to_bits_host_batch(const uint8_t *const *seqs, const size_t *lens, size_t count, uint64_t *const *outs);
CN_API int cn_bits_to_n_host_batch(const uint64_t *const *bits, const size_t *lens, size_t count, uint8_t *const *outs);

/* Asynchronous flavours: the call runs on a library thread and `*req` receives a handle; cn_wait(req) blocks until it is
 * done, returns ITS status (cn_last_error() then describes its failure) and releases the handle -- every request must be
 * waited for exactly once, and the buffers must stay valid and untouched until then.  One encode and one decode in flight
 * load the two PCIe directions at once (33 instead of 26.5 Gnt/s for a stream of round trips on one link). */
typedef struct cn_request cn_request;
CN_API int cn_n_to_bits_host_async(const uint8_t *n, size_t len, uint64_t *out, cn_request **req);
CN_API int cn_bits_to_n_host_async(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out, cn_request **req);
CN_API int cn_wait(cn_request *req);

/* ---- device-resident entry points (the roofline path; buffers already in HBM) ------------------ */
/* `stream` is a cudaStream_t (NULL = default stream); launches are asynchronous on it, on the current
 * device.  d_n may have any alignment (16-byte aligned input takes the fast path); d_out must be 8-byte
 * aligned.  Pinned host pointers are accepted too (zero-copy over PCIe). */
CN_API int cn_encode_device(const void *d_n, size_t len, void *d_out, void *stream);
/* d_bits 8-byte aligned; d_out any alignment (16-byte aligned takes the fast path). */
CN_API int cn_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream);

/* Many independent sequences that already live in HBM, tightly concatenated: sequence i = bytes [offsets[i], offsets[i+1])
 * of d_n and words [word_offsets[i], word_offsets[i+1]) of the packed side, word_offsets[i+1] - word_offsets[i] =
 * cn_words_for_len(len_i) -- per sequence exactly what n_to_bits_lut / bits_to_n_lut produce (src/n_to_bits.rs:34-69).
 * d_offsets / d_word_offsets: count + 1 non-decreasing uint64 in DEVICE memory, first element 0; total_len =
 * offsets[count], total_words = word_offsets[count] (the caller knows them; the library does not read device memory back).
 * One launch, any alignment of d_n / d_out; the device-resident counterpart of the *_host_batch calls. */
CN_API int cn_encode_segmented_device(const void *d_n, const void *d_offsets, const void *d_word_offsets, size_t count,
                                      size_t total_len, size_t total_words, void *d_out, void *stream);
CN_API int cn_decode_segmented_device(const void *d_bits, const void *d_offsets, const void *d_word_offsets, size_t count,
                                      size_t total_len, size_t total_words, void *d_out, void *stream);

/* ---- multi-GPU: encode + assemble in ONE kernel over NVLink peer memory (one process per GPU) ---------------- */
/* Reads the shard once and stores every packed word to `nout` (1..8) destinations: the caller's own assembled buffer
 * and peer buffers mapped with cn_ipc_open.  Each d_outs[k] is a device pointer ALREADY offset to this shard's first
 * word.  d_n must be 32-byte aligned.  Replaces cn_encode_device + an all-gather of the packed shards. */
CN_API int cn_encode_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, void *stream);
/* The same for the base-5 codec (cn_encode2_device + all-gather / gather-to-root in one kernel): d_n 16-byte aligned,
 * every d_outs[k] 32-byte aligned and already offset to this shard's first word (shard starts that are multiples of
 * 108 nucleotides = 4 words keep it so). */
CN_API int cn_encode2_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, void *stream);
/* CUDA IPC plumbing for the above: export a device pointer as a 64-byte handle + offset inside its allocation
 * (send them to the peer process), map it there, unmap it. */
CN_API int cn_ipc_export(void *d_ptr, void *handle64, size_t *offset);
CN_API int cn_ipc_open(const void *handle64, size_t offset, void **d_ptr);
CN_API int cn_ipc_close(void *d_ptr, size_t offset);

/* ---- encode + validation in one pass ---------------------------------------------------------- */
/* The reference maps bytes outside {A,C,G,T,U,a,c,g,t,u} silently (LUT -> 0, SIMD -> (b>>1)&3; its README points to an
 * external validity checker).  These entry points produce exactly the same packed words as cn_n_to_bits_host /
 * cn_encode_device and additionally count the offending bytes, at no extra HBM traffic.
 * Device flavour: *d_invalid_count (8-byte aligned uint64 in device memory) is INCREMENTED; zero it first. */
CN_API int cn_n_to_bits_checked_host(const uint8_t *n, size_t len, uint64_t *out, uint64_t *invalid_count);
CN_API int cn_encode_checked_device(const void *d_n, size_t len, void *d_out, void *d_invalid_count, void *stream);

/* ---- encode with a choice of semantics for bytes outside the alphabet --------------------------------------- */
#define CN_ENC_PLAIN      0   /* (byte >> 1) & 3 for every byte: n_to_bits_{pext,shift,movemask,mul} (src/n_to_bits.rs:85,100) */
#define CN_ENC_COUNT      1   /* same words as PLAIN + count of the bytes outside the alphabet (= the *_checked_* entry points) */
#define CN_ENC_LUT_EXACT  2   /* bytes outside the alphabet (and bytes >= 0x80, an out-of-bounds LUT read in the reference)
                                 encode as 0: bit-exact n_to_bits_lut (BYTE_LUT, src/n_to_bits.rs:8-21, :42) on ALL inputs;
                                 the offending bytes are counted too when a counter is given */
/* invalid_count may be NULL except in mode CN_ENC_COUNT.  Device flavour: *d_invalid_count is INCREMENTED. */
CN_API int cn_n_to_bits_ex_host(const uint8_t *n, size_t len, uint64_t *out, int mode, uint64_t *invalid_count);
CN_API int cn_encode_ex_device(const void *d_n, size_t len, void *d_out, int mode, void *d_invalid_count, void *stream);
/* The same three modes for the base-5 codec; its alphabet is {A,C,G,T,U,N,a,c,g,t,u,n}.  PLAIN looks digits up by the low
 * 3 bits like n_to_bits2_pext's pshufb LUT (src/n_to_bits2.rs:127-136); LUT_EXACT maps everything unlisted to digit 0 like
 * n_to_bits2_lut's BYTE_LUT (src/n_to_bits2.rs:8-23). */
CN_API int cn_n_to_bits2_ex_host(const uint8_t *n, size_t len, uint64_t *out, int mode, uint64_t *invalid_count);
CN_API int cn_encode2_ex_device(const void *d_n, size_t len, void *d_out, int mode, void *d_invalid_count, void *stream);

/* ---- multi-GPU inside ONE process (SURVEY 8b cn_encode_sharded, 8e) ------------------------------------------ */
/* Devices the HOST-SLICE calls fan out over.  With count >= 2, one cn_n_to_bits_host / cn_bits_to_n_host (and _ex_, base-5)
 * call is cut by sequence offset into word-aligned ranges (word w depends only on nucleotides 32w..32w+31,
 * src/n_to_bits.rs:39-42) and every range runs on its own GPU -- count PCIe links instead of one -- driven by a worker
 * thread per device inside the library; the result is identical.  (Calls on PAGEABLE buffers are bounded by the CPU
 * staging copies, not by the links, and use at most CN_FANOUT_PAGEABLE_MAX = 2 of the devices.)  count == 0 restores the
 * single-device behaviour
 * (the device chosen with cn_init, else the current one).  Environment: CN_DEVICES=all or CN_DEVICES=0,1,2 does the same
 * for an unmodified caller. */
CN_API int cn_set_devices(const int *devices, int count);
CN_API int cn_get_devices(int *devices, int capacity, int *count);
/* The shard planner: range [*start, *end) of shard `shard` of `nshards` over total_len nucleotides; boundaries are
 * multiples of `granule` (a multiple of 32 for the 2-bit codec, of 27 for base-5), the ragged tail goes to the last shard. */
CN_API int cn_shard_bounds(size_t total_len, int nshards, int shard, size_t granule, size_t *start, size_t *end);
/* Device-resident shards, one per device, launched from the calling thread (asynchronous; streams may be NULL = every
 * device's default stream, else streams[k] is a cudaStream_t of devices[k]).  Shard k: lens[k] nucleotides at d_n[k] ->
 * cn_words_for_len(lens[k]) words at d_out[k].  After cn_enable_peer_access a d_out[k] (encode) or d_bits[k] (decode) may
 * point into ANOTHER device's memory: the stores / loads then cross NVLink inside the kernel, i.e. encode + gather-to-root
 * and scatter + decode are single kernels with no separate collective. */
CN_API int cn_encode_sharded(int nshards, const int *devices, const void *const *d_n, const size_t *lens,
                             void *const *d_out, void *const *streams);
CN_API int cn_decode_sharded(int nshards, const int *devices, const void *const *d_bits, const size_t *nwords,
                             const size_t *lens, void *const *d_out, void *const *streams);
CN_API int cn_enable_peer_access(const int *devices, int count);
CN_API int cn_synchronize_devices(const int *devices, int count);

/* ---- operations on packed 2-bit words ---------------------------------------------------------- */
/* PARITY UNPINNED: the reference motivates these (README.md:21-23, 45, 415-418) but holds no code for them.  They are
 * defined through the reference's own codec: hamming = number of i < len with bits_to_n_lut(a)[i] != bits_to_n_lut(b)[i];
 * complement = n_to_bits_lut of the decoded sequence with A<->T, C<->G; reverse_complement = the same, reversed.  Only the
 * first len nucleotides (ceil(len/32) words) are read; unused high bits of the last output word are zero;
 * CN_ERR_LENGTH when len > 32 * nwords.  Words 8-byte aligned (32-byte aligned takes the fast path).
 * *d_result (device uint64) is INCREMENTED: zero it first.  reverse_complement cannot run in place. */
CN_API int cn_hamming_device(const void *d_a, const void *d_b, size_t nwords, size_t len, void *d_result, void *stream);
CN_API int cn_complement_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream);
CN_API int cn_reverse_complement_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream);
CN_API int cn_hamming_host(const uint64_t *a, const uint64_t *b, size_t nwords, size_t len, uint64_t *result);
CN_API int cn_complement_host(const uint64_t *bits, size_t nwords, size_t len, uint64_t *out);
CN_API int cn_reverse_complement_host(const uint64_t *bits, size_t nwords, size_t len, uint64_t *out);

/* ---- base-5 codec: {A,C,T/U,G,N} -> digits 0..4, 3 digits -> 7 bits, 27 nucleotides per u64 ------------ */
/* (len / 27) + (len % 27 != 0): words produced by n_to_bits2_* (src/n_to_bits2.rs:38, :121). */
CN_API size_t cn_words2_for_len(size_t len);
/* n_to_bits2_{lut,pext}(n) -- src/n_to_bits2.rs:37, 118.  `out` receives cn_words2_for_len(len) words; the digit
 * of a byte is looked up by its low 3 bits like the reference's pshufb LUT (:127-136). */
CN_API int cn_n_to_bits2_host(const uint8_t *n, size_t len, uint64_t *out);
/* bits_to_n2_{lut,pdep}(bits, len) -- src/n_to_bits2.rs:78, 196.  CN_ERR_LENGTH when len > 27 * nwords (:79-81). */
CN_API int cn_bits_to_n2_host(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out);
/* Many independent sequences in one call, like cn_n_to_bits_host_batch: element i == cn_n_to_bits2_host on sequence i
 * (cn_words2_for_len(lens[i]) words each). */
CN_API int cn_n_to_bits2_host_batch(const uint8_t *const *seqs, const size_t *lens, size_t count, uint64_t *const *outs);
CN_API int cn_bits_to_n2_host_batch(const uint64_t *const *bits, const size_t *lens, size_t count, uint8_t *const *outs);
/* Device-resident pair; the ASCII side takes the tiled fast path when 16-byte aligned. */
CN_API int cn_encode2_device(const void *d_n, size_t len, void *d_out, void *stream);
CN_API int cn_decode2_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream);
/* Synthetic 5-letter data: alphabet 5 = {A,C,G,T,N}, alphabet 12 = {A,C,G,T,U,N,a,c,g,t,u,n}. */
CN_API int cn_generate2_device(void *d_out, size_t offset, size_t len, uint64_t seed, int alphabet, void *stream);

/* ---- harness helpers (tests, bench, the C++/Rust harnesses; not needed by a plain caller) ------- */
/* Deterministic synthetic data, bit-identical to the host generator used by the tests:
 * alphabet 4 = uniform {A,C,G,T}; alphabet 10 = {A,C,G,T,U,a,c,g,t,u}.  `offset` (multiple of 8) is the
 * global index of d_out[0], so shards of one logical sequence can be generated independently. */
CN_API int cn_generate_device(void *d_out, size_t offset, size_t len, uint64_t seed, int alphabet, void *stream);
CN_API int cn_generate_words_device(void *d_out, size_t first_word, size_t nwords, uint64_t seed, void *stream);

CN_API int cn_device_malloc(void **d_ptr, size_t bytes);
CN_API int cn_device_free(void *d_ptr);
CN_API int cn_host_malloc_pinned(void **h_ptr, size_t bytes);
CN_API int cn_host_free_pinned(void *h_ptr);
CN_API int cn_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
CN_API int cn_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
CN_API int cn_stream_synchronize(void *stream);
/* Launch the codec `iters` times back to back on an internal stream, bracketed by CUDA events recorded
 * on that same stream; *ms_total receives the elapsed device time of all iterations. */
CN_API int cn_time_encode_device(const void *d_n, size_t len, void *d_out, int iters, float *ms_total);
CN_API int cn_time_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, int iters, float *ms_total);

/* ---- tuning knobs (defaults are the measured best on B200) ------------------------ */
#define CN_DIR_ENCODE 0
#define CN_DIR_DECODE 1
/* vec: ASCII bytes per thread per memory instruction (16 or 32); unroll: 1,2,4,8; threads: 64,128,256,512 */
CN_API int cn_set_tuning(int direction, int vec, int unroll, int threads);
CN_API int cn_get_tuning(int direction, int *vec, int *unroll, int *threads);
/* Host-slice strategy: 0 = staged (pinned ring + copy engines, chunked and overlapped),
 * 1 = zero-copy (kernel reads/writes pinned host memory directly over PCIe).  chunk_bytes = ASCII bytes
 * per pipeline chunk (0 keeps the current value). */
CN_API int cn_set_host_strategy(int strategy, size_t chunk_bytes);
/* Chunk sizes of the staged pipeline (0 keeps a value): pinned caller buffers are DMA'd in pinned_chunk pieces, pageable
 * ones staged in pageable_chunk pieces.  cn_set_host_strategy's chunk_bytes sets both. */
CN_API int cn_set_host_chunks(size_t pinned_chunk, size_t pageable_chunk);
/* Threads that copy pageable caller memory to / from pinned staging (the waiting caller included).  0 = default:
 * CN_HOST_THREADS, else min(12, 5/8 of the cores).  The pool only grows; a lower value takes effect for new pools only. */
CN_API int cn_set_host_threads(int threads);

#ifdef __cplusplus
}
#endif
#endif /* CUTE_NUCLEOTIDES_CUDA_H */
