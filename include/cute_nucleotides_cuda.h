/*
 * cute_nucleotides_cuda.h -- C ABI of the B200 (sm_100a) implementation of cute-nucleotides' 2-bit
 * nucleotide codec.  This is the drop-in boundary: the entry points are exactly what a Rust `extern "C"`
 * block (rust/src/ffi.rs) binds to put `n_to_bits_cuda` / `bits_to_n_cuda` next
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
 *   - bytes outside {A,C,G,T,U,a,c,g,t,u} are mapped by (byte >> 1) & 3 like the reference's SIMD
 *     encoders (pext/shift/movemask/mul); the reference's variants disagree with each other on such
 *     bytes, so parity is defined on the valid alphabet only.
 *
 * Conventions: every function returns CN_OK (0) or a CN_ERR_* code and never throws or aborts;
 * cn_last_error() returns a thread-local description of the last failure on the calling thread.
 * Outputs are ALWAYS caller-allocated: no memory allocated here is ever handed to the caller to free
 * with a foreign allocator (Rust's Vec owns what Rust allocated).
 * All entry points are thread-safe; the host-slice calls keep per-thread staging state (released when the thread
 * exits).  Input and output buffers of one call must not overlap (the reference's functions cannot alias either: they
 * return a fresh Vec); decoding back over the buffer that was ENCODED earlier is fine (separate calls).
 * The cn_set_* knobs are process-wide harness switches: set them before issuing calls, not concurrently with them.
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

#define CN_ABI_VERSION 1

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

/* ---- device-resident entry points (the roofline path; buffers already in HBM) ------------------ */
/* `stream` is a cudaStream_t (NULL = default stream); launches are asynchronous on it, on the current
 * device.  d_n may have any alignment (16-byte aligned input takes the fast path); d_out must be 8-byte
 * aligned.  Pinned host pointers are accepted too (zero-copy over PCIe). */
CN_API int cn_encode_device(const void *d_n, size_t len, void *d_out, void *stream);
/* d_bits 8-byte aligned; d_out any alignment (16-byte aligned takes the fast path). */
CN_API int cn_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream);

/* ---- multi-GPU: encode + assemble in ONE kernel over NVLink peer memory (one process per GPU) ---------------- */
/* Reads the shard once and stores every packed word to `nout` (1..8) destinations: the caller's own assembled buffer
 * and peer buffers mapped with cn_ipc_open.  Each d_outs[k] is a device pointer ALREADY offset to this shard's first
 * word.  d_n must be 32-byte aligned.  Replaces cn_encode_device + an all-gather of the packed shards. */
CN_API int cn_encode_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, void *stream);
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

/* ---- base-5 codec: {A,C,T/U,G,N} -> digits 0..4, 3 digits -> 7 bits, 27 nucleotides per u64 ------------ */
/* (len / 27) + (len % 27 != 0): words produced by n_to_bits2_* (src/n_to_bits2.rs:38, :121). */
CN_API size_t cn_words2_for_len(size_t len);
/* n_to_bits2_{lut,pext}(n) -- src/n_to_bits2.rs:37, 118.  `out` receives cn_words2_for_len(len) words; the digit
 * of a byte is looked up by its low 3 bits like the reference's pshufb LUT (:127-136). */
CN_API int cn_n_to_bits2_host(const uint8_t *n, size_t len, uint64_t *out);
/* bits_to_n2_{lut,pdep}(bits, len) -- src/n_to_bits2.rs:78, 196.  CN_ERR_LENGTH when len > 27 * nwords (:79-81). */
CN_API int cn_bits_to_n2_host(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out);
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

#ifdef __cplusplus
}
#endif
#endif /* CUTE_NUCLEOTIDES_CUDA_H */
