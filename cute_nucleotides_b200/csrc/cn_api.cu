// cn_api.cu -- the extern "C" boundary declared in include/cute_nucleotides_cuda.h.
//
// Two families of entry points:
//   * device-resident (cn_encode_device / cn_decode_device, cn_launch.cuh): one kernel launch on the caller's
//     stream; this is the roofline path (BASELINE configs 2-5).
//   * host-slice (cn_n_to_bits_host / cn_bits_to_n_host, host_pipeline.cuh): the literal drop-in for the reference's
//     fn(&[u8]) -> Vec<u64> / fn(&[u64], usize) -> Vec<u8> (src/n_to_bits.rs:34, :51); PCIe-bound, so it is a chunked
//     pipeline that keeps both DMA directions and the kernel busy at once.
// There is no CPU fallback anywhere in this library: without a CUDA device every call fails with CN_ERR_CUDA.
#include "multi_device.cuh"
#include "packed_ops.cuh"

namespace {

// internal stream + events for cn_time_*_device
struct Timer {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};
thread_local Timer t_timer;

// The stream and events belong to ONE device: they are rebuilt when the calling thread's current device has changed.
// The private stream has no ordering against whatever produced the buffers, so the device is drained first.
int timer_prepare(Timer &t)
{
    int dev = 0;
    CN_CUDA(cudaGetDevice(&dev));
    if (t.stream && t.device != dev) {
        cudaEventDestroy(t.t0); cudaEventDestroy(t.t1); cudaStreamDestroy(t.stream);
        cudaGetLastError();
        t = Timer{};
    }
    if (!t.stream) {
        CN_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
        CN_CUDA(cudaEventCreate(&t.t0));
        CN_CUDA(cudaEventCreate(&t.t1));
        t.device = dev;
    }
    CN_CUDA(cudaDeviceSynchronize());
    return CN_OK;
}

// the *_ex_host entry points: validate the mode, zero the counter, run the (possibly fanned-out) host call
int ex_host(const Codec &cd, const char *who, const uint8_t *n, size_t len, uint64_t *out, int mode, uint64_t *invalid_count)
{
    if (mode != CN_ENC_PLAIN && mode != CN_ENC_COUNT && mode != CN_ENC_LUT_EXACT) return fail(CN_ERR_ARG, "%s: bad mode %d", who, mode);
    if (mode == CN_ENC_COUNT && !invalid_count) return fail(CN_ERR_ARG, "%s: CN_ENC_COUNT needs a counter", who);
    if (invalid_count) *invalid_count = 0;
    if (len == 0) return CN_OK;
    if (!n || !out) return fail(CN_ERR_ARG, "%s: null pointer", who);
    return host_codec(cd, true, n, reinterpret_cast<uint8_t *>(out), len, HostMode{mode, invalid_count});
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

int cn_n_to_bits_host_batch(const uint8_t *const *seqs, const size_t *lens, size_t count, uint64_t *const *outs)
{
    if (count == 0) return CN_OK;
    if (!seqs || !lens || !outs) return fail(CN_ERR_ARG, "cn_n_to_bits_host_batch: null pointer");
    return host_batch(kCodec2bit, true, reinterpret_cast<const void *const *>(seqs), lens, count, reinterpret_cast<void *const *>(outs));
}

int cn_bits_to_n_host_batch(const uint64_t *const *bits, const size_t *lens, size_t count, uint8_t *const *outs)
{
    if (count == 0) return CN_OK;
    if (!bits || !lens || !outs) return fail(CN_ERR_ARG, "cn_bits_to_n_host_batch: null pointer");
    return host_batch(kCodec2bit, false, reinterpret_cast<const void *const *>(bits), lens, count, reinterpret_cast<void *const *>(outs));
}

int cn_n_to_bits2_host_batch(const uint8_t *const *seqs, const size_t *lens, size_t count, uint64_t *const *outs)
{
    if (count == 0) return CN_OK;
    if (!seqs || !lens || !outs) return fail(CN_ERR_ARG, "cn_n_to_bits2_host_batch: null pointer");
    return host_batch(kCodecBase5, true, reinterpret_cast<const void *const *>(seqs), lens, count, reinterpret_cast<void *const *>(outs));
}

int cn_bits_to_n2_host_batch(const uint64_t *const *bits, const size_t *lens, size_t count, uint8_t *const *outs)
{
    if (count == 0) return CN_OK;
    if (!bits || !lens || !outs) return fail(CN_ERR_ARG, "cn_bits_to_n2_host_batch: null pointer");
    return host_batch(kCodecBase5, false, reinterpret_cast<const void *const *>(bits), lens, count, reinterpret_cast<void *const *>(outs));
}

int cn_n_to_bits_host_async(const uint8_t *n, size_t len, uint64_t *out, cn_request **req)
{
    if (!req) return fail(CN_ERR_ARG, "cn_n_to_bits_host_async: null request pointer");
    *req = nullptr;
    if (len != 0 && (!n || !out)) return fail(CN_ERR_ARG, "cn_n_to_bits_host_async: null pointer");
    return no_throw([&] {
        Request *r = nullptr;
        int rc = submit_async([=] { return len == 0 ? (int)CN_OK : host_codec(kCodec2bit, true, n, reinterpret_cast<uint8_t *>(out), len); }, &r);
        *req = reinterpret_cast<cn_request *>(r);
        return rc;
    });
}

int cn_bits_to_n_host_async(const uint64_t *bits, size_t nwords, size_t len, uint8_t *out, cn_request **req)
{
    if (!req) return fail(CN_ERR_ARG, "cn_bits_to_n_host_async: null request pointer");
    *req = nullptr;
    if (len > (nwords << 5) || (nwords >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);   // where the reference panics, up front
    if (len != 0 && (!bits || !out)) return fail(CN_ERR_ARG, "cn_bits_to_n_host_async: null pointer");
    return no_throw([&] {
        Request *r = nullptr;
        int rc = submit_async([=] { return len == 0 ? (int)CN_OK : host_codec(kCodec2bit, false, reinterpret_cast<const uint8_t *>(bits), out, len); }, &r);
        *req = reinterpret_cast<cn_request *>(r);
        return rc;
    });
}

int cn_wait(cn_request *req)
{
    if (!req) return fail(CN_ERR_ARG, "cn_wait: null request");
    return wait_request(reinterpret_cast<Request *>(req));
}

int cn_n_to_bits_checked_host(const uint8_t *n, size_t len, uint64_t *out, uint64_t *invalid_count)
{
    if (!invalid_count) return fail(CN_ERR_ARG, "cn_n_to_bits_checked_host: null counter");
    *invalid_count = 0;
    if (len == 0) return CN_OK;
    if (!n || !out) return fail(CN_ERR_ARG, "cn_n_to_bits_checked_host: null pointer");
    return host_codec(kCodec2bit, true, n, reinterpret_cast<uint8_t *>(out), len, HostMode{cn::kEncCount, invalid_count});
}

int cn_encode_checked_device(const void *d_n, size_t len, void *d_out, void *d_invalid_count, void *stream)
{
    return encode_device(d_n, len, d_out, cn::kEncCount, static_cast<unsigned long long *>(d_invalid_count), static_cast<cudaStream_t>(stream));
}

int cn_n_to_bits_ex_host(const uint8_t *n, size_t len, uint64_t *out, int mode, uint64_t *invalid_count)
{
    return ex_host(kCodec2bit, "cn_n_to_bits_ex_host", n, len, out, mode, invalid_count);
}
int cn_n_to_bits2_ex_host(const uint8_t *n, size_t len, uint64_t *out, int mode, uint64_t *invalid_count)
{
    return ex_host(kCodecBase5, "cn_n_to_bits2_ex_host", n, len, out, mode, invalid_count);
}
int cn_encode_ex_device(const void *d_n, size_t len, void *d_out, int mode, void *d_invalid_count, void *stream)
{
    return encode_device(d_n, len, d_out, mode, static_cast<unsigned long long *>(d_invalid_count), static_cast<cudaStream_t>(stream));
}
int cn_encode2_ex_device(const void *d_n, size_t len, void *d_out, int mode, void *d_invalid_count, void *stream)
{
    return encode2_device(d_n, len, d_out, mode, static_cast<unsigned long long *>(d_invalid_count), static_cast<cudaStream_t>(stream));
}

int cn_set_devices(const int *devices, int count) { return no_throw([&] { return set_devices(devices, count); }); }
int cn_get_devices(int *devices, int capacity, int *count)
{
    if (!count || capacity < 0 || (capacity > 0 && !devices)) return fail(CN_ERR_ARG, "cn_get_devices: bad arguments");
    return no_throw([&] {
        const std::vector<int> devs = devices_snapshot();
        *count = (int)devs.size();
        for (int i = 0; i < capacity && i < (int)devs.size(); i++) devices[i] = devs[i];
        return (int)CN_OK;
    });
}
int cn_shard_bounds(size_t total_len, int nshards, int shard, size_t granule, size_t *start, size_t *end)
{
    if (nshards < 1 || shard < 0 || shard >= nshards || granule == 0 || !start || !end) return fail(CN_ERR_ARG, "cn_shard_bounds: bad arguments");
    shard_range(total_len, nshards, shard, granule, start, end);
    return CN_OK;
}
int cn_encode_sharded(int nshards, const int *devices, const void *const *d_n, const size_t *lens, void *const *d_out, void *const *streams)
{
    return encode_sharded(nshards, devices, d_n, lens, d_out, streams, cn::kEncPlain, nullptr);
}
int cn_decode_sharded(int nshards, const int *devices, const void *const *d_bits, const size_t *nwords, const size_t *lens,
                      void *const *d_out, void *const *streams)
{
    return decode_sharded(nshards, devices, d_bits, nwords, lens, d_out, streams);
}
int cn_enable_peer_access(const int *devices, int count) { return enable_peer_access(devices, count); }
int cn_synchronize_devices(const int *devices, int count) { return synchronize_devices(devices, count); }

int cn_encode_segmented_device(const void *d_n, const void *d_offsets, const void *d_word_offsets, size_t count, size_t total_len,
                               size_t total_words, void *d_out, void *stream)
{
    return segmented_device(true, d_n, d_offsets, d_word_offsets, count, total_len, total_words, d_out, static_cast<cudaStream_t>(stream));
}
int cn_decode_segmented_device(const void *d_bits, const void *d_offsets, const void *d_word_offsets, size_t count, size_t total_len,
                               size_t total_words, void *d_out, void *stream)
{
    return segmented_device(false, d_bits, d_offsets, d_word_offsets, count, total_len, total_words, d_out, static_cast<cudaStream_t>(stream));
}

int cn_encode_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, void *stream)
{
    return encode_multi_device(d_n, len, d_outs, nout, static_cast<cudaStream_t>(stream));
}
int cn_encode2_multi_device(const void *d_n, size_t len, void *const *d_outs, int nout, void *stream)
{
    return encode2_multi_device(d_n, len, d_outs, nout, static_cast<cudaStream_t>(stream));
}
int cn_ipc_export(void *d_ptr, void *handle64, size_t *offset) { return ipc_export(d_ptr, handle64, offset); }
int cn_ipc_open(const void *handle64, size_t offset, void **d_ptr) { return ipc_open(handle64, offset, d_ptr); }
int cn_ipc_close(void *d_ptr, size_t offset) { return ipc_close(d_ptr, offset); }

int cn_encode_device(const void *d_n, size_t len, void *d_out, void *stream)
{
    return encode_device(d_n, len, d_out, cn::kEncPlain, nullptr, static_cast<cudaStream_t>(stream));
}

int cn_decode_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream)
{
    return decode_device(d_bits, nwords, len, d_out, static_cast<cudaStream_t>(stream));
}

int cn_hamming_device(const void *d_a, const void *d_b, size_t nwords, size_t len, void *d_result, void *stream)
{
    return hamming_device(d_a, d_b, nwords, len, d_result, static_cast<cudaStream_t>(stream));
}
int cn_complement_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream)
{
    return complement_device(d_bits, nwords, len, d_out, false, static_cast<cudaStream_t>(stream));
}
int cn_reverse_complement_device(const void *d_bits, size_t nwords, size_t len, void *d_out, void *stream)
{
    return complement_device(d_bits, nwords, len, d_out, true, static_cast<cudaStream_t>(stream));
}
int cn_hamming_host(const uint64_t *a, const uint64_t *b, size_t nwords, size_t len, uint64_t *result) { return hamming_host(a, b, nwords, len, result); }
int cn_complement_host(const uint64_t *bits, size_t nwords, size_t len, uint64_t *out) { return complement_host(bits, nwords, len, out, false); }
int cn_reverse_complement_host(const uint64_t *bits, size_t nwords, size_t len, uint64_t *out) { return complement_host(bits, nwords, len, out, true); }

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
    return encode2_device(d_n, len, d_out, cn::kEncPlain, nullptr, static_cast<cudaStream_t>(stream));
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
    CN_CUDA(cudaEventRecord(t.t0, t.stream));
    for (int i = 0; i < iters; i++) {
        rc = encode_device(d_n, len, d_out, cn::kEncPlain, nullptr, t.stream);
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
    g_tune[direction].store(pack_tuning(vec, unroll, threads), std::memory_order_relaxed);
    return CN_OK;
}

int cn_get_tuning(int direction, int *vec, int *unroll, int *threads)
{
    if ((direction != CN_DIR_ENCODE && direction != CN_DIR_DECODE) || !vec || !unroll || !threads)
        return fail(CN_ERR_ARG, "cn_get_tuning: bad arguments");
    const Tuning t = load_tuning(direction);
    *vec = t.vec; *unroll = t.unroll; *threads = t.threads;
    return CN_OK;
}

int cn_set_host_strategy(int strategy, size_t chunk_bytes)
{
    if (strategy != 0 && strategy != 1) return fail(CN_ERR_ARG, "cn_set_host_strategy: strategy must be 0 or 1");
    if (chunk_bytes) {
        if (chunk_bytes < 4096 || (chunk_bytes & 4095) || chunk_bytes > ((size_t)1 << 30))
            return fail(CN_ERR_ARG, "cn_set_host_strategy: chunk must be a multiple of 4096 in [4 KiB, 1 GiB]");
        g_host_chunk = chunk_bytes;
        g_host_chunk_pageable = chunk_bytes;
    }
    g_host_strategy = strategy;
    return CN_OK;
}

int cn_set_host_chunks(size_t pinned_chunk, size_t pageable_chunk)
{
    for (size_t c : {pinned_chunk, pageable_chunk})
        if (c && (c < 4096 || (c & 4095) || c > ((size_t)1 << 30)))
            return fail(CN_ERR_ARG, "cn_set_host_chunks: chunks must be multiples of 4096 in [4 KiB, 1 GiB]");
    if (pinned_chunk) g_host_chunk = pinned_chunk;
    if (pageable_chunk) g_host_chunk_pageable = pageable_chunk;
    return CN_OK;
}

int cn_set_host_threads(int threads)
{
    if (threads < 0 || threads > 64) return fail(CN_ERR_ARG, "cn_set_host_threads: 0..64");
    g_host_threads = threads;
    return CN_OK;
}

}  // extern "C"
