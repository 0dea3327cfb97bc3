// multi_device.cuh -- multi-GPU inside ONE process, behind the C ABI (SURVEY 8b `cn_encode_sharded`, 8e "single
// process, one stream / host thread per GPU").
//
// The codec needs no exchange: packed word w depends only on nucleotides 32w..32w+31 (src/n_to_bits.rs:39-42), so a
// call is cut into contiguous, word-aligned ranges and every range is an independent job for one GPU:
//   * host-slice calls (cn_n_to_bits_host & co.): after cn_set_devices({d0..dn}) ONE call fans out over n GPUs --
//     n PCIe links instead of one.  Each device has a persistent worker thread that owns a host_pipeline ring on its
//     GPU; the calling thread runs the last range itself.  Nothing changes for the caller: same pointers, same result.
//   * device-resident shards (cn_encode_sharded / cn_decode_sharded): one launch per device on that device's stream,
//     no host thread needed (launches are asynchronous).  With cn_enable_peer_access a shard's output pointer may live
//     on ANOTHER device -- the kernel's stores then travel over NVLink, which is the fused gather-to-root.
#pragma once
#include "host_pipeline.cuh"

#include <exception>
#include <functional>
#include <map>
#include <new>

namespace {

// contiguous range [start, end) of shard `k` of `n`: boundaries are multiples of `granule`, ranges are balanced to
// within one granule, the ragged tail lands on the last non-empty shard.  Same plan as sharded.shard_bounds (Python).
inline void shard_range(size_t total, int n, int k, size_t granule, size_t *start, size_t *end)
{
    const size_t units = (total + granule - 1) / granule;
    const size_t base = units / (size_t)n, extra = units % (size_t)n;
    const size_t first = (size_t)k * base + ((size_t)k < extra ? (size_t)k : extra);
    const size_t count = base + ((size_t)k < extra ? 1 : 0);
    size_t s = first * granule, e = (first + count) * granule;
    *start = s < total ? s : total;
    *end = e < total ? e : total;
}

// ---- the device set of the host-slice calls -------------------------------------------------------------
std::mutex g_devices_mu;
std::vector<int> g_devices;                 // empty: single-device behaviour (the calling thread's device)
bool g_devices_from_env_done = false;

// CN_DEVICES=all | "0,1,2": lets an unmodified caller (the Rust shim, the C++ harness) fan out without code changes
void devices_from_env_locked()
{
    if (g_devices_from_env_done) return;
    g_devices_from_env_done = true;
    const char *env = std::getenv("CN_DEVICES");
    if (!env || !*env) return;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return; }
    std::vector<int> devs;
    if (!strcmp(env, "all")) {
        for (int d = 0; d < count; d++) devs.push_back(d);
    } else {
        for (const char *p = env; *p;) {
            char *endp = nullptr;
            long d = strtol(p, &endp, 10);
            if (endp == p) break;
            if (d >= 0 && d < count) devs.push_back((int)d);
            p = *endp == ',' ? endp + 1 : endp;
        }
    }
    if (devs.size() > 1) g_devices = devs;
}

std::vector<int> devices_snapshot()
{
    std::lock_guard<std::mutex> lk(g_devices_mu);
    devices_from_env_locked();
    return g_devices;
}

int set_devices(const int *devices, int count)
{
    if (count < 0 || count > 64 || (count > 0 && !devices)) return fail(CN_ERR_ARG, "cn_set_devices: bad arguments");
    int have = 0;
    if (count > 0) CN_CUDA(cudaGetDeviceCount(&have));
    std::vector<int> devs;
    for (int i = 0; i < count; i++) {
        if (devices[i] < 0 || devices[i] >= have) return fail(CN_ERR_ARG, "cn_set_devices: device %d out of range (0..%d)", devices[i], have - 1);
        devs.push_back(devices[i]);
    }
    std::lock_guard<std::mutex> lk(g_devices_mu);
    g_devices_from_env_done = true;          // an explicit call overrides CN_DEVICES
    g_devices = devs;
    return CN_OK;
}

// ---- per-device worker threads ----------------------------------------------------------------------------
struct FanCall {                            // one fanned-out host-slice call
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
};
struct FanJob {                             // one device's share of it
    std::function<int()> work;              // runs on the worker thread with t_device set to the job's device
    bool live = false;
    int rc = CN_OK;
    char err[sizeof t_err] = "";
    FanCall *call = nullptr;
};

int run_fan_job(FanJob &j, int device)
{
    const int saved = t_device;
    t_device = device;
    j.rc = j.work();
    if (j.rc != CN_OK) { strncpy(j.err, t_err, sizeof j.err - 1); j.err[sizeof j.err - 1] = 0; }
    t_device = saved;
    return j.rc;
}

class DeviceWorker {
public:
    // one worker per device, created on first use, never destroyed (it only ever waits on a condition variable)
    static DeviceWorker &get(int device)
    {
        static std::mutex mu;
        static std::map<int, DeviceWorker *> *workers = new std::map<int, DeviceWorker *>();
        std::lock_guard<std::mutex> lk(mu);
        auto it = workers->find(device);
        if (it != workers->end()) return *it->second;
        DeviceWorker *w = new DeviceWorker(device);
        (*workers)[device] = w;
        return *w;
    }
    void submit(FanJob *job)
    {
        { std::lock_guard<std::mutex> lk(mu_); queue_.push_back(job); }
        cv_.notify_one();
    }

private:
    explicit DeviceWorker(int device) : device_(device) { std::thread([this] { run(); }).detach(); }
    void run()
    {
        for (;;) {
            FanJob *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !queue_.empty(); });
                job = queue_.front();
                queue_.pop_front();
            }
            run_fan_job(*job, device_);
            FanCall *call = job->call;
            { std::lock_guard<std::mutex> lk(call->mu); call->pending--; }
            call->cv.notify_all();
        }
    }
    const int device_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<FanJob *> queue_;
};

// below this many nucleotides per device a call is not worth splitting further
constexpr size_t kFanMinPerDevice = (size_t)8 << 20;
const size_t g_fan_pageable_max = [] {
    const char *env = std::getenv("CN_FANOUT_PAGEABLE_MAX");
    long v = env ? std::atol(env) : 2;
    return (size_t)(v < 1 ? 1 : v);
}();

// Job k runs on devs[k]: every job but the last goes to its device's worker, the calling thread runs the last one itself
// and then waits for the others.  The caller's current device is restored; the first failure is reported.
int run_fanned_out(std::vector<FanJob> &jobs, const std::vector<int> &devs)
{
    int saved_dev = -1;
    cudaGetDevice(&saved_dev);
    cudaGetLastError();
    FanCall call;
    const size_t parts = jobs.size();
    for (size_t k = 0; k + 1 < parts; k++)
        if (jobs[k].live) { jobs[k].call = &call; call.pending++; }
    for (size_t k = 0; k + 1 < parts; k++)
        if (jobs[k].live) DeviceWorker::get(devs[k]).submit(&jobs[k]);
    if (jobs[parts - 1].live) run_fan_job(jobs[parts - 1], devs[parts - 1]);
    {
        std::unique_lock<std::mutex> lk(call.mu);
        call.cv.wait(lk, [&] { return call.pending == 0; });
    }
    if (saved_dev >= 0) { cudaSetDevice(saved_dev); cudaGetLastError(); }
    for (FanJob &j : jobs)
        if (j.rc != CN_OK) { snprintf(t_err, sizeof t_err, "%s", j.err); return j.rc; }
    return CN_OK;
}

// The host-slice entry points land here: one device -> host_codec_one on the calling thread; a device set -> fan out.
int host_codec_impl(const Codec &cd, bool encode, const uint8_t *src, uint8_t *dst, size_t len, const HostMode &mode)
{
    if (mode.invalid_out) *mode.invalid_out = 0;
    const std::vector<int> devs = devices_snapshot();
    if (devs.empty()) return host_codec_one(cd, encode, src, dst, len, mode);

    size_t parts = len / kFanMinPerDevice;
    if (parts > devs.size()) parts = devs.size();
    // Pageable caller buffers are staged by CPU copier threads, and those -- not the links -- are the bottleneck: every
    // extra DMA stream takes host memory bandwidth away from them.  Measured with 8 links (profiles/bench_r02_n8*.json,
    // tools/host_ceiling `mix`): 20 Gnt/s over 8 GPUs vs 33 Gnt/s over 2.  So a pageable call uses at most
    // CN_FANOUT_PAGEABLE_MAX (default 2) devices.
    if (parts > g_fan_pageable_max && len >= kFanMinPerDevice) {
        const size_t nwords = cd.words(len);
        if (!is_pinned(src, encode ? len : nwords * 8) || !is_pinned(dst, encode ? nwords * 8 : len)) parts = g_fan_pageable_max;
    }
    if (parts < 1) parts = 1;
    // ranges are multiples of 1 Mi nucleotides (2-bit) / 256 warp tiles (base-5): whole words, whole chunk units, and
    // 128-byte aligned packed offsets
    const size_t granule = cd.group == 32 ? ((size_t)1 << 20) : (size_t)cn::kB5WarpBytes * 256;

    std::vector<FanJob> jobs(parts);
    std::vector<uint64_t> invalid(parts, 0);
    const bool counted = encode && mode.enc_mode != cn::kEncPlain;
    for (size_t k = 0; k < parts; k++) {
        size_t s, e;
        shard_range(len, (int)parts, (int)k, granule, &s, &e);
        const size_t word_off = s / cd.group * 8, part_len = e - s;
        const uint8_t *part_src = src + (encode ? s : word_off);
        uint8_t *part_dst = dst + (encode ? word_off : s);
        const HostMode part_mode{mode.enc_mode, counted ? &invalid[k] : nullptr};
        jobs[k].live = part_len != 0;
        jobs[k].work = [&cd, encode, part_src, part_dst, part_len, part_mode] {
            return host_codec_one(cd, encode, part_src, part_dst, part_len, part_mode);
        };
    }
    int rc = run_fanned_out(jobs, devs);
    if (rc != CN_OK) return rc;
    uint64_t total_invalid = 0;
    for (uint64_t v : invalid) total_invalid += v;
    if (mode.invalid_out) *mode.invalid_out = total_invalid;
    return CN_OK;
}

// Batched small sequences: the batch is cut into runs of consecutive sequences of about equal total length, one run per
// device (the sequences are pageable host memory as a rule, so at most CN_FANOUT_PAGEABLE_MAX devices are used).
int host_batch_impl(const Codec &cd, bool encode, const void *const *in, const size_t *lens, size_t count, void *const *out)
{
    const std::vector<int> devs = devices_snapshot();
    size_t total = 0;
    for (size_t i = 0; i < count; i++) total += lens[i];
    size_t parts = total / kFanMinPerDevice;
    if (parts > devs.size()) parts = devs.size();
    if (parts > g_fan_pageable_max) parts = g_fan_pageable_max;
    if (devs.empty() || parts <= 1) {
        if (devs.empty()) return host_batch_one(cd, encode, in, lens, count, out);
        std::vector<FanJob> one(1);
        one[0].live = true;
        one[0].work = [=, &cd] { return host_batch_one(cd, encode, in, lens, count, out); };
        return run_fanned_out(one, devs);
    }
    std::vector<FanJob> jobs(parts);
    size_t first = 0, acc = 0;
    for (size_t k = 0; k < parts; k++) {
        const size_t target = total / parts * (k + 1);
        size_t last = first;
        while (last < count && (k + 1 == parts || acc + lens[last] <= target)) acc += lens[last++];
        const size_t n = last - first, at = first;
        jobs[k].live = n != 0;
        jobs[k].work = [=, &cd] { return host_batch_one(cd, encode, in + at, lens + at, n, out + at); };
        first = last;
    }
    return run_fanned_out(jobs, devs);
}

// Nothing may unwind across the C ABI: the host paths allocate (job lists, task queues, plans), so an allocation failure is
// turned into a status here.
template <class F> int no_throw(F &&body) noexcept
{
    try { return body(); }
    catch (const std::bad_alloc &) { return fail(CN_ERR_NOMEM, "host memory allocation failed"); }
    catch (const std::exception &e) { return fail(CN_ERR_CUDA, "internal error: %s", e.what()); }
    catch (...) { return fail(CN_ERR_CUDA, "internal error"); }
}
int host_codec(const Codec &cd, bool encode, const uint8_t *src, uint8_t *dst, size_t len, const HostMode &mode = HostMode{}) noexcept
{
    return no_throw([&] { return host_codec_impl(cd, encode, src, dst, len, mode); });
}
int host_batch(const Codec &cd, bool encode, const void *const *in, const size_t *lens, size_t count, void *const *out) noexcept
{
    return no_throw([&] { return host_batch_impl(cd, encode, in, lens, count, out); });
}

// ---- asynchronous host calls ------------------------------------------------------------------------------
// cn_*_host_async hands the call to one of a few library threads (each with its own staging ring) and returns at once;
// cn_wait collects the status.  A single-threaded caller can thus keep an encode and a decode in flight together -- the two
// calls load opposite PCIe directions, so a stream of round trips runs at ~33 Gnt/s on one link instead of 26.5.
struct Request {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    int rc = CN_OK;
    char err[sizeof t_err] = "";
    std::function<int()> work;
    int device = -1;
};

class AsyncPool {
public:
    static AsyncPool &get()
    {
        static AsyncPool *pool = new AsyncPool();            // leaked on purpose, like the other pools
        return *pool;
    }
    void submit(Request *r)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.push_back(r);
            // one thread per request in flight, up to kMax: spawn when more requests are queued than threads are idle
            if ((int)queue_.size() > idle_ && threads_ < kMax) { threads_++; std::thread([this] { run(); }).detach(); }
        }
        cv_.notify_one();
    }

private:
    static constexpr int kMax = 4;
    void run()
    {
        for (;;) {
            Request *r;
            {
                std::unique_lock<std::mutex> lk(mu_);
                idle_++;
                cv_.wait(lk, [&] { return !queue_.empty(); });
                idle_--;
                r = queue_.front();
                queue_.pop_front();
            }
            t_device = r->device;
            int rc = r->work();
            {
                std::lock_guard<std::mutex> lk(r->mu);
                r->rc = rc;
                if (rc != CN_OK) { strncpy(r->err, t_err, sizeof r->err - 1); r->err[sizeof r->err - 1] = 0; }
                r->done = true;
            }
            r->cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Request *> queue_;
    int threads_ = 0, idle_ = 0;
};

// the device the calling thread would use right now: the one chosen with cn_init, else the current one
int calling_device()
{
    if (t_device >= 0) return t_device;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

int submit_async(std::function<int()> work, Request **out)
{
    Request *r = new Request();
    r->work = std::move(work);
    r->device = calling_device();
    *out = r;
    AsyncPool::get().submit(r);
    return CN_OK;
}

int wait_request(Request *r)
{
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->cv.wait(lk, [&] { return r->done; });
    }
    const int rc = r->rc;
    if (rc != CN_OK) snprintf(t_err, sizeof t_err, "%s", r->err);
    delete r;
    return rc;
}

// ---- device-resident shards, one per device, single process ---------------------------------------------
struct DeviceGuard {                        // launches happen on other devices; leave the caller's current device as it was
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (saved >= 0) { cudaSetDevice(saved); cudaGetLastError(); } }
};

int check_shard_args(const char *who, int nshards, const int *devices, const void *a, const void *b, const void *c)
{
    if (nshards < 1 || nshards > 64 || !devices || !a || !b || !c) return fail(CN_ERR_ARG, "%s: bad arguments", who);
    return CN_OK;
}

int encode_sharded(int nshards, const int *devices, const void *const *d_n, const size_t *lens, void *const *d_out,
                   void *const *streams, int mode, void *const *counters)
{
    int rc = check_shard_args("cn_encode_sharded", nshards, devices, d_n, lens, d_out);
    if (rc != CN_OK) return rc;
    DeviceGuard guard;
    for (int k = 0; k < nshards; k++) {
        if (lens[k] == 0) continue;
        CN_CUDA(cudaSetDevice(devices[k]));
        cudaStream_t s = streams ? static_cast<cudaStream_t>(streams[k]) : nullptr;
        rc = encode_device(d_n[k], lens[k], d_out[k], mode, counters ? static_cast<unsigned long long *>(counters[k]) : nullptr, s);
        if (rc != CN_OK) return rc;
    }
    return CN_OK;
}

int decode_sharded(int nshards, const int *devices, const void *const *d_bits, const size_t *nwords, const size_t *lens,
                   void *const *d_out, void *const *streams)
{
    int rc = check_shard_args("cn_decode_sharded", nshards, devices, d_bits, lens, d_out);
    if (rc != CN_OK) return rc;
    if (!nwords) return fail(CN_ERR_ARG, "cn_decode_sharded: bad arguments");
    for (int k = 0; k < nshards; k++)                                   // the reference's panic, before anything is launched
        if (lens[k] > (nwords[k] << 5) || (nwords[k] >> 59) != 0) return fail(CN_ERR_LENGTH, "%s", kPanicText);
    DeviceGuard guard;
    for (int k = 0; k < nshards; k++) {
        if (lens[k] == 0) continue;
        CN_CUDA(cudaSetDevice(devices[k]));
        cudaStream_t s = streams ? static_cast<cudaStream_t>(streams[k]) : nullptr;
        rc = decode_device(d_bits[k], nwords[k], lens[k], d_out[k], s);
        if (rc != CN_OK) return rc;
    }
    return CN_OK;
}

int enable_peer_access(const int *devices, int count)
{
    if (count < 1 || !devices) return fail(CN_ERR_ARG, "cn_enable_peer_access: bad arguments");
    DeviceGuard guard;
    for (int i = 0; i < count; i++) {
        CN_CUDA(cudaSetDevice(devices[i]));
        for (int j = 0; j < count; j++) {
            if (i == j) continue;
            int can = 0;
            CN_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) return fail(CN_ERR_CUDA, "cn_enable_peer_access: device %d cannot access device %d", devices[i], devices[j]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }
            if (e != cudaSuccess) return fail(CN_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", devices[i], devices[j], cudaGetErrorString(e));
        }
    }
    return CN_OK;
}

int synchronize_devices(const int *devices, int count)
{
    if (count < 1 || !devices) return fail(CN_ERR_ARG, "cn_synchronize_devices: bad arguments");
    DeviceGuard guard;
    for (int i = 0; i < count; i++) {
        CN_CUDA(cudaSetDevice(devices[i]));
        CN_CUDA(cudaDeviceSynchronize());
    }
    return CN_OK;
}

}  // namespace
