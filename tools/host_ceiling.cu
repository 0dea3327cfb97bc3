// host_ceiling.cu -- what the HOST side of the box can carry (harness, not product code).
//
// The host-slice entry points (cn_n_to_bits_host / cn_bits_to_n_host) are bounded by three things that have
// nothing to do with the kernels: the PCIe link(s), the host memory system that feeds them, and -- for
// pageable caller buffers -- the staging copies the CPU has to make.  This tool measures each of them alone
// and together on the box it runs on and prints one JSON object per line:
//
//   sys       core count, NUMA nodes, THP mode, pageable-memory-access / host-register attributes per GPU
//   hostmem   T threads x {memcpy, AVX2 non-temporal copy, read-only sum, STREAM triad} over private 256 MiB slices
//   pcie      G GPUs x {H2D only, D2H only, both directions at once} from pinned memory, 1 GiB per direction per GPU
//   mix       the same H2D+D2H traffic while T threads run the staging copy (what a pageable call does)
//   register  cudaHostRegister / cudaHostUnregister throughput on a touched 1 GiB malloc (4 KiB pages and THP)
//
//   regpar    cudaHostRegister+Unregister of disjoint 32 MiB chunks from T threads, with an H2D stream running beside it
//
// Usage: host_ceiling [sys] [hostmem] [pcie] [mix] [register] [regpar]   (default: all)
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::string slurp(const char *path)
{
    std::ifstream f(path);
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    for (auto &c : s) if (c == '"' || c == '\n') c = ' ';
    return s;
}

// ---- a spin barrier so all threads start a timed pass together ------------------------------------------
struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

// ---- copy flavours ----------------------------------------------------------------------------------------
static void nt_copy(void *dst, const void *src, size_t bytes)      // dst 32-byte aligned, bytes multiple of 128
{
    const __m256i *s = static_cast<const __m256i *>(src);
    __m256i *d = static_cast<__m256i *>(dst);
    for (size_t i = 0; i < bytes / 32; i += 4) {
        __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1);
        __m256i c = _mm256_loadu_si256(s + i + 2), e = _mm256_loadu_si256(s + i + 3);
        _mm256_stream_si256(d + i, a); _mm256_stream_si256(d + i + 1, b);
        _mm256_stream_si256(d + i + 2, c); _mm256_stream_si256(d + i + 3, e);
    }
    _mm_sfence();
}
static uint64_t read_sum(const void *src, size_t bytes)
{
    const __m256i *s = static_cast<const __m256i *>(src);
    __m256i acc0 = _mm256_setzero_si256(), acc1 = acc0;
    for (size_t i = 0; i < bytes / 32; i += 2) {
        acc0 = _mm256_xor_si256(acc0, _mm256_loadu_si256(s + i));
        acc1 = _mm256_xor_si256(acc1, _mm256_loadu_si256(s + i + 1));
    }
    acc0 = _mm256_xor_si256(acc0, acc1);
    return (uint64_t)_mm256_extract_epi64(acc0, 0) ^ (uint64_t)_mm256_extract_epi64(acc0, 3);
}
static void triad(double *a, const double *b, const double *c, size_t n)
{
    for (size_t i = 0; i < n; i++) a[i] = b[i] + 3.0 * c[i];
}

static void *big_alloc(size_t bytes, bool thp)
{
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("mmap"); exit(1); }
    madvise(p, bytes, thp ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
    return p;
}

// ---- sys ----------------------------------------------------------------------------------------------------
static void run_sys()
{
    int ngpu = 0;
    cudaGetDeviceCount(&ngpu);
    printf("{\"section\": \"sys\", \"online_cpus\": %ld, \"hardware_concurrency\": %u, \"numa_nodes_online\": \"%s\", "
           "\"thp_enabled\": \"%s\", \"thp_defrag\": \"%s\", \"mem_total_kb\": \"%s\", \"gpus\": %d",
           sysconf(_SC_NPROCESSORS_ONLN), std::thread::hardware_concurrency(), slurp("/sys/devices/system/node/online").c_str(),
           slurp("/sys/kernel/mm/transparent_hugepage/enabled").c_str(), slurp("/sys/kernel/mm/transparent_hugepage/defrag").c_str(),
           slurp("/proc/meminfo").substr(0, 40).c_str(), ngpu);
    printf(", \"per_gpu\": [");
    for (int d = 0; d < ngpu; d++) {
        int pageable = 0, hostreg = 0, uva = 0, pthost = 0, numa = -1;
        cudaDeviceGetAttribute(&pageable, cudaDevAttrPageableMemoryAccess, d);
        cudaDeviceGetAttribute(&pthost, cudaDevAttrPageableMemoryAccessUsesHostPageTables, d);
        cudaDeviceGetAttribute(&hostreg, cudaDevAttrHostRegisterSupported, d);
        cudaDeviceGetAttribute(&uva, cudaDevAttrUnifiedAddressing, d);
        char bus[32] = "";
        cudaDeviceGetPCIBusId(bus, sizeof bus, d);
        std::string lower(bus);
        for (auto &c : lower) c = (char)tolower(c);
        std::string node = slurp(("/sys/bus/pci/devices/" + lower + "/numa_node").c_str());
        if (!node.empty()) numa = atoi(node.c_str());
        std::string speed = slurp(("/sys/bus/pci/devices/" + lower + "/current_link_speed").c_str());
        std::string width = slurp(("/sys/bus/pci/devices/" + lower + "/current_link_width").c_str());
        printf("%s{\"dev\": %d, \"bus\": \"%s\", \"numa_node\": %d, \"link_speed\": \"%s\", \"link_width\": \"%s\", "
               "\"pageable_access\": %d, \"uses_host_page_tables\": %d, \"host_register\": %d, \"uva\": %d}",
               d ? ", " : "", d, bus, numa, speed.c_str(), width.c_str(), pageable, pthost, hostreg, uva);
    }
    printf("]}\n");
    fflush(stdout);
}

// ---- hostmem --------------------------------------------------------------------------------------------------
static void run_hostmem()
{
    const size_t slice = (size_t)256 << 20;
    const int maxt = (int)std::thread::hardware_concurrency();
    std::vector<int> counts;
    for (int t : {1, 2, 4, 6, 8, 12, 16, 24, 32, 48, 64}) if (t <= maxt) counts.push_back(t);
    if (counts.empty() || counts.back() != maxt) counts.push_back(maxt);
    const int tmax = counts.back();
    // per-thread private buffers, touched by the owning thread
    std::vector<uint8_t *> src(tmax), dst(tmax), third(tmax);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < tmax; t++)
            th.emplace_back([&, t] {
                src[t] = (uint8_t *)big_alloc(slice, true); dst[t] = (uint8_t *)big_alloc(slice, true); third[t] = (uint8_t *)big_alloc(slice, true);
                memset(src[t], 1 + t, slice); memset(dst[t], 2, slice); memset(third[t], 3, slice);
            });
        for (auto &x : th) x.join();
    }
    const char *names[4] = {"memcpy", "nt_copy", "read", "triad"};
    for (int T : counts) {
        double gbs[4] = {0, 0, 0, 0};
        for (int kind = 0; kind < 4; kind++) {
            Barrier bar(T + 1);
            std::atomic<uint64_t> sink{0};
            const int reps = 4;
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    for (int r = 0; r < reps + 1; r++) {
                        bar.wait();
                        if (kind == 0) memcpy(dst[t], src[t], slice);
                        else if (kind == 1) nt_copy(dst[t], src[t], slice);
                        else if (kind == 2) sink += read_sum(src[t], slice);
                        else triad((double *)dst[t], (const double *)src[t], (const double *)third[t], slice / 8);
                        bar.wait();
                    }
                });
            double best = 1e30;
            for (int r = 0; r < reps + 1; r++) {
                bar.wait();
                double t0 = now();
                bar.wait();
                double dt = now() - t0;
                if (r > 0 && dt < best) best = dt;
            }
            for (auto &x : th) x.join();
            // bytes of memory traffic counted the STREAM way: copy = 2x, read = 1x, triad = 3x
            double mult = kind == 2 ? 1.0 : (kind == 3 ? 3.0 : 2.0);
            gbs[kind] = mult * slice * T / best / 1e9;
        }
        printf("{\"section\": \"hostmem\", \"threads\": %d, \"slice_mib\": %zu", T, slice >> 20);
        for (int k = 0; k < 4; k++) printf(", \"%s_gbs\": %.1f", names[k], gbs[k]);
        printf(", \"note\": \"traffic counted STREAM-style: copies 2x payload, triad 3x; payload copy rate = copy_gbs / 2\"}\n");
        fflush(stdout);
    }
    for (int t = 0; t < tmax; t++) { munmap(src[t], slice); munmap(dst[t], slice); munmap(third[t], slice); }
}

// ---- pcie -----------------------------------------------------------------------------------------------------
struct Link {
    int dev;
    uint8_t *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
};

static std::vector<Link> make_links(int G, size_t bytes)
{
    std::vector<Link> links(G);
    for (int g = 0; g < G; g++) {
        Link &l = links[g];
        l.dev = g;
        CK(cudaSetDevice(g));
        CK(cudaHostAlloc(&l.h_in, bytes, cudaHostAllocDefault));
        CK(cudaHostAlloc(&l.h_out, bytes, cudaHostAllocDefault));
        memset(l.h_in, 5, bytes); memset(l.h_out, 6, bytes);
        CK(cudaMalloc(&l.d_in, bytes));
        CK(cudaMalloc(&l.d_out, bytes));
        CK(cudaMemset(l.d_out, 7, bytes));
        CK(cudaStreamCreateWithFlags(&l.s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&l.s_out, cudaStreamNonBlocking));
    }
    return links;
}
static void free_links(std::vector<Link> &links)
{
    for (auto &l : links) {
        CK(cudaSetDevice(l.dev));
        cudaFreeHost(l.h_in); cudaFreeHost(l.h_out); cudaFree(l.d_in); cudaFree(l.d_out);
        cudaStreamDestroy(l.s_in); cudaStreamDestroy(l.s_out);
    }
}
// mode bit 0 = H2D, bit 1 = D2H; returns seconds for `reps` passes of `bytes` per direction per link
static double run_dma(std::vector<Link> &links, size_t bytes, int mode, int reps)
{
    const size_t chunk = (size_t)64 << 20;
    for (auto &l : links) { CK(cudaSetDevice(l.dev)); CK(cudaDeviceSynchronize()); }
    double t0 = now();
    for (int r = 0; r < reps; r++)
        for (size_t off = 0; off < bytes; off += chunk)
            for (auto &l : links) {
                CK(cudaSetDevice(l.dev));
                size_t n = bytes - off < chunk ? bytes - off : chunk;
                if (mode & 1) CK(cudaMemcpyAsync(l.d_in + off, l.h_in + off, n, cudaMemcpyHostToDevice, l.s_in));
                if (mode & 2) CK(cudaMemcpyAsync(l.h_out + off, l.d_out + off, n, cudaMemcpyDeviceToHost, l.s_out));
            }
    for (auto &l : links) { CK(cudaSetDevice(l.dev)); CK(cudaStreamSynchronize(l.s_in)); CK(cudaStreamSynchronize(l.s_out)); }
    return now() - t0;
}

static void run_pcie(bool with_mix)
{
    int ngpu = 0;
    CK(cudaGetDeviceCount(&ngpu));
    const size_t bytes = (size_t)1 << 30;
    const int reps = 4;
    for (int G : {1, 2, 4, 8}) {
        if (G > ngpu) break;
        std::vector<Link> links = make_links(G, bytes);
        run_dma(links, bytes, 3, 1);                                   // warm-up
        double th2d = run_dma(links, bytes, 1, reps), td2h = run_dma(links, bytes, 2, reps), tboth = run_dma(links, bytes, 3, reps);
        const double tot = (double)bytes * reps * G / 1e9;
        printf("{\"section\": \"pcie\", \"gpus\": %d, \"h2d_only_gbs\": %.1f, \"d2h_only_gbs\": %.1f, \"both_h2d_gbs\": %.1f, \"both_d2h_gbs\": %.1f, "
               "\"both_total_gbs\": %.1f, \"per_gpu_h2d_only_gbs\": %.1f, \"per_gpu_d2h_only_gbs\": %.1f}\n",
               G, tot / th2d, tot / td2h, tot / tboth, tot / tboth, 2 * tot / tboth, tot / th2d / G, tot / td2h / G);
        fflush(stdout);
        if (with_mix) {
            // the staging copies of a pageable call run on CPU threads WHILE the DMA engines stream: do they fight?
            const size_t slice = (size_t)256 << 20;
            const int maxt = (int)std::thread::hardware_concurrency();
            for (int T : {4, 8, 16}) {
                if (T > maxt) break;
                std::vector<uint8_t *> src(T), dst(T);
                for (int t = 0; t < T; t++) {
                    src[t] = (uint8_t *)big_alloc(slice, true);
                    CK(cudaHostAlloc(&dst[t], slice, cudaHostAllocDefault));
                    memset(src[t], 1, slice);
                }
                std::atomic<bool> stop{false};
                std::atomic<uint64_t> copied{0};
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back([&, t] {
                        while (!stop.load()) { nt_copy(dst[t], src[t], slice); copied += slice; }
                    });
                std::this_thread::sleep_for(std::chrono::milliseconds(50));
                uint64_t c0 = copied.load();
                double t0 = now();
                double tdma = run_dma(links, bytes, 3, reps);
                double dt = now() - t0;
                uint64_t c1 = copied.load();
                stop = true;
                for (auto &x : th) x.join();
                printf("{\"section\": \"mix\", \"gpus\": %d, \"copy_threads\": %d, \"dma_h2d_gbs\": %.1f, \"dma_d2h_gbs\": %.1f, "
                       "\"concurrent_nt_copy_payload_gbs\": %.1f}\n", G, T, tot / tdma, tot / tdma, (double)(c1 - c0) / dt / 1e9);
                fflush(stdout);
                for (int t = 0; t < T; t++) { munmap(src[t], slice); cudaFreeHost(dst[t]); }
            }
        }
        free_links(links);
    }
}

// ---- register ---------------------------------------------------------------------------------------------------
static void run_register()
{
    CK(cudaSetDevice(0));
    CK(cudaFree(nullptr));
    const size_t bytes = (size_t)1 << 30;
    for (int thp = 0; thp < 2; thp++) {
        uint8_t *p = (uint8_t *)big_alloc(bytes, thp != 0);
        memset(p, 1, bytes);
        double t0 = now();
        cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
        double t1 = now();
        if (e != cudaSuccess) {
            printf("{\"section\": \"register\", \"thp\": %d, \"error\": \"%s\"}\n", thp, cudaGetErrorString(e));
            cudaGetLastError();
            munmap(p, bytes);
            continue;
        }
        uint8_t *d = nullptr;
        CK(cudaMalloc(&d, bytes));
        double t2 = now();
        CK(cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice));
        double t3 = now();
        CK(cudaHostUnregister(p));
        double t4 = now();
        // the plain pageable cudaMemcpy for comparison (driver-internal staging)
        CK(cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice));
        double t5 = now();
        CK(cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice));
        double t6 = now();
        printf("{\"section\": \"register\", \"thp\": %d, \"register_gbs\": %.2f, \"h2d_registered_gbs\": %.1f, \"unregister_gbs\": %.2f, "
               "\"register_copy_unregister_gbs\": %.2f, \"pageable_cudaMemcpy_gbs\": %.1f}\n",
               thp, bytes / (t1 - t0) / 1e9, bytes / (t3 - t2) / 1e9, bytes / (t4 - t3) / 1e9, bytes / ((t1 - t0) + (t3 - t2) + (t4 - t3)) / 1e9,
               bytes / (t6 - t5) / 1e9);
        fflush(stdout);
        CK(cudaFree(d));
        munmap(p, bytes);
    }
}


// ---- regpar: does cudaHostRegister scale over threads, and does it disturb a DMA stream that is running? -------
static void run_regpar()
{
    CK(cudaSetDevice(0));
    CK(cudaFree(nullptr));
    const size_t total = (size_t)4 << 30, chunk = (size_t)32 << 20;
    const size_t nchunks = total / chunk;
    for (int thp = 0; thp < 2; thp++) {
        uint8_t *p = (uint8_t *)big_alloc(total, thp != 0);
        {
            std::vector<std::thread> th;
            for (int t = 0; t < 8; t++) th.emplace_back([&, t] { memset(p + (total / 8) * t, 1 + t, total / 8); });
            for (auto &x : th) x.join();
        }
        for (int T : {1, 2, 4, 8}) {
            // T threads register + unregister disjoint chunks; the main thread streams an unrelated pinned buffer H2D
            std::vector<Link> links = make_links(1, (size_t)1 << 30);
            std::atomic<size_t> next{0};
            std::atomic<int> errors{0};
            std::atomic<bool> done{false};
            double reg_s = 0;
            std::vector<std::thread> th;
            double t0 = now();
            for (int t = 0; t < T; t++)
                th.emplace_back([&] {
                    cudaSetDevice(0);
                    for (;;) {
                        size_t c = next.fetch_add(1);
                        if (c >= nchunks) break;
                        if (cudaHostRegister(p + c * chunk, chunk, cudaHostRegisterDefault) != cudaSuccess) { errors++; cudaGetLastError(); continue; }
                        if (cudaHostUnregister(p + c * chunk) != cudaSuccess) { errors++; cudaGetLastError(); }
                    }
                });
            std::thread dma([&] {
                cudaSetDevice(0);
                size_t moved = 0;
                double d0 = now();
                while (!done.load()) {
                    cudaMemcpyAsync(links[0].d_in, links[0].h_in, (size_t)256 << 20, cudaMemcpyHostToDevice, links[0].s_in);
                    cudaStreamSynchronize(links[0].s_in);
                    moved += (size_t)256 << 20;
                }
                reg_s = moved / (now() - d0) / 1e9;
            });
            for (auto &x : th) x.join();
            double dt = now() - t0;
            done = true;
            dma.join();
            printf("{\"section\": \"regpar\", \"thp\": %d, \"threads\": %d, \"chunk_mib\": %zu, \"register_unregister_gbs\": %.1f, "
                   "\"concurrent_h2d_gbs\": %.1f, \"errors\": %d}\n", thp, T, chunk >> 20, total / dt / 1e9, reg_s, errors.load());
            fflush(stdout);
            free_links(links);
        }
        munmap(p, total);
    }
}

int main(int argc, char **argv)
{
    bool all = argc == 1;
    auto want = [&](const char *s) {
        if (all) return true;
        for (int i = 1; i < argc; i++) if (!strcmp(argv[i], s)) return true;
        return false;
    };
    if (want("sys")) run_sys();
    if (want("hostmem")) run_hostmem();
    if (want("pcie") || want("mix")) run_pcie(want("mix"));
    if (want("register")) run_register();
    if (want("regpar")) run_regpar();
    return 0;
}
