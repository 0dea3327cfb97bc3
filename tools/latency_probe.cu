// latency_probe.cu -- where do the ~16 us of a 40 000-nt host call go?  Times, on pinned 40 KB buffers:
//   (a) empty kernel + cudaStreamSynchronize          (b) empty kernel + host polling a pinned flag the kernel writes
//   (c) encode kernel reading/writing pinned memory + cudaStreamSynchronize      (d) same + flag polling (single CTA)
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../cute_nucleotides_b200/csrc/codec_kernels.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
using namespace cn;

__global__ void k_empty() {}
__global__ void k_flag(volatile unsigned *flag, unsigned seq) { if (threadIdx.x == 0) { *flag = seq; } }
// single CTA: encode nvec 32-byte vectors, then publish seq
__global__ void __launch_bounds__(1024) k_encode_flag(const uint8_t *in, uint32_t *out32, size_t nvec, volatile unsigned *flag, unsigned seq)
{
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        u32x8 v = ld_stream32(in + (i << 5));
        st_stream8(out32 + (i << 1), pack16(v.v[0], v.v[1], v.v[2], v.v[3]), pack16(v.v[4], v.v[5], v.v[6], v.v[7]));
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); *flag = seq; }
}
__global__ void __launch_bounds__(256) k_encode(const uint8_t *in, uint32_t *out32, size_t nvec)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < nvec) { u32x8 v = ld_stream32(in + (i << 5)); st_stream8(out32 + (i << 1), pack16(v.v[0], v.v[1], v.v[2], v.v[3]), pack16(v.v[4], v.v[5], v.v[6], v.v[7])); }
}
template <typename F> double med_us(F f, int n = 2000)
{
    std::vector<double> t;
    for (int i = 0; i < n; i++) { auto a = std::chrono::steady_clock::now(); f(i); auto b = std::chrono::steady_clock::now(); t.push_back(std::chrono::duration<double, std::micro>(b - a).count()); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}
int main()
{
    const size_t L = 40000, nvec = L / 32;
    uint8_t *h_in; uint32_t *h_out; unsigned *h_flag;
    CK(cudaHostAlloc(&h_in, L, 0)); CK(cudaHostAlloc(&h_out, L / 4, 0)); CK(cudaHostAlloc(&h_flag, 64, 0));
    memset(h_in, 'A', L); *h_flag = 0;
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    volatile unsigned *vf = h_flag;
    for (int i = 0; i < 100; i++) { k_empty<<<1, 32, 0, s>>>(); } CK(cudaStreamSynchronize(s));
    double a = med_us([&](int) { k_empty<<<1, 32, 0, s>>>(); cudaStreamSynchronize(s); });
    unsigned seq = 0;
    double b = med_us([&](int) { ++seq; k_flag<<<1, 32, 0, s>>>(vf, seq); while (*vf != seq) { } });
    CK(cudaStreamSynchronize(s));
    double c = med_us([&](int) { k_encode<<<(unsigned)((nvec + 255) / 256), 256, 0, s>>>(h_in, h_out, nvec); cudaStreamSynchronize(s); });
    double d = med_us([&](int) { ++seq; k_encode_flag<<<1, 1024, 0, s>>>(h_in, h_out, nvec, vf, seq); while (*vf != seq) { } });
    CK(cudaStreamSynchronize(s));
    std::vector<uint8_t> src(L, 'C'), dst(L / 4);
    double m = med_us([&](int) { memcpy(h_in, src.data(), L); memcpy(dst.data(), h_out, L / 4); });
    printf("{\"empty_launch_sync_us\": %.2f, \"empty_launch_poll_us\": %.2f, \"encode40k_zero_copy_sync_us\": %.2f, \"encode40k_zero_copy_poll_us\": %.2f, \"staging_memcpy_us\": %.2f}\n", a, b, c, d, m);
    return 0;
}
