// hbm_ceiling.cu -- what can the B200 memory system deliver for the codec's two traffic mixes?
//
// The encode kernel reads 4 bytes per byte it writes, the decode kernel writes 4 bytes per byte it reads.
// This harness times compute-free kernels with those mixes (and pure read / write / copy) plus the
// load/store flavours that could matter (128- vs 256-bit, cache policies, TMA bulk copies through
// shared memory), so the codec kernels can be judged against what the memory system can do for that
// mix rather than against a 1:1 copy.  Prints one JSON object per variant.
//
// build: make -C tools     run (GPU box): tools/hbm_ceiling [GiB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../cute_nucleotides_b200/csrc/codec_kernels.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

using namespace cn;

// ---- load / store flavours ----------------------------------------------------------------------
enum { LD_NC_NA = 0, LD_PLAIN = 1, LD_NC = 2, LD_CS = 3 };
enum { ST_NA = 0, ST_PLAIN = 1, ST_CS = 2, ST_WT = 3, ST_EVICT_FIRST = 4 };

template <int F> __device__ __forceinline__ uint4 ld16(const void *p)
{
    uint4 r;
    if (F == LD_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (F == LD_PLAIN) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (F == LD_NC) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int F> __device__ __forceinline__ void st16(void *p, uint4 v, uint64_t pol)
{
    if (F == ST_NA) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (F == ST_PLAIN) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (F == ST_CS) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (F == ST_WT) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// ---- pure read / write / copy --------------------------------------------------------------------
template <int VEC, int U, int T>
__global__ void __launch_bounds__(T) k_read(const uint8_t *__restrict__ in, uint32_t *__restrict__ sink, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint32_t acc = 0;
    if (VEC == 16) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; v[u] = i < nvec ? ld_stream16(in + (i << 4)) : make_uint4(0, 0, 0, 0); }
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    } else {
        u32x8 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) v[u] = ld_stream32(in + (i << 5)); else for (int k = 0; k < 8; k++) v[u].v[k] = 0; }
#pragma unroll
        for (int u = 0; u < U; u++) for (int k = 0; k < 8; k++) acc ^= v[u].v[k];
    }
    if (acc == 0x12345678u) sink[0] = acc;     // never true for our data; keeps the loads alive
}

template <int VEC, int U, int T>
__global__ void __launch_bounds__(T) k_write(uint8_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll
    for (int u = 0; u < U; u++) {
        size_t i = t + (size_t)u * T;
        if (i < nvec) { if (VEC == 16) st_stream16(out + (i << 4), v); else st_stream32(out + (i << 5), v, v); }
    }
}

template <int VEC, int U, int T>
__global__ void __launch_bounds__(T) k_copy(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    if (VEC == 16) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) v[u] = ld_stream16(in + (i << 4)); }
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream16(out + (i << 4), v[u]); }
    } else {
        u32x8 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) v[u] = ld_stream32(in + (i << 5)); }
#pragma unroll
        for (int u = 0; u < U; u++) {
            size_t i = t + (size_t)u * T;
            if (i < nvec) st_stream32(out + (i << 5), make_uint4(v[u].v[0], v[u].v[1], v[u].v[2], v[u].v[3]), make_uint4(v[u].v[4], v[u].v[5], v[u].v[6], v[u].v[7]));
        }
    }
}

// ---- encode-like mix (4 read : 1 write), no arithmetic -----------------------------------------------
template <int LDF, int U, int T>
__global__ void __launch_bounds__(T) k_mix41(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) v[u] = ld16<LDF>(in + (i << 4)); }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream4(out + i, v[u].x ^ v[u].y ^ v[u].z ^ v[u].w); }
}

// ---- decode-like mix (1 read : 4 write), no arithmetic; store flavour selectable ---------------------
template <int STF, int U, int T>
__global__ void __launch_bounds__(T) k_mix14(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint64_t pol = STF == ST_EVICT_FIRST ? policy_evict_first() : 0;
    uint32_t w[U];
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) w[u] = ld_stream4(in + i); }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st16<STF>(out + (i << 4), make_uint4(w[u], w[u] + 1, w[u] + 2, w[u] + 3), pol); }
}

// ---- 256-bit flavours of the mixes ---------------------------------------------------------------------
template <bool NA, int U, int T>
__global__ void __launch_bounds__(T) k_mix41_v32(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    u32x8 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        size_t i = t + (size_t)u * T;
        if (i < nvec) {
            if (NA) v[u] = ld_stream32(in + (i << 5));
            else asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[u].v[0]), "=r"(v[u].v[1]), "=r"(v[u].v[2]), "=r"(v[u].v[3]), "=r"(v[u].v[4]), "=r"(v[u].v[5]), "=r"(v[u].v[6]), "=r"(v[u].v[7]) : "l"(in + (i << 5)));
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream8(out + (i << 1), v[u].v[0] ^ v[u].v[1] ^ v[u].v[2] ^ v[u].v[3], v[u].v[4] ^ v[u].v[5] ^ v[u].v[6] ^ v[u].v[7]); }
}

template <int U, int T, bool L2_256>
__global__ void __launch_bounds__(T) k_mix14_v32(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, size_t nvec)
{
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint2 w[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        size_t i = t + (size_t)u * T;
        if (i < nvec) {
            if (L2_256) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.u32 {%0,%1}, [%2];" : "=r"(w[u].x), "=r"(w[u].y) : "l"(in + (i << 1)));
            else w[u] = ld_stream8(in + (i << 1));
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream32(out + (i << 5), make_uint4(w[u].x, w[u].x + 1, w[u].x + 2, w[u].x + 3), make_uint4(w[u].y, w[u].y + 1, w[u].y + 2, w[u].y + 3)); }
}

// decode-like mix with the (small) input stream pulled into L2 in bursts: blocks are grouped G at a time;
// the first P blocks of a group each prefetch 1/P of the NEXT group's input with cp.async.bulk.prefetch.L2,
// so DRAM sees one concentrated read burst per group instead of reads sprinkled into the write stream.
template <int U, int T>
__global__ void __launch_bounds__(T) k_mix14_pf(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, size_t nvec, unsigned G, unsigned P, unsigned dist)
{
    const size_t tile_in_bytes = (size_t)T * U * 8;                     // input bytes per block (v32: 8 B per thread-step)
    const unsigned g = blockIdx.x / G, r = blockIdx.x % G;
    if (r < P && threadIdx.x == 0) {
        size_t group_bytes = (size_t)G * tile_in_bytes;
        size_t seg = (group_bytes / P) & ~(size_t)15;
        size_t start = (size_t)(g + dist) * group_bytes + (size_t)r * seg;
        size_t total = ((nvec * 8) & ~(size_t)15);
        if (start + seg <= total)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t *>(in) + start), "r"((unsigned)seg) : "memory");
    }
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    uint2 w[U];
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) w[u] = ld_stream8(in + (i << 1)); }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream32(out + (i << 5), make_uint4(w[u].x, w[u].x + 1, w[u].x + 2, w[u].x + 3), make_uint4(w[u].y, w[u].y + 1, w[u].y + 2, w[u].y + 3)); }
}

// encode-like mix with the OUTPUT kept in L2 longer?  Not controllable; instead try bursting reads the same way.
template <int U, int T>
__global__ void __launch_bounds__(T) k_mix41_pf(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, size_t nvec, unsigned G, unsigned P, unsigned dist)
{
    const size_t tile_in_bytes = (size_t)T * U * 32;
    const unsigned g = blockIdx.x / G, r = blockIdx.x % G;
    if (r < P && threadIdx.x == 0) {
        size_t group_bytes = (size_t)G * tile_in_bytes;
        size_t seg = (group_bytes / P) & ~(size_t)15;
        size_t start = (size_t)(g + dist) * group_bytes + (size_t)r * seg;
        if (start + seg <= nvec * 32)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + start), "r"((unsigned)seg) : "memory");
    }
    size_t t = (size_t)blockIdx.x * (T * U) + threadIdx.x;
    u32x8 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) v[u] = ld_stream32(in + (i << 5)); }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = t + (size_t)u * T; if (i < nvec) st_stream8(out + (i << 1), v[u].v[0] ^ v[u].v[1] ^ v[u].v[2] ^ v[u].v[3], v[u].v[4] ^ v[u].v[5] ^ v[u].v[6] ^ v[u].v[7]); }
}

// decode-like and encode-like mixes with the CTA order remapped into S interleaved streams: CTA b works on tile
// (b % S) * (nblocks / S) + b / S, so at any time the chip touches S far-apart regions instead of one linear front.
template <int T>
__global__ void __launch_bounds__(T) k_mix14_remap(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, size_t nvec, unsigned S)
{
    const size_t per = gridDim.x / S;
    const size_t tile = (size_t)(blockIdx.x % S) * per + blockIdx.x / S;
    size_t i = tile * T + threadIdx.x;
    if (blockIdx.x >= per * S) i = (size_t)blockIdx.x * T + threadIdx.x;
    if (i < nvec) {
        uint2 w = ld_stream8(in + (i << 1));
        st_stream32(out + (i << 5), make_uint4(w.x, w.x + 1, w.x + 2, w.x + 3), make_uint4(w.y, w.y + 1, w.y + 2, w.y + 3));
    }
}
template <int T>
__global__ void __launch_bounds__(T) k_mix41_remap(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, size_t nvec, unsigned S)
{
    const size_t per = gridDim.x / S;
    const size_t tile = (size_t)(blockIdx.x % S) * per + blockIdx.x / S;
    size_t i = tile * T + threadIdx.x;
    if (blockIdx.x >= per * S) i = (size_t)blockIdx.x * T + threadIdx.x;
    if (i < nvec) {
        u32x8 v = ld_stream32(in + (i << 5));
        st_stream8(out + (i << 1), v.v[0] ^ v.v[1] ^ v.v[2] ^ v.v[3], v.v[4] ^ v.v[5] ^ v.v[6] ^ v.v[7]);
    }
}

// ---- TMA (cp.async.bulk) variants with the real codec arithmetic --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// encode: persistent CTAs; thread 0 streams TILE-byte ASCII tiles into a STAGES-deep smem ring with
// cp.async.bulk; all T threads pack from smem and store packed u32 straight to global.
template <int TILE, int STAGES, int T>
__global__ void __launch_bounds__(T) k_enc_tma(const uint8_t *__restrict__ in, uint32_t *__restrict__ out32, size_t ntiles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const size_t first = blockIdx.x, stride = gridDim.x;
    size_t my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    // prologue: fill the ring
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && (size_t)s < my; s++) { mbar_expect_tx(&full[s], TILE); bulk_g2s(smem + (size_t)s * TILE, in + (first + (size_t)s * stride) * TILE, TILE, &full[s]); }
    for (size_t k = 0; k < my; k++) {
        const int s = (int)(k % STAGES);
        const unsigned parity = (unsigned)((k / STAGES) & 1);
        mbar_wait(&full[s], parity);
        const uint4 *tile = reinterpret_cast<const uint4 *>(smem + (size_t)s * TILE);
        uint32_t *dst = out32 + (first + k * stride) * (TILE / 16);
#pragma unroll
        for (int j = 0; j < TILE / 16 / T; j++) st_stream4(dst + j * T + threadIdx.x, pack16(tile[j * T + threadIdx.x]));
        __syncthreads();                                   // everyone is done reading stage s
        if (threadIdx.x == 0 && k + STAGES < my) { mbar_expect_tx(&full[s], TILE); bulk_g2s(smem + (size_t)s * TILE, in + (first + (k + STAGES) * stride) * TILE, TILE, &full[s]); }
    }
}

// decode: persistent CTAs; threads load packed u32, expand into a smem tile, thread 0 bulk-stores the tile.
template <int TILE, int STAGES, int T>
__global__ void __launch_bounds__(T) k_dec_tma(const uint32_t *__restrict__ bits32, uint8_t *__restrict__ out, size_t ntiles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const size_t first = blockIdx.x, stride = gridDim.x;
    size_t my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    for (size_t k = 0; k < my; k++) {
        const int s = (int)(k % STAGES);
        const size_t tile_idx = first + k * stride;
        const uint32_t *src = bits32 + tile_idx * (TILE / 16);
        uint32_t w[TILE / 16 / T];
#pragma unroll
        for (int j = 0; j < TILE / 16 / T; j++) w[j] = ld_stream4(src + j * T + threadIdx.x);
        // stage s was last used by the bulk store issued STAGES iterations ago: wait until it has been read
        if (threadIdx.x == 0) bulk_wait_read<STAGES - 1>();
        __syncthreads();
        uint4 *tile = reinterpret_cast<uint4 *>(smem + (size_t)s * TILE);
#pragma unroll
        for (int j = 0; j < TILE / 16 / T; j++) tile[j * T + threadIdx.x] = unpack16(w[j]);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_s2g(out + tile_idx * TILE, tile, TILE); bulk_commit(); }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();
}

// ---- harness -----------------------------------------------------------------------------------------
static cudaEvent_t e0, e1;
template <typename F> static float time_ms(F launch, int iters = 10, int reps = 3)
{
    float best = 1e30f;
    for (int r = 0; r <= reps; r++) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; i++) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms / iters < best) best = ms / iters;
    }
    return best;
}
static void report(const char *name, const char *variant, double bytes, float ms)
{
    printf("{\"kernel\": \"%s\", \"variant\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}\n", name, variant, ms, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
}
static unsigned blocks_for(size_t nvec, int tile) { return (unsigned)((nvec + tile - 1) / tile); }

int main(int argc, char **argv)
{
    const size_t gib = argc > 1 ? (size_t)atoll(argv[1]) : 8;
    const size_t L = gib << 30;                  // bytes on the "big" side
    uint8_t *big, *big2; uint32_t *small;
    CK(cudaMalloc(&big, L)); CK(cudaMalloc(&big2, L)); CK(cudaMalloc(&small, L / 4 + 64));
    CK(cudaMemset(big, 0x41, L)); CK(cudaMemset(big2, 0x43, L)); CK(cudaMemset(small, 0x1B, L / 4));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t n16 = L / 16, n32 = L / 32;

#define RUN(NAME, VAR, BYTES, ...) report(NAME, VAR, BYTES, time_ms([&] { __VA_ARGS__; }))
    // pure read
    RUN("read", "v16 u4 t256", (double)L, (k_read<16, 4, 256><<<blocks_for(n16, 1024), 256>>>(big, small, n16)));
    RUN("read", "v32 u1 t256", (double)L, (k_read<32, 1, 256><<<blocks_for(n32, 256), 256>>>(big, small, n32)));
    RUN("read", "v32 u2 t256", (double)L, (k_read<32, 2, 256><<<blocks_for(n32, 512), 256>>>(big, small, n32)));
    // pure write
    RUN("write", "v16 u4 t256", (double)L, (k_write<16, 4, 256><<<blocks_for(n16, 1024), 256>>>(big, n16)));
    RUN("write", "v32 u1 t256", (double)L, (k_write<32, 1, 256><<<blocks_for(n32, 256), 256>>>(big, n32)));
    RUN("write", "cudaMemsetAsync", (double)L, CK(cudaMemsetAsync(big, 1, L)));
    // copy 1:1
    RUN("copy", "v16 u4 t256", 2.0 * L, (k_copy<16, 4, 256><<<blocks_for(n16, 1024), 256>>>(big, big2, n16)));
    RUN("copy", "v32 u1 t256", 2.0 * L, (k_copy<32, 1, 256><<<blocks_for(n32, 256), 256>>>(big, big2, n32)));
    RUN("copy", "cudaMemcpyAsync d2d", 2.0 * L, CK(cudaMemcpyAsync(big2, big, L, cudaMemcpyDeviceToDevice)));
    // encode-like 4:1
    RUN("mix41", "ld.nc.na u4", 1.25 * L, (k_mix41<LD_NC_NA, 4, 256><<<blocks_for(n16, 1024), 256>>>(big, small, n16)));
    RUN("mix41", "ld.nc.na u2", 1.25 * L, (k_mix41<LD_NC_NA, 2, 256><<<blocks_for(n16, 512), 256>>>(big, small, n16)));
    RUN("mix41", "ld plain u2", 1.25 * L, (k_mix41<LD_PLAIN, 2, 256><<<blocks_for(n16, 512), 256>>>(big, small, n16)));
    RUN("mix41", "ld.nc u2", 1.25 * L, (k_mix41<LD_NC, 2, 256><<<blocks_for(n16, 512), 256>>>(big, small, n16)));
    RUN("mix41", "ld.cs u2", 1.25 * L, (k_mix41<LD_CS, 2, 256><<<blocks_for(n16, 512), 256>>>(big, small, n16)));
    // decode-like 1:4
    RUN("mix14", "st.na u2", 1.25 * L, (k_mix14<ST_NA, 2, 256><<<blocks_for(n16, 512), 256>>>(small, big, n16)));
    RUN("mix14", "st.na u4", 1.25 * L, (k_mix14<ST_NA, 4, 256><<<blocks_for(n16, 1024), 256>>>(small, big, n16)));
    RUN("mix14", "st plain u2", 1.25 * L, (k_mix14<ST_PLAIN, 2, 256><<<blocks_for(n16, 512), 256>>>(small, big, n16)));
    RUN("mix14", "st.cs u2", 1.25 * L, (k_mix14<ST_CS, 2, 256><<<blocks_for(n16, 512), 256>>>(small, big, n16)));
    RUN("mix14", "st.wt u2", 1.25 * L, (k_mix14<ST_WT, 2, 256><<<blocks_for(n16, 512), 256>>>(small, big, n16)));
    RUN("mix14", "st L2::evict_first u2", 1.25 * L, (k_mix14<ST_EVICT_FIRST, 2, 256><<<blocks_for(n16, 512), 256>>>(small, big, n16)));

    // 256-bit mixes
    RUN("mix41_v32", "ld.nc.na.v8 u1", 1.25 * L, (k_mix41_v32<true, 1, 256><<<blocks_for(n32, 256), 256>>>(big, small, n32)));
    RUN("mix41_v32", "ld plain v8 u1", 1.25 * L, (k_mix41_v32<false, 1, 256><<<blocks_for(n32, 256), 256>>>(big, small, n32)));
    RUN("mix41_v32", "ld plain v8 u2", 1.25 * L, (k_mix41_v32<false, 2, 256><<<blocks_for(n32, 512), 256>>>(big, small, n32)));
    RUN("mix14_v32", "st.na.v8 u1", 1.25 * L, (k_mix14_v32<1, 256, false><<<blocks_for(n32, 256), 256>>>(small, big, n32)));
    RUN("mix14_v32", "st.na.v8 u2", 1.25 * L, (k_mix14_v32<2, 256, false><<<blocks_for(n32, 512), 256>>>(small, big, n32)));
    RUN("mix14_v32", "st.na.v8 u1 ld L2::256B", 1.25 * L, (k_mix14_v32<1, 256, true><<<blocks_for(n32, 256), 256>>>(small, big, n32)));
    {
        const unsigned Gs[] = {2048, 8192, 32768}; const unsigned Ps[] = {64, 512}; const unsigned Ds[] = {1, 2};
        for (unsigned G : Gs) for (unsigned P : Ps) for (unsigned D : Ds) {
            char var[96]; snprintf(var, sizeof var, "v32 u1 burst-prefetch G=%u (%.1f MB in) P=%u dist=%u", G, G * 256 * 8 / 1e6, P, D);
            RUN("mix14_pf", var, 1.25 * L, (k_mix14_pf<1, 256><<<blocks_for(n32, 256), 256>>>(small, big, n32, G, P, D)));
        }
        for (unsigned G : Gs) for (unsigned P : Ps) {
            char var[96]; snprintf(var, sizeof var, "v32 u1 burst-prefetch G=%u (%.1f MB in) P=%u dist=1", G, G * 256 * 32 / 1e6, P);
            RUN("mix41_pf", var, 1.25 * L, (k_mix41_pf<1, 256><<<blocks_for(n32, 256), 256>>>(big, small, n32, G, P, 1)));
        }
    }

    for (unsigned S : {1u, 2u, 8u, 64u, 1024u}) {
        char var[64]; snprintf(var, sizeof var, "v32 u1 CTA order remapped into %u streams", S);
        RUN("mix14_remap", var, 1.25 * L, (k_mix14_remap<256><<<blocks_for(n32, 256), 256>>>(small, big, n32, S)));
        RUN("mix41_remap", var, 1.25 * L, (k_mix41_remap<256><<<blocks_for(n32, 256), 256>>>(big, small, n32, S)));
    }

    // TMA variants (real arithmetic).  persistent grid: CTAS per SM x 148
    {
        constexpr int TILE = 16384, T = 256;
        const size_t ntiles = L / TILE;
#define ENC_TMA(STAGES, PERSM) do { \
        auto kern = k_enc_tma<TILE, STAGES, T>; \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * STAGES)); \
        char var[64]; snprintf(var, sizeof var, "tile16K stages%d %dcta/sm", STAGES, PERSM); \
        RUN("enc_tma", var, 1.25 * L, (kern<<<148 * PERSM, T, TILE * STAGES>>>(big, small, ntiles))); } while (0)
        ENC_TMA(2, 4); ENC_TMA(3, 4); ENC_TMA(4, 3); ENC_TMA(6, 2); ENC_TMA(4, 2); ENC_TMA(8, 1);
#define DEC_TMA(STAGES, PERSM) do { \
        auto kern = k_dec_tma<TILE, STAGES, T>; \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * STAGES)); \
        char var[64]; snprintf(var, sizeof var, "tile16K stages%d %dcta/sm", STAGES, PERSM); \
        RUN("dec_tma", var, 1.25 * L, (kern<<<148 * PERSM, T, TILE * STAGES>>>(small, big, ntiles))); } while (0)
        DEC_TMA(2, 4); DEC_TMA(3, 4); DEC_TMA(4, 3); DEC_TMA(2, 6); DEC_TMA(4, 2);
    }
    return 0;
}
