// Micro-benchmarks that ground the recurrence design (run on the B200 box):
//   1. mma.sync tf32 / bf16 issue rates
//   2. tcgen05.mma with A in TMEM (TS form), tiny N: numerics (truncation of raw fp32 -> tf32?) and latency
//   3. DSMEM all-gather step latency inside a cluster of 8 / 16 CTAs
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/ubench scratch/ubench.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ 1. mma.sync rates
template <int KIND>
__global__ void __launch_bounds__(256, 1) mma_rate_kernel(float* out, long long* cyc, int iters) {
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// mixed tf32 / bf16 issue patterns: PAT 0 = TTBB TTBB (the recurrence k-loop), 1 = TBTB, 2 = 8T then 8B
template <int PAT>
__global__ void __launch_bounds__(256, 1) mma_mix_kernel(float* out, long long* cyc, int iters) {
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
#define TMMA(c) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]))
#define BMMA(c) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]))
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (PAT == 0) { TMMA(0); TMMA(1); BMMA(2); BMMA(3); TMMA(0); TMMA(1); BMMA(2); BMMA(3); }
        if (PAT == 1) { TMMA(0); BMMA(2); TMMA(1); BMMA(3); TMMA(0); BMMA(2); TMMA(1); BMMA(3); }
        if (PAT == 2) { TMMA(0); TMMA(1); TMMA(0); TMMA(1); BMMA(2); BMMA(3); BMMA(2); BMMA(3); }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// realistic operand traffic: 32 distinct B fragments and 4 distinct A fragments held in registers
__global__ void __launch_bounds__(256, 1) mma_regs_kernel(const uint32_t* src, float* out, long long* cyc, int iters) {
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4][4], b[32][2];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = src[(threadIdx.x * 16 + i * 4 + j) % 4096];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 2; ++j) b[i][j] = src[(threadIdx.x * 64 + i * 2 + j + 7) % 4096];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#define TM(c, ai, bi) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a[ai][0]), "r"(a[ai][1]), "r"(a[ai][2]), "r"(a[ai][3]), "r"(b[bi][0]), "r"(b[bi][1]))
#define BM(c, ai, bi) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a[ai][0]), "r"(a[ai][1]), "r"(a[ai][2]), "r"(a[ai][3]), "r"(b[bi][0]), "r"(b[bi][1]))
            TM(0, 0, 4 * k + 0); TM(1, 0, 4 * k + 1); BM(2, 1, 4 * k + 2); BM(3, 1, 4 * k + 3);
            TM(0, 2, 4 * k + 1); TM(1, 2, 4 * k + 0); BM(2, 3, 4 * k + 3); BM(3, 3, 4 * k + 2);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ------------------------------------------------------------------ 2. tcgen05 TS
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > (1ll << 31)) __trap();
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)lt << 61;
    return d;
}
__device__ __forceinline__ void umma_ts_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_bf16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// W [128][256] fp32 (raw, low bits set) -> TMEM cols [0,256) ; Wb [128][256] bf16 -> TMEM cols [256,384)
// X [NB][256] fp32 -> smem no-swizzle K-major tf32 operand; Xb bf16 operand.
// D1 (cols 384..384+NB) = W * X^T (tf32), D2 (cols 448..) = Wb * Xb^T (bf16)
// timing: repeat { 32 tf32 MMAs (N=NB) + 16 bf16 MMAs (N=16) ; commit ; wait } iters times
template <int NB>
__global__ void __launch_bounds__(128, 1) ts_kernel(const float* W, const __nv_bfloat16* Wb, const float* X,
                                                    const __nv_bfloat16* Xb, float* D1, float* D2, long long* cyc,
                                                    int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* xs = smem;                         // tf32 operand: 64 k-chunks x NB rows x 16 B
    uint8_t* xbs = smem + 64 * NB * 16;         // bf16 operand: 32 k-chunks x 16 rows x 16 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    // A operands into TMEM
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(W[(size_t)tid * 256 + c0 + j]);
        tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    }
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        const uint32_t* wb = reinterpret_cast<const uint32_t*>(Wb + (size_t)tid * 256);
        for (int j = 0; j < 32; ++j) r[j] = wb[c0 + j];
        tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 256 + c0, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // B operands into smem (canonical no-swizzle K-major)
    for (int i = tid; i < NB * 256; i += 128) {
        int n = i / 256, k = i % 256;
        *reinterpret_cast<float*>(xs + (k / 4) * (NB * 16) + n * 16 + (k % 4) * 4) = X[i];
    }
    for (int i = tid; i < 16 * 256; i += 128) {
        int n = i / 256, k = i % 256;
        *reinterpret_cast<__nv_bfloat16*>(xbs + (k / 8) * (16 * 16) + n * 16 + (k % 8) * 2) = Xb[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t ph = 0;
    long long t_issue = 0, t_total = 0;
    for (int it = 0; it < iters; ++it) {
        if (tid == 0) {
            long long t0 = clock64();
            const uint32_t xa = s_u32(xs), xba = s_u32(xbs);
#pragma unroll
            for (int k = 0; k < 32; ++k)
                umma_ts_tf32(tm + 384, tm + k * 8, make_sdesc(xa + k * 2 * (NB * 16), NB * 16, 128, 0), idesc_tf32, k != 0);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                umma_ts_bf16(tm + 448, tm + 256 + k * 8, make_sdesc(xba + k * 2 * 256, 256, 128, 0), idesc_bf16, k != 0);
            long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, ph);
            long long t2 = clock64();
            t_issue += t1 - t0;
            t_total += t2 - t0;
        }
        ph ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back
    for (int c0 = 0; c0 < NB; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 384 + c0, r);
        for (int j = 0; j < 16; ++j) D1[(size_t)tid * NB + c0 + j] = __uint_as_float(r[j]);
    }
    {
        uint32_t r[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 448, r);
        for (int j = 0; j < 16; ++j) D2[(size_t)tid * 16 + j] = __uint_as_float(r[j]);
    }
    if (tid == 0) { cyc[0] = t_issue / iters; cyc[1] = t_total / iters; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// ------------------------------------------------------------------ 3. DSMEM all-gather
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t a, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t a, float4 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t a) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MODE 0: all threads store, __syncthreads, CS threads arrive.  MODE 1: per-warp arrive (no CTA barrier):
// each warp pushes whole rows of work to peers and lane 0 arrives after __syncwarp (count = warps * CS... see below)
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t bar_cluster, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(s_u32(src)), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bulk_multicast(uint32_t dst_local, const void* gsrc, uint32_t bytes, uint32_t bar_local, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_local), "l"(gsrc), "r"(bytes), "r"(bar_local), "h"(mask) : "memory");
}
__device__ float g_xchg[2 * 256 * 4096];     // [buf][cta][tile floats]

template <int CS, int MODE>
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int tile_bytes, int steps, long long* cyc, float* sink) {
    extern __shared__ __align__(16) uint8_t sm[];
    float* recv = reinterpret_cast<float*>(sm);                       // [2][CS][tile]
    float* mine = recv + 2 * CS * (tile_bytes / 4);                   // [tile]
    __shared__ __align__(8) uint64_t full[2];
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    const int nvec = tile_bytes / 16;
    const int total = nvec * CS;                                      // (vector, peer) pairs
    if (tid == 0) {
        const int cnt = MODE == 0 ? CS : (MODE == 1 ? CS * 8 : 1);
        (void)cnt;
        mbar_init(&full[0], cnt); mbar_init(&full[1], cnt);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (MODE >= 2) { mbar_expect_tx(&full[0], CS * tile_bytes); mbar_expect_tx(&full[1], CS * tile_bytes); }
    }
    for (int i = tid; i < tile_bytes / 4; i += 256) mine[i] = (float)(i + rank);
    __syncthreads();
    cluster_sync_all();
    uint32_t ph[2] = {0, 0};
    float acc = 0.f;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        const uint32_t slot = s_u32(recv + ((size_t)buf * CS + rank) * (tile_bytes / 4));
        if (MODE == 0) {
            for (int i = tid; i < total; i += 256) {
                const int peer = i / nvec, v = i % nvec;
                st_cluster_v4(mapa(slot + v * 16, peer), reinterpret_cast<const float4*>(mine)[v]);
            }
            __syncthreads();
            if (tid < CS) mbar_arrive_cluster(mapa(s_u32(&full[buf]), tid));
        } else if (MODE == 2) {
            // st.async: data + transaction count in one remote operation, no fences, no CTA barrier
            const uint32_t bar = s_u32(&full[buf]);
            for (int i = tid; i < total; i += 256) {
                const int peer = i % CS, v = i / CS;
                st_async_v4(mapa(slot + v * 16, peer), reinterpret_cast<const float4*>(mine)[v], mapa(bar, peer));
            }
        } else if (MODE == 3) {
            // bulk DMA copy of the whole tile to each peer (tile already staged in `mine`)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < CS) dsmem_bulk_copy(mapa(slot, tid), mapa(s_u32(&full[buf]), tid), mine, tile_bytes);
        } else if (MODE == 4) {
            // tile -> global (L2) -> one multicast bulk copy into every CTA of the cluster
            float* g = g_xchg + ((size_t)buf * gridDim.x + blockIdx.x) * (tile_bytes / 4);
            if (tid < nvec) reinterpret_cast<float4*>(g)[tid] = reinterpret_cast<const float4*>(mine)[tid];
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncthreads();
            if (tid == 0) bulk_multicast(slot, g, tile_bytes, s_u32(&full[buf]), (uint16_t)((1u << CS) - 1));
        } else {
            // warp w pushes vectors {w, w+8, ...} of the tile to every peer; lane -> (peer = lane % CS, sub = lane / CS)
            const int w = tid / 32, lane = tid % 32;
            constexpr int SUB = 32 / CS;
            const int peer = lane % CS, sub = lane / CS;
            for (int v = w * SUB + sub; v < nvec; v += 8 * SUB)
                st_cluster_v4(mapa(slot + v * 16, peer), reinterpret_cast<const float4*>(mine)[v]);
            __syncwarp();
            if (lane < CS) mbar_arrive_cluster(mapa(s_u32(&full[buf]), lane));
        }
        {
            const long long tw = clock64();
            while (!mbar_try_wait_cluster(&full[buf], ph[buf])) if (clock64() - tw > (1ll << 31)) __trap();
        }
        ph[buf] ^= 1;
        if (MODE >= 2 && tid == 0) mbar_expect_tx(&full[buf], CS * tile_bytes);   // arm the next phase of this buffer
        // consume something from every slot so the data path is real
        acc += recv[((size_t)buf * CS + (tid % CS)) * (tile_bytes / 4) + (tid / CS) % (tile_bytes / 4)];
        mine[tid % (tile_bytes / 4)] = acc * 1e-9f;
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync_all();
    sink[blockIdx.x * 256 + tid] = acc;
    if (blockIdx.x == 0 && tid == 0) *cyc = (t1 - t0) / steps;
}

template <int CS, int MODE>
void run_dsmem(int nclusters, int tile_bytes, long long* d_cyc, float* d_sink) {
    auto k = dsmem_kernel<CS, MODE>;
    size_t smem = (size_t)(2 * CS + 1) * tile_bytes;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (CS > 8) CK(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CS);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 200 * 1024;       // as much smem as the real kernel: 1 CTA / SM
    (void)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int maxc = 0;
    CK(cudaOccupancyMaxActiveClusters(&maxc, k, &cfg));
    CK(cudaLaunchKernelEx(&cfg, k, tile_bytes, 2000, d_cyc, d_sink));
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("dsmem allgather CS=%2d mode=%d clusters=%2d (max active %d) tile=%5d B: %lld cycles/step\n", CS, MODE, nclusters, maxc, tile_bytes, c);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int NB>
void run_ts() {
    std::vector<float> W(128 * 256), X(NB * 256);
    std::vector<__nv_bfloat16> Wb(128 * 256), Xb(16 * 256);
    srand(1);
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.15f;
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (size_t i = 0; i < Wb.size(); ++i) Wb[i] = __float2bfloat16(W[i] * 0.01f);
    for (size_t i = 0; i < Xb.size(); ++i) Xb[i] = __float2bfloat16(X[i % X.size()]);
    float *dW, *dX, *dD1, *dD2; __nv_bfloat16 *dWb, *dXb; long long* dc;
    CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dWb, Wb.size() * 2)); CK(cudaMalloc(&dXb, Xb.size() * 2));
    CK(cudaMalloc(&dD1, 128 * NB * 4)); CK(cudaMalloc(&dD2, 128 * 16 * 4)); CK(cudaMalloc(&dc, 16));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dWb, Wb.data(), Wb.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dXb, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice));
    size_t smem = 64 * NB * 16 + 32 * 16 * 16 + 1024;
    CK(cudaFuncSetAttribute(ts_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ts_kernel<NB><<<1, 128, smem>>>(dW, dWb, dX, dXb, dD1, dD2, dc, 200);
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * NB), D2(128 * 16); long long c[2];
    CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost));
    // references: truncation and round-to-nearest models of fp32 -> tf32
    auto trunc13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    auto rn13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    double e_tr = 0, e_rn = 0, e_full = 0, mx = 0, e_bf = 0, mxb = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
            double st = 0, sr = 0, sf = 0;
            for (int k = 0; k < 256; ++k) {
                st += (double)trunc13(W[m * 256 + k]) * trunc13(X[n * 256 + k]);
                sr += (double)rn13(W[m * 256 + k]) * rn13(X[n * 256 + k]);
                sf += (double)W[m * 256 + k] * X[n * 256 + k];
            }
            double g = D1[m * NB + n];
            e_tr = fmax(e_tr, fabs(g - st)); e_rn = fmax(e_rn, fabs(g - sr)); e_full = fmax(e_full, fabs(g - sf)); mx = fmax(mx, fabs(sf));
        }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            double s = 0;
            for (int k = 0; k < 256; ++k) s += (double)__bfloat162float(Wb[m * 256 + k]) * __bfloat162float(Xb[n * 256 + k]);
            e_bf = fmax(e_bf, fabs(D2[m * 16 + n] - s)); mxb = fmax(mxb, fabs(s));
        }
    printf("tcgen05 TS NB=%d: issue %lld cyc, issue+complete %lld cyc (32 tf32 N=%d + 16 bf16 N=16 MMAs)\n", NB, c[0], c[1], NB);
    printf("   tf32 max|D|=%.4f err vs trunc-model %.3e, vs rn-model %.3e, vs exact fp32 %.3e\n", mx, e_tr, e_rn, e_full);
    printf("   bf16 max|D|=%.6f err %.3e\n", mxb, e_bf);
    (void)bf16_round;
}

int main() {
    long long* d_cyc; float* d_out;
    CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_out, 148 * 256 * 4 * 4));
    for (int kind = 0; kind < 2; ++kind) {
        const int iters = 2000;
        if (kind == 0) mma_rate_kernel<0><<<148, 256>>>(d_out, d_cyc, iters); else mma_rate_kernel<1><<<148, 256>>>(d_out, d_cyc, iters);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        // 8 warps / CTA = 2 per SMSP, 4 MMAs per iter per warp
        printf("mma.sync %s: %.2f cycles per MMA per SMSP (2 warps/SMSP, 4 independent chains)\n", kind == 0 ? "m16n8k8 tf32" : "m16n8k16 bf16",
               (double)c / (iters * 4 * 2));
    }
    for (int pat = 0; pat < 3; ++pat) {
        const int iters = 2000;
        if (pat == 0) mma_mix_kernel<0><<<148, 256>>>(d_out, d_cyc, iters);
        if (pat == 1) mma_mix_kernel<1><<<148, 256>>>(d_out, d_cyc, iters);
        if (pat == 2) mma_mix_kernel<2><<<148, 256>>>(d_out, d_cyc, iters);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("mma.sync mixed pattern %d: %.2f cycles per MMA per SMSP\n", pat, (double)c / (iters * 8 * 2));
    }
    {
        uint32_t* d_src; CK(cudaMalloc(&d_src, 4096 * 4)); CK(cudaMemset(d_src, 0x3c, 4096 * 4));
        mma_regs_kernel<<<148, 256>>>(d_src, d_out, d_cyc, 500);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("mma.sync 32 distinct B fragments in registers: %.2f cycles per MMA per SMSP\n", (double)c / (500 * 64 * 2));
    }
    run_ts<16>();
    run_ts<32>();
    run_dsmem<16, 3>(7, 1024, d_cyc, d_out);
    run_dsmem<16, 4>(7, 1024, d_cyc, d_out);
    run_dsmem<16, 3>(7, 256, d_cyc, d_out);
    run_dsmem<16, 4>(7, 256, d_cyc, d_out);
    run_dsmem<8, 3>(8, 2048, d_cyc, d_out);
    run_dsmem<8, 4>(8, 2048, d_cyc, d_out);
    run_dsmem<2, 3>(8, 1024, d_cyc, d_out);
    run_dsmem<2, 4>(8, 1024, d_cyc, d_out);
    run_dsmem<16, 0>(7, 1024, d_cyc, d_out);
    run_dsmem<16, 2>(7, 1024, d_cyc, d_out);
    run_dsmem<16, 2>(8, 1024, d_cyc, d_out);
    run_dsmem<16, 2>(8, 256, d_cyc, d_out);
    run_dsmem<8, 2>(15, 1024, d_cyc, d_out);
    run_dsmem<8, 2>(8, 2048, d_cyc, d_out);
    run_dsmem<4, 2>(8, 1024, d_cyc, d_out);
    run_dsmem<2, 2>(8, 1024, d_cyc, d_out);
    run_dsmem<8, 0>(16, 1024, d_cyc, d_out);
    run_dsmem<8, 1>(16, 1024, d_cyc, d_out);
    run_dsmem<8, 0>(8, 2048, d_cyc, d_out);
    run_dsmem<8, 1>(8, 2048, d_cyc, d_out);
    run_dsmem<16, 0>(8, 1024, d_cyc, d_out);
    run_dsmem<16, 1>(8, 1024, d_cyc, d_out);
    run_dsmem<8, 1>(1, 1024, d_cyc, d_out);
    run_dsmem<8, 1>(16, 256, d_cyc, d_out);
    return 0;
}
