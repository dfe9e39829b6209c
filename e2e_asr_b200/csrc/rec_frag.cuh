// Shared device helpers of the register-resident LSTM recurrence kernels (lstm_rec_ws.cu, lstm_rec_h512.cu):
// cluster / mbarrier / bulk-copy PTX wrappers, the 3xTF32 and mixed tf32+bf16 fragment products.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace e2e {
namespace {

struct MParams {
    float* G;
    float* Hout;
    float* Cst;
    const float* Wh;
    const float* dOut;
    const int* lens;
    float* xg;            // forward: global exchange tiles [NS][2][grid][256]
    int B, T, Tp, H, ndir;
    int nslices;          // 16-row batch slices per direction
    long long sb, st;
    long long* dbg;       // optional clock64 stamps of CTA 0 / thread 0: [step][slice][8]
    int carry_c;          // forward, ndir = 1: the cell state entering step 0 is Cst at time -1 (a continued sequence)
};

constexpr int R = 16, UPC = 16, NTH = 256;
constexpr int TILE = R * UPC;                       // floats in one exchanged [16 rows][16 units] tile (1 KB)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 33)) __trap();      // never hang the GPU on a protocol bug
    }
}
// global -> the same shared-memory offset of every CTA in `mask`, completion on each one's mbarrier
__device__ __forceinline__ void bulk_multicast(uint32_t dst_local, const void* gsrc, uint32_t bytes, uint32_t bar_local,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(dst_local), "l"(gsrc), "r"(bytes), "r"(bar_local), "h"(mask) : "memory");
}
// local shared -> shared memory of another CTA of the cluster, completion on that CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t bar_cluster, const void* src,
                                                uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(s_u32(src)), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// 3xTF32: x = hi + lo, hi = x with the low 13 mantissa bits cleared (exactly a TF32 number; one LOP
// instead of a quarter-rate cvt), lo = x - hi (exact; the MMA keeps its top 11 bits: error 2^-21 |x|);
// D += lo*hi + hi*lo + hi*hi in fp32.  Relative error ~2^-20, inside the 1e-4 parity budget.
__device__ __forceinline__ uint32_t cvt_tf32(float x) {     // round-to-nearest TF32 (set-up time only)
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// accurate-enough fast transcendentals: ex2.approx + rcp.approx, absolute error ~2e-7
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

// One k8-tile of the contraction.  `a` is the thread's A fragment exactly as the MMA wants it --
// (row g, k tq), (row g+8, k tq), (row g, k tq+4), (row g+8, k tq+4) -- because the exchanged tiles are
// stored in fragment order ([k-tile][lane][4]): one conflict-free LDS.128, no register shuffling.
// The raw fp32 words serve as the "hi" operand (the tensor core ignores the low 13 mantissa bits).
template <int NTL>
__device__ __forceinline__ void ktile_mma(float (&acc)[NTL][4], float (&accx)[NTL][4], const float4 a,
                                          const uint32_t (&bh)[NTL][2], const uint32_t (&bl)[NTL][2]) {
    uint32_t ah[4], al[4];
    split_tf32(a.x, ah[0], al[0]); split_tf32(a.y, ah[1], al[1]); split_tf32(a.z, ah[2], al[2]); split_tf32(a.w, ah[3], al[3]);
    // products grouped by kind so that dependent MMAs on one accumulator are NTL issues apart
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_tf32(accx[nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_tf32(acc[nt], ah, bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_tf32(accx[nt], ah, bl[nt][0], bl[nt][1]);
}

// Mixed scheme (kMixed): bf16 m16n8k16 issues at the same 8 cycles / SMSP as tf32 m16n8k8, so the two
// cross terms of a k16 pair cost one MMA each instead of two: per k16 and n-tile
//     acc  += tf32(a) * tf32(W)            (2 x m16n8k8, a = raw fp32 words)
//     accx += bf16(a - tf32(a)) * bf16(W)  (1 x m16n8k16)
//     accx += bf16(a) * bf16(W - tf32(W))  (1 x m16n8k16)
// 4 MMAs instead of 6.  Error per product ~2^-19 (bf16 rounding, 2^-9, of an operand of a term that is
// itself 2^-10..2^-12 of the product), the same class as 3xTF32 with a truncating split.
// The bf16 MMA's logical k index (2tq, 2tq+1 | 2tq+8, 2tq+9) is mapped to the physical
// (tq, tq+4 of k8-tile 0 | tq, tq+4 of k8-tile 1), i.e. to exactly the values the thread already holds.
constexpr bool kMixed = true;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// bf16 pair by TRUNCATION (one PRMT on the ALU pipe): F2FP.BF16 runs on the quarter-rate conversion pipe that the
// gate nonlinearities (MUFU) also use -- 64 packs per warp and step made the k-loop conversion-bound (measured).
// Used for the per-step A operand only, where the truncated term is itself <= 2^-10 of the product
// (error <= 2^-18 relative); the resident W fragments keep round-to-nearest packs.
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// a0 / a1: the thread's A fragments of the pair's two k8-tiles; bh0 / bh1: tf32 W of the two tiles;
// wb / wl: bf16 W and bf16 (W - tf32 W) of the pair
template <int NTL>
__device__ __forceinline__ void k16_mma(float (&acc)[NTL][4], float (&accx)[NTL][4], const float4 a0, const float4 a1,
                                        const uint32_t (&bh0)[NTL][2], const uint32_t (&bh1)[NTL][2],
                                        const uint32_t (&wb)[NTL][2], const uint32_t (&wl)[NTL][2]) {
    uint32_t h0[4], l0[4], h1[4], l1[4];
    split_tf32(a0.x, h0[0], l0[0]); split_tf32(a0.y, h0[1], l0[1]); split_tf32(a0.z, h0[2], l0[2]); split_tf32(a0.w, h0[3], l0[3]);
    split_tf32(a1.x, h1[0], l1[0]); split_tf32(a1.y, h1[1], l1[1]); split_tf32(a1.z, h1[2], l1[2]); split_tf32(a1.w, h1[3], l1[3]);
    const uint32_t al[4] = {pack_bf16_trunc(__uint_as_float(l0[0]), __uint_as_float(l0[2])), pack_bf16_trunc(__uint_as_float(l0[1]), __uint_as_float(l0[3])),
                            pack_bf16_trunc(__uint_as_float(l1[0]), __uint_as_float(l1[2])), pack_bf16_trunc(__uint_as_float(l1[1]), __uint_as_float(l1[3]))};
    const uint32_t ab[4] = {pack_bf16_trunc(a0.x, a0.z), pack_bf16_trunc(a0.y, a0.w), pack_bf16_trunc(a1.x, a1.z), pack_bf16_trunc(a1.y, a1.w)};
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_tf32(acc[nt], h0, bh0[nt][0], bh0[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_bf16(accx[nt], al, wb[nt][0], wb[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_tf32(acc[nt], h1, bh1[nt][0], bh1[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_bf16(accx[nt], ab, wl[nt][0], wl[nt][1]);
}

// fp16 split scheme of the FORWARD recurrence: x = hi + lo' 2^-11 with hi = fp16(x) (11 significand bits) and
// lo' = fp16((x - hi) 2^11) -- the residual scaled into fp16's normal range, so |x - hi - lo' 2^-11| <= 2^-23 |x|.
// Per k16 and n-tile THREE m16n8k16 f16 MMAs:  acc += hi_a hi_w;  accx += lo'_a hi_w + hi_a lo'_w;  z = acc + 2^-11 accx
// (the dropped lo lo term is <= 2^-23 of the product: the accuracy class of 3xTF32) instead of the four of the
// tf32 + bf16 scheme above, and no per-step operand conversion in the k-loop: the h_t tiles are exchanged already
// split and packed in fragment order.  Only for operands with a bounded range (|h| < 1, weights): fp16 has 5
// exponent bits, so the backward pass (gradients span many orders of magnitude) keeps the tf32 + bf16 scheme.
constexpr float kF16LoScale = 2048.0f, kF16LoInv = 1.0f / 2048.0f;
__device__ __forceinline__ void split_f16(float x, unsigned short& hi, unsigned short& lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn((x - __half2float(h)) * kF16LoScale);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}
__device__ __forceinline__ uint32_t pack_u16(unsigned short lo, unsigned short hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint4 a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// ah / al: the thread's packed A fragments (hi and lo' halves) of one k16 pair; wh / wl: the resident W fragments
template <int NTL>
__device__ __forceinline__ void k16_mma_f16(float (&acc)[NTL][4], float (&accx)[NTL][4], const uint4 ah, const uint4 al,
                                            const uint32_t (&wh)[NTL][2], const uint32_t (&wl)[NTL][2]) {
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_f16(acc[nt], ah, wh[nt][0], wh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_f16(accx[nt], al, wh[nt][0], wh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) mma_f16(accx[nt], ah, wl[nt][0], wl[nt][1]);
}

}  // namespace
}  // namespace e2e
