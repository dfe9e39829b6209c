// Tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// memory -> tcgen05.mma with the accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]) (+ Z[M,N]) (+ C)
//
// Serves the dense contractions of the path (SURVEY.md 2.3 K1/K4/K7 and their
// dX/dW twins).  Two numeric modes:
//   mode 1 "tf32x3": fp32-accurate.  Each fp32 operand is split by a pre-pass into
//          big = x with the low 13 mantissa bits cleared (exactly a TF32 number) and
//          small = x - big (exact); the kernel accumulates
//          big*big + big*small + small*big in the fp32 TMEM accumulator
//          (kind::tf32), error ~2^-21 relative -- the north-star's 1e-4 budget.
//   mode 2 "bf16":   operands rounded to bf16 by the pre-pass, one kind::f16 MMA.
//   mode 3 "bf16x2": each fp32 operand is split into hi = bf16(x) and lo = bf16(x - hi) (16 significand bits);
//          the kernel accumulates hi*hi + hi*lo + lo*hi with three kind::f16 MMAs: error ~2^-17 relative per
//          product (well inside the 1e-4 budget) at HALF the tensor-pipe cost of 3xTF32, because kind::f16 runs
//          at twice the kind::tf32 rate.
//   mode 4 "f16x2":  each fp32 operand is split into hi = fp16(x) and lo' = fp16((x - hi) * 2^11) -- the residual scaled
//          into fp16's normal range, 22 significand bits in all -- and the kernel keeps TWO TMEM accumulators:
//          main += hi*hi, cross += hi*lo' + lo'*hi, C = main + cross * 2^-11: error ~2^-22 relative, the accuracy
//          class of 3xTF32, at the kind::f16 rate (twice kind::tf32) -- half the tensor-pipe cost.  fp16 has 5 exponent
//          bits, so the A operand (activations, or gradients spanning many orders of magnitude ACROSS rows) is scaled
//          per ROW by a power of two before the split and the output row is scaled back in the epilogue; B must be
//          bounded (weights).  A transposed A (weight-gradient products, whose K runs over the rows) stays on mode 1.
// All four transpose combinations are native: a row-major operand whose
// contiguous dimension is K is a K-major UMMA operand, otherwise an MN-major one
// (instruction-descriptor bits 15/16); TMA boxes always follow the contiguous
// dimension, so nothing is ever transposed in memory.
//
// CTA = one 128x128 output tile (x one K split): warp 0 = TMA producer, warp 1 =
// TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue (each owns the
// 32 TMEM lanes its warp-id%4 may access).  3-stage mbarrier ring, 64 KB/stage.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

namespace e2e {

namespace {

constexpr int TBM = 128, TBN = 128;
constexpr int STAGES = 3;
constexpr int OPER_BYTES = 16384;                  // one 128 x (128 B) operand tile
constexpr int STAGE_BYTES = 4 * OPER_BYTES;        // A_big, A_small, B_big, B_small (bf16: A, -, B, -)
constexpr int TC_THREADS = 192;

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must fault (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 33)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (BF16) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    }
}
// tcgen05.commit: the mbarrier gets one arrival when all previously issued MMAs retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (Blackwell version field = 1).  layout_type: 2 =
// SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks) -- the only
// layout the hardware accepts for MN-major 32-bit (tf32) operands.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                               uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;       // descriptor version
    d |= (uint64_t)layout_type << 61;
    return d;
}

struct TcParams {
    int M, N, K;
    float* C;
    int ldc;
    const float* bias;
    const float* Z;
    int ldz;
    int accumulate;
    int kblocks_per_split;     // K blocks (of BKE elements) handled by one blockIdx.z
    int a_mn_major, b_mn_major;
    const float* row_scale;    // mode 4: C row m is multiplied by row_scale[m] (undoes the per-row operand scaling), or NULL
    float* dbg;                // debug: if set, stage-0 smem (64 KB) and the raw TMEM tile are dumped here
};

// BF16: element = 2 bytes, 64 elements per 128 B, UMMA_K = 16;  TF32: 4 bytes, 32 per 128 B, UMMA_K = 8.
// MODE 1 = 3xTF32, 2 = bf16, 3 = bf16x2 (two bf16 parts per operand, three products), 4 = f16x2 (two fp16 parts, the
// low one scaled by 2^11, cross terms in a second accumulator).
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1, TcParams p) {
    constexpr bool BF16 = MODE >= 2;               // 2-byte operands (bf16 or fp16): kind::f16
    constexpr bool F16X2 = MODE == 4;
    constexpr int TMEM_COLS = F16X2 ? 256 : 128;   // f16x2: main accumulator in columns [0,128), cross terms in [128,256)
    constexpr int PARTS = MODE == 2 ? 1 : 2;       // operand parts staged per k-block
    constexpr int BKE = BF16 ? 64 : 32;            // K elements per stage (one 128-byte swizzle span)
    constexpr int ELT = BF16 ? 2 : 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
    const int kb_total = (p.K + BKE - 1) / BKE;
    const int kb_begin = blockIdx.z * p.kblocks_per_split;
    const int kb_end = min(kb_total, kb_begin + p.kblocks_per_split);
    const int nkb = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        // ================= TMA producer (one elected lane) =================
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                const int k0 = (kb_begin + i) * BKE;
                mbar_expect_tx(&full_bar[s], 2 * PARTS * OPER_BYTES);
#pragma unroll
                for (int part = 0; part < PARTS; ++part) {
                    const CUtensorMap* ma = part ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = part ? &mapB1 : &mapB0;
                    uint8_t* sa = st + part * OPER_BYTES;
                    uint8_t* sb = st + (2 + part) * OPER_BYTES;
                    if (!p.a_mn_major) {
                        tma_load_2d(ma, &full_bar[s], sa, k0, m0);                 // box {BKE (K), 128 rows of M}
                    } else {
                        // stored [K][M]: boxes {128B of M, BKE rows of K}, one per 128-byte M chunk
                        constexpr int CH = 128 / ELT;                               // M elements per chunk
                        for (int c = 0; c < TBM / CH; ++c)
                            tma_load_2d(ma, &full_bar[s], sa + c * (BKE * 128), m0 + c * CH, k0);
                    }
                    if (p.b_mn_major) {
                        constexpr int CH = 128 / ELT;
                        for (int c = 0; c < TBN / CH; ++c)
                            tma_load_2d(mb, &full_bar[s], sb + c * (BKE * 128), n0 + c * CH, k0);
                    } else {
                        tma_load_2d(mb, &full_bar[s], sb, k0, n0);                 // stored [N][K]
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (single thread) =================
        if (lane == 0) {
            // instruction descriptor: D=f32, A/B format, majors, N>>3, M>>4
            constexpr uint32_t FMT = F16X2 ? 0u : (BF16 ? 1u : 2u);      // kind::f16: 0 = F16, 1 = BF16; kind::tf32: 2
            uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) |
                             ((uint32_t)(p.a_mn_major ? 1 : 0) << 15) | ((uint32_t)(p.b_mn_major ? 1 : 0) << 16) |
                             ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            constexpr int UK = BF16 ? 16 : 8;                  // K elements per instruction
            constexpr int KSTEPS = BKE / UK;                   // 4
            // K-major: rows of 128 B, 8-row groups 1024 B apart; advance 32 B per k-step.
            // MN-major: atoms of (128 B of MN) x (8 k rows) = 1024 B; MN chunks BKE*128 B apart (LBO);
            //           8-row k groups 1024 B apart (SBO); advance UK rows = UK*128 B per k-step.
            //           tf32 MN-major uses the 32-byte-atom swizzle whose k groups are 4 rows = 512 B.
            const uint32_t a_lbo = p.a_mn_major ? BKE * 128 : 16, a_sbo = (!BF16 && p.a_mn_major) ? 512 : 1024;
            const uint32_t b_lbo = p.b_mn_major ? BKE * 128 : 16, b_sbo = (!BF16 && p.b_mn_major) ? 512 : 1024;
            const uint32_t a_lt = (!BF16 && p.a_mn_major) ? 1 : 2, b_lt = (!BF16 && p.b_mn_major) ? 1 : 2;
            const uint32_t a_step = p.a_mn_major ? UK * 128 : 32;
            const uint32_t b_step = p.b_mn_major ? UK * 128 : 32;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                    const uint64_t a0 = make_sdesc(st + k * a_step, a_lbo, a_sbo, a_lt);
                    const uint64_t b0 = make_sdesc(st + 2 * OPER_BYTES + k * b_step, b_lbo, b_sbo, b_lt);
                    if (PARTS == 1) {
                        umma<true>(tmem_d, a0, b0, idesc, (i | k) != 0);
                    } else {
                        const uint64_t a1 = make_sdesc(st + OPER_BYTES + k * a_step, a_lbo, a_sbo, a_lt);
                        const uint64_t b1 = make_sdesc(st + 3 * OPER_BYTES + k * b_step, b_lbo, b_sbo, b_lt);
                        if (F16X2) {
                            umma<true>(tmem_d + 128, a1, b0, idesc, (i | k) != 0);   // lo' * hi  -> cross accumulator
                            umma<true>(tmem_d + 128, a0, b1, idesc, 1);              // hi * lo'
                            umma<true>(tmem_d, a0, b0, idesc, (i | k) != 0);         // hi * hi   -> main accumulator
                        } else {
                            umma<BF16>(tmem_d, a1, b0, idesc, (i | k) != 0);      // small * big
                            umma<BF16>(tmem_d, a0, b1, idesc, 1);                 // big * small
                            umma<BF16>(tmem_d, a0, b0, idesc, 1);                 // big * big
                        }
                    }
                }
                umma_commit(&empty_bar[s]);            // frees the smem stage when these MMAs retire
            }
            umma_commit(&tmem_full_bar);               // accumulator complete
        }
    } else {
        // ================= epilogue: TMEM -> registers -> global =================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp % 4;                        // TMEM lane quarter this warp may access
        const int m = m0 + q * 32 + lane;
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            const float* s0 = reinterpret_cast<const float*>(smem);
            for (int i = (warp - 2) * 32 + lane; i < STAGE_BYTES / 4; i += 128) p.dbg[i] = s0[i];
            for (int c0 = 0; c0 < TBN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
                for (int j = 0; j < 32; ++j)
                    p.dbg[STAGE_BYTES / 4 + (q * 32 + lane) * TBN + c0 + j] = __uint_as_float(r[j]);
            }
        }
        const bool split = gridDim.z > 1;
        const bool lead = blockIdx.z == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < TBN; c0 += 32) {
            uint32_t r[32];
            if (nkb > 0) {
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
                if (F16X2) {
                    uint32_t rx[32];
                    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + 128 + c0, rx);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        r[j] = __float_as_uint(fmaf(__uint_as_float(rx[j]), 1.0f / 2048.0f, __uint_as_float(r[j])));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (F16X2 && p.row_scale != nullptr && m < p.M) {
                const float rs = __ldg(p.row_scale + m);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * rs);
            }
            const bool vec = !split && (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                             (n0 + c0 + 32 <= p.N);
            if (m < p.M && vec) {
                float* crow = p.C + (size_t)m * p.ldc + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = __uint_as_float(r[j + e]);
                        if (p.bias) v[e] += __ldg(p.bias + n0 + c0 + j + e);
                        if (p.Z) v[e] += __ldg(p.Z + (size_t)m * p.ldz + n0 + c0 + j + e);
                    }
                    float4* dst = reinterpret_cast<float4*>(crow + j);
                    if (p.accumulate) {
                        float4 o = *dst;
                        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
                    }
                    *dst = make_float4(v[0], v[1], v[2], v[3]);
                }
            } else if (m < p.M) {
                float* crow = p.C + (size_t)m * p.ldc;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    int n = n0 + c0 + j;
                    if (n < p.N) {
                        float v = __uint_as_float(r[j]);
                        if (!split || lead) {
                            if (p.bias) v += __ldg(p.bias + n);
                            if (p.Z) v += __ldg(p.Z + (size_t)m * p.ldz + n);
                        }
                        if (split) atomicAdd(crow + n, v);
                        else crow[n] = p.accumulate ? crow[n] + v : v;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ pre-pass
// big = x with the low 13 mantissa bits cleared (a TF32 value), small = x - big.
__global__ void split_tf32_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx,
                                  float* __restrict__ big, float* __restrict__ small, int ldo) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = rows * (size_t)(ldo / 4);
    if (i >= total) return;
    size_t r = i / (ldo / 4);
    int c = (int)(i % (ldo / 4)) * 4;
    float v[4], b[4], s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = (c + j < cols) ? x[r * ldx + c + j] : 0.f;
        b[j] = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
        s[j] = v[j] - b[j];
    }
    *reinterpret_cast<float4*>(big + r * ldo + c) = make_float4(b[0], b[1], b[2], b[3]);
    *reinterpret_cast<float4*>(small + r * ldo + c) = make_float4(s[0], s[1], s[2], s[3]);
}
__global__ void cvt_bf16_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx,
                                __nv_bfloat16* __restrict__ out, int ldo) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = rows * (size_t)(ldo / 8);
    if (i >= total) return;
    size_t r = i / (ldo / 8);
    int c = (int)(i % (ldo / 8)) * 8;
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16((c + j < cols) ? x[r * ldx + c + j] : 0.f);
    *reinterpret_cast<uint4*>(out + r * ldo + c) = *reinterpret_cast<const uint4*>(o);
}

// hi = bf16(x) (round to nearest), lo = bf16(x - hi): x = hi + lo up to 2^-17 |x|.
__global__ void split_bf16x2_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ldo) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = rows * (size_t)(ldo / 8);
    if (i >= total) return;
    size_t r = i / (ldo / 8);
    int c = (int)(i % (ldo / 8)) * 8;
    __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float v = (c + j < cols) ? x[r * ldx + c + j] : 0.f;
        h[j] = __float2bfloat16(v);
        l[j] = __float2bfloat16(v - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint4*>(hi + r * ldo + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * ldo + c) = *reinterpret_cast<const uint4*>(l);
}

// mode 4: hi = fp16(x * s), lo' = fp16((x * s - hi) * 2^11); s = 1, or a per-row power of two that puts the row's
// largest magnitude into [2^13, 2^14) (one warp per row; row_inv[r] = 1 / s undoes it in the GEMM epilogue).
__device__ __forceinline__ void f16_parts(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn((v - __half2float(h)) * 2048.0f);
}
__global__ void split_f16x2_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx,
                                   __half* __restrict__ hi, __half* __restrict__ lo, int ldo) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = rows * (size_t)(ldo / 8);
    if (i >= total) return;
    size_t r = i / (ldo / 8);
    int c = (int)(i % (ldo / 8)) * 8;
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f16_parts((c + j < cols) ? x[r * ldx + c + j] : 0.f, h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + r * ldo + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * ldo + c) = *reinterpret_cast<const uint4*>(l);
}
__global__ void __launch_bounds__(256)
split_rows_f16_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx, __half* __restrict__ hi,
                      __half* __restrict__ lo, int ldo, float* __restrict__ row_inv) {
    const size_t r = blockIdx.x * (size_t)(blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    float amax = 0.f;
    for (int c = lane; c < cols; c += 32) amax = fmaxf(amax, fabsf(xr[c]));
    amax = warp_max(amax);
    // s = 2^(13 - floor(log2 amax)): amax * s in [2^13, 2^14); zero / denormal rows are left alone
    float s = 1.0f;
    if (amax >= 1.17549435e-38f && amax < 3.0e38f) {
        const int e = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
        const int se = min(max(13 - e, -126), 126);
        s = __uint_as_float((uint32_t)(se + 127) << 23);
    }
    if (lane == 0) row_inv[r] = 1.0f / s;
    for (int c = lane * 8; c < ldo; c += 256) {
        __align__(16) __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f16_parts((c + j < cols) ? xr[c + j] * s : 0.f, h[j], l[j]);
        *reinterpret_cast<uint4*>(hi + r * ldo + c) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(lo + r * ldo + c) = *reinterpret_cast<const uint4*>(l);
    }
}

// The same split with the row held in registers (one read of x instead of two): ITER chunks of 8 columns per lane,
// cols <= ITER * 256, cols % 8 == 0 and 16-byte aligned rows.
template <int ITER>
__global__ void __launch_bounds__(256)
split_rows_f16_reg_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx, __half* __restrict__ hi,
                          __half* __restrict__ lo, int ldo, float* __restrict__ row_inv) {
    const size_t r = blockIdx.x * (size_t)(blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    float v[ITER][8];
    float amax = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int c = lane * 8 + it * 256;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (c < cols) {
            a = __ldg(reinterpret_cast<const float4*>(xr + c));
            b = __ldg(reinterpret_cast<const float4*>(xr + c + 4));
        }
        v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
        v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[it][j]));
    }
    amax = warp_max(amax);
    float s = 1.0f;
    if (amax >= 1.17549435e-38f && amax < 3.0e38f) {
        const int e = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
        const int se = min(max(13 - e, -126), 126);
        s = __uint_as_float((uint32_t)(se + 127) << 23);
    }
    if (lane == 0) row_inv[r] = 1.0f / s;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int c = lane * 8 + it * 256;
        if (c < ldo) {
            __align__(16) __half h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f16_parts(v[it][j] * s, h[j], l[j]);
            *reinterpret_cast<uint4*>(hi + r * ldo + c) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(lo + r * ldo + c) = *reinterpret_cast<const uint4*>(l);
        }
    }
}
// Wide rows (cols <= ITER * 2048): one CTA per row, 8 * ITER columns per thread in registers, amax through shared memory.
template <int ITER>
__global__ void __launch_bounds__(256)
split_rows_f16_cta_kernel(size_t rows, int cols, const float* __restrict__ x, int ldx, __half* __restrict__ hi,
                          __half* __restrict__ lo, int ldo, float* __restrict__ row_inv) {
    __shared__ float wmax[8];
    const size_t r = blockIdx.x;
    const float* xr = x + r * ldx;
    float v[ITER][8];
    float amax = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int c = (threadIdx.x + it * 256) * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (c < cols) {
            a = __ldg(reinterpret_cast<const float4*>(xr + c));
            b = __ldg(reinterpret_cast<const float4*>(xr + c + 4));
        }
        v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
        v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[it][j]));
    }
    amax = warp_max(amax);
    if (threadIdx.x % 32 == 0) wmax[threadIdx.x / 32] = amax;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, wmax[i]);
    float s = 1.0f;
    if (amax >= 1.17549435e-38f && amax < 3.0e38f) {
        const int e = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
        const int se = min(max(13 - e, -126), 126);
        s = __uint_as_float((uint32_t)(se + 127) << 23);
    }
    if (threadIdx.x == 0) row_inv[r] = 1.0f / s;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int c = (threadIdx.x + it * 256) * 8;
        if (c < ldo) {
            __align__(16) __half h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f16_parts(v[it][j] * s, h[j], l[j]);
            *reinterpret_cast<uint4*>(hi + r * ldo + c) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(lo + r * ldo + c) = *reinterpret_cast<const uint4*>(l);
        }
    }
}
// dispatch: register-resident rows when they fit, else the two-pass kernel
static void launch_split_rows_f16(cudaStream_t st, size_t rows, int cols, const float* x, int ldx, __half* hi, __half* lo,
                                  int ldo, float* row_inv) {
    const bool al = cols % 8 == 0 && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0;
    const bool vec = al && ldo <= 2048;
    const unsigned grid = (unsigned)cdiv(rows, 8);
    if (al && ldo > 2048 && ldo <= 4096)
        split_rows_f16_cta_kernel<2><<<(unsigned)rows, 256, 0, st>>>(rows, cols, x, ldx, hi, lo, ldo, row_inv);
    else if (al && ldo > 4096 && ldo <= 8192)
        split_rows_f16_cta_kernel<4><<<(unsigned)rows, 256, 0, st>>>(rows, cols, x, ldx, hi, lo, ldo, row_inv);
    else if (vec && ldo <= 1024)
        split_rows_f16_reg_kernel<4><<<grid, 256, 0, st>>>(rows, cols, x, ldx, hi, lo, ldo, row_inv);
    else if (vec)
        split_rows_f16_reg_kernel<8><<<grid, 256, 0, st>>>(rows, cols, x, ldx, hi, lo, ldo, row_inv);
    else
        split_rows_f16_kernel<<<grid, 256, 0, st>>>(rows, cols, x, ldx, hi, lo, ldo, row_inv);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// 2-D row-major tensor [rows][cols] (cols contiguous, row stride ld elements); box = {box_c, box_r}
// bf16: 0 = fp32 elements, 1 = bf16, 2 = fp16
bool make_map(CUtensorMap* map, int bf16, const void* ptr, size_t rows, size_t cols, size_t ld, int box_c,
              int box_r, bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * (bf16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, bf16 == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                               : (bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 2,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

void* g_ws = nullptr;
size_t g_ws_bytes = 0;
// per-stream scratch (concurrent streams must not share the operand pre-pass buffers)
struct StreamWs { cudaStream_t st; void* ptr; size_t bytes; };
// torch hands out streams from pools of 32 per priority and device: every handle a process can see fits
constexpr int MAX_STREAM_WS = 256;
StreamWs g_stream_ws[MAX_STREAM_WS];
int g_n_stream_ws = 0;
float* g_dbg = nullptr;
long long g_min_work = 1ll << 27;

}  // namespace

void set_workspace(void* ptr, size_t bytes) {
    g_ws = ptr;
    g_ws_bytes = bytes;
}
// returns false when the table is full: a GEMM on that stream would silently share the default scratch with the
// main stream (a data race), so the caller must treat it as an error
bool set_stream_workspace(cudaStream_t st, void* ptr, size_t bytes) {
    for (int i = 0; i < g_n_stream_ws; ++i)
        if (g_stream_ws[i].st == st) { g_stream_ws[i].ptr = ptr; g_stream_ws[i].bytes = bytes; return true; }
    if (g_n_stream_ws >= MAX_STREAM_WS) return false;
    g_stream_ws[g_n_stream_ws++] = StreamWs{st, ptr, bytes};
    return true;
}
int g_max_kblocks = 256;     // k-blocks (of 32 tf32 / 64 16-bit elements) per CTA before split-K kicks in; 0 = off
void set_tc_debug(float* dbg, long long min_work) {
    g_dbg = dbg;
    g_min_work = min_work;
}

// Returns with *handled = false when the shape/alignment is not eligible (caller falls back to FFMA).
// x_lo = x - (x with the low 13 mantissa bits cleared): the "small" half of the 3xTF32 split.  The "big" half
// needs no copy: kind::tf32 ignores the low 13 bits of its operands (measured), so the raw fp32 tensor serves.
__global__ void split_lo_kernel(size_t n4, const float4* __restrict__ x, float4* __restrict__ lo) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        float4 o;
        o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        lo[i] = o;
    }
}
// bf16x2: the 4n bytes of `lo` hold two bf16 planes, hi[n] then lo[n]
__global__ void split_planes_kernel(size_t n8, const float4* __restrict__ x, uint4* __restrict__ hi,
                                    uint4* __restrict__ lo) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = x[2 * i], b = x[2 * i + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = __float2bfloat16(v[j]);
            l[j] = __float2bfloat16(v[j] - __bfloat162float(h[j]));
        }
        hi[i] = *reinterpret_cast<const uint4*>(h);
        lo[i] = *reinterpret_cast<const uint4*>(l);
    }
}
// f16x2: the same two-plane layout with fp16 parts (no scaling: bounded operands -- weights, activations)
__global__ void split_planes_f16_kernel(size_t n8, const float4* __restrict__ x, uint4* __restrict__ hi,
                                        uint4* __restrict__ lo) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = x[2 * i], b = x[2 * i + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __align__(16) __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f16_parts(v[j], h[j], l[j]);
        hi[i] = *reinterpret_cast<const uint4*>(h);
        lo[i] = *reinterpret_cast<const uint4*>(l);
    }
}
// f16x2 split of a contiguous [rows, cols] matrix with per-row power-of-two scaling (gradient operands): planes hi
// then lo' (rows * cols halves each), row_inv[rows] = the factors that undo the scaling
int split_rows_f16(cudaStream_t st, size_t rows, int cols, const float* x, float* planes, float* row_inv) {
    E2E_REQUIRE(cols % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)planes & 15) == 0,
                "split_rows_f16: rows must be 16-byte aligned multiples of 8 elements");
    if (rows == 0) return 0;
    __half* hi = reinterpret_cast<__half*>(planes);
    launch_split_rows_f16(st, rows, cols, x, cols, hi, hi + rows * (size_t)cols, cols, row_inv);
    E2E_LAUNCH_CHECK();
    return 0;
}
int split_lo(cudaStream_t st, int mode, size_t n, const float* x, float* lo) {
    E2E_REQUIRE(mode == 1 || mode == 3 || mode == 4,
                "split_lo: mode %d has no operand split (1 = tf32x3, 3 = bf16x2, 4 = f16x2)", mode);
    E2E_REQUIRE(n % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)lo & 15) == 0,
                "split_lo: buffers must be 16-byte aligned with a multiple of 8 elements");
    if (n == 0) return 0;
    if (mode == 3 || mode == 4) {
        const unsigned grid = (unsigned)min((size_t)148 * 8, (n / 8 + 255) / 256);
        if (mode == 3)
            split_planes_kernel<<<grid, 256, 0, st>>>(n / 8, (const float4*)x, (uint4*)lo, (uint4*)((__nv_bfloat16*)lo + n));
        else
            split_planes_f16_kernel<<<grid, 256, 0, st>>>(n / 8, (const float4*)x, (uint4*)lo, (uint4*)((__half*)lo + n));
        E2E_LAUNCH_CHECK();
        return 0;
    }
    // short-lived CTAs (4 float4 per thread), not a resident grid-stride loop: these splits run on side streams, and a
    // CTA that stays resident for the whole pass keeps its SM from the 16-CTA clusters of the recurrence on the
    // critical path (they need EMPTY SMs) -- with many short CTAs the higher-priority clusters get in as SMs drain
    static int g_split_grid_stride = -1;
    if (g_split_grid_stride < 0) g_split_grid_stride = getenv("E2E_SPLIT_GRID_STRIDE") ? atoi(getenv("E2E_SPLIT_GRID_STRIDE")) : 0;
    const size_t want = g_split_grid_stride ? (size_t)148 * 8 : (n / 4 + 1023) / 1024;
    split_lo_kernel<<<(unsigned)min(want, (size_t)1 << 30), 256, 0, st>>>(n / 4, (const float4*)x, (float4*)lo);
    E2E_LAUNCH_CHECK();
    return 0;
}

// A_lo / B_lo (optional, tf32x3 mode): precomputed split_lo() of the operand with the operand's own layout.  When
// given (and TMA-addressable) the operand's pre-pass is skipped: the tensor maps point at the caller's buffers.
// row_scale (mode 4 with caller-provided A planes): the per-row factors that undo the planes' scaling, or NULL.
int gemm_tc(cudaStream_t st, int mode, int transA, int transB, int M, int N, int K, const float* A, int lda,
            const float* B, int ldb, float* C, int ldc, const float* bias, const float* Z, int ldz, int accumulate,
            bool* handled, const float* A_lo, const float* B_lo, size_t a_plane, size_t b_plane,
            const float* row_scale) {
    *handled = false;
    if (mode < 1 || mode > 4) return 0;
    // f16x2 scales the A operand per output row: a transposed A (K runs over its rows) stays on 3xTF32, whose splits
    // are not interchangeable with the fp16 planes
    if (mode == 4 && transA) { mode = 1; A_lo = nullptr; B_lo = nullptr; row_scale = nullptr; }
    const bool bf16 = mode >= 2;
    // big enough to pay for the pre-pass and to fill tiles; K >= one stage
    if (M < 96 || N < 64 || K < 32 || (long long)M * N * K < g_min_work) return 0;
    void* ws_ptr = g_ws;
    size_t ws_bytes = g_ws_bytes;
    for (int i = 0; i < g_n_stream_ws; ++i)
        if (g_stream_ws[i].st == st) { ws_ptr = g_stream_ws[i].ptr; ws_bytes = g_stream_ws[i].bytes; }
    if (!get_encode() || ws_ptr == nullptr) return 0;
    // stored shapes
    const size_t a_rows = transA ? K : M, a_cols = transA ? M : K;
    const size_t b_rows = transB ? N : K, b_cols = transB ? K : N;
    const int eper = bf16 ? 8 : 4;                                  // elements per 16 bytes
    const size_t a_ld = (a_cols + eper - 1) / eper * eper, b_ld = (b_cols + eper - 1) / eper * eper;
    const size_t esz = bf16 ? 2 : 4, nparts = mode == 2 ? 1 : 2;
    size_t a_bytes = (a_rows * a_ld * esz + 1023) / 1024 * 1024, b_bytes = (b_rows * b_ld * esz + 1023) / 1024 * 1024;
    const size_t rs_bytes = mode == 4 ? ((size_t)M * 4 + 1023) / 1024 * 1024 : 0;       // per-row scales of the pre-pass
    // operands whose split the caller already holds: used in place when TMA can address them
    // (tf32x3: x itself + its fp32 "small" half; bf16x2: the two bf16 planes of split_lo, `plane` elements apart)
    auto direct_ok = [&](const float* x, const float* lo, int ld, size_t plane) {
        if (mode == 1) return lo != nullptr && ld % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)lo & 15) == 0;
        if (mode == 3 || mode == 4)
            return lo != nullptr && ld % 8 == 0 && plane % 8 == 0 && plane > 0 && ((uintptr_t)lo & 15) == 0;
        return false;
    };
    const bool a_direct = direct_ok(A, A_lo, lda, a_plane), b_direct = direct_ok(B, B_lo, ldb, b_plane);
    if (a_direct) a_bytes = 0;
    if (b_direct) b_bytes = 0;
    if (nparts * (a_bytes + b_bytes) + rs_bytes > ws_bytes) return 0;
    uint8_t* ws = (uint8_t*)ws_ptr;
    float* rs_ws = reinterpret_cast<float*>(ws + nparts * (a_bytes + b_bytes));
    void* A0 = ws;
    void* A1 = ws + a_bytes;
    void* B0 = ws + nparts * a_bytes;
    void* B1 = ws + nparts * a_bytes + b_bytes;
    size_t a_ld_eff = a_ld, b_ld_eff = b_ld;
    if (a_direct) { A0 = const_cast<float*>(A); A1 = const_cast<float*>(A_lo); a_ld_eff = lda; }
    if (b_direct) { B0 = const_cast<float*>(B); B1 = const_cast<float*>(B_lo); b_ld_eff = ldb; }
    if (mode == 4) {
        if (a_direct) { A0 = const_cast<float*>(A_lo); A1 = (__half*)A0 + a_plane; }
        else row_scale = nullptr;
        if (b_direct) { B0 = const_cast<float*>(B_lo); B1 = (__half*)B0 + b_plane; }
        if (!a_direct) {       // A is [M][K] here (transA went to mode 1): per-row scaled split, undone in the epilogue
            launch_split_rows_f16(st, a_rows, (int)a_cols, A, lda, (__half*)A0, (__half*)A1, (int)a_ld, rs_ws);
            E2E_LAUNCH_CHECK();
            row_scale = rs_ws;
        }
        if (!b_direct) {
            split_f16x2_kernel<<<cdiv(b_rows * (b_ld / 8), 256), 256, 0, st>>>(b_rows, (int)b_cols, B, ldb, (__half*)B0, (__half*)B1, (int)b_ld);
            E2E_LAUNCH_CHECK();
        }
    } else if (mode == 3) {
        if (a_direct) { A0 = const_cast<float*>(A_lo); A1 = (__nv_bfloat16*)A0 + a_plane; }
        if (b_direct) { B0 = const_cast<float*>(B_lo); B1 = (__nv_bfloat16*)B0 + b_plane; }
        if (!a_direct) {
            split_bf16x2_kernel<<<cdiv(a_rows * (a_ld / 8), 256), 256, 0, st>>>(a_rows, (int)a_cols, A, lda, (__nv_bfloat16*)A0, (__nv_bfloat16*)A1, (int)a_ld);
            E2E_LAUNCH_CHECK();
        }
        if (!b_direct) {
            split_bf16x2_kernel<<<cdiv(b_rows * (b_ld / 8), 256), 256, 0, st>>>(b_rows, (int)b_cols, B, ldb, (__nv_bfloat16*)B0, (__nv_bfloat16*)B1, (int)b_ld);
            E2E_LAUNCH_CHECK();
        }
    } else if (bf16) {
        cvt_bf16_kernel<<<cdiv(a_rows * (a_ld / 8), 256), 256, 0, st>>>(a_rows, (int)a_cols, A, lda, (__nv_bfloat16*)A0, (int)a_ld);
        E2E_LAUNCH_CHECK();
        cvt_bf16_kernel<<<cdiv(b_rows * (b_ld / 8), 256), 256, 0, st>>>(b_rows, (int)b_cols, B, ldb, (__nv_bfloat16*)B0, (int)b_ld);
        E2E_LAUNCH_CHECK();
    } else {
        if (!a_direct) {
            split_tf32_kernel<<<cdiv(a_rows * (a_ld / 4), 256), 256, 0, st>>>(a_rows, (int)a_cols, A, lda, (float*)A0, (float*)A1, (int)a_ld);
            E2E_LAUNCH_CHECK();
        }
        if (!b_direct) {
            split_tf32_kernel<<<cdiv(b_rows * (b_ld / 4), 256), 256, 0, st>>>(b_rows, (int)b_cols, B, ldb, (float*)B0, (float*)B1, (int)b_ld);
            E2E_LAUNCH_CHECK();
        }
    }
    const int bke = bf16 ? 64 : 32;           // elements per 128-byte span
    // A: K-major when the stored matrix is [M][K] (not transposed); B: K-major when stored [N][K] (transposed)
    const int a_mn = transA ? 1 : 0, b_mn = transB ? 0 : 1;
    CUtensorMap mA0, mA1, mB0, mB1;
    bool ok = true;
    const bool a32 = !bf16 && a_mn, b32 = !bf16 && b_mn;      // tf32 MN-major: 32-byte-atom swizzle
    const int dt = mode == 4 ? 2 : (bf16 ? 1 : 0);
    ok &= make_map(&mA0, dt, A0, a_rows, a_cols, a_ld_eff, bke, a_mn ? bke : TBM, a32);
    ok &= make_map(&mB0, dt, B0, b_rows, b_cols, b_ld_eff, bke, b_mn ? bke : TBN, b32);
    if (nparts == 2) {
        ok &= make_map(&mA1, dt, A1, a_rows, a_cols, a_ld_eff, bke, a_mn ? bke : TBM, a32);
        ok &= make_map(&mB1, dt, B1, b_rows, b_cols, b_ld_eff, bke, b_mn ? bke : TBN, b32);
    } else {
        mA1 = mA0;
        mB1 = mB0;
    }
    if (!ok) return 0;

    TcParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias; p.Z = Z; p.ldz = ldz;
    p.accumulate = accumulate; p.a_mn_major = a_mn; p.b_mn_major = b_mn; p.dbg = g_dbg;
    p.row_scale = mode == 4 ? row_scale : nullptr;
    const int gm = cdiv(M, TBM), gn = cdiv(N, TBN), tiles = gm * gn;
    const int kb_total = cdiv(K, bke);
    int splits = 1;
    const int nsm = sm_count();
    if (tiles * 2 <= nsm && kb_total >= 16) splits = min(nsm / tiles, kb_total / 8);
    // very long K (the weight-gradient products: K = all frames of the batch): keep a CTA's life short.  These GEMMs run
    // on side streams next to the cluster recurrences of the critical path, whose 16-CTA clusters need 16 EMPTY SMs of a
    // GPC at once -- they wait for resident GEMM CTAs to retire, so a CTA that loops over all of K delays them by its
    // whole run time.
    {
        static bool env_read = false;
        if (!env_read) {
            env_read = true;
            if (const char* e = getenv("E2E_MAX_KBLOCKS")) g_max_kblocks = atoi(e);
        }
    }
    if (g_max_kblocks > 0 && kb_total > 2 * g_max_kblocks) splits = max(splits, cdiv(kb_total, g_max_kblocks));
    if (splits < 1) splits = 1;
    p.kblocks_per_split = cdiv(kb_total, splits);
    splits = cdiv(kb_total, p.kblocks_per_split);
    if (splits > 1 && !accumulate) {
        if (ldc == N) E2E_CHECK_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
        else E2E_CHECK_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
    }
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    dim3 grid(gn, gm, splits);
    if (mode == 4) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tc_kernel<4><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB0, mB1, p);
    } else if (mode == 3) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tc_kernel<3><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB0, mB1, p);
    } else if (mode == 2) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tc_kernel<2><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB0, mB1, p);
    } else {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB0, mB1, p);
    }
    E2E_LAUNCH_CHECK();
    *handled = true;
    return 0;
}

}  // namespace e2e
