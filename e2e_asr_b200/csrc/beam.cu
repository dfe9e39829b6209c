// Float64 decoder-step kernels for beam search (reference beam_search.py:137-221).
//
// The reference scores hypotheses in float64 (its zero states are np.zeros, so every
// GEMV after the embedding is promoted) while weights, embeddings,
// encoder states and enc.AttnW stay float32.  These kernels keep exactly that dtype
// flow on the device -- fp32 operands are widened at load, all arithmetic is fp64 --
// and are batched over every live hypothesis of every utterance.
#include "../../include/e2e_asr_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace e2e {

// C[M,N] (f64) = A[M,K] (f64) . B[K,N] (f32 weights) + bias[N] (f32)
// Register-tiled DFMA kernel: CTA = 64 x 64 output tile, 64 threads, 8 x 8 accumulators per thread (one k-step =
// 8 + 8 shared-memory operands for 64 FMAs), several CTAs resident per SM to hide the global-load latency.
// Thread (ty, tx) owns rows {16 i + 2 ty + e} and columns {16 j + 2 tx + e} (i, j < 4; e < 2): a quarter warp reads
// 128 contiguous bytes of the B tile (conflict-free LDS.128) and one broadcast address of the A tile.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(64)
gemm_f64_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                double* __restrict__ C, int ldc, const float* __restrict__ bias) {
    __shared__ __align__(16) double As[GT][GK + 1];      // [m][k]
    __shared__ __align__(16) double Bs[GK][GT];          // [k][n]
    const int tid = threadIdx.x, tx = tid % 8, ty = tid / 8;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < K; k0 += GK) {
        // A tile: 64 rows x 16 k (each row 128 contiguous bytes); B tile: 16 k x 64 columns of fp32
#pragma unroll 4
        for (int it = 0; it < GT * GK / 64; ++it) {
            const int i = it * 64 + tid, r = i / GK, k = i % GK;
            As[r][k] = (m0 + r < M && k0 + k < K) ? A[(size_t)(m0 + r) * lda + k0 + k] : 0.0;
        }
#pragma unroll 4
        for (int it = 0; it < GK * GT / 64; ++it) {
            const int i = it * 64 + tid, k = i / GT, c = i % GT;
            Bs[k][c] = (k0 + k < K && n0 + c < N) ? (double)B[(size_t)(k0 + k) * ldb + n0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[2 * i] = As[16 * i + 2 * ty][k];
                a[2 * i + 1] = As[16 * i + 2 * ty + 1][k];
                const double2 bv = *reinterpret_cast<const double2*>(&Bs[k][16 * i + 2 * tx]);
                b[2 * i] = bv.x;
                b[2 * i + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + 16 * (i / 2) + 2 * ty + (i % 2);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + 16 * (j / 2) + 2 * tx + (j % 2);
            if (n < N) C[(size_t)m * ldc + n] = acc[i][j] + (bias ? (double)bias[n] : 0.0);
        }
    }
}

// Same product with a 2-stage cp.async pipeline (needs 16-byte-aligned rows: lda % 2 == 0, ldb % 4 == 0, K % 16 == 0):
// 128 threads, thread (ty, tx) owns rows {16 i + 2 ty + e} (i < 4) and columns {32 j + 2 tx + e} (j < 2); the next
// k-tile streams into the other shared-memory stage while this one is multiplied; B stays fp32 in shared memory and is
// widened at the read.  Out-of-range rows / columns are zero-filled by cp.async's src-size operand.
constexpr int PT = 64, PK = 16, PAS = PK + 2;      // A row stride in doubles: 18 (16-byte aligned rows, banks spread)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(src_bytes) : "memory");
}
template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<double> { typedef double2 type; };
// TB = double: the weights were widened once by the caller (e2e_gemm_f64d) -- no per-k conversions in the inner loop
// (F2F shares the FP64 pipe with the FMAs); (double)float is exact, so the results are bit-identical.
template <typename TB>
__global__ void __launch_bounds__(128)
gemm_f64_pipe_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const TB* __restrict__ B, int ldb,
                     double* __restrict__ C, int ldc, const float* __restrict__ bias) {
    constexpr int BV = 16 / (int)sizeof(TB);                 // B elements per 16-byte chunk
    __shared__ __align__(16) double As[2][PT][PAS];
    __shared__ __align__(16) TB Bs[2][PK][PT];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * PT, n0 = blockIdx.x * PT;
    auto issue = [&](int stage, int k0) {
        // A: 64 rows x 16 doubles = 512 chunks of 16 B; B: 16 rows x 64 floats = 256 chunks of 16 B
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = it * 128 + tid, r = i / 8, c = (i % 8) * 2;
            const bool ok = m0 + r < M;
            cp_async16(&As[stage][r][c], A + (size_t)(ok ? m0 + r : 0) * lda + k0 + c, ok ? 16 : 0);
        }
#pragma unroll
        for (int it = 0; it < PK * PT / BV / 128; ++it) {
            const int i = it * 128 + tid, k = i / (PT / BV), c = (i % (PT / BV)) * BV;
            const int rem = N - (n0 + c);                       // columns left in this row of B
            const int bytes = rem >= BV ? 16 : (rem > 0 ? rem * (int)sizeof(TB) : 0);
            cp_async16(&Bs[stage][k][c], B + (size_t)(k0 + k) * ldb + (rem > 0 ? n0 + c : 0), bytes);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int nk = K / PK;
    issue(0, 0);
    for (int t = 0; t < nk; ++t) {
        if (t + 1 < nk) {
            issue((t + 1) & 1, (t + 1) * PK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int s = t & 1;
#pragma unroll
        for (int k = 0; k < PK; ++k) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[2 * i] = As[s][16 * i + 2 * ty][k];
                a[2 * i + 1] = As[s][16 * i + 2 * ty + 1][k];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const typename Pair<TB>::type bv = *reinterpret_cast<const typename Pair<TB>::type*>(&Bs[s][k][32 * j + 2 * tx]);
                b[2 * j] = (double)bv.x;
                b[2 * j + 1] = (double)bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + 16 * (i / 2) + 2 * ty + (i % 2);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 32 * (j / 2) + 2 * tx + (j % 2);
            if (n < N) C[(size_t)m * ldc + n] = acc[i][j] + (bias ? (double)bias[n] : 0.0);
        }
    }
}

int g_f64_mma = 1;      // e2e_gemm_f64d: 1 = FP64 tensor-core kernel, 0 = register-tiled DFMA kernel (e2e_set_f64_mma, tests)
// FP64 tensor-core version (mma.sync m8n8k4 f64) of the aligned product with float64 weights: the register-tiled DFMA
// kernels above spend three shared-memory wavefronts per four FMA issues (measured 15.6 TFLOP/s, the LSU as busy as the
// FP64 pipe); a DMMA takes its 8x4 / 4x8 fragments with ONE 8-byte load per lane for 256 FMAs.  CTA = 64 x 64 tile,
// 4 warps of 32 x 32 (4 x 4 m8n8 tiles, 32 accumulator doubles per lane), 2-stage cp.async pipeline over k-tiles of 16;
// row strides 18 (A, [m][k]) and 68 (B, [k][n]) doubles spread the fragment loads evenly over the banks (two
// wavefronts per 256-byte load, the minimum).  The four products of a k4 block are summed inside the tensor core, so
// the result differs from the sequential-FMA kernels in the last bits (not in accuracy).
constexpr int BS2 = PT + 4;
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
// Operands of the DMMA product beyond C = A.B + bias (all optional):
//  * K-concatenated A: columns [0, K1) of the product's A come from A, columns [K1, K) from A2 (K1 % 16 == 0) -- the
//    reference's np.concatenate([x, h]) . W (basic_lstm.py:17, beam_search.py:186,194) without materialising the
//    concatenation;
//  * a row addend Z[zrow[m], :] (float64): the part of a product that depends only on the row's TOKEN -- emb[tok] . Wx
//    + b of the LM-LSTM -- is a table computed once per model, so the step's product keeps only K = H;
//  * an LSTM epilogue (EPI = 1): with the 4H gate columns interleaved in blocks of 32 -- column 32 q + 8 g + i holds
//    gate g (i, j, f, o) of unit 8 q + i -- the lane that owns rows {8 i + lr} and columns {8 j + 2 lc + e} of a warp
//    tile holds all four gates of two units, so BasicLSTM.__call__ (basic_lstm.py:14-23) runs on the accumulators and
//    the pre-activations never go to memory.  Same formulas as lstm_step_f64_kernel.
struct GemmF64Ext {
    const double* A2;
    int lda2, K1;
    const double* Z;
    int ldz;
    const long long* zrow;
    const double* c_prev;      // EPI = 1: [M, H] in, c_out / h_out [M, H] resp. [M, ldh] out
    double* c_out;
    double* h_out;
    int ldh, H;
};
// CTA = (2 WM) x 64 tile, 4 warps of WM x 32 (WM / 8 x 4 m8n8 tiles), k-tiles of TPK, 2-stage cp.async pipeline.
// ncu (profiles/r4_ncu_beam_summary.txt): with 64-row tiles and k-tiles of 16 the DMMA pipe is ~77 % busy while an SM
// has work, but only ~70 % of the SMs' time is covered -- 640 CTAs against 592 resident slots leave a tail on 48 SMs.
// 32-row tiles halve the tail; k-tiles of 32 give them the 64 DMMAs per barrier pair of the 64-row tile.  Row strides
// TPK + 4 (A, [m][k]) and 68 (B, [k][n]) doubles: both fragment loads are bank-conflict free.
template <int WM, int TPK, int EPI>
__global__ void __launch_bounds__(128)
gemm_f64_mma_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                    double* __restrict__ C, int ldc, const float* __restrict__ bias, GemmF64Ext x) {
    constexpr int TM = 2 * WM, MI = WM / 8, AS = TPK + 4;
    extern __shared__ __align__(16) double smem_d[];
    double* As = smem_d;                                  // [2][TM][AS]
    double* Bs = smem_d + 2 * TM * AS;                    // [2][TPK][BS2]
    const int tid = threadIdx.x, lane = tid % 32, w = tid / 32, wm = w / 2, wn = w % 2;
    const int lr = lane / 4, lc = lane % 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * PT;
    auto issue = [&](int stage, int k0) {
        const bool second = k0 >= x.K1;
        const double* Ab = second ? x.A2 : A;
        const int ld = second ? x.lda2 : lda, kk0 = second ? k0 - x.K1 : k0;
#pragma unroll
        for (int it = 0; it < TM * TPK / 256; ++it) {
            const int i = it * 128 + tid, r = i / (TPK / 2), c = (i % (TPK / 2)) * 2;
            const bool ok = m0 + r < M;
            cp_async16(&As[(stage * TM + r) * AS + c], Ab + (size_t)(ok ? m0 + r : 0) * ld + kk0 + c, ok ? 16 : 0);
        }
#pragma unroll
        for (int it = 0; it < TPK / 4; ++it) {
            const int i = it * 128 + tid, k = i / 32, c = (i % 32) * 2;
            const int rem = N - (n0 + c);
            const int bytes = rem >= 2 ? 16 : (rem > 0 ? 8 : 0);
            cp_async16(&Bs[(stage * TPK + k) * BS2 + c], B + (size_t)(k0 + k) * ldb + (rem > 0 ? n0 + c : 0), bytes);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int nk = K / TPK;
    issue(0, 0);
    for (int t = 0; t < nk; ++t) {
        if (t + 1 < nk) {
            issue((t + 1) & 1, (t + 1) * TPK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* as = As + (t & 1) * TM * AS + (WM * wm + lr) * AS + lc;
        const double* bs = Bs + (t & 1) * TPK * BS2 + lc * BS2 + 32 * wn + lr;
#pragma unroll
        for (int kk = 0; kk < TPK / 4; ++kk) {
            double a[MI], b[4];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[8 * i * AS + 4 * kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[4 * kk * BS2 + 8 * j];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + WM * wm + 8 * i + lr;
        if (m >= M) continue;
        const double* zr = x.Z ? x.Z + (size_t)x.zrow[m] * x.ldz : nullptr;
        if (EPI == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + 32 * wn + 8 * j + 2 * lc + e;
                    if (n < N) {
                        double v = acc[i][j][e] + (bias ? (double)bias[n] : 0.0);
                        if (zr) v += zr[n];
                        C[(size_t)m * ldc + n] = v;
                    }
                }
        } else {
            // N = 4H is a multiple of 64: no column bounds.  Column n0 + 32 wn + 8 g + (2 lc + e) = gate g of unit u
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nb = n0 + 32 * wn + 2 * lc + e;
                const int u = (nb / 32) * 8 + 2 * lc + e;
                double g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = acc[i][j][e] + (bias ? (double)bias[nb + 8 * j] : 0.0);
                    if (zr) g[j] += zr[nb + 8 * j];
                }
                const double si = 1.0 / (1.0 + exp(-g[0]));
                const double tj = tanh(g[1]);
                const double sf = 1.0 / (1.0 + exp(-(g[2] + 1.0)));
                const double so = 1.0 / (1.0 + exp(-g[3]));
                const double cn = x.c_prev[(size_t)m * x.H + u] * sf + si * tj;
                x.c_out[(size_t)m * x.H + u] = cn;
                x.h_out[(size_t)m * x.ldh + u] = so * tanh(cn);
            }
        }
    }
}

int g_f64_tile = -1;     // E2E_F64_TILE: 0 = automatic (default), 1 = k-tiles of 16 only (64- / 32-row tiles by CTA count)
template <int WM, int TPK, int EPI>
static int launch_f64_mma_t(cudaStream_t st, int M, int N, int K, const double* A, int lda, const double* B, int ldb,
                            double* C, int ldc, const float* bias, const GemmF64Ext& x) {
    constexpr int SMEM = (2 * (2 * WM) * (TPK + 4) + 2 * TPK * BS2) * (int)sizeof(double);
    static bool attr_set = false;
    if (!attr_set && SMEM > 48 * 1024) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gemm_f64_mma_kernel<WM, TPK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    gemm_f64_mma_kernel<WM, TPK, EPI><<<dim3(cdiv(N, PT), cdiv(M, 2 * WM)), 128, SMEM, st>>>(M, N, K, A, lda, B, ldb, C, ldc,
                                                                                         bias, x);
    E2E_LAUNCH_CHECK();
    return 0;
}
// the same product with the weights already widened to float64 (aligned operands only: K % 16 == 0, even lda / ldb)
static int launch_f64_mma(cudaStream_t st, int M, int N, int K, const double* A, int lda, const double* B, int ldb,
                          double* C, int ldc, const float* bias, const GemmF64Ext& x, bool lstm) {
    if (g_f64_tile < 0) g_f64_tile = getenv("E2E_F64_TILE") ? atoi(getenv("E2E_F64_TILE")) : 0;
    if (g_f64_tile == 0 && K % 32 == 0 && x.K1 % 32 == 0) {
        if (lstm) return launch_f64_mma_t<16, 32, 1>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
        return launch_f64_mma_t<16, 32, 0>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
    }
    // 64-row tiles unless they leave fewer than two CTAs per SM
    const bool small = (long long)cdiv(M, 64) * cdiv(N, PT) < 2 * 148;
    if (lstm) {
        if (small) return launch_f64_mma_t<16, 16, 1>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
        return launch_f64_mma_t<32, 16, 1>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
    }
    if (small) return launch_f64_mma_t<16, 16, 0>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
    return launch_f64_mma_t<32, 16, 0>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x);
}
int gemm_f64d(cudaStream_t st, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
              int ldc, const float* bias) {
    if (M <= 0 || N <= 0) return 0;
    E2E_REQUIRE(K > 0 && K % PK == 0 && lda % 2 == 0 && ldb % 2 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0,
                "gemm_f64d: K %% 16 == 0 and 16-byte aligned rows required (K=%d lda=%d ldb=%d)", K, lda, ldb);
    if (g_f64_mma) {
        GemmF64Ext x = {};
        x.K1 = K;
        return launch_f64_mma(st, M, N, K, A, lda, B, ldb, C, ldc, bias, x, false);
    }
    gemm_f64_pipe_kernel<double><<<dim3(cdiv(N, PT), cdiv(M, PT)), 128, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias);
    E2E_LAUNCH_CHECK();
    return 0;
}
// C = [A1 | A2] . B + bias + Z[zrow]  (A2 / Z optional), resp. the LSTM step on that product (c_out, h_out; no C)
int gemm_f64d_cat(cudaStream_t st, int M, int N, int K1, int K2, const double* A1, int lda1, const double* A2, int lda2,
                  const double* B, int ldb, double* C, int ldc, const float* bias, const double* Z, int ldz,
                  const long long* zrow, const double* c_prev, double* c_out, double* h_out, int ldh) {
    if (M <= 0 || N <= 0) return 0;
    const bool lstm = c_out != nullptr;
    E2E_REQUIRE(K1 > 0 && K1 % PK == 0 && K2 >= 0 && K2 % PK == 0 && lda1 % 2 == 0 && ldb % 2 == 0 &&
                    ((uintptr_t)A1 & 15) == 0 && ((uintptr_t)B & 15) == 0,
                "gemm_f64d_cat: K1, K2 %% 16 == 0 and 16-byte aligned rows required (K1=%d K2=%d lda1=%d ldb=%d)", K1, K2,
                lda1, ldb);
    E2E_REQUIRE(K2 == 0 || (A2 != nullptr && lda2 % 2 == 0 && ((uintptr_t)A2 & 15) == 0),
                "gemm_f64d_cat: the second A operand must be 16-byte aligned with an even row stride (lda2=%d)", lda2);
    E2E_REQUIRE(Z == nullptr || zrow != nullptr, "gemm_f64d_cat: a row addend needs its row indices");
    E2E_REQUIRE(!lstm || (N % 64 == 0 && c_prev != nullptr && h_out != nullptr),
                "gemm_f64d_cat: the LSTM epilogue needs N = 4H with H %% 16 == 0 (N=%d), c_prev and h_out", N);
    E2E_REQUIRE(lstm || C != nullptr, "gemm_f64d_cat: no output");
    GemmF64Ext x = {};
    x.A2 = A2; x.lda2 = lda2; x.K1 = K1;
    x.Z = Z; x.ldz = ldz; x.zrow = zrow;
    x.c_prev = c_prev; x.c_out = c_out; x.h_out = h_out; x.ldh = ldh; x.H = N / 4;
    return launch_f64_mma(st, M, N, K1 + K2, A1, lda1, B, ldb, C, ldc, bias, x, lstm);
}

int gemm_f64(cudaStream_t st, int M, int N, int K, const double* A, int lda, const float* B, int ldb, double* C,
             int ldc, const float* bias) {
    if (M <= 0 || N <= 0) return 0;
    const bool aligned = K > 0 && K % PK == 0 && lda % 2 == 0 && ldb % 4 == 0 && ((uintptr_t)A & 15) == 0 &&
                         ((uintptr_t)B & 15) == 0;
    if (aligned)
        gemm_f64_pipe_kernel<float><<<dim3(cdiv(N, PT), cdiv(M, PT)), 128, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias);
    else
        gemm_f64_kernel<<<dim3(cdiv(N, GT), cdiv(M, GT)), 64, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias);
    E2E_LAUNCH_CHECK();
    return 0;
}

// BasicLSTM.__call__ (basic_lstm.py:14-23) on pre-activations z [n,4H] (i|j|f|o); returns (new_c, new_h)
__global__ void lstm_step_f64_kernel(int n, int H, const double* __restrict__ z, const double* __restrict__ c_prev,
                                     double* __restrict__ c_out, double* __restrict__ h_out, int ldh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    int r = i / H, u = i % H;
    const double* g = z + (size_t)r * 4 * H;
    double si = 1.0 / (1.0 + exp(-g[u]));
    double tj = tanh(g[H + u]);
    double sf = 1.0 / (1.0 + exp(-(g[2 * H + u] + 1.0)));
    double so = 1.0 / (1.0 + exp(-g[3 * H + u]));
    double cn = c_prev[(size_t)r * H + u] * sf + si * tj;
    c_out[(size_t)r * H + u] = cn;
    h_out[(size_t)r * ldh + u] = so * tanh(cn);
}
int lstm_step_f64(cudaStream_t st, int n, int H, const double* z, const double* c_prev, double* c_out, double* h_out,
                  int ldh) {
    if (n <= 0) return 0;
    lstm_step_f64_kernel<<<cdiv(n * H, 256), 256, 0, st>>>(n, H, z, c_prev, c_out, h_out, ldh);
    E2E_LAUNCH_CHECK();
    return 0;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1): one exp and one division instead of libm's tanh (the attention scores take 14 M of
// them per decoding step at 256 utterances x beam 10 and are bound by it).  Absolute error < 3e-16, saturates correctly
// (exp -> inf gives 1, exp -> 0 gives -1); the relative accuracy libm keeps near 0 is not needed: a score is a sum of
// v_a tanh(.) over enc . AttnW values that are themselves float32 products (beam_search.py:148).
__device__ __forceinline__ double tanh_exp(double x) { return 1.0 - 2.0 / (exp(2.0 * x) + 1.0); }

// calc_attention (beam_search.py:150-159): one CTA per hypothesis; the utterance's encoder rows are
// [row_off, row_off + T) of enc / HF (already length-sliced: NO mask, beam_search.py:155).
__global__ void __launch_bounds__(256)
attn_beam_f64_kernel(int A, int D, const float* __restrict__ HF, const float* __restrict__ enc,
                     const int* __restrict__ row_off, const int* __restrict__ Tlen, const double* __restrict__ y,
                     const float* __restrict__ v, double* __restrict__ ctx, int ldctx) {
    extern __shared__ double sm[];
    double* y_s = sm;          // [A]
    double* s_s = sm + A;      // [T]
    __shared__ double red[32];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nw = 8;
    const int off = row_off[r], T = Tlen[r];
    for (int a = tid; a < A; a += 256) y_s[a] = y[(size_t)r * A + a];
    __syncthreads();
    for (int tau = warp; tau < T; tau += nw) {
        double p = 0.0;
        for (int a = lane; a < A; a += 32)
            p += tanh_exp((double)HF[(size_t)(off + tau) * A + a] + y_s[a]) * (double)v[a];
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0) s_s[tau] = p;
    }
    __syncthreads();
    double mx = -INFINITY;
    for (int tau = tid; tau < T; tau += 256) mx = fmax(mx, s_s[tau]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nw; ++w) mx = fmax(mx, red[w]);
    __syncthreads();
    double sum = 0.0;
    for (int tau = tid; tau < T; tau += 256) {
        double e = exp(s_s[tau] - mx);
        s_s[tau] = e;
        sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.0;
    for (int w = 0; w < nw; ++w) sum += red[w];
    for (int d = tid; d < D; d += 256) {
        double c = 0.0;
        for (int tau = 0; tau < T; ++tau) c = fma(s_s[tau] / sum, (double)enc[(size_t)(off + tau) * D + d], c);
        ctx[(size_t)r * ldctx + d] = c;
    }
}
int attn_beam_f64(cudaStream_t st, int n, int A, int D, int Tmax, const float* HF, const float* enc,
                  const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx) {
    if (n <= 0) return 0;
    size_t smem = sizeof(double) * (A + Tmax);
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_beam_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_beam_f64_kernel<<<n, 256, smem, st>>>(A, D, HF, enc, row_off, Tlen, y, v, ctx, ldctx);
    E2E_LAUNCH_CHECK();
    return 0;
}

// The same read-out for hypotheses stored in groups: rows u*beam .. u*beam+beam-1 belong to utterance u and share its
// encoder rows (row_off / Tlen of the group's first row).  One CTA per utterance: every HF / enc element is loaded ONCE
// and used by all `beam` hypotheses (the per-row kernel re-reads the utterance's 100+ KB per hypothesis).  Every
// hypothesis goes through exactly the arithmetic of attn_beam_f64_kernel -- same lane partition and shuffle trees of
// the scores, same warp-chunked softmax sums, same sequential read-out -- so the results are bit-identical.
constexpr int MAXB = 16;
__global__ void __launch_bounds__(256)
attn_beam_group_f64_kernel(int beam, int A, int D, int Tmax, const float* __restrict__ HF, const float* __restrict__ enc,
                           const int* __restrict__ row_off, const int* __restrict__ Tlen, const double* __restrict__ y,
                           const float* __restrict__ v, double* __restrict__ ctx, int ldctx) {
    extern __shared__ double sm[];
    double* y_s = sm;                       // [beam][A]
    double* v_s = y_s + beam * A;           // [A]
    double* s_s = v_s + A;                  // [beam][Tmax]  scores, then exp, then alpha
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nw = 8;
    const int r0 = u * beam;
    const int off = row_off[r0], T = Tlen[r0];
    for (int i = tid; i < beam * A; i += 256) y_s[i] = y[(size_t)(r0 + i / A) * A + i % A];
    for (int a = tid; a < A; a += 256) v_s[a] = (double)v[a];
    __syncthreads();
    // scores s[r][tau] = sum_a tanh(HF[tau][a] + y[r][a]) v[a]
    for (int tau = warp; tau < T; tau += nw) {
        const float* hrow = HF + (size_t)(off + tau) * A;
        for (int r = 0; r < beam; ++r) {
            double p = 0.0;
            // four independent exp / division chains per lane in flight (the evaluation is latency bound), summed in the
            // order of the per-row kernel: a = lane, lane + 32, ...
            for (int a0 = lane; a0 < A; a0 += 128) {
                double t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int a = a0 + 32 * j;
                    t[j] = a < A ? tanh_exp((double)__ldg(hrow + a) + y_s[r * A + a]) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (a0 + 32 * j < A) p = fma(t[j], v_s[a0 + 32 * j], p);
            }
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            if (lane == 0) s_s[r * Tmax + tau] = p;
        }
    }
    __syncthreads();
    // softmax per hypothesis: one warp each; chunk c of 32 positions = warp c of the per-row kernel (T <= 256)
    for (int r = warp; r < beam; r += nw) {
        double* sr = s_s + r * Tmax;
        double mx = -INFINITY;
        for (int tau = lane; tau < T; tau += 32) mx = fmax(mx, sr[tau]);
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sum = 0.0;
        for (int c = 0; c < nw; ++c) {
            const int tau = 32 * c + lane;
            double e = 0.0;
            if (tau < T) { e = exp(sr[tau] - mx); sr[tau] = e; }
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            sum += e;
        }
        __syncwarp();
        for (int tau = lane; tau < T; tau += 32) sr[tau] = sr[tau] / sum;
    }
    __syncthreads();
    // read-out ctx[r][d] = sum_tau alpha[r][tau] enc[tau][d]
    for (int d = tid; d < D; d += 256) {
        double c[MAXB];
#pragma unroll
        for (int r = 0; r < MAXB; ++r) c[r] = 0.0;
        for (int tau0 = 0; tau0 < T; tau0 += 8) {       // eight encoder rows in flight per thread, summed in row order
            float ev[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ev[i] = tau0 + i < T ? __ldg(enc + (size_t)(off + tau0 + i) * D + d) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (tau0 + i < T) {
                    const double e = (double)ev[i];
#pragma unroll
                    for (int r = 0; r < MAXB; ++r)
                        if (r < beam) c[r] = fma(s_s[r * Tmax + tau0 + i], e, c[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < MAXB; ++r)
            if (r < beam) ctx[(size_t)(r0 + r) * ldctx + d] = c[r];
    }
}
int attn_beam_group_f64(cudaStream_t st, int N, int beam, int A, int D, int Tmax, const float* HF, const float* enc,
                        const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx) {
    if (N <= 0 || beam <= 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)beam * A + A + (size_t)beam * Tmax);
    if (beam > MAXB || Tmax > 256 || smem > 200 * 1024)      // the per-row kernel serves any shape
        return attn_beam_f64(st, N * beam, A, D, Tmax, HF, enc, row_off, Tlen, y, v, ctx, ldctx);
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_beam_group_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_beam_group_f64_kernel<<<N, 256, smem, st>>>(beam, A, D, Tmax, HF, enc, row_off, Tlen, y, v, ctx, ldctx);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- the same read-out from exponentials ----------------------------------------------------------------------------
// tanh(h + y) = 1 - 2 / (exp(2h) exp(2y) + 1): exp(2 HF) depends only on the utterance (a table filled once per
// decode, e2e_exp2x_f64) and exp(2 y) only on (hypothesis, a) -- `beam` x A values per step and utterance -- so the
// T x beam x A inner loop of the scores is a multiplication, an addition and a division instead of an exp and a
// division (the loop is bound by the FP64 pipe: ~3x fewer FP64 instructions).  Both exponents are clamped to
// [-300, 300], so the product neither overflows to inf * 0 nor underflows to 0 * inf; inside the clamp the result
// differs from tanh_exp(h + y) by the rounding of one product (absolute error < 3e-16 like tanh_exp itself), outside
// it tanh is +-1 to machine precision unless the two arguments cancel to within a few units -- pre-activations of
// magnitude 150, which a trained attention layer does not produce.
constexpr double EXP2X_CLAMP = 300.0;
__device__ __forceinline__ double exp2x(double x) { return exp(fmin(fmax(2.0 * x, -EXP2X_CLAMP), EXP2X_CLAMP)); }
__global__ void exp2x_f64_kernel(size_t n, const float* __restrict__ x, double* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = exp2x((double)x[i]);
}
int exp2x_f64(cudaStream_t st, size_t n, const float* x, double* out) {
    if (n == 0) return 0;
    exp2x_f64_kernel<<<(unsigned)min((n + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(n, x, out);
    E2E_LAUNCH_CHECK();
    return 0;
}
// 1 / x for a normal positive x: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~2^-20) + two Newton steps -- within an ulp of
// the quotient, without the range checks and the slow-path branch of the generic division (x = exp(.) exp(.) + 1 lies
// in [1, 1e261] here), so the chains of neighbouring elements interleave.
__device__ __forceinline__ double rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// attn_beam_group_f64_kernel with EHF = exp2x(HF) in place of HF; scores summed in the same order, softmax and
// read-out identical.  16 warps per utterance (the evaluation is latency bound: dependent FP64 chains), the scores of
// all hypotheses of a frame kept in registers so that their shuffle reductions overlap.
constexpr int ATT_E_THREADS = 512;
__global__ void __launch_bounds__(ATT_E_THREADS, 2)
attn_beam_group_e_f64_kernel(int beam, int A, int D, int Tmax, const double* __restrict__ EHF,
                             const float* __restrict__ enc, const int* __restrict__ row_off,
                             const int* __restrict__ Tlen, const double* __restrict__ y, const float* __restrict__ v,
                             double* __restrict__ ctx, int ldctx) {
    extern __shared__ double sm[];
    double* y_s = sm;                       // [beam][A]  exp(2 y)
    double* v_s = y_s + beam * A;           // [A]
    double* s_s = v_s + A;                  // [beam][Tmax]  scores, then exp, then alpha
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nwarps = ATT_E_THREADS / 32;
    const int r0 = u * beam;
    const int off = row_off[r0], T = Tlen[r0];
    for (int i = tid; i < beam * A; i += ATT_E_THREADS) y_s[i] = exp2x(y[(size_t)(r0 + i / A) * A + i % A]);
    for (int a = tid; a < A; a += ATT_E_THREADS) v_s[a] = (double)v[a];
    __syncthreads();
    // scores s[r][tau] = sum_a (1 - 2 / (EHF[tau][a] exp(2 y[r][a]) + 1)) v[a], a = lane, lane + 32, ... in that order
    for (int tau = warp; tau < T; tau += nwarps) {
        const double* hrow = EHF + (size_t)(off + tau) * A;
        double p[MAXB];
#pragma unroll
        for (int r = 0; r < MAXB; ++r) p[r] = 0.0;
        for (int a0 = lane; a0 < A; a0 += 128) {
            double eh[4], vv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int a = a0 + 32 * j;
                eh[j] = a < A ? __ldg(hrow + a) : 0.0;
                vv[j] = a < A ? v_s[a] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < MAXB; ++r) {
                if (r < beam) {
                    double t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int a = a0 + 32 * j;
                        t[j] = a < A ? fma(-2.0, rcp_pos(fma(eh[j], y_s[r * A + a], 1.0)), 1.0) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (a0 + 32 * j < A) p[r] = fma(t[j], vv[j], p[r]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < MAXB; ++r)
                if (r < beam) p[r] += __shfl_xor_sync(0xffffffffu, p[r], o);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < MAXB; ++r)
                if (r < beam) s_s[r * Tmax + tau] = p[r];
        }
    }
    __syncthreads();
    // softmax per hypothesis: one warp each; eight chunks of 32 positions, summed chunk by chunk like the per-row
    // kernel's eight warps (T <= 256)
    for (int r = warp; r < beam; r += nwarps) {
        double* sr = s_s + r * Tmax;
        double mx = -INFINITY;
        for (int tau = lane; tau < T; tau += 32) mx = fmax(mx, sr[tau]);
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sum = 0.0;
        for (int c = 0; c < 8; ++c) {
            const int tau = 32 * c + lane;
            double e = 0.0;
            if (tau < T) { e = exp(sr[tau] - mx); sr[tau] = e; }
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            sum += e;
        }
        __syncwarp();
        for (int tau = lane; tau < T; tau += 32) sr[tau] = sr[tau] / sum;
    }
    __syncthreads();
    // read-out ctx[r][d] = sum_tau alpha[r][tau] enc[tau][d]
    for (int d = tid; d < D; d += ATT_E_THREADS) {
        double c[MAXB];
#pragma unroll
        for (int r = 0; r < MAXB; ++r) c[r] = 0.0;
        for (int tau0 = 0; tau0 < T; tau0 += 8) {
            float ev[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ev[i] = tau0 + i < T ? __ldg(enc + (size_t)(off + tau0 + i) * D + d) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (tau0 + i < T) {
                    const double e = (double)ev[i];
#pragma unroll
                    for (int r = 0; r < MAXB; ++r)
                        if (r < beam) c[r] = fma(s_s[r * Tmax + tau0 + i], e, c[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < MAXB; ++r)
            if (r < beam) ctx[(size_t)(r0 + r) * ldctx + d] = c[r];
    }
}
int attn_beam_group_e_f64(cudaStream_t st, int N, int beam, int A, int D, int Tmax, const double* EHF, const float* enc,
                          const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx) {
    if (N <= 0 || beam <= 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)beam * A + A + (size_t)beam * Tmax);
    E2E_REQUIRE(beam <= MAXB && Tmax <= 256 && smem <= 200 * 1024,
                "attn_beam_group_e_f64: beam <= %d, Tmax <= 256 required (beam=%d Tmax=%d)", MAXB, beam, Tmax);
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_beam_group_e_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_beam_group_e_f64_kernel<<<N, ATT_E_THREADS, smem, st>>>(beam, A, D, Tmax, EHF, enc, row_off, Tlen, y, v, ctx, ldctx);
    E2E_LAUNCH_CHECK();
    return 0;
}

// get_top_k tail (beam_search.py:196-214): combined = log(softmax(dec)) + lm_weight*log(softmax(lm));
// the k largest entries per row (as a set, like np.argpartition), written in descending score order.
__global__ void __launch_bounds__(256)
logsoftmax_topk_f64_kernel(int V, const double* __restrict__ logits, const double* __restrict__ lm_logits,
                           double lm_weight, const int* __restrict__ krow, int kmax, int* __restrict__ out_idx,
                           double* __restrict__ out_val, double* __restrict__ scratch) {
    __shared__ double red[32];
    __shared__ int redi[32];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nw = 8;
    const double* x = logits + (size_t)r * V;
    double* comb = scratch + (size_t)r * V;
    auto block_max = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double m = red[0];
        for (int w = 1; w < nw; ++w) m = fmax(m, red[w]);
        __syncthreads();
        return m;
    };
    auto block_sum = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w];
        __syncthreads();
        return s;
    };
    double mx = -INFINITY;
    for (int v = tid; v < V; v += 256) mx = fmax(mx, x[v]);
    mx = block_max(mx);
    double s = 0.0;
    for (int v = tid; v < V; v += 256) s += exp(x[v] - mx);
    s = block_sum(s);
    for (int v = tid; v < V; v += 256) comb[v] = log(exp(x[v] - mx) / s);
    if (lm_logits != nullptr) {
        const double* xl = lm_logits + (size_t)r * V;
        double ml = -INFINITY;
        for (int v = tid; v < V; v += 256) ml = fmax(ml, xl[v]);
        ml = block_max(ml);
        double sl = 0.0;
        for (int v = tid; v < V; v += 256) sl += exp(xl[v] - ml);
        sl = block_sum(sl);
        for (int v = tid; v < V; v += 256) comb[v] += lm_weight * log(exp(xl[v] - ml) / sl);
    }
    __syncthreads();
    const int k = krow[r];
    for (int j = 0; j < kmax; ++j) {
        if (j >= k) {
            if (tid == 0) { out_idx[(size_t)r * kmax + j] = -1; out_val[(size_t)r * kmax + j] = -INFINITY; }
            continue;
        }
        double best = -INFINITY;
        int bi = 0x7fffffff;
        for (int v = tid; v < V; v += 256) {
            double c = comb[v];
            if (c > best || (c == best && v < bi)) { best = c; bi = v; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { red[warp] = best; redi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nw; ++w)
                if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
            out_idx[(size_t)r * kmax + j] = bi;
            out_val[(size_t)r * kmax + j] = best;
            comb[bi] = -INFINITY;
        }
        __syncthreads();
    }
}
// The same row with its entries in registers (V <= 256 PER): the k selection rounds of the kernel above re-read the
// combined scores from global memory and take two CTA barriers and a one-thread scan each; here a round is a register
// scan, a warp shuffle tree and ONE barrier (the warp results alternate between two shared buffers, and every thread
// scans the eight of them itself).  Sums, maxima and tie rules are those of the kernel above, so the output is
// bit-identical.
template <int PER>
__global__ void __launch_bounds__(256)
logsoftmax_topk_reg_kernel(int V, const double* __restrict__ logits, const double* __restrict__ lm_logits,
                           double lm_weight, const int* __restrict__ krow, int kmax, int* __restrict__ out_idx,
                           double* __restrict__ out_val) {
    __shared__ double red[2][8];
    __shared__ int redi[2][8];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32, nw = 8;
    auto block_max = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[0][warp] = v;
        __syncthreads();
        double m = red[0][0];
        for (int w = 1; w < nw; ++w) m = fmax(m, red[0][w]);
        __syncthreads();
        return m;
    };
    auto block_sum = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[0][warp] = v;
        __syncthreads();
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[0][w];
        __syncthreads();
        return s;
    };
    auto log_softmax = [&](const double* x, double (&out)[PER]) {
        double mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int v = tid + 256 * i;
            out[i] = v < V ? x[v] : -INFINITY;
            mx = fmax(mx, out[i]);
        }
        mx = block_max(mx);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (tid + 256 * i < V) { out[i] = exp(out[i] - mx); s += out[i]; }       // kept: not evaluated twice
        s = block_sum(s);
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (tid + 256 * i < V) out[i] = log(out[i] / s);
    };
    double comb[PER];
    log_softmax(logits + (size_t)r * V, comb);
    if (lm_logits != nullptr) {
        double lm[PER];
        log_softmax(lm_logits + (size_t)r * V, lm);
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (tid + 256 * i < V) comb[i] += lm_weight * lm[i];
    }
    const int k = krow[r];
    for (int j = 0; j < kmax; ++j) {
        if (j >= k) {
            if (tid == 0) { out_idx[(size_t)r * kmax + j] = -1; out_val[(size_t)r * kmax + j] = -INFINITY; }
            continue;
        }
        double best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int v = tid + 256 * i;
            if (v < V && (comb[i] > best || (comb[i] == best && v < bi))) { best = comb[i]; bi = v; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { red[j & 1][warp] = best; redi[j & 1][warp] = bi; }
        __syncthreads();
        best = red[j & 1][0];
        bi = redi[j & 1][0];
        for (int w = 1; w < nw; ++w) {
            const double ob = red[j & 1][w];
            const int oi = redi[j & 1][w];
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (tid == 0) { out_idx[(size_t)r * kmax + j] = bi; out_val[(size_t)r * kmax + j] = best; }
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (tid + 256 * i == bi) comb[i] = -INFINITY;            // taken
    }
}
int logsoftmax_topk_f64(cudaStream_t st, int n, int V, const double* logits, const double* lm_logits,
                        double lm_weight, const int* krow, int kmax, int* out_idx, double* out_val, double* scratch) {
    if (n <= 0) return 0;
    if (V <= 256 * 4)
        logsoftmax_topk_reg_kernel<4><<<n, 256, 0, st>>>(V, logits, lm_logits, lm_weight, krow, kmax, out_idx, out_val);
    else if (V <= 256 * 8)
        logsoftmax_topk_reg_kernel<8><<<n, 256, 0, st>>>(V, logits, lm_logits, lm_weight, krow, kmax, out_idx, out_val);
    else
        logsoftmax_topk_f64_kernel<<<n, 256, 0, st>>>(V, logits, lm_logits, lm_weight, krow, kmax, out_idx, out_val, scratch);
    E2E_LAUNCH_CHECK();
    return 0;
}

// rows of a float32 table widened to float64: out[r, :E] = (double) emb[ids[r], :]
__global__ void embed_gather_f64_kernel(int n, int E, const float* __restrict__ emb, const long long* __restrict__ ids,
                                        double* __restrict__ out, int ldo) {
    int r = blockIdx.x;
    for (int e = threadIdx.x; e < E; e += blockDim.x) out[(size_t)r * ldo + e] = (double)emb[(size_t)ids[r] * E + e];
}
int embed_gather_f64(cudaStream_t st, int n, int E, const float* emb, const long long* ids, double* out, int ldo) {
    if (n <= 0) return 0;
    embed_gather_f64_kernel<<<n, 128, 0, st>>>(n, E, emb, ids, out, ldo);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- candidate merge of one decoding step, all utterances (beam_search.py:255-266 for the GO step, :294-329) ----
// Hypotheses live in FIXED slots: utterance u owns rows [u*beam, (u+1)*beam), live ones first.  One CTA (one warp) per
// utterance: the candidates are val[row][j] + score[row] over the live rows and j < k (k = the utterance's current
// beam size), in (row, j) order like the reference's concatenation; the k best are taken -- the reference's
// np.argpartition(all_scores, -k)[-k:] as a SET; the order inside the set only decides float64 ties -- here in
// descending score order, ties to the lower candidate index.  A candidate that emits EOS retires to the utterance's
// final list (k shrinks), the others become the new live rows with back-pointers; the histories are reconstructed
// from par_hist / tok_hist on the host at the end.  `step` is read from device memory so that the launch can be
// replayed from a CUDA graph.
__global__ void __launch_bounds__(32)
beam_merge_kernel(e2e_beam_merge_args a) {
    extern __shared__ double sm_d[];
    const int u = blockIdx.x, lane = threadIdx.x, beam = a.beam;
    const int step = *a.step;
    double* csc = sm_d;                                  // [beam*beam] candidate scores
    int* ctok = reinterpret_cast<int*>(csc + beam * beam);
    int* crow = ctok + beam * beam;
    const int k = a.k_u[u];
    const int base = u * beam;
    // gather the candidates (sequential over rows: at most beam rows)
    int nc = 0;
    for (int slot = 0; slot < beam; ++slot) {
        const int row = base + slot;
        if (!a.alive[row]) continue;
        const double sc = a.score[row];
        for (int j = lane; j < k; j += 32) {
            csc[nc + j] = a.out_val[(size_t)row * beam + j] + sc;
            ctok[nc + j] = a.out_idx[(size_t)row * beam + j];
            crow[nc + j] = row;
        }
        nc += k;
    }
    __syncwarp();
    const double pen = step > 0 ? a.word_ins_penalty * (double)(step + 1) : 0.0;
    int n_new = 0, k_left = k, nfin = a.fin_cnt[u];
    for (int i = 0; i < k; ++i) {
        double best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) {
            const double v = csc[c];
            if (ctok[c] >= 0 && (v > best || (v == best && c < bi))) { best = v; bi = c; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (bi == 0x7fffffff) break;                     // fewer candidates than k (cannot happen for k <= beam)
        const int tok = ctok[bi], prow = crow[bi];
        __syncwarp();
        if (lane == 0) {
            ctok[bi] = -1;                               // taken
            const double sc = best + pen;
            if (tok == a.eos_id) {
                if (nfin < beam) {
                    a.fin_step[base + nfin] = step;
                    a.fin_row[base + nfin] = prow;
                    a.fin_score[base + nfin] = sc;
                }
                ++nfin;
                --k_left;
            } else {
                const int row = base + n_new;
                a.new_tok[row] = tok;
                a.new_score[row] = sc;
                a.parent[row] = prow;
                a.par_hist[(size_t)step * a.R + row] = prow;
                a.tok_hist[(size_t)step * a.R + row] = tok;
                ++n_new;
            }
        }
        n_new = __shfl_sync(0xffffffffu, n_new, 0);
        k_left = __shfl_sync(0xffffffffu, k_left, 0);
        nfin = __shfl_sync(0xffffffffu, nfin, 0);
        __syncwarp();
    }
    // dead slots: finite dummies (their rows still flow through the batched decoder step)
    for (int slot = n_new + lane; slot < beam; slot += 32) {
        const int row = base + slot;
        a.new_tok[row] = 0;
        a.new_score[row] = 0.0;
        a.parent[row] = base;
        a.par_hist[(size_t)step * a.R + row] = -1;
        a.tok_hist[(size_t)step * a.R + row] = -1;
    }
    for (int slot = lane; slot < beam; slot += 32) {
        a.new_alive[base + slot] = slot < n_new ? 1 : 0;
        a.krow[base + slot] = slot < n_new ? k_left : 0;
    }
    if (lane == 0) {
        a.k_u[u] = k_left;
        a.fin_cnt[u] = nfin;
        if (k_left > 0) atomicAdd(a.n_live, k_left);
    }
}

int beam_merge(cudaStream_t st, const e2e_beam_merge_args* a) {
    if (a->N <= 0) return 0;
    E2E_REQUIRE(a->beam >= 1 && a->beam <= 64, "beam_merge: beam size %d out of [1, 64]", a->beam);
    const size_t smem = (size_t)a->beam * a->beam * (sizeof(double) + 2 * sizeof(int));
    beam_merge_kernel<<<a->N, 32, smem, st>>>(*a);
    E2E_LAUNCH_CHECK();
    return 0;
}

// out[r, :] = in[parent[r], :] for several float64 state matrices at once (the back-pointer gather of a beam step:
// BeamEntry state_list / context_vec of the parent hypothesis, beam_search.py:313-318)
__global__ void beam_gather_kernel(int R, const int* __restrict__ parent, int nmat, e2e_beam_gather_args g) {
    const int r = blockIdx.x;
    const int p = parent[r];
    for (int m = 0; m < nmat; ++m) {
        const double* src = g.src[m] + (size_t)p * g.width[m];
        double* dst = g.dst[m] + (size_t)r * g.width[m];
        for (int i = threadIdx.x; i < g.width[m]; i += blockDim.x) dst[i] = src[i];
    }
}
int beam_gather(cudaStream_t st, int R, const int* parent, const e2e_beam_gather_args* g) {
    if (R <= 0) return 0;
    E2E_REQUIRE(g->nmat >= 0 && g->nmat <= 8, "beam_gather: %d matrices (max 8)", g->nmat);
    beam_gather_kernel<<<R, 128, 0, st>>>(R, parent, g->nmat, *g);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
