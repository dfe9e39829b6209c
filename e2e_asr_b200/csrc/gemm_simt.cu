// fp32 CUDA-core GEMM (FFMA), the exact-fp32 path of the dense contractions
// (encoder/decoder input projections, AttnW, Attn/Output projections and their
// dX/dW twins).  The tensor-core paths live in
// gemm_tc.cu; this kernel also serves every shape those cannot take (tiny test
// sizes, unaligned leading dimensions, the skinny per-step decoder products).
//
//   C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]) (+ Z[M,N]) (+ C if accumulate)
//
// 128x128x16 or 64x64x16 CTA tile, 256 threads, 8x8 / 4x4 register tile per
// thread, operands staged K-major in shared memory so the inner product reads
// float4s.  Split-K (atomicAdd into C) is used when the output has too few
// tiles to fill the 148 SMs (weight gradients: small outputs over T*B rows;
// per-step decoder products with M = batch).
#include "common.cuh"

namespace e2e {

constexpr int BK = 16, NT = 256;

template <int BMt, int BNt, bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(NT)
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                 const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                 const float* __restrict__ bias, const float* __restrict__ Z, int ldz,
                 int accumulate, int ksplit_len) {
    constexpr int TM = BMt / 16, TN = BNt / 16;
    constexpr int PA = BMt / 64, PB = BNt / 64;      // float4 loads per thread per k-tile
    __shared__ __align__(16) float As[2][BK][BMt + 4];
    __shared__ __align__(16) float Bs[2][BK][BNt + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BMt, n0 = blockIdx.x * BNt;
    const int kbeg = blockIdx.z * ksplit_len;
    const int kend = min(K, kbeg + ksplit_len);
    const int tx = tid % 16, ty = tid / 16;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // element (m,k) of op(A): TA ? A[k*lda+m] : A[m*lda+k];  (k,n) of op(B): TB ? B[n*ldb+k] : B[k*ldb+n]
    float ra[PA * 4], rb[PB * 4];
    auto load_op = [&](const float* __restrict__ P, int ld, bool kcontig, int r0, int R, int BT, int npass,
                       float* reg, int k0) {
        if (kcontig) {   // memory contiguous along k: thread -> (r = tid/4 + 64h, k = (tid%4)*4)
            for (int h = 0; h < npass; ++h) {
                int r = r0 + tid / 4 + 64 * h, k = k0 + (tid % 4) * 4;
                if (VEC && r < R && k + 3 < kend) {
                    float4 v = *reinterpret_cast<const float4*>(P + (size_t)r * ld + k);
                    reg[h * 4 + 0] = v.x; reg[h * 4 + 1] = v.y; reg[h * 4 + 2] = v.z; reg[h * 4 + 3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j)
                        reg[h * 4 + j] = (r < R && k + j < kend) ? P[(size_t)r * ld + k + j] : 0.f;
                }
            }
        } else {         // contiguous along the M/N index: thread -> (k = tid/(BT/4) + kpp*h, r = (tid%(BT/4))*4)
            const int tpr = BT / 4, kpp = NT / tpr;
            for (int h = 0; h < npass; ++h) {
                int k = k0 + tid / tpr + kpp * h, r = r0 + (tid % tpr) * 4;
                if (VEC && k < kend && r + 3 < R) {
                    float4 v = *reinterpret_cast<const float4*>(P + (size_t)k * ld + r);
                    reg[h * 4 + 0] = v.x; reg[h * 4 + 1] = v.y; reg[h * 4 + 2] = v.z; reg[h * 4 + 3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j)
                        reg[h * 4 + j] = (k < kend && r + j < R) ? P[(size_t)k * ld + r + j] : 0.f;
                }
            }
        }
    };
    auto load_tiles = [&](int k0) {
        load_op(A, lda, !TA, m0, M, BMt, PA, ra, k0);
        load_op(B, ldb, TB, n0, N, BNt, PB, rb, k0);
    };
    auto store_tiles = [&](int buf) {
        if (!TA) {
#pragma unroll
            for (int h = 0; h < PA; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) As[buf][(tid % 4) * 4 + j][tid / 4 + 64 * h] = ra[h * 4 + j];
        } else {
            constexpr int tpr = BMt / 4, kpp = NT / tpr;
#pragma unroll
            for (int h = 0; h < PA; ++h)
                *reinterpret_cast<float4*>(&As[buf][tid / tpr + kpp * h][(tid % tpr) * 4]) =
                    make_float4(ra[h * 4], ra[h * 4 + 1], ra[h * 4 + 2], ra[h * 4 + 3]);
        }
        if (TB) {
#pragma unroll
            for (int h = 0; h < PB; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[buf][(tid % 4) * 4 + j][tid / 4 + 64 * h] = rb[h * 4 + j];
        } else {
            constexpr int tpr = BNt / 4, kpp = NT / tpr;
#pragma unroll
            for (int h = 0; h < PB; ++h)
                *reinterpret_cast<float4*>(&Bs[buf][tid / tpr + kpp * h][(tid % tpr) * 4]) =
                    make_float4(rb[h * 4], rb[h * 4 + 1], rb[h * 4 + 2], rb[h * 4 + 3]);
        }
    };

    int nk = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;
    if (nk > 0) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int it = 0; it < nk; ++it) {
        int buf = it & 1;
        if (it + 1 < nk) load_tiles(kbeg + (it + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][64 * h + ty * 4]);
                a[h * 4] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][64 * h + tx * 4]);
                b[h * 4] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (it + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (n >= N) continue;
            float v = acc[i][j];
            float* c = C + (size_t)m * ldc + n;
            if (!split || blockIdx.z == 0) {
                if (bias != nullptr) v += bias[n];
                if (Z != nullptr) v += Z[(size_t)m * ldz + n];
            }
            if (split) atomicAdd(c, v);
            else *c = accumulate ? *c + v : v;
        }
    }
}

__global__ void zero_rows_kernel(float* C, int M, int N, int ldc) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)M * N) C[(i / N) * ldc + (i % N)] = 0.f;
}

template <int BMt, int BNt>
static void launch_simt(cudaStream_t st, dim3 grid, int transA, int transB, bool vec, int M, int N, int K,
                        const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                        const float* Z, int ldz, int accumulate, int klen) {
#define LAUNCH(TA_, TB_, V_)                                                                             \
    gemm_simt_kernel<BMt, BNt, TA_, TB_, V_><<<grid, NT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, \
                                                                 Z, ldz, accumulate, klen)
    if (!transA && !transB) { if (vec) LAUNCH(false, false, true); else LAUNCH(false, false, false); }
    else if (transA && !transB) { if (vec) LAUNCH(true, false, true); else LAUNCH(true, false, false); }
    else if (!transA && transB) { if (vec) LAUNCH(false, true, true); else LAUNCH(false, true, false); }
    else { if (vec) LAUNCH(true, true, true); else LAUNCH(true, true, false); }
#undef LAUNCH
}

int gemm_simt(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda,
              const float* B, int ldb, float* C, int ldc, const float* bias, const float* Z, int ldz,
              int accumulate) {
    if (M <= 0 || N <= 0) return 0;
    bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    const int nsm = sm_count();
    bool big = (long long)cdiv(M, 128) * cdiv(N, 128) >= nsm;
    int bm = big ? 128 : 64, bn = big ? 128 : 64;
    int gm = cdiv(M, bm), gn = cdiv(N, bn);
    int tiles = gm * gn;
    int splits = 1;
    if (tiles * 2 <= nsm && K >= 256) {
        splits = min(nsm / tiles, K / 64);
        if (splits < 1) splits = 1;
    }
    int klen = K > 0 ? cdiv(cdiv(K, splits), BK) * BK : BK;
    splits = K > 0 ? cdiv(K, klen) : 1;
    if (splits > 1 && !accumulate)
        zero_rows_kernel<<<cdiv((long long)M * N, 256), 256, 0, st>>>(C, M, N, ldc);
    dim3 grid(gn, gm, splits);
    if (big) launch_simt<128, 128>(st, grid, transA, transB, vec, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate, klen);
    else launch_simt<64, 64>(st, grid, transA, transB, vec, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate, klen);
    E2E_LAUNCH_CHECK();
    return 0;
}

// out[n] (+)= sum_m X[m, n]   (bias gradients).  Thread = 4 adjacent columns (float4 loads),
// rows strided over a 2-D grid with 4 loads in flight per thread; partial sums by atomicAdd.
__global__ void colsum_kernel(int M, int N, const float* __restrict__ X, int ldx, float* __restrict__ out,
                              int accumulate, int rows_per_block, int vec) {
    const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
    if (vec) {
        int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (n >= N) return;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
        int m = m0;
        for (; m + 3 < m1; m += 4) {
            float4 a = __ldg(reinterpret_cast<const float4*>(X + (size_t)m * ldx + n));
            float4 b = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m + 1) * ldx + n));
            float4 c = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m + 2) * ldx + n));
            float4 d = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m + 3) * ldx + n));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
            s2.x += c.x; s2.y += c.y; s2.z += c.z; s2.w += c.w;
            s3.x += d.x; s3.y += d.y; s3.z += d.z; s3.w += d.w;
        }
        for (; m < m1; ++m) {
            float4 a = __ldg(reinterpret_cast<const float4*>(X + (size_t)m * ldx + n));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        }
        float r[4] = {(s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y), (s0.z + s1.z) + (s2.z + s3.z),
                      (s0.w + s1.w) + (s2.w + s3.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (gridDim.y > 1) atomicAdd(out + n + j, r[j]);
            else out[n + j] = accumulate ? out[n + j] + r[j] : r[j];
        }
    } else {
        int n = blockIdx.x * blockDim.x + threadIdx.x;
        if (n >= N) return;
        float s = 0.f;
        for (int m = m0; m < m1; ++m) s += X[(size_t)m * ldx + n];
        if (gridDim.y > 1) atomicAdd(out + n, s);
        else out[n] = accumulate ? out[n] + s : s;
    }
}

int colsum(cudaStream_t st, int M, int N, const float* X, int ldx, float* out, int accumulate) {
    if (N <= 0) return 0;
    const int vec = (N % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const int cols_per_block = vec ? 128 * 4 : 128;
    int gx = cdiv(N, cols_per_block);
    int gy = 1;
    if (M > 256) gy = min(cdiv(M, 64), max(1, 8 * sm_count() / gx));
    int rpb = M > 0 ? cdiv(M, gy) : 1;
    gy = M > 0 ? cdiv(M, rpb) : 1;
    if (gy > 1 && !accumulate) E2E_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
    colsum_kernel<<<dim3(gx, gy), 128, 0, st>>>(M, N, X, ldx, out, accumulate, rpb, vec);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
