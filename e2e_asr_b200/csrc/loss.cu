// Losses: the reference's masked, length-normalised sequence cross-entropy
// (losses.py:6-35) and the builder-defined auxiliary CTC (TF-1.x
// tf.nn.ctc_loss semantics, blank = C-1; no reference code exists for it).
#include "common.cuh"

namespace e2e {

// ---- row log-sum-exp -------------------------------------------------------
// one warp per row; lse[row] = log sum_v exp(x[row, v])
__global__ void row_lse_kernel(int rows, int V, const float* __restrict__ x, int ldx, float* __restrict__ lse) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const float* r = x + (size_t)row * ldx;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, r[v]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(r[v] - mx);
    s = warp_sum(s);
    if (lane == 0) lse[row] = mx + logf(s);
}

int row_lse(cudaStream_t st, int rows, int V, const float* x, int ldx, float* lse) {
    if (rows <= 0) return 0;
    row_lse_kernel<<<cdiv(rows, 8), 256, 0, st>>>(rows, V, x, ldx, lse);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- sequence cross-entropy (losses.py:18-35) ------------------------------
// logits [(U*B), V] time-major rows (t*B+b); targets [U][B] (row stride ldt in
// elements, so a [U+1,B] decoder-input tensor shifted by one row can be passed);
// cost_row = (lse - logit[target]) * [t < len_b] / (len_b * B).
__global__ void ce_rowcost_kernel(int U, int B, int V, const float* __restrict__ logits,
                                  const long long* __restrict__ targets, const int* __restrict__ lens,
                                  float* __restrict__ lse, float* __restrict__ cost) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= U * B) return;
    int t = row / B, b = row % B;
    const float* r = logits + (size_t)row * V;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, r[v]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(r[v] - mx);
    s = warp_sum(s);
    if (lane == 0) {
        float l = mx + logf(s);
        lse[row] = l;
        int len = lens[b];
        cost[row] = (t < len) ? (l - r[targets[row]]) / ((float)len * (float)B) : 0.f;
    }
}

// deterministic single-block sum: out[0] = sum x[0..n)
__global__ void sum_kernel(int n, const float* __restrict__ x, float* __restrict__ out, float scale) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = warp_sum(s);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) tot += red[w];
        out[0] = tot * scale;
    }
}

int ce_fwd(cudaStream_t st, int U, int B, int V, const float* logits, const long long* targets, const int* lens,
           float* lse, float* cost, float* loss) {
    if (U * B <= 0) return 0;
    ce_rowcost_kernel<<<cdiv(U * B, 8), 256, 0, st>>>(U, B, V, logits, targets, lens, lse, cost);
    E2E_LAUNCH_CHECK();
    sum_kernel<<<1, 1024, 0, st>>>(U * B, cost, loss, 1.0f);
    E2E_LAUNCH_CHECK();
    return 0;
}

// dlogits = g * [t<len]/(len*B) * (softmax - onehot), g read from device memory
__global__ void ce_bwd_kernel(int U, int B, int V, const float* __restrict__ logits,
                              const long long* __restrict__ targets, const int* __restrict__ lens,
                              const float* __restrict__ lse, const float* __restrict__ gscale,
                              float* __restrict__ dlogits) {
    int row = blockIdx.x;
    int t = row / B, b = row % B;
    int len = lens[b];
    float* d = dlogits + (size_t)row * V;
    if (t >= len) {
        for (int v = threadIdx.x; v < V; v += blockDim.x) d[v] = 0.f;
        return;
    }
    float w = gscale[0] / ((float)len * (float)B);
    const float* r = logits + (size_t)row * V;
    float l = lse[row];
    int tg = (int)targets[row];
    for (int v = threadIdx.x; v < V; v += blockDim.x) d[v] = w * (expf(r[v] - l) - (v == tg ? 1.f : 0.f));
}

int ce_bwd(cudaStream_t st, int U, int B, int V, const float* logits, const long long* targets, const int* lens,
           const float* lse, const float* gscale, float* dlogits) {
    if (U * B <= 0) return 0;
    ce_bwd_kernel<<<U * B, 256, 0, st>>>(U, B, V, logits, targets, lens, lse, gscale, dlogits);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- CTC: scaled forward-backward in three phases ------------------------------
// logits rows are addressed as (b*sb + t*st)*C (so batch-major encoder states
// projected by a GEMM need no transpose); lse_rows holds the per-row softmax
// normaliser.  Extended label sequence: s even -> blank (= C-1), s odd -> label s/2;
// S = 2L+1 states.
//
//   A  ctc_emit_kernel   (parallel, HBM bound)  Y[b][t][s] = softmax_t(ext(s)): the S emissions a frame needs are
//                        gathered ONCE into a compact row of pitch SP = 32*SPL floats, so that
//   B  ctc_sweep_kernel  (sequential, latency bound) walks contiguous, prefetchable rows: the alpha and the beta
//                        recursion are independent of each other and run on two warps of one CTA, concurrently;
//                        8 utterances share a CTA (16 warps), so a batch of 64 occupies 8 SMs instead of 64 -- the
//                        sweep runs beside the encoder recurrences, whose cluster CTAs each need a whole SM.
//                        Lane l owns the SPL consecutive states [l*SPL, (l+1)*SPL); neighbours come by shuffles.
//   C  ctc_grad_kernel   (parallel, HBM bound)  gamma_t(s) and the gradient row, one warp per (utterance, frame).
//
// Numerics: alpha_t and beta_t are kept in the LINEAR domain, renormalised to sum 1 at every frame (Rabiner
// scaling): no log/exp in the recursions, fp32 relative error stays ~1e-7 per frame instead of the ~1e-5 absolute
// error a float log-space recursion accumulates (the north-star says "log space"; the scaled linear domain computes
// the same quantity more accurately in fp32).  log p = sum_t log c_t (double accumulator).
// The state occupancy gamma_t(s) = alpha_t(s) beta_t(s) / (p y_t(s)) sums to 1 over s for every t, so it is obtained
// by normalising alpha^ beta^ / y per frame (products in double to survive underflow), whatever the scales:
//   grad_t(k) = softmax_t(k) - sum_{s in lab(k)} gamma_t(s),  zero for t >= len.
// Infeasible labels give loss = +inf (TF raises); callers supply feasible labels.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int CTC_ROWS_PER_CTA = 8;        // (utterance, frame) rows per CTA of the parallel kernels

static int ctc_spl(int S_max) {
    const int opts[] = {1, 2, 4, 8, 12, 16, 24, 32};
    for (int o : opts)
        if (S_max <= 32 * o) return o;
    return 0;
}

// Phase A: one warp per (b, t) row.
__global__ void __launch_bounds__(32 * CTC_ROWS_PER_CTA)
ctc_emit_kernel(int T, int B, int C, long long sb, long long stt, const float* __restrict__ logits,
                const float* __restrict__ lse_rows, const int* __restrict__ in_lens,
                const long long* __restrict__ labels, int ldl, const int* __restrict__ label_lens,
                float* __restrict__ Y, int SP) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long long r = (long long)blockIdx.x * CTC_ROWS_PER_CTA + warp;
    if (r >= (long long)B * T) return;
    const int b = (int)(r / T), t = (int)(r % T);
    if (t >= min(in_lens[b], T)) return;
    const int S = 2 * label_lens[b] + 1, blank = C - 1;
    const float* row = logits + ((size_t)b * sb + (size_t)t * stt) * C;
    const float nz = lse_rows[(size_t)b * sb + (size_t)t * stt];
    const float yb = expf(row[blank] - nz);
    float* y = Y + ((size_t)b * T + t) * SP;
    for (int s = lane; s < SP; s += 32) {
        float v = 0.f;
        if (s < S) v = (s & 1) ? expf(row[(int)labels[(size_t)b * ldl + s / 2]] - nz) : yb;
        y[s] = v;
    }
}

template <int SPL>
__device__ __forceinline__ void ctc_load_row(float (&d)[SPL], const float* __restrict__ p) {
    if constexpr (SPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < SPL / 4; ++i) {
            const float4 v = reinterpret_cast<const float4*>(p)[i];
            d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
        }
    } else if constexpr (SPL == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        d[0] = v.x; d[1] = v.y;
    } else {
        d[0] = p[0];
    }
}
template <int SPL>
__device__ __forceinline__ void ctc_store_row(double* __restrict__ p, const double (&d)[SPL]) {
    if constexpr (SPL % 2 == 0) {
#pragma unroll
        for (int i = 0; i < SPL / 2; ++i) reinterpret_cast<double2*>(p)[i] = make_double2(d[2 * i], d[2 * i + 1]);
    } else {
        p[0] = d[0];
    }
}
__device__ __forceinline__ double shfl_up_d(double v, int n) { return __shfl_up_sync(0xffffffffu, v, n); }
__device__ __forceinline__ double shfl_down_d(double v, int n) { return __shfl_down_sync(0xffffffffu, v, n); }

// Phase B: warp 2u = alpha sweep, warp 2u+1 = beta sweep of utterance blockIdx.x*UPC + u.
// DEPTH frames of emissions are kept in flight in registers.  The scaled alpha^ / beta^ are DOUBLES: within one frame
// the states that matter for gamma = alpha beta / (p y) can sit 1e-35 below the frame's largest alpha (long, tightly
// constrained label sequences: T = 1000, S = 1001 at cfg-4), which a float flushes to zero.
template <int SPL, int DEPTH, int UPC>
__global__ void __launch_bounds__(64 * UPC)
ctc_sweep_kernel(int T, int B, const int* __restrict__ in_lens, const long long* __restrict__ labels, int ldl,
                 const int* __restrict__ label_lens, const float* __restrict__ Y, double* __restrict__ Aw,
                 double* __restrict__ Bw, int SP, float* __restrict__ loss_b) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x * UPC + warp / 2;
    if (b >= B) return;
    const bool beta = warp & 1;
    const int Tb = min(in_lens[b], T), L = label_lens[b];
    const int S = 2 * L + 1;
    // sk bit i: state s = lane*SPL+i may be entered from s-2 (alpha) / state s+2 may be entered from s (beta)
    unsigned sk = 0u;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int s = lane * SPL + i + (beta ? 2 : 0);
        if (s < S && (s & 1) && s >= 3 && labels[(size_t)b * ldl + s / 2] != labels[(size_t)b * ldl + s / 2 - 1])
            sk |= 1u << i;
    }
    const size_t base = (size_t)b * T * SP + (size_t)lane * SPL;
    const float* y_b = Y + base;
    double* out_b = (beta ? Bw : Aw) + base;
    float ring[DEPTH][SPL];
    double a[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) a[i] = 0.0;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) ring[d][i] = 0.f;
        const int t = beta ? Tb - 1 - d : d;
        if (d < Tb) ctc_load_row<SPL>(ring[d], y_b + (size_t)t * SP);
    }
    double logp = 0.0;
    for (int n0 = 0; n0 < Tb; n0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int n = n0 + d;                      // n-th frame of this sweep
            if (n < Tb) {
                const int t = beta ? Tb - 1 - n : n;
                float y[SPL];
#pragma unroll
                for (int i = 0; i < SPL; ++i) y[i] = ring[d][i];
                if (n + DEPTH < Tb) ctc_load_row<SPL>(ring[d], y_b + (size_t)(beta ? t - DEPTH : t + DEPTH) * SP);
                double lsum = 0.0;
                if (!beta) {
                    double p1 = shfl_up_d(a[SPL - 1], 1);
                    double p2 = SPL >= 2 ? shfl_up_d(a[SPL >= 2 ? SPL - 2 : 0], 1) : shfl_up_d(a[0], 2);
                    if (lane == 0) { p1 = 0.0; p2 = 0.0; }
                    if (SPL == 1 && lane == 1) p2 = 0.0;
                    // in place, from the highest state down: a[i] still holds frame n-1 for every index below i
#pragma unroll
                    for (int i = SPL - 1; i >= 0; --i) {
                        const int s = lane * SPL + i;
                        double v;
                        if (n == 0) {
                            v = (s <= 1) ? (double)y[i] : 0.0;
                        } else {
                            const double q1 = i >= 1 ? a[i >= 1 ? i - 1 : 0] : p1;
                            const double q2 = i >= 2 ? a[i >= 2 ? i - 2 : 0] : ((i == 1 && SPL >= 2) ? p1 : p2);
                            v = (a[i] + q1 + (((sk >> i) & 1u) ? q2 : 0.0)) * (double)y[i];
                        }
                        a[i] = v;
                        lsum += v;
                    }
                } else {
                    double n1 = shfl_down_d(a[0], 1);
                    double n2 = SPL >= 2 ? shfl_down_d(a[SPL >= 2 ? 1 : 0], 1) : shfl_down_d(a[0], 2);
                    if (lane == 31) { n1 = 0.0; n2 = 0.0; }
                    if (SPL == 1 && lane == 30) n2 = 0.0;
                    // in place, from the lowest state up
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const int s = lane * SPL + i;
                        double v;
                        if (n == 0) {
                            v = (s == S - 1 || s == S - 2) ? (double)y[i] : 0.0;
                        } else {
                            const double q1 = (i + 1 < SPL) ? a[(i + 1 < SPL) ? i + 1 : 0] : n1;
                            const double q2 = (i + 2 < SPL) ? a[(i + 2 < SPL) ? i + 2 : 0]
                                                            : ((i + 2 == SPL && SPL >= 2) ? n1 : n2);
                            v = (a[i] + q1 + (((sk >> i) & 1u) ? q2 : 0.0)) * (double)y[i];
                        }
                        a[i] = v;
                        lsum += v;
                    }
                }
                // normaliser: any positive factor works as long as log p counts the factor that was applied
                const float cf = warp_sum((float)lsum);
                double inv;
                if (cf > 1e-30f) {
                    const float invf = 1.0f / cf;
                    inv = (double)invf;
                    if (!beta) logp -= (double)logf(invf);
                } else {                                // frame mass below float range (or zero: infeasible labels)
                    const double c = warp_sum_d(lsum);
                    inv = c > 0.0 ? 1.0 / c : 0.0;
                    if (!beta) logp += log(c);
                }
#pragma unroll
                for (int i = 0; i < SPL; ++i) a[i] *= inv;
                ctc_store_row<SPL>(out_b + (size_t)t * SP, a);
            }
        }
    }
    if (!beta) {   // p = (alpha_{T-1}(S-1) + alpha_{T-1}(S-2)) * prod c_t
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int s = lane * SPL + i;
            if (Tb > 0 && (s == S - 1 || (s == S - 2 && S >= 2))) mine += a[i];
        }
        mine = warp_sum_d(mine);
        logp += log(mine);
        if (lane == 0) loss_b[b] = (float)(-logp);
    }
}

// Phase C: one warp per (utterance, frame) row, fully parallel:
//   gamma(s) = alpha^(s) beta^(s) / y(s), normalised over s;
//   grad[k] = out_scale * (softmax(k) - sum_{s: ext(s)=k} gamma(s)),   0 for t >= len.
__global__ void __launch_bounds__(32 * CTC_ROWS_PER_CTA)
ctc_grad_kernel(int T, int B, int C, long long sb, long long stt, const float* __restrict__ logits,
                const float* __restrict__ lse_rows, const int* __restrict__ in_lens,
                const long long* __restrict__ labels, int ldl, const int* __restrict__ label_lens,
                const float* __restrict__ Y, const double* __restrict__ Aw, const double* __restrict__ Bw, int SP,
                float* __restrict__ grad, float out_scale) {
    extern __shared__ float occ_all[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const long long r = (long long)blockIdx.x * CTC_ROWS_PER_CTA + warp;
    if (r >= (long long)B * T) return;
    const int b = (int)(r / T), t = (int)(r % T);
    float* occ = occ_all + (size_t)warp * C;
    float* grow = grad + ((size_t)b * sb + (size_t)t * stt) * C;
    if (t >= min(in_lens[b], T)) {
        for (int k = lane; k < C; k += 32) grow[k] = 0.f;
        return;
    }
    const int S = 2 * label_lens[b] + 1, blank = C - 1;
    for (int k = lane; k < C; k += 32) occ[k] = 0.f;
    const size_t off = ((size_t)b * T + t) * SP;
    double zs = 0.0;
    for (int s = lane; s < S; s += 32) {
        const float y = Y[off + s];
        if (y > 0.f) zs += Aw[off + s] * Bw[off + s] / (double)y;
    }
    zs = warp_sum_d(zs);
    const double invz = zs > 0.0 ? 1.0 / zs : 0.0;
    __syncwarp();
    for (int s = lane; s < S; s += 32) {
        const float y = Y[off + s];
        if (y > 0.f) {
            const float v = (float)(Aw[off + s] * Bw[off + s] / (double)y * invz);
            const int e = (s & 1) ? (int)labels[(size_t)b * ldl + s / 2] : blank;
            if (v != 0.f) atomicAdd(&occ[e], v);
        }
    }
    __syncwarp();
    const float* row = logits + ((size_t)b * sb + (size_t)t * stt) * C;
    const float nz = lse_rows[(size_t)b * sb + (size_t)t * stt];
    for (int k = lane; k < C; k += 32) grow[k] = out_scale * (expf(row[k] - nz) - occ[k]);
}

size_t ctc_workspace_floats(int T, int B, int max_label_len) {
    const int spl = ctc_spl(2 * max_label_len + 1);
    return spl ? (size_t)5 * B * T * 32 * spl : 0;        // Y (float) + alpha^, beta^ (double)
}

int ctc_fwd_grad(cudaStream_t st, int T, int B, int C, long long sb, long long stt, const float* logits,
                 const float* lse_rows, const int* in_lens, const long long* labels, int ldl,
                 const int* label_lens, int max_label_len, float* ws, float* loss_b, float* grad,
                 float out_scale) {
    if (B <= 0 || T <= 0) return 0;
    const int S_max = 2 * max_label_len + 1;
    const int spl = ctc_spl(S_max);
    const size_t smem = sizeof(float) * C * CTC_ROWS_PER_CTA;
    E2E_REQUIRE(spl > 0, "ctc: label length %d too long (max 511)", max_label_len);
    E2E_REQUIRE(smem <= 200 * 1024, "ctc: %d classes do not fit shared memory", C);
    const int SP = 32 * spl;
    float* Y = ws;
    double* Aw = reinterpret_cast<double*>(ws + (size_t)B * T * SP);
    double* Bw = Aw + (size_t)B * T * SP;
    E2E_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "ctc: workspace must be 16-byte aligned");
    const int row_ctas = cdiv((long long)B * T, CTC_ROWS_PER_CTA);
    ctc_emit_kernel<<<row_ctas, 32 * CTC_ROWS_PER_CTA, 0, st>>>(T, B, C, sb, stt, logits, lse_rows, in_lens, labels,
                                                               ldl, label_lens, Y, SP);
    E2E_LAUNCH_CHECK();
    // utterances per CTA: 8 (16 warps) while the state registers allow 128 per thread, else 4
#define CTC_CASE(SPL_, D_, UPC_)                                                                              \
    ctc_sweep_kernel<SPL_, D_, UPC_><<<cdiv(B, UPC_), 64 * UPC_, 0, st>>>(T, B, in_lens, labels, ldl, label_lens, Y, \
                                                                          Aw, Bw, SP, loss_b);
    switch (spl) {
        case 1: CTC_CASE(1, 8, 8) break;
        case 2: CTC_CASE(2, 8, 8) break;
        case 4: CTC_CASE(4, 8, 8) break;
        case 8: CTC_CASE(8, 6, 8) break;
        case 12: CTC_CASE(12, 4, 4) break;
        case 16: CTC_CASE(16, 4, 4) break;
        case 24: CTC_CASE(24, 2, 4) break;
        default: CTC_CASE(32, 2, 4) break;
    }
#undef CTC_CASE
    E2E_LAUNCH_CHECK();
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_grad_kernel<<<row_ctas, 32 * CTC_ROWS_PER_CTA, smem, st>>>(T, B, C, sb, stt, logits, lse_rows, in_lens, labels,
                                                                  ldl, label_lens, Y, Aw, Bw, SP, grad, out_scale);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
