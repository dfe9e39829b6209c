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

// ---- CTC: one warp per utterance, scaled forward-backward, fused gradient -----
// logits rows are addressed as (b*sb + t*st)*C (so batch-major encoder states
// projected by a GEMM need no transpose); lse_rows holds the per-row softmax
// normaliser.  Lane l owns the SPL consecutive extended-label states
// [l*SPL, (l+1)*SPL); neighbours are exchanged with shuffles.
//
// Numerics: alpha_t and beta_t are kept in the LINEAR domain, renormalised to sum
// 1 at every frame (Rabiner scaling): no log/exp in the recursions, fp32 relative
// error stays ~1e-7 per frame instead of the ~1e-5 absolute error a float log-space
// recursion accumulates.  log p = sum_t log c_t (double accumulator).  The state
// occupancy gamma_t(s) = alpha_t(s) beta_t(s) / (p y_t(s)) sums to 1 over s for
// every t, so it is obtained by normalising alpha^ beta^ / y per frame (products in
// double to survive underflow):
//   grad_t(k) = softmax_t(k) - sum_{s in lab(k)} gamma_t(s),  zero for t >= len.
// alpha^ is spilled to `alpha_ws` [B][T][S_max] for the beta/gradient sweep.
// Infeasible labels give loss = +inf (TF raises); callers supply feasible labels.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Kernel A: the two sequential sweeps.  Only the S gathered emissions per frame
// are touched here (the next frame's are prefetched while the current one is
// processed); gamma_t(s) overwrites alpha^_t(s) in `ws` [B][T][S_max].
template <int SPL>
__global__ void __launch_bounds__(32)
ctc_sweep_kernel(int T, int B, int C, long long sb, long long stt, const float* __restrict__ logits,
                 const float* __restrict__ lse_rows, const int* __restrict__ in_lens,
                 const long long* __restrict__ labels, int ldl, const int* __restrict__ label_lens,
                 float* __restrict__ ws, int S_max, float* __restrict__ loss_b) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int Tb = min(in_lens[b], T), L = label_lens[b];
    const int S = 2 * L + 1, blank = C - 1;
    int ext[SPL];
    bool skip[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        int s = lane * SPL + i;
        int e = blank;
        bool sk = false;
        if (s < S && (s & 1)) {
            e = (int)labels[(size_t)b * ldl + s / 2];
            if (s >= 3) sk = e != (int)labels[(size_t)b * ldl + s / 2 - 1];
        }
        ext[i] = e;
        skip[i] = sk;
    }
    float* aw = ws + (size_t)b * T * S_max;
    auto emissions = [&](int t, float (&y)[SPL]) {
        const float* row = logits + ((size_t)b * sb + (size_t)t * stt) * C;
        const float nz = lse_rows[(size_t)b * sb + (size_t)t * stt];
#pragma unroll
        for (int i = 0; i < SPL; ++i) y[i] = (lane * SPL + i < S) ? expf(row[ext[i]] - nz) : 0.f;
    };
    float a[SPL], y[SPL], yn[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) { a[i] = 0.f; yn[i] = 0.f; }
    double logp = 0.0;
    if (Tb > 0) emissions(0, yn);
    // ---- alpha sweep
    for (int t = 0; t < Tb; ++t) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) y[i] = yn[i];
        if (t + 1 < Tb) emissions(t + 1, yn);
        float prev1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
        float prev2 = __shfl_up_sync(0xffffffffu, a[SPL >= 2 ? SPL - 2 : 0], 1);
        if (SPL == 1) prev2 = __shfl_up_sync(0xffffffffu, a[0], 2);
        if (lane == 0) { prev1 = 0.f; prev2 = 0.f; }
        if (SPL == 1 && lane == 1) prev2 = 0.f;
        float na[SPL];
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            int s = lane * SPL + i;
            float v;
            if (t == 0) {
                v = (s <= 1) ? y[i] : 0.f;
            } else {
                float p1 = i >= 1 ? a[i >= 1 ? i - 1 : 0] : prev1;
                float p2 = i >= 2 ? a[i >= 2 ? i - 2 : 0] : (i == 1 ? prev1 : prev2);
                if (SPL == 1) { p1 = prev1; p2 = prev2; }
                v = (a[i] + p1 + (skip[i] ? p2 : 0.f)) * y[i];
            }
            na[i] = v;
            lsum += v;
        }
        float c = warp_sum(lsum);
        float inv = c > 0.f ? 1.0f / c : 0.f;
        logp += (double)logf(c);
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            a[i] = na[i] * inv;
            int s = lane * SPL + i;
            if (s < S) aw[(size_t)t * S_max + s] = a[i];
        }
    }
    {   // p = (alpha_{T-1}(S-1) + alpha_{T-1}(S-2)) * prod c_t
        float mine = 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            int s = lane * SPL + i;
            if (Tb > 0 && (s == S - 1 || (s == S - 2 && S >= 2))) mine += a[i];
        }
        mine = warp_sum(mine);
        logp += (double)logf(mine);
    }
    if (lane == 0) loss_b[b] = (float)(-logp);
    // ---- beta sweep; gamma replaces alpha^ in ws
    float bt[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) bt[i] = 0.f;
    if (Tb > 0) emissions(Tb - 1, yn);
    for (int t = Tb - 1; t >= 0; --t) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) y[i] = yn[i];
        float al[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            int s = lane * SPL + i;
            al[i] = (s < S) ? aw[(size_t)t * S_max + s] : 0.f;
        }
        if (t > 0) emissions(t - 1, yn);
        float nxt1 = __shfl_down_sync(0xffffffffu, bt[0], 1);
        float nxt2 = __shfl_down_sync(0xffffffffu, bt[SPL >= 2 ? 1 : 0], 1);
        if (SPL == 1) nxt2 = __shfl_down_sync(0xffffffffu, bt[0], 2);
        bool nskip1 = __shfl_down_sync(0xffffffffu, (int)skip[0], 1);
        bool nskip2 = __shfl_down_sync(0xffffffffu, (int)skip[SPL >= 2 ? 1 : 0], 1);
        if (SPL == 1) nskip2 = __shfl_down_sync(0xffffffffu, (int)skip[0], 2);
        if (lane == 31) { nxt1 = 0.f; nxt2 = 0.f; nskip1 = false; nskip2 = false; }
        if (SPL == 1 && lane == 30) { nxt2 = 0.f; nskip2 = false; }
        float nb[SPL];
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            int s = lane * SPL + i;
            float v;
            if (t == Tb - 1) {
                v = (s == S - 1 || s == S - 2) ? y[i] : 0.f;
            } else {
                float n1 = (i + 1 < SPL) ? bt[(i + 1 < SPL) ? i + 1 : 0] : nxt1;
                float n2;
                bool sk2;
                if (SPL == 1) { n1 = nxt1; n2 = nxt2; sk2 = nskip2; }
                else if (i + 2 < SPL) { n2 = bt[(i + 2 < SPL) ? i + 2 : 0]; sk2 = skip[(i + 2 < SPL) ? i + 2 : 0]; }
                else if (i + 2 == SPL) { n2 = nxt1; sk2 = nskip1; }
                else { n2 = nxt2; sk2 = nskip2; }
                v = (bt[i] + n1 + (sk2 ? n2 : 0.f)) * y[i];
            }
            nb[i] = v;
            lsum += v;
        }
        float d = warp_sum(lsum);
        float invd = d > 0.f ? 1.0f / d : 0.f;
        double gm[SPL];
        double zs = 0.0;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            bt[i] = nb[i] * invd;
            double g = 0.0;
            if (y[i] > 0.f) g = (double)al[i] * (double)bt[i] * (double)(1.0f / y[i]);
            gm[i] = g;
            zs += g;
        }
        zs = warp_sum_d(zs);
        // zs is O(1e-38 .. 1): scale by its float reciprocal of the renormalised value
        const double invz = zs > 0.0 ? 1.0 / zs : 0.0;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            int s = lane * SPL + i;
            if (s < S) aw[(size_t)t * S_max + s] = (float)(gm[i] * invz);
        }
    }
}

// Kernel B: one warp per (utterance, frame) row, fully parallel:
//   grad[k] = out_scale * (softmax(k) - sum_{s: ext(s)=k} gamma(s)),   0 for t >= len.
constexpr int CTC_ROWS_PER_CTA = 8;
__global__ void __launch_bounds__(32 * CTC_ROWS_PER_CTA)
ctc_grad_kernel(int T, int B, int C, long long sb, long long stt, const float* __restrict__ logits,
                const float* __restrict__ lse_rows, const int* __restrict__ in_lens,
                const long long* __restrict__ labels, int ldl, const int* __restrict__ label_lens,
                const float* __restrict__ ws, int S_max, float* __restrict__ grad, float out_scale) {
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
    __syncwarp();
    const float* g = ws + ((size_t)b * T + t) * S_max;
    for (int s = lane; s < S; s += 32) {
        float v = g[s];
        int e = (s & 1) ? (int)labels[(size_t)b * ldl + s / 2] : blank;
        if (v != 0.f) atomicAdd(&occ[e], v);
    }
    __syncwarp();
    const float* row = logits + ((size_t)b * sb + (size_t)t * stt) * C;
    const float nz = lse_rows[(size_t)b * sb + (size_t)t * stt];
    for (int k = lane; k < C; k += 32) grow[k] = out_scale * (expf(row[k] - nz) - occ[k]);
}

int ctc_fwd_grad(cudaStream_t st, int T, int B, int C, long long sb, long long stt, const float* logits,
                 const float* lse_rows, const int* in_lens, const long long* labels, int ldl,
                 const int* label_lens, int max_label_len, float* alpha_ws, float* loss_b, float* grad,
                 float out_scale) {
    if (B <= 0 || T <= 0) return 0;
    int S_max = 2 * max_label_len + 1;
    size_t smem = sizeof(float) * C * CTC_ROWS_PER_CTA;
    E2E_REQUIRE(S_max <= 32 * 32, "ctc: label length %d too long (max 511)", max_label_len);
    E2E_REQUIRE(smem <= 200 * 1024, "ctc: %d classes do not fit shared memory", C);
#define CTC_CASE(SPL_)                                                                                   \
    ctc_sweep_kernel<SPL_><<<B, 32, 0, st>>>(T, B, C, sb, stt, logits, lse_rows, in_lens, labels, ldl, \
                                             label_lens, alpha_ws, S_max, loss_b);
    if (S_max <= 32) CTC_CASE(1)
    else if (S_max <= 64) CTC_CASE(2)
    else if (S_max <= 128) CTC_CASE(4)
    else if (S_max <= 256) CTC_CASE(8)
    else if (S_max <= 512) CTC_CASE(16)
    else CTC_CASE(32)
#undef CTC_CASE
    E2E_LAUNCH_CHECK();
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_grad_kernel<<<cdiv((long long)B * T, CTC_ROWS_PER_CTA), 32 * CTC_ROWS_PER_CTA, smem, st>>>(
        T, B, C, sb, stt, logits, lse_rows, in_lens, labels, ldl, label_lens, alpha_ws, S_max, grad, out_scale);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
