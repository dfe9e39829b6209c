// Attention-decoder step kernels (teacher-forced training and greedy/beam
// inference): MLP-attention read-out, decoder-LSTM pointwise step, embedding
// gather/scatter.  Follows reference attn_decoder.py:76-166 in the step order of
// SURVEY.md A.4; the query is the LSTM CELL state c (decoder.py:79-80).
//
// Decoder tensors are time-major [U][B][.] (row = t*B + b), the layout of the
// reference's emitted logits (attn_decoder.py:170).
#include "common.cuh"

namespace e2e {

// ---- embedding ----------------------------------------------------------
__global__ void embed_gather_kernel(int n, int E, const float* __restrict__ emb, const long long* __restrict__ ids,
                                    float* __restrict__ out) {
    int row = blockIdx.x;
    if (row >= n) return;
    const float* src = emb + (size_t)ids[row] * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) out[(size_t)row * E + e] = src[e];
}

__global__ void embed_scatter_add_kernel(int n, int E, float* __restrict__ demb, const long long* __restrict__ ids,
                                         const float* __restrict__ dout, int ldd) {
    int row = blockIdx.x;
    if (row >= n) return;
    float* dst = demb + (size_t)ids[row] * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, dout[(size_t)row * ldd + e]);
}

int embed_gather(cudaStream_t st, int n, int E, const float* emb, const long long* ids, float* out) {
    if (n <= 0) return 0;
    embed_gather_kernel<<<n, 128, 0, st>>>(n, E, emb, ids, out);
    E2E_LAUNCH_CHECK();
    return 0;
}
int embed_scatter_add(cudaStream_t st, int n, int E, float* demb, const long long* ids, const float* dout, int ldd) {
    if (n <= 0) return 0;
    embed_scatter_add_kernel<<<n, 128, 0, st>>>(n, E, demb, ids, dout, ldd);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- decoder LSTM pointwise step ----------------------------------------
// gates_pre [B][4H] TF column order i|j|f|o (basic_lstm.py:17); cprev = committed
// c_{t-1}.  Writes acts (sig i, tanh j, sig(f+1), sig o), c_new (into cat[:, :H]),
// and the committed state for step t+1: live rows take (c_new, h_new), finished
// rows keep (c, h) (raw_rnn copies state through for finished rows).
__global__ void dec_pointwise_fwd_kernel(int B, int H, int t, const float* __restrict__ gates_pre,
                                         const float* __restrict__ cprev, const float* __restrict__ hprev, int ldh,
                                         const int* __restrict__ lens, float* __restrict__ acts,
                                         float* __restrict__ cnew_out, int ldc, float* __restrict__ c_next,
                                         float* __restrict__ h_next, int ldhn, float* __restrict__ h_new_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    int b = i / H, u = i % H;
    const float* g = gates_pre + (size_t)b * 4 * H;
    float si = sigmoidf_acc(g[u]);
    float tj = tanhf(g[H + u]);
    float sf = sigmoidf_acc(g[2 * H + u] + 1.0f);
    float so = sigmoidf_acc(g[3 * H + u]);
    float cp = cprev[(size_t)b * H + u];
    float cn = cp * sf + si * tj;
    float hn = tanhf(cn) * so;
    float* a = acts + (size_t)b * 4 * H;
    a[u] = si; a[H + u] = tj; a[2 * H + u] = sf; a[3 * H + u] = so;
    cnew_out[(size_t)b * ldc + u] = cn;
    if (h_new_out) h_new_out[(size_t)b * H + u] = hn;
    bool live = t < lens[b];
    if (c_next) c_next[(size_t)b * H + u] = live ? cn : cp;
    if (h_next) h_next[(size_t)b * ldhn + u] = live ? hn : hprev[(size_t)b * ldh + u];
}

int dec_pointwise_fwd(cudaStream_t st, int B, int H, int t, const float* gates_pre, const float* cprev,
                      const float* hprev, int ldh, const int* lens, float* acts, float* cnew_out, int ldc,
                      float* c_next, float* h_next, int ldhn, float* h_new_out) {
    dec_pointwise_fwd_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(B, H, t, gates_pre, cprev, hprev, ldh, lens, acts,
                                                               cnew_out, ldc, c_next, h_next, ldhn, h_new_out);
    E2E_LAUNCH_CHECK();
    return 0;
}

// Backward of the step for live rows; finished rows carry no gradient (their
// logits are zeroed and masked out of the loss, losses.py:24-28).
//   dc_in : grad wrt c_new from attention query + AttnProjection (dcat[:, :H])
//   dc_carry / dh_carry : grads wrt the committed (c, h)_t coming from step t+1
// Outputs dz [B][4H] (TF column order) and the new dc_carry.
__global__ void dec_pointwise_bwd_kernel(int B, int H, int t, const float* __restrict__ acts,
                                         const float* __restrict__ cnew, int ldc, const float* __restrict__ cprev,
                                         const float* __restrict__ dc_in, int lddc, float* __restrict__ dc_carry,
                                         const float* __restrict__ dh_carry, int lddh, const int* __restrict__ lens,
                                         float* __restrict__ dz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    int b = i / H, u = i % H;
    float* z = dz + (size_t)b * 4 * H;
    if (t >= lens[b]) {
        z[u] = 0.f; z[H + u] = 0.f; z[2 * H + u] = 0.f; z[3 * H + u] = 0.f;
        dc_carry[(size_t)b * H + u] = 0.f;
        return;
    }
    const float* a = acts + (size_t)b * 4 * H;
    float si = a[u], tj = a[H + u], sf = a[2 * H + u], so = a[3 * H + u];
    float tc = tanhf(cnew[(size_t)b * ldc + u]);
    float dh = dh_carry ? dh_carry[(size_t)b * lddh + u] : 0.f;
    float dc = dc_in[(size_t)b * lddc + u] + dc_carry[(size_t)b * H + u] + dh * so * (1.f - tc * tc);
    z[u] = dc * tj * si * (1.f - si);
    z[H + u] = dc * si * (1.f - tj * tj);
    z[2 * H + u] = dc * cprev[(size_t)b * H + u] * sf * (1.f - sf);
    z[3 * H + u] = dh * tc * so * (1.f - so);
    dc_carry[(size_t)b * H + u] = dc * sf;
}

int dec_pointwise_bwd(cudaStream_t st, int B, int H, int t, const float* acts, const float* cnew, int ldc,
                      const float* cprev, const float* dc_in, int lddc, float* dc_carry, const float* dh_carry,
                      int lddh, const int* lens, float* dz) {
    dec_pointwise_bwd_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(B, H, t, acts, cnew, ldc, cprev, dc_in, lddc,
                                                               dc_carry, dh_carry, lddh, lens, dz);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- attention read-out (attn_decoder.py:77-93) ----------------------------
// One CTA per batch row.  s_tau = sum_a v_a tanh(HF[b,tau,a] + y[b,a]);
// alpha = softmax over ALL positions, then * mask, then renormalised (:85-88) --
// which equals the softmax restricted to tau < enc_len; ctx = sum alpha*enc.
// HF/enc rows are (b*Tp + tau).  alpha is written for tau < Tn (zeros past len).
constexpr int ATT_THREADS = 256;

// ex2.approx + rcp.approx: absolute error ~2e-7, inside the 1e-4 parity budget (the persistent kernels use the same)
__device__ __forceinline__ float att_tanh(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

// VEC: A % 4 == 0, D % 4 == 0 and 16-byte aligned rows -- scores read HF with one float4 per lane, the read-out holds a
// float4 of ctx per thread and splits the time axis over ATT_THREADS / (D / 4) thread groups (independent loads in flight
// instead of one dependent chain per column).
template <bool VEC>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(int Tn, int Tp, int A, int D, const float* __restrict__ HF, const float* __restrict__ enc,
                const int* __restrict__ enc_len, const float* __restrict__ y, const float* __restrict__ v,
                float* __restrict__ alpha, float* __restrict__ ctx, int ldctx) {
    extern __shared__ __align__(16) float sm[];
    float* y_s = sm;            // [A]
    float* v_s = sm + A;        // [A]
    float* s_s = sm + 2 * A;    // [Tn]
    __shared__ float red[32];
    __shared__ __align__(16) float part_s[ATT_THREADS * 4];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int nw = ATT_THREADS / 32;
    const int len = min(enc_len[b], Tn);
    for (int a = tid; a < A; a += ATT_THREADS) { y_s[a] = y[(size_t)b * A + a]; v_s[a] = v[a]; }
    __syncthreads();
    const float* HFb = HF + (size_t)b * Tp * A;
    for (int tau = warp; tau < len; tau += nw) {
        float p = 0.f;
        if (VEC) {
            for (int a4 = lane; a4 < A / 4; a4 += 32) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(HFb + (size_t)tau * A) + a4);
                const float4 yy = *reinterpret_cast<const float4*>(y_s + 4 * a4), vv = *reinterpret_cast<const float4*>(v_s + 4 * a4);
                p += vv.x * att_tanh(h.x + yy.x) + vv.y * att_tanh(h.y + yy.y) + vv.z * att_tanh(h.z + yy.z) + vv.w * att_tanh(h.w + yy.w);
            }
        } else {
            for (int a = lane; a < A; a += 32) p += v_s[a] * att_tanh(HFb[(size_t)tau * A + a] + y_s[a]);
        }
        p = warp_sum(p);
        if (lane == 0) s_s[tau] = p;
    }
    __syncthreads();
    // masked softmax over tau < len
    float mx = -INFINITY;
    for (int tau = tid; tau < len; tau += ATT_THREADS) mx = fmaxf(mx, s_s[tau]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int tau = tid; tau < len; tau += ATT_THREADS) {
        float e = expf(s_s[tau] - mx);
        s_s[tau] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < nw; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int tau = tid; tau < Tn; tau += ATT_THREADS) {
        float al = tau < len ? s_s[tau] * inv : 0.f;
        if (tau < len) s_s[tau] = al;
        alpha[(size_t)b * Tn + tau] = al;
    }
    __syncthreads();
    const float* encb = enc + (size_t)b * Tp * D;
    if (VEC && D / 4 <= ATT_THREADS) {
        const int d4n = D / 4, ng = ATT_THREADS / d4n;
        const int d4 = tid % d4n, gi = tid / d4n;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < ng) {
#pragma unroll 4
            for (int tau = gi; tau < len; tau += ng) {
                const float4 e = __ldg(reinterpret_cast<const float4*>(encb + (size_t)tau * D) + d4);
                const float al = s_s[tau];
                c.x = fmaf(al, e.x, c.x); c.y = fmaf(al, e.y, c.y); c.z = fmaf(al, e.z, c.z); c.w = fmaf(al, e.w, c.w);
            }
        }
        if (ng > 1) {
            *reinterpret_cast<float4*>(part_s + 4 * tid) = c;
            __syncthreads();
            if (gi == 0) {
                for (int g2 = 1; g2 < ng; ++g2) {
                    const float4 o = *reinterpret_cast<const float4*>(part_s + 4 * (g2 * d4n + d4));
                    c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
                }
            }
        }
        if (gi == 0) {
            float* cr = ctx + (size_t)b * ldctx + 4 * d4;
            cr[0] = c.x; cr[1] = c.y; cr[2] = c.z; cr[3] = c.w;
        }
    } else {
        for (int d = tid; d < D; d += ATT_THREADS) {
            float c = 0.f;
            for (int tau = 0; tau < len; ++tau) c = fmaf(s_s[tau], encb[(size_t)tau * D + d], c);
            ctx[(size_t)b * ldctx + d] = c;
        }
    }
}

int attn_fwd(cudaStream_t st, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
             const int* enc_len, const float* y, const float* v, float* alpha, float* ctx, int ldctx) {
    size_t smem = sizeof(float) * (2 * A + Tn);
    if (smem > 44 * 1024) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const bool vec = A % 4 == 0 && D % 4 == 0 && (((uintptr_t)HF | (uintptr_t)enc) & 15) == 0;
    if (vec) attn_fwd_kernel<true><<<B, ATT_THREADS, smem, st>>>(Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, ctx, ldctx);
    else attn_fwd_kernel<false><<<B, ATT_THREADS, smem, st>>>(Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, ctx, ldctx);
    E2E_LAUNCH_CHECK();
    return 0;
}

// Backward of one read-out.  dctx [B][D] (row stride lddctx) is the total gradient
// wrt ctx_t.  Accumulates dHF[b] and denc[b] over decoder steps (exclusive to this
// CTA: no atomics), writes dy [B][A] and accumulates dv_part [B][A].
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(int Tn, int Tp, int A, int D, const float* __restrict__ HF, const float* __restrict__ enc,
                const int* __restrict__ enc_len, const float* __restrict__ y, const float* __restrict__ v,
                const float* __restrict__ alpha, const float* __restrict__ dctx, int lddctx,
                float* __restrict__ dHF, float* __restrict__ denc, float* __restrict__ dy,
                float* __restrict__ dv_part) {
    extern __shared__ float sm[];
    float* y_s = sm;              // [A]
    float* v_s = sm + A;          // [A]
    float* dctx_s = sm + 2 * A;   // [D]
    float* ds_s = dctx_s + D;     // [Tn]  (da, then ds)
    float* acc_s = ds_s + Tn;     // [nw][2][A] cross-warp partials
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int nw = ATT_THREADS / 32;
    const int len = min(enc_len[b], Tn);
    for (int a = tid; a < A; a += ATT_THREADS) { y_s[a] = y[(size_t)b * A + a]; v_s[a] = v[a]; }
    for (int d = tid; d < D; d += ATT_THREADS) dctx_s[d] = dctx[(size_t)b * lddctx + d];
    __syncthreads();
    const float* al = alpha + (size_t)b * Tn;
    const float* encb = enc + (size_t)b * Tp * D;
    float* dencb = denc + (size_t)b * Tp * D;
    // da_tau = dctx . enc_tau ; denc_tau += alpha_tau * dctx
    float part = 0.f;
    for (int tau = warp; tau < len; tau += nw) {
        float a_t = al[tau];
        float p = 0.f;
        for (int d = lane; d < D; d += 32) {
            float dc = dctx_s[d];
            p = fmaf(dc, encb[(size_t)tau * D + d], p);
            dencb[(size_t)tau * D + d] += a_t * dc;
        }
        p = warp_sum(p);
        if (lane == 0) { ds_s[tau] = p; part += a_t * p; }
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    float dot = 0.f;
    for (int w = 0; w < nw; ++w) dot += red[w];
    // ds_tau = alpha_tau (da_tau - sum alpha da)
    for (int tau = tid; tau < len; tau += ATT_THREADS) ds_s[tau] = al[tau] * (ds_s[tau] - dot);
    __syncthreads();
    const float* HFb = HF + (size_t)b * Tp * A;
    float* dHFb = dHF + (size_t)b * Tp * A;
    // per-thread columns a = lane + 32 j; warps stride over tau
    for (int a0 = 0; a0 < A; a0 += 32) {
        int a = a0 + lane;
        float dya = 0.f, dva = 0.f;
        if (a < A) {
            for (int tau = warp; tau < len; tau += nw) {
                float th = tanhf(HFb[(size_t)tau * A + a] + y_s[a]);
                float ds = ds_s[tau];
                float dpre = ds * v_s[a] * (1.f - th * th);
                dHFb[(size_t)tau * A + a] += dpre;
                dya += dpre;
                dva = fmaf(ds, th, dva);
            }
            acc_s[(warp * 2 + 0) * A + a] = dya;
            acc_s[(warp * 2 + 1) * A + a] = dva;
        }
    }
    __syncthreads();
    for (int a = tid; a < A; a += ATT_THREADS) {
        float dya = 0.f, dva = 0.f;
        for (int w = 0; w < nw; ++w) { dya += acc_s[(w * 2) * A + a]; dva += acc_s[(w * 2 + 1) * A + a]; }
        dy[(size_t)b * A + a] = dya;
        dv_part[(size_t)b * A + a] += dva;
    }
}

int attn_bwd(cudaStream_t st, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
             const int* enc_len, const float* y, const float* v, const float* alpha, const float* dctx, int lddctx,
             float* dHF, float* denc, float* dy, float* dv_part) {
    size_t smem = sizeof(float) * (2 * A + D + Tn + (ATT_THREADS / 32) * 2 * A);
    if (smem > 48 * 1024)
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<<<B, ATT_THREADS, smem, st>>>(Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, dctx, lddctx, dHF,
                                                  denc, dy, dv_part);
    E2E_LAUNCH_CHECK();
    return 0;
}

// Per-step part of the attention backward for the step-by-step decoder loop: da, ds and dy only.  The sums over the
// decoder steps (dHF, denc, dv) are formed AFTER the loop from the stored ds_t, alpha_t, y_t and dctx_t
// (dec_deferred_attn_grads, decoder_persist.cu) instead of a read-modify-write of dHF[b] and denc[b] in every step.
template <bool VEC>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_step_kernel(int Tn, int Tp, int A, int D, const float* __restrict__ HF, const float* __restrict__ enc,
                     const int* __restrict__ enc_len, const float* __restrict__ y, const float* __restrict__ v,
                     const float* __restrict__ alpha, const float* __restrict__ dctx, int lddctx,
                     float* __restrict__ ds_out, float* __restrict__ dy) {
    extern __shared__ __align__(16) float sm[];
    float* y_s = sm;              // [A]
    float* v_s = sm + A;          // [A]
    float* dctx_s = sm + 2 * A;   // [D]
    float* ds_s = dctx_s + D;     // [Tn]  (da, then ds)
    float* acc_s = ds_s + Tn;     // [groups][A] partial dy
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int nw = ATT_THREADS / 32;
    const int len = min(enc_len[b], Tn);
    for (int a = tid; a < A; a += ATT_THREADS) { y_s[a] = y[(size_t)b * A + a]; v_s[a] = v[a]; }
    for (int d = tid; d < D; d += ATT_THREADS) dctx_s[d] = dctx[(size_t)b * lddctx + d];
    __syncthreads();
    const float* al = alpha + (size_t)b * Tn;
    const float* encb = enc + (size_t)b * Tp * D;
    // da_tau = dctx . enc_tau
    float part = 0.f;
    for (int tau = warp; tau < len; tau += nw) {
        float p = 0.f;
        if (VEC) {
            const float4* er = reinterpret_cast<const float4*>(encb + (size_t)tau * D);
#pragma unroll 4
            for (int d4 = lane; d4 < D / 4; d4 += 32) {
                const float4 e = __ldg(er + d4), c = *reinterpret_cast<const float4*>(dctx_s + 4 * d4);
                p = fmaf(c.x, e.x, p); p = fmaf(c.y, e.y, p); p = fmaf(c.z, e.z, p); p = fmaf(c.w, e.w, p);
            }
        } else {
            for (int d = lane; d < D; d += 32) p = fmaf(dctx_s[d], encb[(size_t)tau * D + d], p);
        }
        p = warp_sum(p);
        if (lane == 0) { ds_s[tau] = p; part += al[tau] * p; }
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    float dot = 0.f;
    for (int w = 0; w < nw; ++w) dot += red[w];
    // ds_tau = alpha_tau (da_tau - sum alpha da)
    for (int tau = tid; tau < Tn; tau += ATT_THREADS) {
        const float ds = tau < len ? al[tau] * (ds_s[tau] - dot) : 0.f;
        if (tau < len) ds_s[tau] = ds;
        ds_out[(size_t)b * Tn + tau] = ds;
    }
    __syncthreads();
    // dy_a = sum_tau ds_tau v_a (1 - th^2): the time axis split over ATT_THREADS / A thread groups
    const float* HFb = HF + (size_t)b * Tp * A;
    const int groups = A <= ATT_THREADS ? ATT_THREADS / A : 1;
    for (int idx = tid; idx < groups * A; idx += ATT_THREADS) {
        const int a = idx % A, gi = idx / A;
        const float ya = y_s[a];
        float dya = 0.f;
#pragma unroll 4
        for (int tau = gi; tau < len; tau += groups) {
            const float th = att_tanh(HFb[(size_t)tau * A + a] + ya);
            dya = fmaf(ds_s[tau], 1.f - th * th, dya);
        }
        acc_s[idx] = dya * v_s[a];
    }
    __syncthreads();
    for (int a = tid; a < A; a += ATT_THREADS) {
        float dya = 0.f;
        for (int gi = 0; gi < groups; ++gi) dya += acc_s[gi * A + a];
        dy[(size_t)b * A + a] = dya;
    }
}

int attn_bwd_step(cudaStream_t st, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
                  const int* enc_len, const float* y, const float* v, const float* alpha, const float* dctx, int lddctx,
                  float* ds_out, float* dy) {
    const int groups = A <= ATT_THREADS ? ATT_THREADS / A : 1;
    size_t smem = sizeof(float) * (2 * A + D + Tn + (size_t)groups * A);
    const bool vec = D % 4 == 0 && ((uintptr_t)enc & 15) == 0 && A % 4 == 0;
    if (smem > 44 * 1024) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        E2E_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (vec) attn_bwd_step_kernel<true><<<B, ATT_THREADS, smem, st>>>(Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, dctx, lddctx, ds_out, dy);
    else attn_bwd_step_kernel<false><<<B, ATT_THREADS, smem, st>>>(Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, dctx, lddctx, ds_out, dy);
    E2E_LAUNCH_CHECK();
    return 0;
}

// zero the logits of finished rows (raw_rnn emits zeros for them)
__global__ void mask_rows_kernel(int U, int B, int V, float* __restrict__ logits, const int* __restrict__ lens) {
    int row = blockIdx.x;
    int t = row / B, b = row % B;
    if (t < lens[b]) return;
    for (int v = threadIdx.x; v < V; v += blockDim.x) logits[(size_t)row * V + v] = 0.f;
}
int mask_rows(cudaStream_t st, int U, int B, int V, float* logits, const int* lens) {
    if (U * B <= 0) return 0;
    mask_rows_kernel<<<U * B, 128, 0, st>>>(U, B, V, logits, lens);
    E2E_LAUNCH_CHECK();
    return 0;
}

// greedy next-token: argmax over V (first maximum, numpy/tf.argmax tie rule)
__global__ void argmax_rows_kernel(int V, const float* __restrict__ x, int ldx, long long* __restrict__ out) {
    int row = blockIdx.x;
    const float* r = x + (size_t)row * ldx;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float f = r[v];
        if (f > best || (f == best && v < bi)) { best = f; bi = v; }
    }
    __shared__ float sb[32];
    __shared__ int si[32];
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (threadIdx.x % 32 == 0) { sb[threadIdx.x / 32] = best; si[threadIdx.x / 32] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; ++w)
            if (sb[w] > best || (sb[w] == best && si[w] < bi)) { best = sb[w]; bi = si[w]; }
        out[row] = bi;
    }
}
int argmax_rows(cudaStream_t st, int rows, int V, const float* x, int ldx, long long* out) {
    if (rows <= 0) return 0;
    argmax_rows_kernel<<<rows, 128, 0, st>>>(V, x, ldx, out);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
