// Small memory-bound helpers: input staging (pad / frame stacking / striding),
// gradient-norm reduction and clipping (tf.clip_by_global_norm,
// seq2seq_model.py:148-151), elementwise utilities.
#include "common.cuh"

namespace e2e {

// Seq2SeqModel.get_batch frame stacking (seq2seq_model.py:164-183) and the
// encoder's initial striding (encoder.py:149-153), fused with the copy into the
// zero-padded [B][Tp][F*stack] staging buffer the layer-1 GEMM reads:
//   out[b][t][k*F + f] = in[b][(t*stride) + k][f]   (0 past the end / past Tp)
__global__ void prepare_input_kernel(int B, int T, int F, int Tp, int stack, int stride,
                                     const float* __restrict__ in, float* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int FS = F * stack;
    size_t total = (size_t)B * Tp * FS;
    if (i >= total) return;
    int c = (int)(i % FS);
    int t = (int)((i / FS) % Tp);
    int b = (int)(i / ((size_t)FS * Tp));
    int k = c / F, f = c % F;
    int ts = t * stride + k;
    int Tout = (T + stride - 1) / stride;
    out[i] = (t < Tout && ts < T) ? in[((size_t)b * T + ts) * F + f] : 0.f;
}

int prepare_input(cudaStream_t st, int B, int T, int F, int Tp, int stack, int stride, const float* in, float* out) {
    size_t total = (size_t)B * Tp * F * stack;
    if (total == 0) return 0;
    prepare_input_kernel<<<cdiv(total, 256), 256, 0, st>>>(B, T, F, Tp, stack, stride, in, out);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- sum of squares (deterministic two-stage) ------------------------------
constexpr int SUMSQ_BLOCKS = 296;

__global__ void sumsq_partial_kernel(size_t n, const float* __restrict__ x, float* __restrict__ partials) {
    __shared__ float red[32];
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = x[i];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) tot += red[w];
        partials[blockIdx.x] = tot;
    }
}
__global__ void sumsq_final_kernel(int nparts, const float* __restrict__ partials, float* __restrict__ out,
                                   float sign, int accumulate) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partials[i];
    s = warp_sum(s);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) tot += red[w];
        out[0] = (accumulate ? out[0] : 0.f) + sign * tot;
    }
}

// out[0] (+)= sign * sum x^2 ; partials: >= 296 floats of scratch
int sumsq(cudaStream_t st, size_t n, const float* x, float* partials, float* out, float sign, int accumulate) {
    sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, st>>>(n, x, partials);
    E2E_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 256, 0, st>>>(SUMSQ_BLOCKS, partials, out, sign, accumulate);
    E2E_LAUNCH_CHECK();
    return 0;
}

// tf.clip_by_global_norm: x *= pre * clip / max(sqrt(sumsq), clip); norm_out[0] = sqrt(sumsq).
// pre: the 1/n of a data-parallel SUM of rank gradients (sumsq is then the caller's sum of squares / n^2), folded in so
// that averaging costs no pass of its own.  err (may be NULL): the persistent kernels' barrier-timeout flag -- a step
// whose recurrence gave up on a barrier has garbage gradients: they are zeroed and the norm reads NaN.
__global__ void clip_scale_kernel(size_t n, float* __restrict__ x, const float* __restrict__ sq, float clip,
                                  float* __restrict__ norm_out, float pre, const int* __restrict__ err) {
    float norm = sqrtf(fmaxf(sq[0], 0.f));
    float scale = pre * clip / fmaxf(norm, clip);
    if (err != nullptr && *err != 0) { scale = 0.f; norm = __int_as_float(0x7fc00000); }
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) norm_out[0] = norm;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] *= scale;
}
int clip_by_norm(cudaStream_t st, size_t n, float* x, const float* sq, float clip, float* norm_out, float pre,
                 const int* err) {
    clip_scale_kernel<<<SUMSQ_BLOCKS, 256, 0, st>>>(n, x, sq, clip, norm_out, pre, err);
    E2E_LAUNCH_CHECK();
    return 0;
}

// x[i] = a * x[i] * (s ? s[0] : 1)
__global__ void scale_kernel(size_t n, float* __restrict__ x, const float* __restrict__ s, float a) {
    float f = a * (s ? s[0] : 1.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] *= f;
}
int scale_inplace(cudaStream_t st, size_t n, float* x, const float* s, float a) {
    if (n == 0) return 0;
    scale_kernel<<<min((size_t)SUMSQ_BLOCKS * 4, (n + 255) / 256), 256, 0, st>>>(n, x, s, a);
    E2E_LAUNCH_CHECK();
    return 0;
}

// y[i] += a * x[i]
__global__ void axpy_kernel(size_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = fmaf(a, x[i], y[i]);
}
int axpy(cudaStream_t st, size_t n, float a, const float* x, float* y) {
    if (n == 0) return 0;
    axpy_kernel<<<min((size_t)SUMSQ_BLOCKS * 4, (n + 255) / 256), 256, 0, st>>>(n, a, x, y);
    E2E_LAUNCH_CHECK();
    return 0;
}

// mean over n of x (loss reduction), scaled
__global__ void mean_kernel(int n, const float* __restrict__ x, float* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = warp_sum(s);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) tot += red[w];
        out[0] = tot / (float)n;
    }
}
int mean_vec(cudaStream_t st, int n, const float* x, float* out) {
    mean_kernel<<<1, 256, 0, st>>>(n, x, out);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- Adam (tf.train.AdamOptimizer defaults, seq2seq_model.py:137,153-155) on the flat buffers ----
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr_t m / (sqrt(v) + eps),  lr_t = lr sqrt(1-b2^t)/(1-b1^t)
// (TF's "epsilon hat" form: eps is added to sqrt(v), the bias corrections are folded into lr_t on the host).
__global__ void adam_kernel(size_t n4, float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                            float4* __restrict__ v, float lr_t, float b1, float b2, float eps) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        mm.x = b1 * mm.x + (1.f - b1) * gg.x; mm.y = b1 * mm.y + (1.f - b1) * gg.y;
        mm.z = b1 * mm.z + (1.f - b1) * gg.z; mm.w = b1 * mm.w + (1.f - b1) * gg.w;
        vv.x = b2 * vv.x + (1.f - b2) * gg.x * gg.x; vv.y = b2 * vv.y + (1.f - b2) * gg.y * gg.y;
        vv.z = b2 * vv.z + (1.f - b2) * gg.z * gg.z; vv.w = b2 * vv.w + (1.f - b2) * gg.w * gg.w;
        pp.x -= lr_t * mm.x / (sqrtf(vv.x) + eps); pp.y -= lr_t * mm.y / (sqrtf(vv.y) + eps);
        pp.z -= lr_t * mm.z / (sqrtf(vv.z) + eps); pp.w -= lr_t * mm.w / (sqrtf(vv.w) + eps);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}
int adam_update(cudaStream_t st, size_t n, float* p, const float* g, float* m, float* v, float lr_t, float b1,
                float b2, float eps) {
    E2E_REQUIRE(n % 4 == 0, "adam: the flat buffers are padded to multiples of 4 floats (got %zu)", n);
    if (n == 0) return 0;
    adam_kernel<<<min((size_t)SUMSQ_BLOCKS * 4, (n / 4 + 255) / 256), 256, 0, st>>>(
        n / 4, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, lr_t, b1, b2, eps);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- output dropout of the recurrent cells (DropoutWrapper(output_keep_prob), encoder.py:50-52, decoder.py:60-63)
// y = x * keep_mask / keep with keep_mask ~ Bernoulli(keep) from the counter-based Philox4x32-10 generator:
// element 4q+j takes word j of philox(counter = (q, offset, 0, 0), key = (seed_lo, seed_hi)); keep iff
// word * 2^-32 < keep.  Stateless, so forward and backward (dy = dout * same mask) regenerate it.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__global__ void dropout_kernel(size_t n, const float* __restrict__ x, float* __restrict__ y, float keep,
                               unsigned long long seed, unsigned offset, size_t first4,
                               const unsigned long long* __restrict__ seed_dev) {
    const float inv = 1.0f / keep;
    if (seed_dev != nullptr) seed = *seed_dev;      // the key of THIS replay of a captured step
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q * 4 < n; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        const size_t qg = q + first4;       // counter = global quad index: a slice of a buffer draws the buffer's mask
        philox4x32_10((uint32_t)qg, offset, (uint32_t)(qg >> 32), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t i = q * 4 + j;
            if (i < n) y[i] = ((float)r[j] * 2.3283064365386963e-10f < keep) ? x[i] * inv : 0.f;
        }
    }
}
int dropout(cudaStream_t st, size_t n, const float* x, float* y, float keep, unsigned long long seed,
            unsigned offset, size_t first, const unsigned long long* seed_dev) {
    E2E_REQUIRE(keep > 0.f && keep <= 1.f, "dropout: keep probability %f out of (0, 1]", keep);
    E2E_REQUIRE(first % 4 == 0, "dropout: the slice must start at a multiple of 4 elements (got %zu)", first);
    if (n == 0) return 0;
    dropout_kernel<<<min((size_t)SUMSQ_BLOCKS * 8, (n / 4 + 256) / 256), 256, 0, st>>>(n, x, y, keep, seed, offset,
                                                                                        first / 4, seed_dev);
    E2E_LAUNCH_CHECK();
    return 0;
}

// ---- scheduled sampling: tf.multinomial(logits, 1) (decoder.py:155-180) with the builder-defined Philox stream ----
// Row r draws u = word0(philox(counter = (first_row + r, offset, 0, 0), key = seed)) * 2^-32 and takes the first
// index whose inclusive cumulative sum of exp(logit - max) (float64, index order) exceeds u * total.
// One warp per row; lane l scans the contiguous chunk [l*ch, (l+1)*ch).
__global__ void sample_rows_kernel(int rows, int V, const float* __restrict__ logits, int ldl, unsigned long long seed,
                                   unsigned offset, unsigned first_row, long long* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (r >= rows) return;
    const float* x = logits + (size_t)r * ldl;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    const int ch = (V + 31) / 32, i0 = lane * ch, i1 = min(V, i0 + ch);
    double loc = 0.0;
    for (int i = i0; i < i1; ++i) loc += exp((double)x[i] - (double)mx);
    // inclusive scan of the lane sums (in lane = index order)
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const double total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t w[4];
    philox4x32_10(first_row + (unsigned)r, offset, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const double target = (double)w[0] * 2.3283064365386963e-10 * total;
    double cum = inc - loc;
    int pick = V;                              // first index with cum > target, V if none in this chunk
    for (int i = i0; i < i1; ++i) {
        cum += exp((double)x[i] - (double)mx);
        if (cum > target) { pick = i; break; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    if (lane == 0) out[r] = pick < V ? pick : V - 1;
}
int sample_rows(cudaStream_t st, int rows, int V, const float* logits, int ldl, unsigned long long seed,
                unsigned offset, unsigned first_row, long long* out) {
    if (rows <= 0) return 0;
    sample_rows_kernel<<<cdiv(rows, 4), 128, 0, st>>>(rows, V, logits, ldl, seed, offset, first_row, out);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
