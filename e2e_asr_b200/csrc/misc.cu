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

// tf.clip_by_global_norm: x *= clip / max(sqrt(sumsq), clip); norm_out[0] = sqrt(sumsq)
__global__ void clip_scale_kernel(size_t n, float* __restrict__ x, const float* __restrict__ sq, float clip,
                                  float* __restrict__ norm_out) {
    float norm = sqrtf(fmaxf(sq[0], 0.f));
    float scale = clip / fmaxf(norm, clip);
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) norm_out[0] = norm;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] *= scale;
}
int clip_by_norm(cudaStream_t st, size_t n, float* x, const float* sq, float clip, float* norm_out) {
    clip_scale_kernel<<<SUMSQ_BLOCKS, 256, 0, st>>>(n, x, sq, clip, norm_out);
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

}  // namespace e2e
