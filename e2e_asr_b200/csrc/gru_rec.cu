// GRU recurrence of the encoder (reference encoder.py:48: tf.nn.rnn_cell.GRUCell when use_lstm=False, run by
// (bidirectional_)dynamic_rnn, encoder.py:77-89).
//
//   [r, u] = sigmoid(Gg[t] + h . Wg_h)        Gg = x . gates/kernel[:I] + gates/bias      (batched GEMM, before)
//   c      = tanh(Gc[t] + (r*h) . Wc_h)       Gc = x . candidate/kernel[:I] + candidate/bias
//   h'     = u*h + (1-u)*c ; rows with t >= len emit zeros and keep their state.
//
// Utterances are independent, so a CTA owns R = 4 batch rows of one direction and walks the time axis alone: no
// inter-CTA synchronisation at all.  The recurrent kernels (H x 2H and H x H fp32) are streamed from L2 every step,
// each element feeding the CTA's 4 rows; h, r*h and the gate vector live in shared memory.  This is the functional
// GRU path (a few microseconds per step, L2-bandwidth bound); the LSTM path of the benchmarked configurations has
// the cluster-resident kernels of lstm_rec_ws.cu.
//
// The buffers are reused like the LSTM's gate buffer: Gg / Gc hold the x-projections on entry to the forward pass,
// the activations (r | u) and c after it, and d(gate pre-activations) / d(candidate pre-activation) after the
// backward pass.
#include "common.cuh"

namespace e2e {

namespace {

constexpr int GR = 4;          // batch rows per CTA
constexpr int GTH = 256;       // threads per CTA

struct GruParams {
    float* Gg;                 // [B*Tp, nd*2H]
    float* Gc;                 // [B*Tp, nd*H]
    float* out;                // [B*Tp, nd*H]   h_t (zero for t >= len)
    float* RH;                 // [B*Tp, nd*H]   r_t * h_{t-1} (fwd: written; zero for t >= len)
    const float* dout;         // [B*Tp, nd*H]   (bwd)
    const float* Wg;           // fwd: Wg_h [nd][H][2H];  bwd: Wg_h^T [nd][2H][H]
    const float* Wc;           // fwd: Wc_h [nd][H][H];   bwd: Wc_h^T [nd][H][H]
    const int* lens;
    int B, T, Tp, H, nd;
};

__global__ void __launch_bounds__(GTH) gru_fwd_kernel(GruParams p) {
    extern __shared__ float sm[];
    const int H = p.H, d = blockIdx.y, tid = threadIdx.x;
    float* h_s = sm;                       // [GR][H]
    float* rh_s = h_s + GR * H;            // [GR][H]
    float* u_s = rh_s + GR * H;            // [GR][H]
    __shared__ int len_s[GR];
    const int b0 = blockIdx.x * GR;
    if (tid < GR) len_s[tid] = b0 + tid < p.B ? p.lens[b0 + tid] : 0;
    for (int i = tid; i < GR * H; i += GTH) h_s[i] = 0.f;
    __syncthreads();
    const float* Wg = p.Wg + (size_t)d * H * 2 * H;
    const float* Wc = p.Wc + (size_t)d * H * H;
    const int ldg = p.nd * 2 * H, ldc = p.nd * H;
    for (int s = 0; s < p.T; ++s) {
        const int t = d == 0 ? s : p.T - 1 - s;
        bool any = false;
#pragma unroll
        for (int r = 0; r < GR; ++r) any |= t < len_s[r];
        if (!any) continue;                                  // uniform across the CTA
        // ---- gates: columns j of [r | u]
        for (int j = tid; j < 2 * H; j += GTH) {
            float acc[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r)
                acc[r] = t < len_s[r] ? p.Gg[((size_t)(b0 + r) * p.Tp + t) * ldg + d * 2 * H + j] : 0.f;
#pragma unroll 8
            for (int k = 0; k < H; ++k) {
                const float w = __ldg(Wg + (size_t)k * 2 * H + j);
#pragma unroll
                for (int r = 0; r < GR; ++r) acc[r] = fmaf(h_s[r * H + k], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                if (t >= len_s[r]) continue;
                const float g = sigmoidf_acc(acc[r]);
                p.Gg[((size_t)(b0 + r) * p.Tp + t) * ldg + d * 2 * H + j] = g;
                if (j < H) rh_s[r * H + j] = g * h_s[r * H + j];
                else u_s[r * H + j - H] = g;
            }
        }
        __syncthreads();
        // ---- candidate and state update: columns j of c
        float hn[GR][2];                                      // up to 2 columns per thread (H <= 512)
        int nj = 0;
        for (int j = tid; j < H; j += GTH, ++nj) {
            float acc[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r)
                acc[r] = t < len_s[r] ? p.Gc[((size_t)(b0 + r) * p.Tp + t) * ldc + d * H + j] : 0.f;
#pragma unroll 8
            for (int k = 0; k < H; ++k) {
                const float w = __ldg(Wc + (size_t)k * H + j);
#pragma unroll
                for (int r = 0; r < GR; ++r) acc[r] = fmaf(rh_s[r * H + k], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                hn[r][nj] = h_s[r * H + j];
                if (t >= len_s[r]) continue;
                const size_t row = (size_t)(b0 + r) * p.Tp + t;
                const float c = tanhf(acc[r]), u = u_s[r * H + j];
                const float v = u * h_s[r * H + j] + (1.f - u) * c;
                p.Gc[row * ldc + d * H + j] = c;
                p.RH[row * ldc + d * H + j] = rh_s[r * H + j];
                p.out[row * ldc + d * H + j] = v;
                hn[r][nj] = v;
            }
        }
        __syncthreads();                                      // every read of h_s / rh_s of this step is done
        nj = 0;
        for (int j = tid; j < H; j += GTH, ++nj)
#pragma unroll
            for (int r = 0; r < GR; ++r) h_s[r * H + j] = hn[r][nj];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(GTH) gru_bwd_kernel(GruParams p) {
    extern __shared__ float sm[];
    const int H = p.H, d = blockIdx.y, tid = threadIdx.x;
    float* dh_s = sm;                      // [GR][H]   d loss / d h carried to the previous step
    float* dc_s = dh_s + GR * H;           // [GR][H]   d candidate pre-activation
    float* dg_s = dc_s + GR * H;           // [GR][2H]  d gate pre-activations (r | u)
    __shared__ int len_s[GR];
    const int b0 = blockIdx.x * GR;
    if (tid < GR) len_s[tid] = b0 + tid < p.B ? p.lens[b0 + tid] : 0;
    for (int i = tid; i < GR * H; i += GTH) dh_s[i] = 0.f;
    __syncthreads();
    const float* WgT = p.Wg + (size_t)d * 2 * H * H;     // [2H][H]
    const float* WcT = p.Wc + (size_t)d * H * H;         // [H][H]
    const int ldg = p.nd * 2 * H, ldc = p.nd * H;
    // frames past an utterance's length carry no gradient: their rows still hold the forward x-projections
    for (int r = 0; r < GR; ++r) {
        if (b0 + r >= p.B) break;
        for (int t = len_s[r]; t < p.Tp; ++t) {
            const size_t row = (size_t)(b0 + r) * p.Tp + t;
            for (int j = tid; j < 2 * H; j += GTH) p.Gg[row * ldg + d * 2 * H + j] = 0.f;
            for (int j = tid; j < H; j += GTH) p.Gc[row * ldc + d * H + j] = 0.f;
        }
    }
    // the forward pass walked s = 0..T-1 (t = s, or T-1-s for the bw direction): walk it backwards
    for (int s = p.T - 1; s >= 0; --s) {
        const int t = d == 0 ? s : p.T - 1 - s;
        const int tp = d == 0 ? t - 1 : t + 1;                // where h_{prev} of this step was emitted
        bool any = false;
#pragma unroll
        for (int r = 0; r < GR; ++r) any |= t < len_s[r];
        if (!any) continue;
        float keep[GR][2], rr[GR][2], hp[GR][2];              // dh_t * u, r, h_prev for this thread's columns
        int nj = 0;
        // ---- A: pointwise through h' = u h + (1-u) c
        for (int k = tid; k < H; k += GTH, ++nj) {
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                keep[r][nj] = 0.f; rr[r][nj] = 0.f; hp[r][nj] = 0.f;
                dc_s[r * H + k] = 0.f;
                dg_s[r * 2 * H + H + k] = 0.f;
                if (t >= len_s[r]) continue;
                const size_t row = (size_t)(b0 + r) * p.Tp + t;
                const float rv = p.Gg[row * ldg + d * 2 * H + k], uv = p.Gg[row * ldg + d * 2 * H + H + k];
                const float cv = p.Gc[row * ldc + d * H + k];
                const bool has_prev = d == 0 ? t > 0 : t + 1 < len_s[r];
                const float hv = has_prev ? p.out[((size_t)(b0 + r) * p.Tp + tp) * ldc + d * H + k] : 0.f;
                const float dh = p.dout[row * ldc + d * H + k] + dh_s[r * H + k];
                dc_s[r * H + k] = dh * (1.f - uv) * (1.f - cv * cv);
                dg_s[r * 2 * H + H + k] = dh * (hv - cv) * uv * (1.f - uv);
                keep[r][nj] = dh * uv; rr[r][nj] = rv; hp[r][nj] = hv;
            }
        }
        __syncthreads();
        // ---- B: d(r*h) = dcand . Wc_h^T ; gate r
        nj = 0;
        for (int k = tid; k < H; k += GTH, ++nj) {
            float acc[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r) acc[r] = 0.f;
#pragma unroll 8
            for (int j = 0; j < H; ++j) {
                const float w = __ldg(WcT + (size_t)j * H + k);
#pragma unroll
                for (int r = 0; r < GR; ++r) acc[r] = fmaf(dc_s[r * H + j], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                dg_s[r * 2 * H + k] = t < len_s[r] ? acc[r] * hp[r][nj] * rr[r][nj] * (1.f - rr[r][nj]) : 0.f;
                keep[r][nj] += acc[r] * rr[r][nj];
            }
        }
        __syncthreads();
        // ---- C: dh_prev += dgate . Wg_h^T ; store the pre-activation gradients
        nj = 0;
        for (int k = tid; k < H; k += GTH, ++nj) {
            float acc[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r) acc[r] = keep[r][nj];
#pragma unroll 8
            for (int j = 0; j < 2 * H; ++j) {
                const float w = __ldg(WgT + (size_t)j * H + k);
#pragma unroll
                for (int r = 0; r < GR; ++r) acc[r] = fmaf(dg_s[r * 2 * H + j], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                if (t >= len_s[r]) continue;
                const size_t row = (size_t)(b0 + r) * p.Tp + t;
                p.Gg[row * ldg + d * 2 * H + k] = dg_s[r * 2 * H + k];
                p.Gg[row * ldg + d * 2 * H + H + k] = dg_s[r * 2 * H + H + k];
                p.Gc[row * ldc + d * H + k] = dc_s[r * H + k];
                keep[r][nj] = acc[r];
            }
        }
        __syncthreads();                                      // all reads of dh_s / dg_s / dc_s of this step are done
        nj = 0;
        for (int k = tid; k < H; k += GTH, ++nj)
#pragma unroll
            for (int r = 0; r < GR; ++r)
                if (t < len_s[r]) dh_s[r * H + k] = keep[r][nj];
        __syncthreads();
    }
}

}  // namespace

int gru_rec(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int nd, float* Gg, float* Gc, float* out,
            float* RH, const float* dout, const float* Wg, const float* Wc, const int* lens) {
    E2E_REQUIRE(H >= 1 && H <= 2 * GTH, "gru_rec: hidden size %d not in [1, %d]", H, 2 * GTH);
    E2E_REQUIRE(nd == 1 || nd == 2, "gru_rec: %d directions", nd);
    E2E_REQUIRE(T <= Tp, "gru_rec: T=%d exceeds the padded length %d", T, Tp);
    if (B <= 0 || T <= 0) return 0;
    GruParams p;
    p.Gg = Gg; p.Gc = Gc; p.out = out; p.RH = RH; p.dout = dout; p.Wg = Wg; p.Wc = Wc; p.lens = lens;
    p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.nd = nd;
    const size_t smem = sizeof(float) * (bwd ? 4 : 3) * GR * H;
    dim3 grid(cdiv(B, GR), nd);
    if (bwd) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gru_bwd_kernel<<<grid, GTH, smem, st>>>(p);
    } else {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gru_fwd_kernel<<<grid, GTH, smem, st>>>(p);
    }
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
