// Pointwise halves of single decoder-cell steps, for the GENERAL decoder path (decoder.py:49-82: stacked cells via
// MultiRNNCell when num_layers_dec > 1, GRUCell when use_lstm=False).  The matrix halves are e2e_gemm calls; autograd
// composes the steps (ops.py: attn_decoder_stepwise).  The benchmarked single-layer LSTM decoder does not come here:
// it has the persistent kernels of decoder_persist.cu.
//
// TF BasicLSTMCell: z = (i | j | f | o) blocks of H; c' = c sig(f + 1) + sig(i) tanh(j); h' = tanh(c') sig(o).
// TF GRUCell: zg = (r | u); rh = sig(r) h; then zc from [x, rh]; h' = u h + (1 - u) tanh(zc).
#include "common.cuh"

namespace e2e {

namespace {

__global__ void lstm_point_fwd_kernel(int n, int H, const float* __restrict__ z, const float* __restrict__ c_prev,
                                      float* __restrict__ c_new, float* __restrict__ h_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int r = i / H, u = i % H;
    const float* g = z + (size_t)r * 4 * H;
    const float si = sigmoidf_acc(g[u]), tj = tanhf(g[H + u]), sf = sigmoidf_acc(g[2 * H + u] + 1.f),
                so = sigmoidf_acc(g[3 * H + u]);
    const float c = c_prev[i] * sf + si * tj;
    c_new[i] = c;
    h_new[i] = tanhf(c) * so;
}
__global__ void lstm_point_bwd_kernel(int n, int H, const float* __restrict__ z, const float* __restrict__ c_prev,
                                      const float* __restrict__ c_new, const float* __restrict__ dc_new,
                                      const float* __restrict__ dh_new, float* __restrict__ dz,
                                      float* __restrict__ dc_prev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int r = i / H, u = i % H;
    const float* g = z + (size_t)r * 4 * H;
    float* d = dz + (size_t)r * 4 * H;
    const float si = sigmoidf_acc(g[u]), tj = tanhf(g[H + u]), sf = sigmoidf_acc(g[2 * H + u] + 1.f),
                so = sigmoidf_acc(g[3 * H + u]);
    const float tc = tanhf(c_new[i]);
    const float dh = dh_new ? dh_new[i] : 0.f;
    const float dc = (dc_new ? dc_new[i] : 0.f) + dh * so * (1.f - tc * tc);
    d[u] = dc * tj * si * (1.f - si);
    d[H + u] = dc * si * (1.f - tj * tj);
    d[2 * H + u] = dc * c_prev[i] * sf * (1.f - sf);
    d[3 * H + u] = dh * tc * so * (1.f - so);
    dc_prev[i] = dc * sf;
}
__global__ void gru_gate_fwd_kernel(int n, int H, const float* __restrict__ zg, const float* __restrict__ h_prev,
                                    float* __restrict__ rh, float* __restrict__ u_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int r = i / H, k = i % H;
    rh[i] = sigmoidf_acc(zg[(size_t)r * 2 * H + k]) * h_prev[i];
    u_out[i] = sigmoidf_acc(zg[(size_t)r * 2 * H + H + k]);
}
__global__ void gru_gate_bwd_kernel(int n, int H, const float* __restrict__ zg, const float* __restrict__ h_prev,
                                    const float* __restrict__ drh, const float* __restrict__ du,
                                    float* __restrict__ dzg, float* __restrict__ dh_prev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int r = i / H, k = i % H;
    const float rv = sigmoidf_acc(zg[(size_t)r * 2 * H + k]), uv = sigmoidf_acc(zg[(size_t)r * 2 * H + H + k]);
    const float a = drh ? drh[i] : 0.f, b = du ? du[i] : 0.f;
    dzg[(size_t)r * 2 * H + k] = a * h_prev[i] * rv * (1.f - rv);
    dzg[(size_t)r * 2 * H + H + k] = b * uv * (1.f - uv);
    dh_prev[i] = a * rv;
}
__global__ void gru_out_fwd_kernel(int n, int H, const float* __restrict__ zc, const float* __restrict__ u,
                                   const float* __restrict__ h_prev, float* __restrict__ h_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    h_new[i] = u[i] * h_prev[i] + (1.f - u[i]) * tanhf(zc[i]);
}
__global__ void gru_out_bwd_kernel(int n, int H, const float* __restrict__ zc, const float* __restrict__ u,
                                   const float* __restrict__ h_prev, const float* __restrict__ dh_new,
                                   float* __restrict__ dzc, float* __restrict__ du, float* __restrict__ dh_prev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const float c = tanhf(zc[i]), d = dh_new[i];
    dzc[i] = d * (1.f - u[i]) * (1.f - c * c);
    du[i] = d * (h_prev[i] - c);
    dh_prev[i] = d * u[i];
}

}  // namespace

#define CELL_LAUNCH(kernel, ...)                                          \
    do {                                                                  \
        if (n <= 0 || H <= 0) return 0;                                   \
        kernel<<<cdiv((long long)n * H, 256), 256, 0, st>>>(__VA_ARGS__); \
        E2E_LAUNCH_CHECK();                                               \
        return 0;                                                         \
    } while (0)

int lstm_point_fwd(cudaStream_t st, int n, int H, const float* z, const float* c_prev, float* c_new, float* h_new) {
    CELL_LAUNCH(lstm_point_fwd_kernel, n, H, z, c_prev, c_new, h_new);
}
int lstm_point_bwd(cudaStream_t st, int n, int H, const float* z, const float* c_prev, const float* c_new,
                   const float* dc_new, const float* dh_new, float* dz, float* dc_prev) {
    CELL_LAUNCH(lstm_point_bwd_kernel, n, H, z, c_prev, c_new, dc_new, dh_new, dz, dc_prev);
}
int gru_gate_fwd(cudaStream_t st, int n, int H, const float* zg, const float* h_prev, float* rh, float* u) {
    CELL_LAUNCH(gru_gate_fwd_kernel, n, H, zg, h_prev, rh, u);
}
int gru_gate_bwd(cudaStream_t st, int n, int H, const float* zg, const float* h_prev, const float* drh,
                 const float* du, float* dzg, float* dh_prev) {
    CELL_LAUNCH(gru_gate_bwd_kernel, n, H, zg, h_prev, drh, du, dzg, dh_prev);
}
int gru_out_fwd(cudaStream_t st, int n, int H, const float* zc, const float* u, const float* h_prev, float* h_new) {
    CELL_LAUNCH(gru_out_fwd_kernel, n, H, zc, u, h_prev, h_new);
}
int gru_out_bwd(cudaStream_t st, int n, int H, const float* zc, const float* u, const float* h_prev,
                const float* dh_new, float* dzc, float* du, float* dh_prev) {
    CELL_LAUNCH(gru_out_bwd_kernel, n, H, zc, u, h_prev, dh_new, dzc, du, dh_prev);
}

}  // namespace e2e
