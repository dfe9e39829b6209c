// extern "C" surface of libe2e_asr_b200.so (declared in include/e2e_asr_b200.h)
// and the host-side sequencing of the decoder loop.
#include <stdarg.h>
#include <string.h>

#include "../../include/e2e_asr_b200.h"
#include "common.cuh"

namespace e2e {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// implemented in the other translation units
int gemm_simt(cudaStream_t, int, int, int, int, int, const float*, int, const float*, int, float*, int,
              const float*, const float*, int, int);
int gemm_tc(cudaStream_t, int mode, int, int, int, int, int, const float*, int, const float*, int, float*, int,
            const float*, const float*, int, int, bool*, const float*, const float*, size_t, size_t, const float*);
int split_rows_f16(cudaStream_t, size_t, int, const float*, float*, float*);
int split_lo(cudaStream_t, int, size_t, const float*, float*);
void set_workspace(void*, size_t);
bool set_stream_workspace(cudaStream_t, void*, size_t);
extern int g_rec_mode;
extern long long* g_rec_dbg;
extern int g_dec_p2p;
void set_tc_debug(float*, long long);
int colsum(cudaStream_t, int, int, const float*, int, float*, int);
int lstm_rec(cudaStream_t, bool, int, int, int, int, int, long long, long long, float*, float*, float*,
             const float*, const float*, const int*, void*, size_t, int*);
size_t lstm_rec_h512_workspace(int B, int ndir, bool bwd);
int lstm_rec_fwd_carry(cudaStream_t, int, int, int, int, long long, long long, float*, float*, float*, const float*,
                       const int*, void*, size_t);
int lstm_pack_weights(cudaStream_t, int, int, const float*, const float*, float*, int, int, float*, float*);
int lstm_unpack_grads(cudaStream_t, int, int, float*, float*, const float*, int, int, const float*, const float*, int);
int prepare_input(cudaStream_t, int, int, int, int, int, int, const float*, float*);
int embed_gather(cudaStream_t, int, int, const float*, const long long*, float*);
int embed_scatter_add(cudaStream_t, int, int, float*, const long long*, const float*, int);
int dec_pointwise_fwd(cudaStream_t, int, int, int, const float*, const float*, const float*, int, const int*,
                      float*, float*, int, float*, float*, int, float*);
int dec_pointwise_bwd(cudaStream_t, int, int, int, const float*, const float*, int, const float*, const float*,
                      int, float*, const float*, int, const int*, float*);
int attn_fwd(cudaStream_t, int, int, int, int, int, const float*, const float*, const int*, const float*,
             const float*, float*, float*, int);
int attn_bwd(cudaStream_t, int, int, int, int, int, const float*, const float*, const int*, const float*,
             const float*, const float*, const float*, int, float*, float*, float*, float*);
int mask_rows(cudaStream_t, int, int, int, float*, const int*);
int argmax_rows(cudaStream_t, int, int, const float*, int, long long*);
int row_lse(cudaStream_t, int, int, const float*, int, float*);
int ce_fwd(cudaStream_t, int, int, int, const float*, const long long*, const int*, float*, float*, float*);
int ce_bwd(cudaStream_t, int, int, int, const float*, const long long*, const int*, const float*, const float*,
           float*);
int ctc_fwd_grad(cudaStream_t, int, int, int, long long, long long, const float*, const float*, const int*,
                 const long long*, int, const int*, int, float*, float*, float*, float);
size_t ctc_workspace_floats(int, int, int);
int beam_merge(cudaStream_t, const e2e_beam_merge_args*);
int beam_gather(cudaStream_t, int, const int*, const e2e_beam_gather_args*);
int sumsq(cudaStream_t, size_t, const float*, float*, float*, float, int);
int clip_by_norm(cudaStream_t, size_t, float*, const float*, float, float*, float, const int*);
int scale_inplace(cudaStream_t, size_t, float*, const float*, float);
int mean_vec(cudaStream_t, int, const float*, float*);
int axpy(cudaStream_t, size_t, float, const float*, float*);
int adam_update(cudaStream_t, size_t, float*, const float*, float*, float*, float, float, float, float);
int dropout(cudaStream_t, size_t, const float*, float*, float, unsigned long long, unsigned, size_t,
            const unsigned long long*);
int sample_rows(cudaStream_t, int, int, const float*, int, unsigned long long, unsigned, unsigned, long long*);
int gru_rec(cudaStream_t, bool, int, int, int, int, int, float*, float*, float*, float*, const float*, const float*,
            const float*, const int*);
int lstm_point_fwd(cudaStream_t, int, int, const float*, const float*, float*, float*);
int lstm_point_bwd(cudaStream_t, int, int, const float*, const float*, const float*, const float*, const float*, float*,
                   float*);
int gru_gate_fwd(cudaStream_t, int, int, const float*, const float*, float*, float*);
int gru_gate_bwd(cudaStream_t, int, int, const float*, const float*, const float*, const float*, float*, float*);
int gru_out_fwd(cudaStream_t, int, int, const float*, const float*, const float*, float*);
int gru_out_bwd(cudaStream_t, int, int, const float*, const float*, const float*, const float*, float*, float*, float*);
int attn_bwd(cudaStream_t, int, int, int, int, int, const float*, const float*, const int*, const float*, const float*,
             const float*, const float*, int, float*, float*, float*, float*);
int dec_persist(cudaStream_t, bool, const e2e_dec_persist_args*, float*, float*, float*);
int dec_persist_fits(const e2e_dec_persist_args*);
int dec_deferred_attn_grads(cudaStream_t, const e2e_dec_persist_args&, float*, float*, float*);
int attn_bwd_step(cudaStream_t, int, int, int, int, int, const float*, const float*, const int*, const float*, const float*,
                  const float*, const float*, int, float*, float*);
int gemm_f64(cudaStream_t, int, int, int, const double*, int, const float*, int, double*, int, const float*);
int gemm_f64d(cudaStream_t, int, int, int, const double*, int, const double*, int, double*, int, const float*);
int gemm_f64d_cat(cudaStream_t, int, int, int, int, const double*, int, const double*, int, const double*, int, double*,
                  int, const float*, const double*, int, const long long*, const double*, double*, double*, int);
int exp2x_f64(cudaStream_t, size_t, const float*, double*);
int attn_beam_group_e_f64(cudaStream_t, int, int, int, int, int, const double*, const float*, const int*, const int*,
                          const double*, const float*, double*, int);
extern int g_f64_mma;
int lstm_step_f64(cudaStream_t, int, int, const double*, const double*, double*, double*, int);
int attn_beam_group_f64(cudaStream_t, int, int, int, int, int, const float*, const float*, const int*, const int*,
                        const double*, const float*, double*, int);
int attn_beam_f64(cudaStream_t, int, int, int, int, const float*, const float*, const int*, const int*, const double*,
                  const float*, double*, int);
int logsoftmax_topk_f64(cudaStream_t, int, int, const double*, const double*, double, const int*, int, int*, double*,
                        double*);
int embed_gather_f64(cudaStream_t, int, int, const float*, const long long*, double*, int);

static int gemm_any(cudaStream_t st, int mode, int tA, int tB, int M, int N, int K, const float* A, int lda,
                    const float* B, int ldb, float* C, int ldc, const float* bias, const float* Z, int ldz,
                    int accumulate, const float* A_lo = nullptr, const float* B_lo = nullptr, size_t a_plane = 0,
                    size_t b_plane = 0, const float* row_scale = nullptr) {
    if (mode != 0) {
        bool handled = false;
        int rc = gemm_tc(st, mode, tA, tB, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate, &handled, A_lo,
                         B_lo, a_plane, b_plane, row_scale);
        if (rc) return rc;
        if (handled) return 0;
    }
    return gemm_simt(st, tA, tB, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate);
}

}  // namespace e2e

using namespace e2e;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int e2e_version(void) { return 1; }
/* 0 = not capturing, 1 = capture active, 2 = capture invalidated, -1 = query failed */
int e2e_capture_status(void* stream) {
    cudaStreamCaptureStatus s;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &s) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (int)s;
}
unsigned long long e2e_launch_count(int reset) {
    unsigned long long n = e2e::g_launches;
    if (reset) e2e::g_launches = 0;
    return n;
}
const char* e2e_last_error(void) { return e2e::g_err; }
int e2e_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("no CUDA device: e2e_asr_b200 has no CPU fallback");
        return -1;
    }
    return sm_count();
}

int e2e_gemm(void* stream, int mode, int transA, int transB, int M, int N, int K, const float* A, int lda,
             const float* B, int ldb, float* C, int ldc, const float* bias, const float* Z, int ldz,
             int accumulate) {
    return gemm_any(ST(stream), mode, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate);
}
int e2e_gemm_lo(void* stream, int mode, int transA, int transB, int M, int N, int K, const float* A, const float* A_lo,
                int lda, const float* B, const float* B_lo, int ldb, float* C, int ldc, const float* bias,
                const float* Z, int ldz, int accumulate, size_t a_plane, size_t b_plane, const float* a_row_scale) {
    return gemm_any(ST(stream), mode, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, Z, ldz, accumulate, A_lo,
                    B_lo, a_plane, b_plane, a_row_scale);
}
int e2e_split_rows_f16(void* stream, size_t rows, int cols, const float* x, float* planes, float* row_inv) {
    return split_rows_f16(ST(stream), rows, cols, x, planes, row_inv);
}
int e2e_split_lo(void* stream, int mode, size_t n, const float* x, float* lo) {
    return split_lo(ST(stream), mode, n, x, lo);
}
int e2e_gru_rec_fwd(void* stream, int B, int T, int Tp, int H, int ndir, float* Gg, float* Gc, float* out, float* RH,
                    const float* Wg_h, const float* Wc_h, const int* lens) {
    return gru_rec(ST(stream), false, B, T, Tp, H, ndir, Gg, Gc, out, RH, nullptr, Wg_h, Wc_h, lens);
}
int e2e_gru_rec_bwd(void* stream, int B, int T, int Tp, int H, int ndir, float* Gg, float* Gc, const float* out,
                    const float* dout, const float* Wg_hT, const float* Wc_hT, const int* lens) {
    return gru_rec(ST(stream), true, B, T, Tp, H, ndir, Gg, Gc, const_cast<float*>(out), nullptr, dout, Wg_hT, Wc_hT,
                   lens);
}
int e2e_lstm_point_fwd(void* stream, int n, int H, const float* z, const float* c_prev, float* c_new, float* h_new) {
    return lstm_point_fwd(ST(stream), n, H, z, c_prev, c_new, h_new);
}
int e2e_lstm_point_bwd(void* stream, int n, int H, const float* z, const float* c_prev, const float* c_new,
                       const float* dc_new, const float* dh_new, float* dz, float* dc_prev) {
    return lstm_point_bwd(ST(stream), n, H, z, c_prev, c_new, dc_new, dh_new, dz, dc_prev);
}
int e2e_gru_gate_fwd(void* stream, int n, int H, const float* zg, const float* h_prev, float* rh, float* u) {
    return gru_gate_fwd(ST(stream), n, H, zg, h_prev, rh, u);
}
int e2e_gru_gate_bwd(void* stream, int n, int H, const float* zg, const float* h_prev, const float* drh,
                     const float* du, float* dzg, float* dh_prev) {
    return gru_gate_bwd(ST(stream), n, H, zg, h_prev, drh, du, dzg, dh_prev);
}
int e2e_gru_out_fwd(void* stream, int n, int H, const float* zc, const float* u, const float* h_prev, float* h_new) {
    return gru_out_fwd(ST(stream), n, H, zc, u, h_prev, h_new);
}
int e2e_gru_out_bwd(void* stream, int n, int H, const float* zc, const float* u, const float* h_prev,
                    const float* dh_new, float* dzc, float* du, float* dh_prev) {
    return gru_out_bwd(ST(stream), n, H, zc, u, h_prev, dh_new, dzc, du, dh_prev);
}
int e2e_attn_bwd(void* stream, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
                 const int* enc_len, const float* y, const float* v, const float* alpha, const float* dctx,
                 int lddctx, float* dHF, float* denc, float* dy, float* dv_part) {
    return attn_bwd(ST(stream), B, Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, dctx, lddctx, dHF, denc, dy, dv_part);
}
int e2e_set_tc_debug(float* dbg, long long min_work) {
    set_tc_debug(dbg, min_work);
    return 0;
}
int e2e_set_rec_mode(int mode) {
    g_rec_mode = mode;
    return 0;
}
int e2e_set_dec_sync(int p2p) {
    g_dec_p2p = p2p ? 1 : 0;
    return 0;
}
int e2e_set_rec_debug(long long* dbg) {
    g_rec_dbg = dbg;
    return 0;
}
int e2e_set_stream_workspace(void* stream, void* ptr, size_t bytes) {
    E2E_REQUIRE(set_stream_workspace(ST(stream), ptr, bytes),
                "e2e_set_stream_workspace: the per-stream scratch table is full (a GEMM on this stream would race "
                "with the default scratch)");
    return 0;
}
int e2e_set_workspace(void* ptr, size_t bytes) {
    set_workspace(ptr, bytes);
    return 0;
}
int e2e_colsum(void* stream, int M, int N, const float* X, int ldx, float* out, int accumulate) {
    return colsum(ST(stream), M, N, X, ldx, out, accumulate);
}
int e2e_lstm_pack_weights(void* stream, int I, int H, const float* kernel, const float* bias, float* Wx, int ldwx,
                          int col0, float* Wh, float* bias_packed) {
    return lstm_pack_weights(ST(stream), I, H, kernel, bias, Wx, ldwx, col0, Wh, bias_packed);
}
int e2e_lstm_unpack_grads(void* stream, int I, int H, float* dkernel, float* dbias, const float* dWx, int ldwx,
                          int col0, const float* dWh, const float* dbias_packed, int accumulate) {
    return lstm_unpack_grads(ST(stream), I, H, dkernel, dbias, dWx, ldwx, col0, dWh, dbias_packed, accumulate);
}
int e2e_lstm_rec_fwd(void* stream, int B, int T, int Tp, int H, int ndir, long long sb, long long st, float* G,
                     float* Hout, float* Cst, const float* Wh, const int* lens, void* ctr_ws, size_t ctr_ws_bytes,
                     int* err_flag) {
    return lstm_rec(ST(stream), false, B, T, Tp, H, ndir, sb, st, G, Hout, Cst, Wh, nullptr, lens, ctr_ws,
                    ctr_ws_bytes, err_flag);
}
size_t e2e_lstm_rec_workspace_bytes(int B, int H, int ndir) {
    size_t need = (size_t)4 * ndir * ((B + 3) / 4);
    const size_t tiles = (size_t)4096 * ndir * 16 * ((B + 15) / 16);
    if (tiles > need) need = tiles;
    if (H == 512) {
        const size_t wide = lstm_rec_h512_workspace(B, ndir, true);
        if (wide > need) need = wide;
    }
    return need;
}
int e2e_lstm_rec_fwd_carry(void* stream, int B, int T, int Tp, int H, long long sb, long long st, float* G, float* Hout,
                           float* Cst, const float* Wh, const int* lens, void* ctr_ws, size_t ctr_ws_bytes,
                           int* err_flag) {
    (void)err_flag;
    return lstm_rec_fwd_carry(ST(stream), B, T, Tp, H, sb, st, G, Hout, Cst, Wh, lens, ctr_ws, ctr_ws_bytes);
}
int e2e_lstm_rec_bwd(void* stream, int B, int T, int Tp, int H, int ndir, long long sb, long long st, float* G,
                     const float* Cst, const float* Wh, const float* dOut, const int* lens, void* ctr_ws,
                     size_t ctr_ws_bytes, int* err_flag) {
    return lstm_rec(ST(stream), true, B, T, Tp, H, ndir, sb, st, G, nullptr, const_cast<float*>(Cst), Wh, dOut,
                    lens, ctr_ws, ctr_ws_bytes, err_flag);
}
int e2e_prepare_input(void* stream, int B, int T, int F, int Tp, int stack, int stride, const float* in,
                      float* out) {
    return prepare_input(ST(stream), B, T, F, Tp, stack, stride, in, out);
}
int e2e_embed_gather(void* stream, int n, int E, const float* emb, const long long* ids, float* out) {
    return embed_gather(ST(stream), n, E, emb, ids, out);
}
int e2e_embed_scatter_add(void* stream, int n, int E, float* demb, const long long* ids, const float* dout,
                          int ldd) {
    return embed_scatter_add(ST(stream), n, E, demb, ids, dout, ldd);
}

int e2e_decoder_loop_fwd(void* stream, const e2e_dec_loop_fwd_args* a) {
    cudaStream_t st = ST(stream);
    const int B = a->B, U = a->U, E = a->E, Hd = a->Hd, A = a->A, D = a->D;
    const int XH = E + Hd, CAT = Hd + D, mode = a->gemm_mode;
    for (int t = 0; t < U; ++t) {
        float* xh_t = a->xh + (size_t)t * B * XH;
        float* cat_t = a->cat + (size_t)t * B * CAT;
        const float* cat_p = a->cat + (size_t)(t - 1) * B * CAT;
        // xin_t = pre_t + ctx_{t-1} . in_k[Hd:]              (attn_decoder.py:157-158)
        int rc = gemm_any(st, mode, 0, 0, B, E, t == 0 ? 0 : D, t == 0 ? a->enc : cat_p + Hd, CAT,
                          a->in_k + (size_t)Hd * E, E, xh_t, XH, nullptr, a->pre + (size_t)t * B * E, E, 0);
        if (rc) return rc;
        // gates = [xin, h] . dec_k + dec_b                   (raw_rnn body -> BasicLSTMCell)
        rc = gemm_any(st, mode, 0, 0, B, 4 * Hd, XH, xh_t, XH, a->dec_k, 4 * Hd, a->gates_tmp, 4 * Hd, a->dec_b,
                      nullptr, 0, 0);
        if (rc) return rc;
        bool last = t + 1 == U;
        rc = dec_pointwise_fwd(st, B, Hd, t, a->gates_tmp, a->cprev + (size_t)t * B * Hd, xh_t + E, XH, a->lens,
                               a->acts + (size_t)t * B * 4 * Hd, cat_t, CAT,
                               last ? nullptr : a->cprev + (size_t)(t + 1) * B * Hd,
                               last ? nullptr : a->xh + (size_t)(t + 1) * B * XH + E, XH, nullptr);
        if (rc) return rc;
        // y = c_new . q_k + q_b (query is the cell state, attn_decoder.py:114)
        float* y_t = a->y + (size_t)t * B * A;
        rc = gemm_any(st, mode, 0, 0, B, A, Hd, cat_t, CAT, a->q_k, A, y_t, A, a->q_b, nullptr, 0, 0);
        if (rc) return rc;
        rc = attn_fwd(st, B, a->Tn, a->Tp, A, D, a->HF, a->enc, a->enc_len, y_t, a->attn_v,
                      a->alpha + (size_t)t * B * a->Tn, cat_t + Hd, CAT);
        if (rc) return rc;
    }
    return 0;
}

int e2e_decoder_loop_bwd(void* stream, const e2e_dec_loop_bwd_args* g) {
    cudaStream_t st = ST(stream);
    const e2e_dec_loop_fwd_args* a = &g->f;
    const int B = a->B, U = a->U, E = a->E, Hd = a->Hd, A = a->A, D = a->D;
    const int XH = E + Hd, CAT = Hd + D, mode = a->gemm_mode;
    for (int t = U - 1; t >= 0; --t) {
        float* dcat_t = g->dcat + (size_t)t * B * CAT;
        const float* cat_t = a->cat + (size_t)t * B * CAT;
        float* dy_t = g->dy + (size_t)t * B * A;
        int rc = attn_bwd_step(st, B, a->Tn, a->Tp, A, D, a->HF, a->enc, a->enc_len, a->y + (size_t)t * B * A,
                               a->attn_v, a->alpha + (size_t)t * B * a->Tn, dcat_t + Hd, CAT,
                               g->ds + (size_t)t * B * a->Tn, dy_t);
        if (rc) return rc;
        // d c_new += dy . q_k^T
        rc = gemm_any(st, mode, 0, 1, B, Hd, A, dy_t, A, a->q_k, A, dcat_t, CAT, nullptr, nullptr, 0, 1);
        if (rc) return rc;
        float* dz_t = g->dgates + (size_t)t * B * 4 * Hd;
        bool last = t + 1 == U;
        rc = dec_pointwise_bwd(st, B, Hd, t, a->acts + (size_t)t * B * 4 * Hd, cat_t, CAT,
                               a->cprev + (size_t)t * B * Hd, dcat_t, CAT, g->dc_carry,
                               last ? nullptr : g->dxh + (size_t)(t + 1) * B * XH + E, XH, a->lens, dz_t);
        if (rc) return rc;
        // (dxin | dh_prev) = dz . dec_k^T
        float* dxh_t = g->dxh + (size_t)t * B * XH;
        rc = gemm_any(st, mode, 0, 1, B, XH, 4 * Hd, dz_t, 4 * Hd, a->dec_k, 4 * Hd, dxh_t, XH, nullptr, nullptr, 0,
                      0);
        if (rc) return rc;
        // d ctx_{t-1} += dxin . in_k[Hd:]^T
        if (t > 0) {
            rc = gemm_any(st, mode, 0, 1, B, D, E, dxh_t, XH, a->in_k + (size_t)Hd * E, E,
                          g->dcat + (size_t)(t - 1) * B * CAT + Hd, CAT, nullptr, nullptr, 0, 1);
            if (rc) return rc;
        }
    }
    // sums over the steps: denc += sum_t alpha_t dctx_t (dcat now holds the total d ctx_t), dHF, dv_part
    e2e_dec_persist_args p = {};
    p.B = B; p.U = U; p.Hd = Hd; p.A = A; p.D = D; p.Tn = a->Tn; p.Tp = a->Tp;
    p.attn_v = a->attn_v; p.HF = a->HF; p.enc = a->enc; p.enc_len = a->enc_len; p.lens = a->lens;
    p.y = a->y; p.alpha = a->alpha; p.dcat = g->dcat; p.ds = g->ds;
    return dec_deferred_attn_grads(st, p, g->denc, g->dHF, g->dv_part);
}

int e2e_decoder_persist_fits(const e2e_dec_persist_args* a) { return dec_persist_fits(a); }
int e2e_decoder_persist_fwd(void* stream, const e2e_dec_persist_args* a) {
    return dec_persist(ST(stream), false, a, nullptr, nullptr, nullptr);
}
int e2e_decoder_persist_bwd(void* stream, const e2e_dec_persist_args* a, float* denc, float* dHF, float* dv_part) {
    return dec_persist(ST(stream), true, a, denc, dHF, dv_part);
}
int e2e_attn_fwd(void* stream, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
                 const int* enc_len, const float* y, const float* v, float* alpha, float* ctx, int ldctx) {
    return attn_fwd(ST(stream), B, Tn, Tp, A, D, HF, enc, enc_len, y, v, alpha, ctx, ldctx);
}
int e2e_dec_pointwise_fwd(void* stream, int B, int H, int t, const float* gates_pre, const float* cprev,
                          const float* hprev, int ldh, const int* lens, float* acts, float* cnew_out, int ldc,
                          float* c_next, float* h_next, int ldhn, float* h_new_out) {
    return dec_pointwise_fwd(ST(stream), B, H, t, gates_pre, cprev, hprev, ldh, lens, acts, cnew_out, ldc, c_next,
                             h_next, ldhn, h_new_out);
}
int e2e_mask_rows(void* stream, int U, int B, int V, float* logits, const int* lens) {
    return mask_rows(ST(stream), U, B, V, logits, lens);
}
int e2e_argmax_rows(void* stream, int rows, int V, const float* x, int ldx, long long* out) {
    return argmax_rows(ST(stream), rows, V, x, ldx, out);
}
int e2e_ce_fwd(void* stream, int U, int B, int V, const float* logits, const long long* targets, const int* lens,
               float* lse, float* cost, float* loss) {
    return ce_fwd(ST(stream), U, B, V, logits, targets, lens, lse, cost, loss);
}
int e2e_ce_bwd(void* stream, int U, int B, int V, const float* logits, const long long* targets, const int* lens,
               const float* lse, const float* gscale, float* dlogits) {
    return ce_bwd(ST(stream), U, B, V, logits, targets, lens, lse, gscale, dlogits);
}
int e2e_row_lse(void* stream, int rows, int V, const float* x, int ldx, float* lse) {
    return row_lse(ST(stream), rows, V, x, ldx, lse);
}
int e2e_ctc_fwd_grad(void* stream, int T, int B, int C, long long sb, long long stt, const float* logits,
                     const float* lse_rows, const int* in_lens, const long long* labels, int ldl,
                     const int* label_lens, int max_label_len, float* ws, float* loss_b, float* grad,
                     float out_scale) {
    return ctc_fwd_grad(ST(stream), T, B, C, sb, stt, logits, lse_rows, in_lens, labels, ldl, label_lens,
                        max_label_len, ws, loss_b, grad, out_scale);
}
size_t e2e_ctc_workspace_floats(int T, int B, int max_label_len) { return ctc_workspace_floats(T, B, max_label_len); }
int e2e_beam_merge(void* stream, const e2e_beam_merge_args* a) { return beam_merge(ST(stream), a); }
int e2e_beam_gather(void* stream, int R, const int* parent, const e2e_beam_gather_args* g) {
    return beam_gather(ST(stream), R, parent, g);
}
int e2e_sumsq(void* stream, size_t n, const float* x, float* partials296, float* out, float sign, int accumulate) {
    return sumsq(ST(stream), n, x, partials296, out, sign, accumulate);
}
int e2e_clip_by_norm(void* stream, size_t n, float* x, const float* sq, float clip, float* norm_out, float pre_scale,
                     const int* err_flag) {
    return clip_by_norm(ST(stream), n, x, sq, clip, norm_out, pre_scale, err_flag);
}
int e2e_scale(void* stream, size_t n, float* x, const float* dev_scalar, float a) {
    return scale_inplace(ST(stream), n, x, dev_scalar, a);
}
int e2e_mean(void* stream, int n, const float* x, float* out) { return mean_vec(ST(stream), n, x, out); }
int e2e_axpy(void* stream, size_t n, float a, const float* x, float* y) { return axpy(ST(stream), n, a, x, y); }
int e2e_adam(void* stream, size_t n, float* param, const float* grad, float* m, float* v, float lr_t, float beta1,
             float beta2, float eps) {
    return adam_update(ST(stream), n, param, grad, m, v, lr_t, beta1, beta2, eps);
}
int e2e_dropout(void* stream, size_t n, const float* x, float* y, float keep, unsigned long long seed,
                unsigned offset, size_t first, const unsigned long long* seed_dev) {
    return dropout(ST(stream), n, x, y, keep, seed, offset, first, seed_dev);
}
int e2e_sample_rows(void* stream, int rows, int V, const float* logits, int ldl, unsigned long long seed,
                    unsigned offset, unsigned first_row, long long* out) {
    return sample_rows(ST(stream), rows, V, logits, ldl, seed, offset, first_row, out);
}

int e2e_set_f64_mma(int on) {
    g_f64_mma = on;
    return 0;
}
int e2e_gemm_f64d(void* stream, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                  int ldc, const float* bias) {
    return gemm_f64d(ST(stream), M, N, K, A, lda, B, ldb, C, ldc, bias);
}
int e2e_gemm_f64(void* stream, int M, int N, int K, const double* A, int lda, const float* B, int ldb, double* C,
                 int ldc, const float* bias) {
    return gemm_f64(ST(stream), M, N, K, A, lda, B, ldb, C, ldc, bias);
}
int e2e_gemm_f64d_cat(void* stream, int M, int N, int K1, int K2, const double* A1, int lda1, const double* A2,
                      int lda2, const double* B, int ldb, double* C, int ldc, const float* bias, const double* Z,
                      int ldz, const long long* zrow) {
    return gemm_f64d_cat(ST(stream), M, N, K1, K2, A1, lda1, A2, lda2, B, ldb, C, ldc, bias, Z, ldz, zrow, nullptr,
                         nullptr, nullptr, 0);
}
int e2e_gemm_f64d_lstm(void* stream, int M, int H, int K1, int K2, const double* A1, int lda1, const double* A2,
                       int lda2, const double* B, int ldb, const float* bias, const double* Z, int ldz,
                       const long long* zrow, const double* c_prev, double* c_out, double* h_out, int ldh) {
    if (c_out == nullptr) {
        set_error("e2e_gemm_f64d_lstm: c_out is NULL");
        return 1;
    }
    return gemm_f64d_cat(ST(stream), M, 4 * H, K1, K2, A1, lda1, A2, lda2, B, ldb, nullptr, 0, bias, Z, ldz, zrow,
                         c_prev, c_out, h_out, ldh);
}
int e2e_exp2x_f64(void* stream, size_t n, const float* x, double* out) { return exp2x_f64(ST(stream), n, x, out); }
int e2e_attn_beam_group_e_f64(void* stream, int N, int beam, int A, int D, int Tmax, const double* EHF,
                              const float* enc, const int* row_off, const int* Tlen, const double* y, const float* v,
                              double* ctx, int ldctx) {
    return attn_beam_group_e_f64(ST(stream), N, beam, A, D, Tmax, EHF, enc, row_off, Tlen, y, v, ctx, ldctx);
}
int e2e_lstm_step_f64(void* stream, int n, int H, const double* z, const double* c_prev, double* c_out, double* h_out,
                      int ldh) {
    return lstm_step_f64(ST(stream), n, H, z, c_prev, c_out, h_out, ldh);
}
int e2e_attn_beam_group_f64(void* stream, int N, int beam, int A, int D, int Tmax, const float* HF, const float* enc,
                            const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx) {
    return attn_beam_group_f64(ST(stream), N, beam, A, D, Tmax, HF, enc, row_off, Tlen, y, v, ctx, ldctx);
}
int e2e_attn_beam_f64(void* stream, int n, int A, int D, int Tmax, const float* HF, const float* enc,
                      const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx) {
    return attn_beam_f64(ST(stream), n, A, D, Tmax, HF, enc, row_off, Tlen, y, v, ctx, ldctx);
}
int e2e_logsoftmax_topk_f64(void* stream, int n, int V, const double* logits, const double* lm_logits,
                            double lm_weight, const int* krow, int kmax, int* out_idx, double* out_val,
                            double* scratch) {
    return logsoftmax_topk_f64(ST(stream), n, V, logits, lm_logits, lm_weight, krow, kmax, out_idx, out_val, scratch);
}
int e2e_embed_gather_f64(void* stream, int n, int E, const float* emb, const long long* ids, double* out, int ldo) {
    return embed_gather_f64(ST(stream), n, E, emb, ids, out, ldo);
}

}  // extern "C"
