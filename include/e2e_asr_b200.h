/* e2e_asr_b200 -- C ABI of the B200-native kernels behind the reference's Python
 * model API (shtoshni/e2e_asr).
 *
 * The reference has no FFI / plugin registry: its device
 * maths are TensorFlow-1.x library ops called from Python.  Each entry point
 * below therefore names the reference call site (file:line under
 * /root/reference) whose TF op(s) it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch
 *     allocations), fp32 unless stated; `ids`/`targets`/`labels` are int64 as in
 *     the reference's batches (speech_dataset.py:17-24); `lens` arrays are int32.
 *   - `stream` is a cudaStream_t; all work is enqueued on it; no host sync, no
 *     allocation inside (scratch comes in through explicit workspace arguments).
 *   - return value 0 = ok; non-zero = error, message from e2e_last_error().
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef E2E_ASR_B200_H
#define E2E_ASR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int e2e_version(void);
/* CUDA-graph capture state of a stream (debug aid for GraphedStep): 0 none, 1 active, 2 invalidated, -1 error */
int e2e_capture_status(void* stream);
const char* e2e_last_error(void);
int e2e_sm_count(void);
/* number of kernels this library has launched (host counter); reset != 0 clears it */
unsigned long long e2e_launch_count(int reset);

/* Dense contraction  C[M,N] = op(A) op(B) (+bias[N]) (+Z[M,N]) (+C)   row-major.
 * Replaces every `_linear` / matmul / 1x1 conv2d on the path: attn_decoder.py:73
 * (hidden_features), :80 (Attention), :117 (AttnProjection), :122-125
 * (OutputProjection), :151 (SimpleProjection), :158 (InputProjection); the
 * x-part of BasicLSTMCell's [x,h].K (encoder.py:77-81); and their gradients
 * (tf.gradients, seq2seq_model.py:148).
 * mode: 0 = fp32 FFMA (exact fp32), 1 = 3xTF32 tcgen05 (fp32-accurate tensor
 * core), 2 = bf16 tcgen05, 3 = bf16x2 tcgen05 (operands as hi + lo bf16 pairs, three kind::f16 products:
 * ~2^-17 relative error per product at half the tensor-pipe cost of mode 1), 4 = f16x2 tcgen05 (operands as
 * hi = fp16(x) and lo' = fp16((x - hi) 2^11), cross terms in a second TMEM accumulator: ~2^-22 relative error --
 * the accuracy class of mode 1 -- at half its tensor-pipe cost; the A operand is scaled per row by a power of two
 * before the split (fp16 has 5 exponent bits) and the output row scaled back; B must be bounded (weights); a
 * transposed A runs mode 1).  Shapes a tensor-core mode cannot take fall back to mode 0 (still on the GPU). */
int e2e_gemm(void* stream, int mode, int transA, int transB, int M, int N, int K,
             const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             const float* bias, const float* Z, int ldz, int accumulate);
/* e2e_gemm with the caller holding the operand split of an operand (either X_lo may be NULL), made once by
 * e2e_split_lo(mode, n, X, buf) over the whole n-element buffer X lives in and shared by every product X enters
 * (one split of dz serves the dX, dW_x and dW_h products of a layer).  The operand's pre-pass is then skipped:
 *   mode 1: buf[i] = X[i] - tf32_trunc(X[i]) (fp32, X's layout); X_lo points at the element matching X's first
 *           element; needs ld % 4 == 0 and 16-byte aligned pointers (the tensor core ignores the low 13 mantissa
 *           bits of the raw fp32 "big" half, so X itself is the other part); x_plane is ignored;
 *   mode 3: buf holds two bf16 planes of n elements, hi = bf16(X) then lo = bf16(X - hi), each in X's layout;
 *           X_lo points at the hi-plane element matching X's first element and x_plane = n (elements between
 *           the planes); needs ld % 8 == 0, x_plane % 8 == 0 and a 16-byte aligned X_lo;
 *   mode 4: as mode 3 with fp16 planes hi = fp16(X s), lo' = fp16((X s - hi) 2^11): e2e_split_lo(4, ...) makes them
 *           with s = 1 (bounded operands); e2e_split_rows_f16 makes the A planes of a contiguous [rows, cols]
 *           matrix with a power-of-two s per row (gradients) and writes the factors 1/s to row_inv, to be passed
 *           as a_row_scale (NULL = unscaled planes): C row m is multiplied by a_row_scale[m].
 * n must be a multiple of 8 and the buffers 16-byte aligned. */
int e2e_gemm_lo(void* stream, int mode, int transA, int transB, int M, int N, int K, const float* A,
                const float* A_lo, int lda, const float* B, const float* B_lo, int ldb, float* C, int ldc,
                const float* bias, const float* Z, int ldz, int accumulate, size_t a_plane, size_t b_plane,
                const float* a_row_scale);
int e2e_split_lo(void* stream, int mode, size_t n, const float* x, float* lo);
int e2e_split_rows_f16(void* stream, size_t rows, int cols, const float* x, float* planes, float* row_inv);

/* Scratch for the tensor-core modes' operand pre-pass (TF32 big/small split or
 * bf16 copies): a caller-owned device buffer; GEMMs whose operands do not fit run
 * the FFMA kernel.  NOT a stream argument: plain (ptr, bytes). */
int e2e_set_workspace(void* ptr, size_t bytes);
/* scratch used instead for GEMMs enqueued on `stream` (streams that run concurrently must not share one) */
int e2e_set_stream_workspace(void* stream, void* ptr, size_t bytes);
/* test hook: dump buffer for the tensor-core kernel (or NULL) and the minimum M*N*K it takes */
int e2e_set_tc_debug(float* dbg, long long min_work);

/* out[N] (+)= column sums of X[M,N]  (bias gradients) */
int e2e_colsum(void* stream, int M, int N, const float* X, int ldx, float* out, int accumulate);

/* TF BasicLSTMCell variable layout <-> kernel layout.  `kernel` is the TF variable
 * [(I+H),4H] with gate-blocked columns i|j|f|o (basic_lstm.py:17; Appendix B of
 * SURVEY.md); packed: Wx[I][ldwx] at column col0 (gate-interleaved [unit][4]),
 * Wh[H][H][4], bias_packed[col0 ...]. */
int e2e_lstm_pack_weights(void* stream, int I, int H, const float* kernel, const float* bias,
                          float* Wx, int ldwx, int col0, float* Wh, float* bias_packed);
int e2e_lstm_unpack_grads(void* stream, int I, int H, float* dkernel, float* dbias,
                          const float* dWx, int ldwx, int col0, const float* dWh,
                          const float* dbias_packed, int accumulate);

/* Persistent LSTM recurrence over all timesteps of one (bi)directional layer.
 * Replaces tf.nn.bidirectional_dynamic_rnn / dynamic_rnn with sequence_length
 * (encoder.py:77-89) and the decoder's lm_cell loop (attn_decoder.py:148).
 * Row (b,t) of every [.,.,x] buffer is b*sb + t*st.  G [rows][ndir][H][4]: in =
 * x-projection+bias, out = gate activations (fwd) / d pre-activations (bwd).
 * Hout [rows][ndir*H] must be zero-initialised (rows with t >= len stay 0).
 * ctr_ws: >= e2e_lstm_rec_workspace_bytes(B, H, ndir) bytes of scratch (exchange tiles of the cluster kernels; with
 * less, a slower kernel that needs only 4*ndir*ceil(B/4) bytes serves the call); err_flag: device int set to 1 if a
 * step barrier ever times out. */
/* 0 (default): fastest eligible kernel -- the warp-specialised register-resident multicast-cluster
 * recurrence for H in {128, 256}, its H = 512 form (W_hh hi plane in registers, lo plane in shared memory; 64 KB of
 * workspace per CTA for the backward reduce-scatter), else the L2-exchange kernel with per-group global counters;
 * 1: always the L2-exchange kernel; 5 / 6: warp-specialised kernel with 1 / 2 slices per cluster; 7 / 8: its forward /
 * backward pass on the tf32 + bf16 products instead of the fp16 split scheme (test hooks). */
int e2e_set_rec_mode(int mode);
/* test hook: device buffer (>= 5*T int64) receiving per-step clock64 stamps of the forward cluster kernel, or NULL */
int e2e_set_rec_debug(long long* dbg);
/* test hook: 1 (default) = the persistent decoder kernels synchronise per 16-row block with point-to-point counters,
 * 0 = three grid-wide barriers per step */
int e2e_set_dec_sync(int p2p);
size_t e2e_lstm_rec_workspace_bytes(int B, int H, int ndir);
int e2e_lstm_rec_fwd(void* stream, int B, int T, int Tp, int H, int ndir, long long sb, long long st,
                     float* G, float* Hout, float* Cst, const float* Wh, const int* lens,
                     void* ctr_ws, size_t ctr_ws_bytes, int* err_flag);
int e2e_lstm_rec_bwd(void* stream, int B, int T, int Tp, int H, int ndir, long long sb, long long st,
                     float* G, const float* Cst, const float* Wh, const float* dOut, const int* lens,
                     void* ctr_ws, size_t ctr_ws_bytes, int* err_flag);
/* e2e_lstm_rec_fwd for a forward-only layer (ndir = 1, H in {128, 256}) that CONTINUES a sequence: the cell state
 * entering step 0 is Cst at time -1 (the row before the pointer passed in); the caller adds h_{-1} . Wh into the
 * pre-activations G of step 0.  Lets the decoder's LM-LSTM run in segments (scheduled sampling). */
int e2e_lstm_rec_fwd_carry(void* stream, int B, int T, int Tp, int H, long long sb, long long st,
                           float* G, float* Hout, float* Cst, const float* Wh, const int* lens,
                           void* ctr_ws, size_t ctr_ws_bytes, int* err_flag);

/* Seq2SeqModel.get_batch frame stacking (seq2seq_model.py:164-183) + initial
 * striding (encoder.py:149-153) + zero padding to Tp rows per utterance. */
int e2e_prepare_input(void* stream, int B, int T, int F, int Tp, int stack, int stride,
                      const float* in, float* out);

/* embedding_lookup (decoder.py:101) and its IndexedSlices gradient */
int e2e_embed_gather(void* stream, int n, int E, const float* emb, const long long* ids, float* out);
int e2e_embed_scatter_add(void* stream, int n, int E, float* demb, const long long* ids,
                          const float* dout, int ldd);

/* The sequential part of AttnDecoder.__call__ (attn_decoder.py:76-166, raw_rnn
 * loop; step order of SURVEY.md A.4) under teacher forcing, after the
 * state-independent parts (embedding, LM-LSTM, the lm_output half of
 * InputProjection) have been batched over all steps by the caller. */
typedef struct {
    int B, U, E, Hd, A, D, Tn, Tp;     /* Tn: attention length, Tp: rows per utterance in HF/enc */
    int gemm_mode;
    const float* in_k;     /* InputProjection/kernel  [Hd+D, E] */
    const float* dec_k;    /* basic_lstm_cell_1/kernel [E+Hd, 4Hd] */
    const float* dec_b;    /* [4Hd] */
    const float* q_k;      /* Attention/kernel [Hd, A] */
    const float* q_b;      /* [A] */
    const float* attn_v;   /* AttnV [A] */
    const float* pre;      /* [U,B,E]  lm_output . in_k[:Hd] + in_b */
    const float* HF;       /* [B,Tp,A] hidden_features */
    const float* enc;      /* [B,Tp,D] encoder states */
    const int* enc_len;    /* [B] */
    const int* lens;       /* [B] target lengths */
    float* xh;             /* [U,B,E+Hd]  (xin_t | committed h_{t-1}); zero-initialised */
    float* cprev;          /* [U,B,Hd]    committed c_{t-1};           zero-initialised */
    float* acts;           /* [U,B,4Hd] */
    float* cat;            /* [U,B,Hd+D]  (c_new_t | ctx_t) */
    float* y;              /* [U,B,A] */
    float* alpha;          /* [U,B,Tn] */
    float* gates_tmp;      /* [B,4Hd] */
} e2e_dec_loop_fwd_args;
int e2e_decoder_loop_fwd(void* stream, const e2e_dec_loop_fwd_args* a);

typedef struct {
    e2e_dec_loop_fwd_args f;   /* same weights / saved tensors as the forward */
    float* dcat;           /* [U,B,Hd+D] in: d(c_new|ctx) from AttnProjection; updated in place */
    float* dgates;         /* [U,B,4Hd] out */
    float* dxh;            /* [U,B,E+Hd] out: (dxin_t | dh_{t-1}) */
    float* dy;             /* [U,B,A] out */
    float* dv_part;        /* [B*Tn,A] out: per (row, position) partial d attn_v (sum the rows) */
    float* dHF;            /* [B,Tp,A] zero-initialised; rows of valid positions are written after the loop */
    float* denc;           /* [B,Tp,D] accumulated into (after the loop: sum_t alpha_t dctx_t) */
    float* dc_carry;       /* [B,Hd]   zero-initialised scratch */
    float* ds;             /* [U,B,Tn] out: d(pre-softmax scores) per step, kept for the sums over the steps */
} e2e_dec_loop_bwd_args;
int e2e_decoder_loop_bwd(void* stream, const e2e_dec_loop_bwd_args* a);

/* Persistent version of the same loop: all U steps in ONE cooperative launch per
 * direction (three grid-synchronous phases per step), with InputProjection's ctx
 * half folded into the decoder-LSTM kernel: W_ch = [in_k[Hd:] . Wx ; Wh] with
 * gate-interleaved columns, pre_g = (lm_out . in_k[:Hd] + in_b) . Wx + b.
 * Buffers marked (z) must be zero-initialised by the caller. */
typedef struct {
    int B, U, Hd, A, D, Tn, Tp;
    int t0, t1;            /* forward only: run steps [t0, t1) (t1 <= 0: all U); the state entering t0 is read from the
                              step buffers (cat[t0-1], hprev[t0], cprev[t0]) a previous launch wrote */
    const float* W_ch;     /* [D+Hd, 4Hd] */
    const float* pre_g;    /* [U,B,4Hd] */
    const float* q_k;      /* [Hd, A] */
    const float* q_b;      /* [A] */
    const float* attn_v;   /* [A] */
    const float* HF;       /* [B,Tp,A] */
    const float* enc;      /* [B,Tp,D] */
    const int* enc_len;
    const int* lens;
    float* cat;            /* [U,B,Hd+D]  (c_new | ctx) */
    float* hprev;          /* [U,B,Hd] committed h_{t-1} (z) */
    float* cprev;          /* [U,B,Hd] committed c_{t-1} (z) */
    float* acts;           /* [U,B,4Hd] gate activations, interleaved */
    float* y;              /* [U,B,A] */
    float* alpha;          /* [U,B,Tn] */
    /* backward only */
    float* dcat;           /* [U,B,Hd+D] in: AttnProjection gradient; ctx half accumulates d ctx */
    float* dz;             /* [U,B,4Hd] out: d gate pre-activations, interleaved */
    float* dch;            /* [U,B,D+Hd] out: (d ctx_{t-1} | d h_{t-1}) produced at step t */
    float* dy;             /* [U,B,A] out */
    float* ds;             /* [U,B,Tn] out: d attention scores */
    float* dc_carry;       /* [B,Hd] (z) */
    unsigned* ctr;         /* 1 counter of scratch */
    int* err;              /* barrier-timeout flag */
} e2e_dec_persist_args;
/* 1 if the shapes in `a` (B, U, Hd, A, D, Tn, Tp; pointers unused) fit the persistent kernels' shared memory,
 * else 0: the caller then runs the per-step kernels (e2e_decoder_loop_fwd/bwd). */
int e2e_decoder_persist_fits(const e2e_dec_persist_args* a);
int e2e_decoder_persist_fwd(void* stream, const e2e_dec_persist_args* a);
/* also accumulates denc [B,Tp,D] += sum_t alpha_t dctx_t, writes dHF [B,Tp,A] (z) and dv_part [B*Tn, A] */
int e2e_decoder_persist_bwd(void* stream, const e2e_dec_persist_args* a, float* denc, float* dHF, float* dv_part);

/* single kernels of the loop, exposed for inference (greedy / beam) and tests */
int e2e_attn_fwd(void* stream, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
                 const int* enc_len, const float* y, const float* v, float* alpha, float* ctx, int ldctx);
/* backward of e2e_attn_fwd for one decoder step: dHF [B,Tp,A], denc [B,Tp,D] and dv_part [B,A] are ACCUMULATED
 * (zero them first), dy [B,A] is written */
int e2e_attn_bwd(void* stream, int B, int Tn, int Tp, int A, int D, const float* HF, const float* enc,
                 const int* enc_len, const float* y, const float* v, const float* alpha, const float* dctx,
                 int lddctx, float* dHF, float* denc, float* dy, float* dv_part);
/* Pointwise halves of single cell steps for the general decoder (decoder.py:49-82: MultiRNNCell stacks when
 * num_layers_dec > 1, GRUCell when use_lstm=False); contiguous [n, *] buffers; the matrix halves are e2e_gemm.
 *   LSTM (BasicLSTMCell): z [n,4H] = (i|j|f|o); c' = c sig(f+1) + sig(i) tanh(j); h' = tanh(c') sig(o);
 *        backward takes dc', dh' (either may be NULL = zero) and writes dz, dc.
 *   GRU (GRUCell): gate: zg [n,2H] = (r|u) -> rh = sig(r) h, u = sig(u); out: h' = u h + (1-u) tanh(zc). */
int e2e_lstm_point_fwd(void* stream, int n, int H, const float* z, const float* c_prev, float* c_new, float* h_new);
int e2e_lstm_point_bwd(void* stream, int n, int H, const float* z, const float* c_prev, const float* c_new,
                       const float* dc_new, const float* dh_new, float* dz, float* dc_prev);
int e2e_gru_gate_fwd(void* stream, int n, int H, const float* zg, const float* h_prev, float* rh, float* u);
int e2e_gru_gate_bwd(void* stream, int n, int H, const float* zg, const float* h_prev, const float* drh,
                     const float* du, float* dzg, float* dh_prev);
int e2e_gru_out_fwd(void* stream, int n, int H, const float* zc, const float* u, const float* h_prev, float* h_new);
int e2e_gru_out_bwd(void* stream, int n, int H, const float* zc, const float* u, const float* h_prev,
                    const float* dh_new, float* dzc, float* du, float* dh_prev);
int e2e_dec_pointwise_fwd(void* stream, int B, int H, int t, const float* gates_pre, const float* cprev,
                          const float* hprev, int ldh, const int* lens, float* acts, float* cnew_out,
                          int ldc, float* c_next, float* h_next, int ldhn, float* h_new_out);
int e2e_mask_rows(void* stream, int U, int B, int V, float* logits, const int* lens);
int e2e_argmax_rows(void* stream, int rows, int V, const float* x, int ldx, long long* out);

/* LossUtils.cross_entropy_loss (losses.py:6-35): fwd writes lse[U*B], cost[U*B]
 * and loss[0]; bwd writes dlogits = gscale[0] * dloss/dlogits. */
int e2e_ce_fwd(void* stream, int U, int B, int V, const float* logits, const long long* targets,
               const int* lens, float* lse, float* cost, float* loss);
int e2e_ce_bwd(void* stream, int U, int B, int V, const float* logits, const long long* targets,
               const int* lens, const float* lse, const float* gscale, float* dlogits);

/* Auxiliary CTC on a lower encoder layer (north-star requirement; the reference
 * only keeps the hook, encoder.py:143-144,160-161 / seq2seq_model.py:104).
 * TF-1.x tf.nn.ctc_loss semantics, blank = C-1.  logits row (b,t) = b*sb + t*st.
 * Writes loss_b[B] (= -log p) and grad = out_scale * d(loss_b)/dlogits.
 * ws: caller-provided scratch of e2e_ctc_workspace_floats(T, B, max_label_len) floats
 * (gathered emissions, scaled alpha, scaled beta; three launches: emit, sweep, grad). */
int e2e_row_lse(void* stream, int rows, int V, const float* x, int ldx, float* lse);
int e2e_ctc_fwd_grad(void* stream, int T, int B, int C, long long sb, long long st, const float* logits,
                     const float* lse_rows, const int* in_lens, const long long* labels, int ldl,
                     const int* label_lens, int max_label_len, float* ws, float* loss_b,
                     float* grad, float out_scale);
size_t e2e_ctc_workspace_floats(int T, int B, int max_label_len);

/* Beam search decoder step in float64 (beam_search.py:137-221; dtype flow of SURVEY.md
 * A.6: fp32 weights / embeddings / encoder states widened at load, fp64 arithmetic),
 * batched over every live hypothesis.  BasicLSTM.__call__ (basic_lstm.py:14-23),
 * calc_attention (beam_search.py:150-159, no length mask), get_top_k's
 * log-softmax + lm_weight term + top-k (beam_search.py:196-214). */
/* e2e_gemm_f64 with the float32 weights widened to float64 once by the caller ((double)float is exact), on the FP64
 * tensor cores (mma.sync m8n8k4 f64; the products of a k4 block are summed inside the tensor core, so the last bits
 * differ from e2e_gemm_f64's sequential FMAs).  K % 16 == 0, even lda / ldb, 16-byte aligned A and B.
 * e2e_set_f64_mma(0) selects the register-tiled DFMA kernel instead (bit-identical to e2e_gemm_f64; test hook). */
int e2e_set_f64_mma(int on);
int e2e_gemm_f64d(void* stream, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                  int ldc, const float* bias);
int e2e_gemm_f64(void* stream, int M, int N, int K, const double* A, int lda, const float* B, int ldb,
                 double* C, int ldc, const float* bias);
/* C[M,N] = [A1 | A2] . B + bias + Z[zrow[m], :] on the FP64 tensor cores: A1 [M,K1] and A2 [M,K2] (optional, K2 = 0)
 * are the two halves of the reference's np.concatenate([x, h]) (basic_lstm.py:17; beam_search.py:186,194), never
 * materialised; Z (optional, float64 [*, ldz], row zrow[m] per output row) carries the part of a product that depends
 * only on the row's token -- emb[tok] . Wx + b of the LM-LSTM, a table computed once per model.  B float64 [K1+K2, N];
 * K1, K2 % 16 == 0, even row strides, 16-byte aligned operands. */
int e2e_gemm_f64d_cat(void* stream, int M, int N, int K1, int K2, const double* A1, int lda1, const double* A2,
                      int lda2, const double* B, int ldb, double* C, int ldc, const float* bias, const double* Z,
                      int ldz, const long long* zrow);
/* The same product followed by BasicLSTM.__call__ (basic_lstm.py:14-23) in the epilogue: (c_out, h_out) from c_prev and
 * the pre-activations, which are never written.  B, bias and Z hold the 4H gate columns INTERLEAVED in blocks of 32:
 * column 32 q + 8 g + i = gate g (i, j, f, o) of unit 8 q + i (e2e_asr_b200.beam_search.lstm_gate_perm).  H % 16 == 0.
 * Same formulas as e2e_lstm_step_f64 on the same sums. */
int e2e_gemm_f64d_lstm(void* stream, int M, int H, int K1, int K2, const double* A1, int lda1, const double* A2,
                       int lda2, const double* B, int ldb, const float* bias, const double* Z, int ldz,
                       const long long* zrow, const double* c_prev, double* c_out, double* h_out, int ldh);
int e2e_lstm_step_f64(void* stream, int n, int H, const double* z, const double* c_prev, double* c_out,
                      double* h_out, int ldh);
/* e2e_attn_beam_f64 for hypotheses stored in groups of `beam` rows per utterance (rows u*beam .. u*beam+beam-1 share
 * row_off / Tlen of the group's first row): one CTA per utterance reads its encoder rows once for all hypotheses;
 * bit-identical results (calc_attention, beam_search.py:150-159). */
int e2e_attn_beam_group_f64(void* stream, int N, int beam, int A, int D, int Tmax, const float* HF, const float* enc,
                            const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx, int ldctx);
int e2e_attn_beam_f64(void* stream, int n, int A, int D, int Tmax, const float* HF, const float* enc,
                      const int* row_off, const int* Tlen, const double* y, const float* v, double* ctx,
                      int ldctx);
/* calc_attention from exponentials: out[i] = exp(clamp(2 x[i], -300, 300)) in float64 (EHF = the table of
 * enc . AttnW, filled once per decode), and e2e_attn_beam_group_f64 with tanh(h + y) evaluated as
 * 1 - 2 / (EHF * exp(2 y) + 1) -- one multiplication and one division per (frame, hypothesis, a) instead of an exp and
 * a division; same summation order; differs from the tanh form by the rounding of one product (< 3e-16 absolute).
 * beam <= 16, Tmax <= 256. */
int e2e_exp2x_f64(void* stream, size_t n, const float* x, double* out);
int e2e_attn_beam_group_e_f64(void* stream, int N, int beam, int A, int D, int Tmax, const double* EHF,
                              const float* enc, const int* row_off, const int* Tlen, const double* y, const float* v,
                              double* ctx, int ldctx);
int e2e_logsoftmax_topk_f64(void* stream, int n, int V, const double* logits, const double* lm_logits,
                            double lm_weight, const int* krow, int kmax, int* out_idx, double* out_val,
                            double* scratch);
int e2e_embed_gather_f64(void* stream, int n, int E, const float* emb, const long long* ids, double* out, int ldo);

/* Candidate merge of one beam-search step for all utterances on the device (beam_search.py:255-266, 294-329).
 * Hypotheses live in fixed slots: utterance u owns rows [u*beam, (u+1)*beam), live ones first; R = N*beam.
 * Per utterance: candidates val[row][j] + score[row] (live rows, j < k_u[u]) -> the k_u[u] best (the reference's
 * np.argpartition(., -k)[-k:] as a set) -> EOS candidates retire to the final list (k_u shrinks), the others become
 * the new live rows.  *step is read on the device (the launch is replayed from a CUDA graph); word_ins_penalty *
 * (step + 1) is added from step 1 on (beam_search.py:321-322; the step-0 candidates keep the bare score, :258-260). */
typedef struct {
    int N, beam, R, eos_id;
    double word_ins_penalty;
    const int* step;           /* [1] device: index of this decoding step */
    const int* out_idx;        /* [R, beam] top tokens per row (e2e_logsoftmax_topk_f64) */
    const double* out_val;     /* [R, beam] their log-probabilities */
    const double* score;       /* [R] score of the hypothesis in each row */
    const int* alive;          /* [R] 1 = the row holds a live hypothesis */
    int* k_u;                  /* [N] current beam size per utterance (in/out) */
    long long* new_tok;        /* [R] out: token fed to the next step */
    double* new_score;         /* [R] out */
    int* parent;               /* [R] out: row (of this step's input rows) each new row extends */
    int* new_alive;            /* [R] out */
    int* krow;                 /* [R] out: k of the row's utterance for live rows, 0 for dead ones */
    int* par_hist;             /* [max_steps, R] out: row `step` = parent (or -1) */
    int* tok_hist;             /* [max_steps, R] out: row `step` = token (or -1) */
    int* fin_cnt;              /* [N] number of finished hypotheses (in/out) */
    int* fin_step;             /* [R] step at which the f-th finished hypothesis of u (index u*beam+f) emitted EOS */
    int* fin_row;              /* [R] its parent row */
    double* fin_score;         /* [R] its score */
    int* n_live;               /* [1] += number of live hypotheses after this step (caller zeroes it) */
} e2e_beam_merge_args;
int e2e_beam_merge(void* stream, const e2e_beam_merge_args* a);

/* dst[m][r, :] = src[m][parent[r], :] for up to 8 float64 state matrices (back-pointer gather of a beam step) */
typedef struct {
    int nmat;
    int width[8];
    const double* src[8];
    double* dst[8];
} e2e_beam_gather_args;
int e2e_beam_gather(void* stream, int R, const int* parent, const e2e_beam_gather_args* g);

/* tf.clip_by_global_norm (seq2seq_model.py:150-151) on the flat gradient buffer:
 * sumsq: out (+)= sign * sum x^2; clip: x *= pre_scale * clip / max(sqrt(sumsq), clip), norm_out = sqrt(sumsq).
 * pre_scale = 1/n folds the averaging of a data-parallel SUM of n rank gradients into the clipping pass (sumsq is
 * then taken with sign = 1/n^2).  err_flag (may be NULL): the persistent kernels' barrier-timeout flag; when set the
 * gradients are zeroed and norm_out reads NaN instead of garbage being applied. */
int e2e_sumsq(void* stream, size_t n, const float* x, float* partials296, float* out, float sign, int accumulate);
int e2e_clip_by_norm(void* stream, size_t n, float* x, const float* sumsq, float clip, float* norm_out,
                     float pre_scale, const int* err_flag);
int e2e_scale(void* stream, size_t n, float* x, const float* dev_scalar, float a);
int e2e_mean(void* stream, int n, const float* x, float* out);
int e2e_axpy(void* stream, size_t n, float a, const float* x, float* y);   /* y += a*x */

/* tf.train.AdamOptimizer.apply_gradients (seq2seq_model.py:137,153-155) on the flat parameter / gradient /
 * moment buffers (n a multiple of 4): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps),
 * with lr_t = lr sqrt(1-b2^t) / (1-b1^t) computed by the caller. */
int e2e_adam(void* stream, size_t n, float* param, const float* grad, float* m, float* v, float lr_t, float beta1,
             float beta2, float eps);

/* GRU recurrence of an encoder layer (encoder.py:48 GRUCell under (bidirectional_)dynamic_rnn, encoder.py:77-89),
 * batch-major rows n = b*Tp + t, directions side by side in the feature axis (fw | bw):
 *   Gg [B*Tp, ndir*2H]: x . gates/kernel[:I] + gates/bias on entry (r | u per direction); (r | u) activations after
 *                       the forward pass; d(gate pre-activations) after the backward pass (zero for t >= len);
 *   Gc [B*Tp, ndir*H] : x . candidate/kernel[:I] + candidate/bias -> c -> d(candidate pre-activation);
 *   out [B*Tp, ndir*H]: h_t, must be zero-initialised (frames t >= len stay zero);  RH: r_t * h_{t-1}, likewise;
 *   Wg_h [ndir][H][2H], Wc_h [ndir][H][H]: the recurrent halves of the TF kernels; the backward pass takes their
 *   transposes Wg_hT [ndir][2H][H], Wc_hT [ndir][H][H].  H <= 512.  One CTA per 4 utterances and direction. */
int e2e_gru_rec_fwd(void* stream, int B, int T, int Tp, int H, int ndir, float* Gg, float* Gc, float* out, float* RH,
                    const float* Wg_h, const float* Wc_h, const int* lens);
int e2e_gru_rec_bwd(void* stream, int B, int T, int Tp, int H, int ndir, float* Gg, float* Gc, const float* out,
                    const float* dout, const float* Wg_hT, const float* Wc_hT, const int* lens);

/* DropoutWrapper(output_keep_prob=keep) on a recurrent cell's outputs (encoder.py:50-52, decoder.py:60-63):
 * with i = first + (index into x), y = x / keep if philox4x32_10(counter = (i/4, offset, 0, 0), key = seed)[i%4]
 * * 2^-32 < keep else 0 (`first`, a multiple of 4, lets a slice of a buffer draw the buffer's mask).
 * Stateless: the backward pass calls it again on the upstream gradient with the same (seed, offset).
 * seed_dev (may be NULL): device address of the key; when given it REPLACES `seed`, so a step captured in a CUDA
 * graph draws a fresh mask at every replay (the host rewrites the device word before launching the graph). */
int e2e_dropout(void* stream, size_t n, const float* x, float* y, float keep, unsigned long long seed,
                unsigned offset, size_t first, const unsigned long long* seed_dev);

/* Scheduled sampling (decoder.py:155-180, tf.multinomial(logits, 1)): out[r] = first index whose cumulative
 * exp(logit - max) (float64, index order) exceeds u_r * total, u_r = word 0 of
 * philox4x32_10(counter = (first_row + r, offset, 0, 0), key = seed) * 2^-32. */
int e2e_sample_rows(void* stream, int rows, int V, const float* logits, int ldl, unsigned long long seed,
                    unsigned offset, unsigned first_row, long long* out);

#ifdef __cplusplus
}
#endif
#endif /* E2E_ASR_B200_H */
