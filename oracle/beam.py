"""CPU oracle for decoding (TEST INFRASTRUCTURE): Python-3 restatement of the
reference's NumPy beam search (beam_search.py:137-338) and of the greedy id
extraction (eval_model.py:84-87,253-254).

Pinned: tests/golden/beam_*.npz hold outputs of the reference's own
beam_search.py executed in this container through a Python-2 semantics shim
(tests/golden/gen_golden.py); tests/test_oracle_golden.py checks this file
against them (ids bit-exact, scores to 1e-12).

dtype flow follows SURVEY.md A.6: weights/embeddings/enc are float32, the zero
initial states are float64 (`np.zeros`, beam_search.py:236-244), so after step 0
every GEMV is float64; `enc @ AttnW` stays float32 (beam_search.py:148).
"""
import numpy as np

from .model import softmax, sigmoid

GO_ID, EOS_ID = 1, 2


def basic_lstm(x, state, w, b):
    """reference basic_lstm.py:14-23 (returns (new_c, new_h))."""
    c, h = state
    x_h = np.concatenate((x, h), axis=0)
    i, j, f, o = np.split(np.matmul(x_h, w) + b, 4)
    f_gate = sigmoid(f + 1)
    new_c = np.multiply(c, f_gate) + np.multiply(sigmoid(i), np.tanh(j))
    new_h = np.multiply(sigmoid(o), np.tanh(new_c))
    return (new_c, new_h)


class DecParams(object):
    """Weight name -> array mapping of BeamSearch.map_dec_variables /
    map_lm_variables (beam_search.py:53-134)."""

    def __init__(self, weights, task="char"):
        p = "model/rnn_decoder_%s/" % task
        g = lambda k: np.asarray(weights[p + k])
        self.lm_lstm_w, self.lm_lstm_b = g("rnn/basic_lstm_cell/kernel"), g("rnn/basic_lstm_cell/bias")
        self.dec_lstm_w, self.dec_lstm_b = g("rnn/basic_lstm_cell_1/kernel"), g("rnn/basic_lstm_cell_1/bias")
        self.attn_dec_w, self.attn_dec_b = g("rnn/Attention/kernel"), g("rnn/Attention/bias")
        self.inp_w, self.inp_b = g("rnn/InputProjection/kernel"), g("rnn/InputProjection/bias")
        self.attn_proj_w, self.attn_proj_b = g("rnn/AttnProjection/kernel"), g("rnn/AttnProjection/bias")
        self.out_w, self.out_b = g("rnn/OutputProjection/kernel"), g("rnn/OutputProjection/bias")
        if p + "rnn/SimpleProjection/kernel" in weights:
            self.simple_w, self.simple_b = g("rnn/SimpleProjection/kernel"), g("rnn/SimpleProjection/bias")
        else:
            self.simple_w = self.simple_b = None
        self.attn_enc_w = np.squeeze(g("AttnW"))               # beam_search.py:94
        self.attn_v = g("AttnV")
        self.embedding = g("decoder/embedding")


def make_step_fn(p, enc, lm_weight=0.0, lp=None):
    """top_k_setup_with_lm (beam_search.py:163-221); lp = the LM checkpoint's variables (map_lm_variables,
    beam_search.py:111-134), by default the decoder's own (lm_path == same checkpoint)."""
    lp = p if lp is None else lp
    if enc.ndim == 3:
        enc = np.squeeze(enc, axis=0)
    attn_enc_term = np.matmul(enc, p.attn_enc_w)               # :148 (float32)

    def attention(dec_state):                                  # :150-159, no mask
        attn_dec_term = np.matmul(dec_state, p.attn_dec_w) + p.attn_dec_b
        attn_sum = np.tanh(attn_enc_term + attn_dec_term)
        attn_logits = np.squeeze(np.matmul(attn_sum, p.attn_v))
        attn_probs = softmax(attn_logits)
        return np.matmul(attn_probs, enc), attn_probs

    def get_top_k(x, x_lm, state_list, context_vec, beam_size):
        dec_state, dec_lm_state, lm_state = state_list
        dec_lm_state = basic_lstm(x, dec_lm_state, p.lm_lstm_w, p.lm_lstm_b)
        dec_lm_output = dec_lm_state[1]
        if p.simple_w is not None:
            dec_lm_output = np.matmul(dec_lm_output, p.simple_w) + p.simple_b
        x_dec = np.matmul(np.concatenate((dec_lm_output, context_vec), axis=0), p.inp_w) + p.inp_b
        dec_state = basic_lstm(x_dec, dec_state, p.dec_lstm_w, p.dec_lstm_b)
        context_vec, _ = attention(dec_state[0])               # query = new_c (:193)
        proj = np.matmul(np.concatenate((dec_state[0], context_vec), axis=0), p.attn_proj_w) + p.attn_proj_b
        log_dec = np.log(softmax(np.matmul(proj, p.out_w) + p.out_b))
        lm_state = basic_lstm(x_lm, lm_state, lp.lm_lstm_w, lp.lm_lstm_b)   # LM branch always runs (:200)
        lm_output = lm_state[1]
        if lp.simple_w is not None:
            lm_output = np.matmul(lm_output, lp.simple_w) + lp.simple_b
        log_lm = np.log(softmax(np.matmul(lm_output, lp.out_w) + lp.out_b))
        combined = log_dec + lm_weight * log_lm
        top = np.argpartition(combined, -beam_size)[-beam_size:]
        return top, combined[top], combined[top], [dec_state, dec_lm_state, lm_state], context_vec, combined

    return get_top_k


def beam_search(weights, enc, beam_size=4, lm_weight=0.0, word_ins_penalty=0, task="char",
                return_score=False, lm_weights=None):
    """BeamSearch.__call__ (beam_search.py:224-338), SURVEY.md A.7.  lm_weights: the variables of a separate LM
    checkpoint (search_params.lm_path, beam_search.py:45-46); None = the decoder's own."""
    p = DecParams(weights, task)
    lp = p if lm_weights is None else DecParams(lm_weights, task)
    step = make_step_fn(p, enc, lm_weight, lp)
    x = p.embedding[GO_ID]
    x_lm0 = lp.embedding[GO_ID]
    hs = p.dec_lstm_w.shape[1] // 4
    ls = p.lm_lstm_w.shape[1] // 4
    zero_dec = (np.zeros(hs), np.zeros(hs))
    zero_lm = (np.zeros(ls), np.zeros(ls))
    zero_attn = np.zeros(enc.shape[-1])
    k = beam_size
    live, final = [], []
    zero_lm2 = (np.zeros(lp.lm_lstm_w.shape[1] // 4), np.zeros(lp.lm_lstm_w.shape[1] // 4))
    top, ms, _, states, ctx, _ = step(x, x_lm0, [zero_dec, zero_lm, zero_lm2], zero_attn, k)
    for idx in range(top.shape[0]):
        ent = ([int(top[idx])], states, ctx, ms[idx])
        if top[idx] == EOS_ID:
            final.append(ent)
            k -= 1
        else:
            live.append(ent)
    step_count = 1
    while step_count < 120 and k > 0:
        nstates, nctx, scores, mscores, indices = [], [], [], [], []
        for seq, st, cx, sc in live:
            e = p.embedding[seq[-1]]
            top, ms, ts, st2, cx2, _ = step(e, lp.embedding[seq[-1]], st, cx, k)
            nstates.append(st2); nctx.append(cx2)
            indices.append(top); scores.append(ts + sc); mscores.append(ms + sc)
        all_scores = np.concatenate(scores)
        all_m = np.concatenate(mscores)
        all_idx = np.concatenate(indices)
        sel = np.argpartition(all_scores, -k)[-k:]
        parents = sel // k                                      # :306 (py2 int division)
        new_live = []
        for i in range(k):                                      # bound fixed before k shrinks (:310)
            pi = int(parents[i])
            seq = live[pi][0] + [int(all_idx[sel[i]])]
            ent = (seq, nstates[pi], nctx[pi], all_m[sel[i]] + word_ins_penalty * len(seq))
            if seq[-1] == EOS_ID:
                final.append(ent)
                k -= 1
            else:
                new_live.append(ent)
        live = new_live
        step_count += 1
    final += live
    best = max(final, key=lambda e: e[3])                       # no length normalisation (:336)
    ids = np.stack(best[0], axis=0)
    return (ids, best[3]) if return_score else ids


def greedy_ids_from_logits(logits, batch_size):
    """eval_model.py:84-87: argmax over V, reshape [T,B], transpose -> [B,T]."""
    return np.argmax(logits, axis=1).reshape(-1, batch_size).T


def cut_at_eos(row):
    """eval_model.py:253-254."""
    row = list(int(v) for v in row)
    return row[:row.index(EOS_ID)] if EOS_ID in row else row
