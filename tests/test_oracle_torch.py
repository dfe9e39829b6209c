"""Independent cross-checks of the (TF-unpinned) parts of the oracle:
torch-CPU autograd for every hand-derived gradient, torch.nn.LSTM with packed
sequences for the dynamic_rnn length semantics, F.ctc_loss for CTC, and float64
central finite differences."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from e2e_asr_b200 import synth
from oracle import model as om

torch.set_num_threads(4)


def _t(a, grad=False):
    return torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=grad)


def torch_lstm_cell(x, c, h, k, b):
    z = torch.cat([x, h], dim=1) @ k + b
    i, j, f, o = z.chunk(4, dim=1)
    c2 = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
    return c2, torch.tanh(c2) * torch.sigmoid(o)


def torch_train_step(W, batch, cfg, tasks, num_layers, ctc_tasks, avg=True):
    """Straightforward torch re-implementation (loops + autograd), written from
    SURVEY.md Appendix A, not from oracle/model.py's backward."""
    x = _t(batch["logmel"])
    lens = torch.tensor(batch["logmel_len"])
    depth_of = dict(num_layers)
    for t_, (d_, _) in ctc_tasks.items():
        depth_of[t_] = d_
    max_depth = max(depth_of.values())
    states, lens_d = {}, {}
    res = 1
    H = cfg.H
    for i in range(max_depth):
        l = i + 1
        outs = []
        B, T, _ = x.shape
        for d in ("fw", "bw"):
            k = W["model/encoder/RNNLayer%d/bidirectional_rnn/%s/basic_lstm_cell/kernel" % (l, d)]
            b = W["model/encoder/RNNLayer%d/bidirectional_rnn/%s/basic_lstm_cell/bias" % (l, d)]
            rows = []
            for bi in range(B):           # per-row, literally reverse_sequence + rnn + reverse_sequence
                n = int(lens[bi])
                seq = x[bi, :n]
                if d == "bw":
                    seq = seq.flip(0)
                c = torch.zeros(1, H, dtype=torch.float64)
                h = torch.zeros(1, H, dtype=torch.float64)
                hs = []
                for t in range(n):
                    c, h = torch_lstm_cell(seq[t:t + 1], c, h, k, b)
                    hs.append(h)
                o = torch.cat(hs, 0) if hs else torch.zeros(0, H, dtype=torch.float64)
                if d == "bw":
                    o = o.flip(0)
                rows.append(torch.cat([o, torch.zeros(T - n, H, dtype=torch.float64)], 0))
            outs.append(torch.stack(rows, 0))
        out = torch.cat(outs, 2)
        states[l], lens_d[l] = out, lens
        if i != max_depth - 1 and res < 8:
            if int(lens.max()) % 2:
                out = torch.cat([out, torch.zeros(B, 1, out.shape[2], dtype=torch.float64)], 1)
            x = out.reshape(B, out.shape[1] // 2, out.shape[2] * 2)
            lens = (lens + 1) // 2
            res *= 2
        else:
            x = out
    losses = {}
    for task in tasks:
        p = "model/rnn_decoder_%s/" % task
        enc, el = states[num_layers[task]], lens_d[num_layers[task]]
        B, Tn, D = enc.shape
        dec_inp = torch.tensor(batch[task]).T
        sl = torch.tensor(batch[task + "_len"])
        Umax = int(sl.max())
        HF = enc @ W[p + "AttnW"].reshape(D, -1)
        mask = (torch.arange(Tn)[None, :] < el[:, None]).double()
        Hd, Hl = cfg.Hd, cfg.Hl
        c = torch.zeros(B, Hd, dtype=torch.float64); h = torch.zeros_like(c)
        cl = torch.zeros(B, Hl, dtype=torch.float64); hl = torch.zeros_like(cl)
        ctx = torch.zeros(B, D, dtype=torch.float64)
        logits = []
        for t in range(Umax):
            u = W[p + "decoder/embedding"][dec_inp[t]]
            cl, hl = torch_lstm_cell(u, cl, hl, W[p + "rnn/basic_lstm_cell/kernel"], W[p + "rnn/basic_lstm_cell/bias"])
            m = hl
            if Hl != Hd:
                m = hl @ W[p + "rnn/SimpleProjection/kernel"] + W[p + "rnn/SimpleProjection/bias"]
            xin = torch.cat([m, ctx], 1) @ W[p + "rnn/InputProjection/kernel"] + W[p + "rnn/InputProjection/bias"]
            c2, h2 = torch_lstm_cell(xin, c, h, W[p + "rnn/basic_lstm_cell_1/kernel"], W[p + "rnn/basic_lstm_cell_1/bias"])
            y = c2 @ W[p + "rnn/Attention/kernel"] + W[p + "rnn/Attention/bias"]
            s = (torch.tanh(HF + y[:, None, :]) * W[p + "AttnV"]).sum(2)
            a = torch.softmax(s, 1) * mask
            a = a / a.sum(1, keepdim=True)
            ctx = (a[:, :, None] * enc).sum(1)
            proj = torch.cat([c2, ctx], 1) @ W[p + "rnn/AttnProjection/kernel"] + W[p + "rnn/AttnProjection/bias"]
            lg = proj @ W[p + "rnn/OutputProjection/kernel"] + W[p + "rnn/OutputProjection/bias"]
            live = (t < sl)[:, None]
            logits.append(torch.where(live, lg, torch.zeros_like(lg)))
            c = torch.where(live, c2, c)
            h = torch.where(live, h2, h)
        logits = torch.stack(logits, 0)                                  # [U,B,V]
        targets = dec_inp[1:Umax + 1]
        cost = F.cross_entropy(logits.reshape(Umax * B, -1), targets.reshape(-1), reduction="none").reshape(Umax, B)
        w_ = (torch.arange(Umax)[:, None] < sl[None, :]).double()
        losses[task] = ((w_ * cost).sum(0) / sl.double()).mean()
    for task, (d_, vocab) in ctc_tasks.items():
        st = states[d_]
        lg = st @ W["model/ctc_%s/kernel" % task] + W["model/ctc_%s/bias" % task]
        lp = F.log_softmax(lg, 2).transpose(0, 1)
        lb = F.ctc_loss(lp, torch.tensor(batch[task]), lens_d[d_], torch.tensor(batch[task + "_len"]),
                        blank=vocab, reduction="none", zero_infinity=False)
        losses[task] = lb.mean()
    n = len(losses)
    total = sum(losses.values()) * ((1.0 / n) if avg else 1.0)
    return losses, total, states


@pytest.mark.parametrize("cname,ctc", [("tiny", True), ("tiny_b", True), ("tiny", False)])
def test_train_step_grads_vs_torch_autograd(cname, ctc):
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ctc_tasks = cfg.ctc if ctc else {}
    out = om.train_step(w, batch, num_layers={"char": 4}, ctc_tasks=ctc_tasks, max_gradient_norm=1.0)
    W = {k: _t(v, True) for k, v in w.items() if ctc or "ctc_" not in k}
    losses, total, states = torch_train_step(W, batch, cfg, ("char",), {"char": 4}, ctc_tasks)
    total.backward()
    assert abs(float(total) - out["total_loss"]) < 1e-10
    for k in losses:
        assert abs(float(losses[k]) - out["losses"][k]) < 1e-10
    for d_ in out["states"]:
        np.testing.assert_allclose(out["states"][d_], states[d_].detach().numpy(), rtol=1e-10, atol=1e-12)
    sq = 0.0
    for k, v in W.items():
        g = v.grad.numpy()
        np.testing.assert_allclose(out["grads"][k], g, rtol=1e-8, atol=1e-11, err_msg=k)
        sq += float((g ** 2).sum())
    assert abs(out["dense_norm"] - np.sqrt(sq)) < 1e-9
    # TF IndexedSlices norm >= is a different number from the dense one when tokens repeat
    scale = 1.0 / max(out["norm"], 1.0)
    for k in out["grads"]:
        np.testing.assert_allclose(out["clipped"][k], out["grads"][k] * scale, rtol=1e-12)


def test_birnn_matches_torch_nn_lstm_packed():
    """dynamic_rnn length semantics against torch.nn.LSTM with
    pack_padded_sequence; gate order i,j,f,o -> i,f,g,o and forget bias +1."""
    rng = np.random.Generator(np.random.PCG64(7))
    B, T, I, H = 5, 13, 6, 7
    X = rng.standard_normal((B, T, I))
    lens = np.array([13, 9, 1, 4, 12])
    for b in range(B):
        X[b, lens[b]:] = 0
    ks = [rng.uniform(-0.5, 0.5, (I + H, 4 * H)) for _ in range(2)]
    bs = [rng.uniform(-0.5, 0.5, (4 * H,)) for _ in range(2)]
    out, _ = om.birnn_layer_fwd(X, lens, ks[0], bs[0], ks[1], bs[1])
    lstm = torch.nn.LSTM(I, H, batch_first=True, bidirectional=True).double()

    def reorder(m):  # columns i,j,f,o -> i,f,g(j),o
        i, j, f, o = np.split(m, 4, axis=-1)
        return np.concatenate([i, f, j, o], axis=-1)
    with torch.no_grad():
        for d, sfx in enumerate(("", "_reverse")):
            k = reorder(ks[d]); b = reorder(bs[d].copy())
            b[H:2 * H] += 1.0
            getattr(lstm, "weight_ih_l0" + sfx).copy_(_t(k[:I].T))
            getattr(lstm, "weight_hh_l0" + sfx).copy_(_t(k[I:].T))
            getattr(lstm, "bias_ih_l0" + sfx).copy_(_t(b))
            getattr(lstm, "bias_hh_l0" + sfx).zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(_t(X), torch.tensor(lens), batch_first=True,
                                                         enforce_sorted=False)
        o, _ = lstm(packed)
        o, _ = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=T)
    np.testing.assert_allclose(out, o.numpy(), rtol=1e-10, atol=1e-12)


def test_ctc_vs_torch_and_repeats():
    rng = np.random.Generator(np.random.PCG64(8))
    T, B, C = 12, 4, 5
    logits = rng.standard_normal((T, B, C)) * 2
    in_lens = np.array([12, 7, 5, 9])
    labels = np.array([[0, 0, 1, 3], [2, 2, 2, 0], [1, 0, 0, 0], [3, 1, 3, 0]])
    lab_lens = np.array([4, 3, 1, 3])
    loss, grad = om.ctc_loss(logits, in_lens, labels, lab_lens)
    lt = _t(logits, True)
    lb = F.ctc_loss(F.log_softmax(lt, 2), torch.tensor(labels), torch.tensor(in_lens), torch.tensor(lab_lens),
                    blank=C - 1, reduction="none")
    lb.sum().backward()
    np.testing.assert_allclose(loss, lb.detach().numpy(), rtol=1e-10)
    np.testing.assert_allclose(grad, lt.grad.numpy(), rtol=1e-8, atol=1e-10)
    assert np.all(grad[7:, 1] == 0)


def test_finite_differences_float64():
    cfg = synth.get_config("tiny")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    kw = dict(num_layers={"char": 4}, ctc_tasks=cfg.ctc)
    out = om.train_step(w, batch, **kw)
    rng = np.random.Generator(np.random.PCG64(1))
    eps = 1e-6
    names = sorted(w.keys())
    for k in names:
        idx = tuple(int(rng.integers(0, s)) for s in w[k].shape)
        wp = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wm = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wp[k][idx] += eps
        wm[k][idx] -= eps
        fd = (om.train_step(wp, batch, want_grads=False, **kw)["total_loss"]
              - om.train_step(wm, batch, want_grads=False, **kw)["total_loss"]) / (2 * eps)
        assert abs(fd - out["grads"][k][idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx, fd, out["grads"][k][idx])


def test_cross_entropy_matches_reference_formula():
    rng = np.random.Generator(np.random.PCG64(2))
    U, B, V = 5, 3, 7
    logits = rng.standard_normal((U * B, V))
    targets = rng.integers(0, V, (U, B))
    sl = np.array([5, 2, 3])
    loss, dl = om.cross_entropy_loss(logits, targets, sl)
    lt = _t(logits, True)
    cost = F.cross_entropy(lt, torch.tensor(targets.reshape(-1)), reduction="none").reshape(U, B)
    w_ = (torch.arange(U)[:, None] < torch.tensor(sl)[None, :]).double()
    ref = ((w_ * cost).sum(0) / torch.tensor(sl).double()).mean()
    ref.backward()
    assert abs(float(ref) - loss) < 1e-12
    np.testing.assert_allclose(dl, lt.grad.numpy(), rtol=1e-10, atol=1e-14)


# ---------------------------------------------------------------- dropout / Adam
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32_10 (kat_vectors)."""
    r = om.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = om.philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(x[0]) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = om.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    m = om.dropout_mask(200003, 0.9, 77, 2)
    assert abs((m > 0).mean() - 0.9) < 5e-3 and abs(m.mean() - 1.0) < 1e-2
    assert not np.array_equal(m, om.dropout_mask(200003, 0.9, 77, 3))      # streams differ
    np.testing.assert_array_equal(m, om.dropout_mask(200003, 0.9, 77, 2))  # stateless


def test_finite_differences_with_dropout():
    """Masks are a fixed function of (seed, stream, index): the dropped graph is still differentiable."""
    cfg = synth.get_config("tiny")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    kw = dict(num_layers={"char": 4}, ctc_tasks=cfg.ctc, out_prob=0.7, out_prob_dec=0.6, dropout_seed=5)
    out = om.train_step(w, batch, **kw)
    base = om.train_step(w, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc)
    assert abs(out["total_loss"] - base["total_loss"]) > 1e-3        # dropout is really applied
    rng = np.random.Generator(np.random.PCG64(3))
    eps = 1e-6
    for k in sorted(w.keys()):
        idx = tuple(int(rng.integers(0, s)) for s in w[k].shape)
        wp = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wm = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wp[k][idx] += eps
        wm[k][idx] -= eps
        fd = (om.train_step(wp, batch, want_grads=False, **kw)["total_loss"]
              - om.train_step(wm, batch, want_grads=False, **kw)["total_loss"]) / (2 * eps)
        assert abs(fd - out["grads"][k][idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx, fd, out["grads"][k][idx])


def test_adam_matches_tf_formula():
    """tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); p -= lr_t m / (sqrt(v) + eps)."""
    rng = np.random.Generator(np.random.PCG64(4))
    p = {"a": rng.standard_normal((3, 4)), "b": rng.standard_normal(5)}
    st = {}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v_ = {k: np.zeros_like(v) for k, v in p.items()}
    ref = {k: x.copy() for k, x in p.items()}
    for t in range(1, 4):
        g = {k: rng.standard_normal(x.shape) for k, x in p.items()}
        p = om.adam_step(p, g, st, 1e-3)
        for k in ref:
            m[k] = 0.9 * m[k] + 0.1 * g[k]
            v_[k] = 0.999 * v_[k] + 0.001 * g[k] ** 2
            ref[k] = ref[k] - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m[k] / (np.sqrt(v_[k]) + 1e-8)
            np.testing.assert_allclose(p[k], ref[k], rtol=1e-12)
    # first step moves every coordinate by ~lr * sign(g)
    assert st["t"] == 3


def test_unidirectional_encoder_forward_vs_torch_and_finite_differences():
    """bi_dir=False (encoder.py:86-89, tf.nn.dynamic_rnn): the forward-only layer against torch.nn.LSTM on packed
    sequences, and the whole forward-only step against float64 central differences."""
    rng = np.random.Generator(np.random.PCG64(11))
    B, T, I, H = 4, 11, 5, 6
    X = rng.standard_normal((B, T, I))
    lens = np.array([11, 7, 2, 9])
    for b in range(B):
        X[b, lens[b]:] = 0
    k = rng.uniform(-0.5, 0.5, (I + H, 4 * H))
    bvec = rng.uniform(-0.5, 0.5, (4 * H,))
    out, _ = om.birnn_layer_fwd(X, lens, k, bvec)
    assert out.shape == (B, T, H)
    lstm = torch.nn.LSTM(I, H, batch_first=True).double()

    def reorder(m):
        i, j, f, o = np.split(m, 4, axis=-1)
        return np.concatenate([i, f, j, o], axis=-1)
    with torch.no_grad():
        kk, bb = reorder(k), reorder(bvec.copy())
        bb[H:2 * H] += 1.0
        lstm.weight_ih_l0.copy_(_t(kk[:I].T)); lstm.weight_hh_l0.copy_(_t(kk[I:].T))
        lstm.bias_ih_l0.copy_(_t(bb)); lstm.bias_hh_l0.zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(_t(X), torch.tensor(lens), batch_first=True,
                                                         enforce_sorted=False)
        o, _ = lstm(packed)
        o, _ = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=T)
    np.testing.assert_allclose(out, o.numpy(), rtol=1e-10, atol=1e-12)

    cfg = synth.get_config("tiny_uni")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, enc_params={"bi_dir": False})
    res = om.train_step(w, batch, **kw)
    assert res["states"][cfg.L].shape[2] == cfg.H
    eps = 1e-6
    for name in sorted(w.keys()):
        idx = tuple(int(rng.integers(0, s)) for s in w[name].shape)
        wp = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wm = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wp[name][idx] += eps
        wm[name][idx] -= eps
        fd = (om.train_step(wp, batch, want_grads=False, **kw)["total_loss"]
              - om.train_step(wm, batch, want_grads=False, **kw)["total_loss"]) / (2 * eps)
        assert abs(fd - res["grads"][name][idx]) < 1e-6 * max(1.0, abs(fd)), (name, idx, fd)


def test_gru_encoder_oracle_vs_torch_autograd_and_finite_differences():
    """use_lstm=False (encoder.py:48, tf.nn.rnn_cell.GRUCell): one direction of the oracle's GRU layer against torch
    autograd on the TF-1.x cell formula ([r, u] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc);
    h' = u h + (1 - u) c; frozen state and zero output past the length), then the whole GRU-encoder step against
    float64 central differences."""
    rng = np.random.Generator(np.random.PCG64(5))
    B, T, I, H = 4, 9, 5, 6
    X = rng.standard_normal((B, T, I))
    lens = np.array([9, 5, 1, 7])
    gk, gb = rng.uniform(-0.5, 0.5, (I + H, 2 * H)), rng.uniform(-0.5, 0.5, 2 * H) + 1.0
    ck, cb = rng.uniform(-0.5, 0.5, (I + H, H)), rng.uniform(-0.5, 0.5, H)
    dout = rng.standard_normal((B, T, H))
    for rev in (False, True):
        out, cache = om._gru_dir_fwd(X, gk, gb, ck, cb, lens, rev)
        dX, (dgk, dgb, dck, dcb) = om._gru_dir_bwd(dout, cache)
        tX, tgk, tgb, tck, tcb = [_t(a).requires_grad_(True) for a in (X, gk, gb, ck, cb)]
        outs = [None] * T
        h = torch.zeros(B, H, dtype=torch.float64)
        for t in (range(T - 1, -1, -1) if rev else range(T)):
            m = torch.tensor(t < lens)[:, None]
            g = torch.sigmoid(torch.cat([tX[:, t], h], 1) @ tgk + tgb)
            r, u = g[:, :H], g[:, H:]
            c = torch.tanh(torch.cat([tX[:, t], r * h], 1) @ tck + tcb)
            hn = u * h + (1 - u) * c
            outs[t] = torch.where(m, hn, torch.zeros_like(hn))
            h = torch.where(m, hn, h)
        tout = torch.stack(outs, 1)
        np.testing.assert_allclose(out, tout.detach().numpy(), rtol=1e-12, atol=1e-13)
        (tout * _t(dout)).sum().backward()
        for got, want in ((dX, tX), (dgk, tgk), (dgb, tgb), (dck, tck), (dcb, tcb)):
            np.testing.assert_allclose(got, want.grad.numpy(), rtol=1e-10, atol=1e-12)

    cfg = synth.get_config("tiny_gru")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, enc_params={"use_lstm": False})
    res = om.train_step(w, batch, **kw)
    eps = 1e-6
    for name in sorted(w.keys()):
        idx = tuple(int(rng.integers(0, s)) for s in w[name].shape)
        wp = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wm = {n: v.astype(np.float64).copy() for n, v in w.items()}
        wp[name][idx] += eps
        wm[name][idx] -= eps
        fd = (om.train_step(wp, batch, want_grads=False, **kw)["total_loss"]
              - om.train_step(wm, batch, want_grads=False, **kw)["total_loss"]) / (2 * eps)
        assert abs(fd - res["grads"][name][idx]) < 1e-6 * max(1.0, abs(fd)), (name, idx, fd)


def test_general_decoder_oracle_consistency_and_finite_differences():
    """oracle attn_decoder_general (MultiRNNCell stacks, GRU cells; decoder.py:49-82): bit-identical to the pinned
    single-LSTM restatement for num_layers_dec=1, and float64 central differences for a 2-layer LSTM decoder, a GRU
    decoder and a 2-layer GRU decoder."""
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    a = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    b = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc,
                      dec_params={"num_layers_dec": 1, "use_lstm": True})
    assert a["total_loss"] == b["total_loss"] and a["norm"] == b["norm"]
    assert all(np.array_equal(a["grads"][k], b["grads"][k]) for k in a["grads"])
    # ... also with output dropout on (the single cell's mask stream is shared by both restatements)
    kwd = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, out_prob_dec=0.7, dropout_seed=9)
    a = om.train_step(w, batch, **kwd)
    b = om.train_step(w, batch, dec_params={"num_layers_dec": 1, "use_lstm": True}, **kwd)
    assert a["total_loss"] == b["total_loss"]
    assert all(np.array_equal(a["grads"][k], b["grads"][k]) for k in a["grads"])
    rng = np.random.Generator(np.random.PCG64(3))
    eps = 1e-6
    for cname, keep in (("tiny_dec2", 1.0), ("tiny_decgru", 1.0), ("tiny_decgru2", 1.0), ("tiny_dec2", 0.6)):
        cfg = synth.get_config(cname)
        w = synth.make_weights(cfg, bias_noise=0.1)
        batch = synth.make_batch(cfg)
        kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, out_prob_dec=keep, dropout_seed=4,
                  dec_params={"num_layers_dec": cfg.get("dec_layers", 1), "use_lstm": cfg.get("dec_lstm", True)})
        res = om.train_step(w, batch, **kw)
        for name in sorted(k for k in w if "rnn_decoder" in k):
            idx = tuple(int(rng.integers(0, s)) for s in w[name].shape)
            wp = {n: v.astype(np.float64).copy() for n, v in w.items()}
            wm = {n: v.astype(np.float64).copy() for n, v in w.items()}
            wp[name][idx] += eps
            wm[name][idx] -= eps
            fd = (om.train_step(wp, batch, want_grads=False, **kw)["total_loss"]
                  - om.train_step(wm, batch, want_grads=False, **kw)["total_loss"]) / (2 * eps)
            assert abs(fd - res["grads"][name][idx]) < 1e-6 * max(1.0, abs(fd)), (cname, name, idx, fd)
