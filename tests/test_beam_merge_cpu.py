"""Host side of the batched beam search: the vectorised candidate merge must reproduce the reference's per-utterance
loop (beam_search.py:255-266, 294-329: concatenate, np.argpartition(-k)[-k:] in that order, back-pointers idx // k,
EOS retires a beam slot) step by step, including the order of the surviving rows."""
import numpy as np
import pytest

from e2e_asr_b200.beam_search import merge_candidates
from e2e_asr_b200.data_utils import EOS_ID


def loop_merge(utt, seqs, scores, k_u, final, idx_h, val_h, step, wip):
    """Per-utterance restatement of beam_search.py:294-329 on the batched row layout."""
    n = len(utt)
    new_utt, new_seqs, new_scores, new_tok, parents = [], [], [], [], []
    r = 0
    while r < n:
        u = utt[r]
        r1 = r
        while r1 < n and utt[r1] == u:
            r1 += 1
        k = k_u[u]
        if step == 0:
            cand_tokens = idx_h[r, :k]
            sel = np.arange(k)
            par = np.zeros(k, np.int64)
            model_scores = val_h[r, :k]
        else:
            all_scores = np.concatenate([val_h[i, :k] + scores[i] for i in range(r, r1)])
            cand_tokens = np.concatenate([idx_h[i, :k] for i in range(r, r1)])
            sel = np.argpartition(all_scores, -k)[-k:]
            par = sel // k
            model_scores = all_scores
        for j in range(k):
            pr = r + int(par[j])
            t_new = int(cand_tokens[sel[j]])
            seq = seqs[pr] + [t_new]
            # step 0 keeps the bare model score (beam_search.py:258-260); later steps add the penalty (:321-322)
            sc = float(model_scores[sel[j]]) + (wip * len(seq) if step else 0.0)
            if t_new == EOS_ID:
                final[u].append((seq, sc))
                k_u[u] -= 1
            else:
                new_utt.append(u); new_seqs.append(seq); new_scores.append(sc)
                new_tok.append(t_new); parents.append(pr)
        r = r1
    return new_utt, new_seqs, new_scores, new_tok, parents


@pytest.mark.parametrize("N,beam,V,seed,ties", [(1, 4, 9, 0, False), (7, 3, 6, 1, False), (40, 10, 12, 2, False),
                                               (16, 5, 7, 3, True)])
def test_vectorised_merge_equals_reference_loop(N, beam, V, seed, ties):
    rng = np.random.default_rng(seed)
    wip = 0.25
    # loop state
    utt_l, seqs_l, scores_l = list(range(N)), [[] for _ in range(N)], [0.0] * N
    k_l, final_l = [beam] * N, [[] for _ in range(N)]
    # vectorised state
    utt = np.arange(N)
    hist = np.zeros((N, 0), np.int64)
    scores = np.zeros(N)
    k_u = np.full(N, beam, np.int64)
    final = [[] for _ in range(N)]
    for step in range(12):
        n = len(utt)
        if n == 0:
            break
        val_h = np.log(rng.dirichlet(np.ones(V), size=n))
        if ties:
            val_h = np.round(val_h, 1)                   # many exactly equal candidate scores
        idx_h = np.argsort(-val_h, axis=1, kind="stable")[:, :beam].astype(np.int32)
        val_h = np.take_along_axis(val_h, idx_h.astype(np.int64), 1)
        nu, ns, nsc, nt, par = loop_merge(utt_l, seqs_l, scores_l, k_l, final_l, idx_h, val_h, step, wip)
        new, finished = merge_candidates(utt, scores, k_u, idx_h, val_h, step, wip)
        for u, pr, sc in finished:
            final[u].append((list(hist[pr]) + [EOS_ID], sc))
        hist = np.concatenate([hist[new["parent"]], new["tok"][:, None]], axis=1)
        assert list(new["utt"]) == nu and list(new["parent"]) == par and list(new["tok"]) == nt
        assert list(new["score"]) == nsc                 # bit-exact float64
        assert [list(h) for h in hist] == ns
        assert list(k_u) == k_l
        assert [[(list(map(int, s)), c) for s, c in f] for f in final] == final_l
        utt_l, seqs_l, scores_l = nu, ns, nsc
        utt, scores = new["utt"], new["score"]
    assert any(len(f) for f in final)                    # EOS really retired some hypotheses


def _loop_best_sequences(ph, th, fc, fs, fr, fsc, alive_h, score_h, beam):
    """The per-utterance Python loop `best_sequences` replaced (first maximum over finals then leftovers, one
    back-pointer walk per utterance)."""
    steps_done, R = th.shape
    N = R // beam

    def backtrack(t, row):
        seq = []
        while t >= 0:
            seq.append(int(th[t, row]))
            row = int(ph[t, row])
            t -= 1
        return seq[::-1]

    outs, scs = [], []
    for u in range(N):
        best, best_sc = None, None
        for f in range(int(fc[u])):
            i = u * beam + f
            if best is None or fsc[i] > best_sc:
                best, best_sc = (int(fs[i]) - 1, int(fr[i]), True), float(fsc[i])
        for slot in range(beam):
            row = u * beam + slot
            if alive_h[row] and (best is None or score_h[row] > best_sc):
                best, best_sc = (steps_done - 1, row, False), float(score_h[row])
        outs.append(np.asarray(backtrack(best[0], best[1]) + ([EOS_ID] if best[2] else []), np.int64))
        scs.append(best_sc)
    return outs, scs


@pytest.mark.parametrize("seed", range(6))
def test_best_sequences_matches_per_utterance_loop(seed):
    from e2e_asr_b200.beam_search import best_sequences
    rng = np.random.default_rng(seed)
    N, beam, S = 17, [1, 4, 10][seed % 3], [1, 9, 40][seed % 3]
    R = N * beam
    base = (np.arange(R) // beam * beam).astype(np.int32)
    ph = (rng.integers(0, beam, (S, R)) + base[None, :]).astype(np.int32)
    th = rng.integers(3, 50, (S, R)).astype(np.int32)
    fc = rng.integers(0, beam + 1, N).astype(np.int32)
    alive = rng.integers(0, 2, R).astype(np.int32)
    for u in range(N):                          # every utterance keeps at least one candidate
        if fc[u] == 0 and not alive[u * beam:(u + 1) * beam].any():
            alive[u * beam] = 1
    fs = rng.integers(0, S + 1, R).astype(np.int32)          # 0: EOS was the first token
    fr = (rng.integers(0, beam, R) + base).astype(np.int32)
    # coarse scores: ties between finals and leftovers must resolve to the FIRST maximum
    fsc = rng.integers(-3, 3, R).astype(np.float64)
    score = rng.integers(-3, 3, R).astype(np.float64)
    got, got_sc = best_sequences(ph, th, fc, fs, fr, fsc, alive, score, beam)
    want, want_sc = _loop_best_sequences(ph, th, fc, fs, fr, fsc, alive, score, beam)
    assert got_sc == want_sc
    for g, w in zip(got, want):
        assert g.dtype == np.int64 and np.array_equal(g, w)
