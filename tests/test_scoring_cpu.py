"""Host-side scoring: word filtering against golden vectors produced by executing the reference's
own data_utils.get_relevant_words, id -> sentence conversion, and the word-error counter."""
import json
import os
import random

from e2e_asr_b200 import scoring
from e2e_asr_b200.data_utils import EOS_ID


def test_get_relevant_words_matches_reference_golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "relevant_words.json")
    for case in json.load(open(path)):
        words, rel = scoring.get_relevant_words(case["in"])
        assert words == case["words"] and rel == case["rel"], case["in"]


def test_wp_array_to_sent_cuts_at_eos_and_joins_pieces():
    vocab = [b"<pad>", b"<go>", b"<eos>", "▁he".encode("utf-8"), b"llo", "▁world".encode("utf-8"), b"!"]
    assert scoring.wp_array_to_sent([3, 4, 5, EOS_ID, 6, 6], vocab) == "hello world"
    assert scoring.wp_array_to_sent([3, 4], vocab, normalizer=str.upper) == "HELLO"
    assert scoring.wp_array_to_sent([EOS_ID, 3], vocab) == ""


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def test_word_errors_distance_and_operation_counts():
    assert scoring.word_errors([], []) == (0, 0, 0, 0)
    assert scoring.word_errors(["a", "b"], ["a", "b"]) == (0, 0, 0, 0)
    assert scoring.word_errors(["a"], ["a", "b", "c"]) == (2, 2, 0, 0)          # two reference words missing
    assert scoring.word_errors(["a", "x", "y"], ["a"]) == (2, 0, 2, 0)          # two extra hypothesis words
    assert scoring.word_errors(["a", "x", "c"], ["a", "b", "c"]) == (1, 0, 0, 1)
    rnd = random.Random(0)
    for _ in range(200):
        a = [rnd.choice("abcd") for _ in range(rnd.randint(0, 9))]
        b = [rnd.choice("abcd") for _ in range(rnd.randint(0, 9))]
        dist, ins, dele, subs = scoring.word_errors(a, b)
        assert dist == _levenshtein(a, b) == ins + dele + subs
        assert len(a) - dele + ins == len(b)


def test_wer_scorer_accumulates_like_the_reference_evaluator():
    sc = scoring.WerScorer()
    assert sc.add("uh the cat sat", "the cat sat") == 0                          # fillers are not scored
    assert sc.add("the dog sat down", "the cat sat") == 2
    assert (sc.total_errors, sc.total_words) == (2, 6) and abs(sc.wer - 2 / 6) < 1e-12
    assert (sc.ins_errs, sc.del_errs, sc.sub_errs) == (0, 1, 1)
