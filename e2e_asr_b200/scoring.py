"""Host-side scoring either side of the hot path: turning decoded ids into words and
counting word errors as the reference's evaluator does (eval_model.py:84-98, 218-229, 250-258; data_utils.py:17-33).
Pure Python on small lists; nothing here touches the GPU."""
from .data_utils import EOS_ID

# data_utils.py:17-18
IGNORED_WORDS = ["[noise]", "[laughter]", "[vocalized-noise]", "uh", "um", "eh", "mm", "hm",
                 "ah", "huh", "ha", "er", "oof", "hee", "ach", "eee", "ew"]


def get_relevant_words(char_str):
    """data_utils.get_relevant_words (data_utils.py:20-33): "<sp>" -> space, split; drop the ignored filler words and
    partial words (trailing "-").  Returns (all words, relevant words)."""
    char_str = char_str.replace("<sp>", " ")
    words = char_str.split()
    rel_words = []
    for word in words:
        if word in IGNORED_WORDS:
            continue
        elif len(word) > 0 and word[-1] == "-":
            continue
        else:
            rel_words.append(word)
    return words, rel_words


def wp_array_to_sent(wp_array, reverse_char_vocab, normalizer=None):
    """eval_model.wp_array_to_sent (eval_model.py:250-258): cut at the first EOS, join the pieces, the word-piece
    marker U+2581 becomes a space, strip, normalise."""
    ids = [int(i) for i in wp_array]
    if EOS_ID in ids:
        ids = ids[:ids.index(EOS_ID)]
    pieces = [p.decode("utf-8") if isinstance(p, bytes) else str(p) for p in (reverse_char_vocab[i] for i in ids)]
    sent = "".join(pieces).replace("▁", " ").strip()
    return normalizer(sent) if normalizer is not None else sent


def word_errors(hyp_words, ref_words):
    """Levenshtein alignment turning `hyp_words` into `ref_words` (eval_model.py:219: ed(decoded_words, gold_words)).
    Returns (distance, insertions, deletions, substitutions) with the reference's opcode meaning: "insert" = words of
    the reference missing from the hypothesis, "delete" = extra hypothesis words, "replace" = substitutions.  The
    distance is unique; the split depends on which optimal alignment is taken: like the third-party `edit_distance`
    package the reference imports (SequenceMatcher with lowest_cost_action: equal / replace when the diagonal is
    cheapest, else insert, else delete -- restated from the published source, the package is absent here, so the
    tie-breaking is unpinned; the total is not affected)."""
    n, m = len(hyp_words), len(ref_words)
    d = [[0] * (m + 1) for _ in range(n + 1)]
    for i in range(1, n + 1):
        d[i][0] = i
    for j in range(1, m + 1):
        d[0][j] = j
    for i in range(1, n + 1):
        hi = hyp_words[i - 1]
        row, prev = d[i], d[i - 1]
        for j in range(1, m + 1):
            sub = prev[j - 1] + (0 if hi == ref_words[j - 1] else 1)
            row[j] = min(sub, prev[j] + 1, row[j - 1] + 1)
    i, j, ins, dele, subs = n, m, 0, 0, 0
    while i > 0 or j > 0:
        if i > 0 and j > 0 and d[i][j] == d[i - 1][j - 1] + (0 if hyp_words[i - 1] == ref_words[j - 1] else 1):
            subs += hyp_words[i - 1] != ref_words[j - 1]
            i, j = i - 1, j - 1
        elif j > 0 and d[i][j] == d[i][j - 1] + 1:
            ins += 1
            j -= 1
        else:
            dele += 1
            i -= 1
    return d[n][m], ins, dele, subs


class WerScorer(object):
    """Accumulates word errors over utterances as eval_model.py:94-98 / 218-229 do: errors = edit distance between
    the relevant decoded words and the relevant gold words, words = number of relevant gold words."""

    def __init__(self):
        self.total_errors = self.total_words = self.ins_errs = self.del_errs = self.sub_errs = 0

    def add(self, decoded_sentence, gold_sentence):
        _, decoded_words = get_relevant_words(decoded_sentence)
        _, gold_words = get_relevant_words(gold_sentence)
        dist, ins, dele, subs = word_errors(decoded_words, gold_words)
        self.total_errors += dist
        self.total_words += len(gold_words)
        self.ins_errs += ins
        self.del_errs += dele
        self.sub_errs += subs
        return dist

    def add_ids(self, decoded_ids, gold_ids, reverse_char_vocab, normalizer=None):
        return self.add(wp_array_to_sent(decoded_ids, reverse_char_vocab, normalizer),
                        wp_array_to_sent(gold_ids, reverse_char_vocab, normalizer))

    @property
    def wer(self):
        return self.total_errors / float(max(self.total_words, 1))
