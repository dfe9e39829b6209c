"""Drop-in boundary: every `class_params()` key / default and every argparse flag (option strings,
destination, default, type) of Encoder, Decoder, AttnDecoder, Seq2SeqModel and BeamSearch equals the reference's own --
golden values produced by executing the reference classes (tests/golden/gen_params_golden.py)."""
import argparse
import json
import os

import pytest

import e2e_asr_b200 as pkg

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_params.json")))
# the one deliberate deviation: the reference's -lm_path flag defaults to an absolute path on its author's cluster
# (beam_search.py:345-346); here it defaults to "" (= no separate LM checkpoint), like class_params()["lm_path"]
SITE_SPECIFIC_DEFAULTS = {("BeamSearch", "lm_path")}
# builder additions on top of the reference's keys (documented in INTEGRATION.md)
EXTRA_KEYS = {"Seq2SeqModel": {"ctc_tasks", "apply_updates", "dropout_seed", "tf_indexed_slices_norm",
                               "overlap_weight_grads"}}


def plain(v):
    if isinstance(v, dict):
        return {k: plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    return v


@pytest.mark.parametrize("cls", sorted(GOLD))
def test_call_signatures_extend_the_reference(cls):
    """Constructor / call / helper signatures: the reference's parameters come first, same names, same defaults; only
    optional keyword parameters (variables=, device=, reducer=, task=) are appended."""
    import inspect
    C = getattr(pkg, cls)
    for meth, ref_params in GOLD[cls]["signatures"].items():
        f = C.__dict__.get(meth)
        if f is None:                      # inherited (e.g. AttnDecoder.get_state from Decoder)
            f = getattr(C, meth)
        f = f.__func__ if isinstance(f, (staticmethod, classmethod)) else f
        ours = [[n, None if prm.default is inspect.Parameter.empty else repr(prm.default)]
                for n, prm in inspect.signature(f).parameters.items()]
        assert ours[:len(ref_params)] == ref_params, (cls, meth, ours, ref_params)
        assert all(d is not None for _, d in ours[len(ref_params):]), (cls, meth, ours)


@pytest.mark.parametrize("cls", sorted(c for c in GOLD if "class_params" in GOLD[c]))
def test_class_params_equal_the_reference(cls):
    ours = plain(dict(getattr(pkg, cls).class_params()))
    ref = GOLD[cls]["class_params"]
    extra = set(ours) - set(ref)
    assert extra <= EXTRA_KEYS.get(cls, set()), extra
    for k, v in ref.items():
        assert k in ours, (cls, k)
        assert ours[k] == v, (cls, k, ours[k], v)


@pytest.mark.parametrize("cls", sorted(c for c in GOLD if "flags" in GOLD[c]))
def test_argparse_flags_equal_the_reference(cls):
    p = argparse.ArgumentParser()
    getattr(pkg, cls).add_parse_options(p)
    ours = {a.dest: a for a in p._actions if a.dest != "help"}
    ref = GOLD[cls]["flags"]
    assert set(ours) == set(ref), (sorted(ours), sorted(ref))
    for dest, spec in ref.items():
        a = ours[dest]
        assert a.option_strings == spec["opts"], (cls, dest)
        if (cls, dest) in SITE_SPECIFIC_DEFAULTS:
            assert a.default == "" and spec["default"].startswith("/")
        else:
            assert a.default == spec["default"], (cls, dest, a.default, spec["default"])
        assert getattr(a.type, "__name__", None) == spec["type"], (cls, dest)


@pytest.mark.parametrize("cls", sorted(c for c in GOLD if "updated" in GOLD[c]))
def test_get_updated_params_type_rule_equals_the_reference(cls):
    """BaseParams.get_updated_params (base_params.py:21-28): an option replaces a default only if it has exactly the
    default's Python type (an int offered for a float default is ignored), unknown options are dropped."""
    ours = plain(dict(getattr(pkg, cls).get_updated_params(GOLD[cls]["update_options"])))
    ref = GOLD[cls]["updated"]
    for k, v in ref.items():
        assert ours[k] == v and type(ours[k]) is type(v), (cls, k, ours[k], v)
    assert "bogus" not in ours
