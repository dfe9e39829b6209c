"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from e2e_asr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return entry.build()


def test_exports_every_declared_symbol(libpath):
    header = open(os.path.join(ROOT, "include", "e2e_asr_b200.h")).read()
    declared = set(re.findall(r"\b(e2e_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    handle = ctypes.CDLL(libpath)
    for sym in sorted(declared):
        assert hasattr(handle, sym), sym
    assert declared == set(_lib.exported_symbols())


def test_struct_layout_matches_header():
    # 9 ints (padded to 40 bytes) + 18 pointers; bwd adds 8 pointers
    assert ctypes.sizeof(_lib.DecLoopFwdArgs) == 40 + 18 * 8
    assert ctypes.sizeof(_lib.DecLoopBwdArgs) == 40 + 27 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(libpath):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _lib.call("e2e_mean", 1, None, None)
    assert _lib.lib().e2e_sm_count() == -1


def test_binding_signatures_match_the_header():
    """Every ctypes signature string in _lib._SIGS has one letter per parameter of the header's declaration, and the
    letter's class (pointer / integer / float / size) matches the declared C type."""
    header = open(os.path.join(ROOT, "include", "e2e_asr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    decls = dict(re.findall(r"\bint\s+(e2e_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    kinds = {"p": "ptr", "i": "int", "I": "uint", "l": "longlong", "z": "size", "f": "float", "d": "double",
             "Q": "ulonglong"}

    def classify(param):
        p = " ".join(param.split())
        if "*" in p:
            return "ptr"
        for key, pat in (("ulonglong", r"\bunsigned long long\b"), ("longlong", r"\blong long\b"),
                         ("size", r"\bsize_t\b"), ("uint", r"\bunsigned\b"), ("float", r"\bfloat\b"),
                         ("double", r"\bdouble\b"), ("int", r"\bint\b")):
            if re.search(pat, p):
                return key
        raise AssertionError("unclassified parameter %r" % param)

    checked = 0
    for name, sig in _lib._SIGS.items():
        assert name in decls, name
        params = [p for p in decls[name].split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(sig), (name, len(params), sig)
        for letter, param in zip(sig, params):
            assert kinds[letter] == classify(param), (name, letter, param.strip())
        checked += 1
    assert checked >= 40


def test_every_entry_point_is_documented_in_integration_md():
    header = open(os.path.join(ROOT, "include", "e2e_asr_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(e2e_[a-z0-9_]+)\s*\(", header))
    assert not [f for f in sorted(declared) if f not in doc]
