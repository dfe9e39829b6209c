"""tests/golden/reference_params.json: the reference's own `class_params()` dictionaries and argparse flags
(names, option strings, defaults), obtained by EXECUTING its classes with `tensorflow` / `bunch` stubbed (they are only
touched at call time).  Run in the build container only; the committed JSON is what tests/test_params_cpu.py uses."""
import argparse
import builtins
import importlib
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from e2e_asr_b200.base_params import Bunch  # noqa: E402


class _Any(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Any(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


class _TfFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Every `tensorflow...` module path resolves to a permissive stub."""

    def find_spec(self, fullname, path, target=None):
        if fullname == "tensorflow" or fullname.startswith("tensorflow."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Any(spec.name)

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _TfFinder())
    b = types.ModuleType("bunch")
    b.Bunch = Bunch
    sys.modules["bunch"] = b
    builtins.xrange = range
    sys.path.insert(0, "/root/reference")

    def plain(v):
        if isinstance(v, dict):
            return {k: plain(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [plain(x) for x in v]
        return v

    import inspect
    out = {}
    for mod, cls in [("encoder", "Encoder"), ("decoder", "Decoder"), ("attn_decoder", "AttnDecoder"),
                     ("seq2seq_model", "Seq2SeqModel"), ("beam_search", "BeamSearch"), ("losses", "LossUtils"),
                     ("basic_lstm", "BasicLSTM")]:
        C = getattr(importlib.import_module(mod), cls)
        entry = {"class_params": plain(dict(C.class_params()))} if hasattr(C, "class_params") else {}
        sigs = {}
        for meth in ("__init__", "__call__", "get_state", "get_batch", "cross_entropy_loss"):
            f = C.__dict__.get(meth)
            f = f.__func__ if isinstance(f, (staticmethod, classmethod)) else f
            if f is not None and inspect.isfunction(f):
                sigs[meth] = [[n, None if prm.default is inspect.Parameter.empty else repr(prm.default)]
                              for n, prm in inspect.signature(f).parameters.items()]
        entry["signatures"] = sigs
        if hasattr(C, "get_updated_params") and hasattr(C, "class_params"):
            # base_params.py:21-28: an option overrides a default only when it has exactly the default's Python type
            opts = {"hidden_size": 128, "out_prob": 1, "use_lstm": True, "hidden_size_dec": 64.0, "samp_prob": 0.25,
                    "beam_size": 7, "lm_weight": 1, "avg": False, "learning_rate": 3, "bogus": 5, "emb_size": 32}
            entry["updated"] = plain(dict(C.get_updated_params(opts)))
            entry["update_options"] = opts
        if hasattr(C, "add_parse_options"):
            p = argparse.ArgumentParser()
            C.add_parse_options(p)
            entry["flags"] = {a.dest: {"opts": a.option_strings, "default": a.default,
                                       "type": getattr(a.type, "__name__", None)}
                              for a in p._actions if a.dest != "help"}
        out[cls] = entry
    with open(os.path.join(HERE, "reference_params.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return out


if __name__ == "__main__":
    for k, v in main().items():
        print(k, len(v.get("class_params", {})), len(v.get("flags", {})), sorted(v["signatures"]))
