"""Variable store: the TF-1.x variable-scope mechanics the reference relies on,
reduced to what the hot path needs.

The reference creates its weights lazily through `tf.get_variable` inside nested
variable scopes and exchanges them by NAME (checkpoints, beam_search.py:56-98;
SURVEY.md Appendix B).  `VariableStore.get` keeps that contract: first use
creates the variable with the reference's initialiser, later uses return the same
tensor.  All variables live in ONE flat fp32 device buffer with a parallel flat
gradient buffer, so gradient clipping and the data-parallel allreduce are single
passes over contiguous memory.
"""
import math

import numpy as np
import torch


class VariableStore(object):
    def __init__(self, device="cuda", seed=4321, capacity=64 * 1024 * 1024):
        self.device = torch.device(device)
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.specs = {}          # name -> (offset, shape)
        self.vars = {}           # name -> leaf tensor (view into flat)
        self.order = []
        self.capacity = capacity
        self.used = 0
        self.flat = torch.zeros((capacity,), dtype=torch.float32, device=self.device)
        self.gflat = torch.zeros((capacity,), dtype=torch.float32, device=self.device)
        self.preloaded = {}

    # -- creation ---------------------------------------------------------
    def load(self, weights):
        """Provide values (numpy, keyed by TF variable name) used instead of the
        random initialisers when the variables are first requested; variables
        that already exist are overwritten.  `weights` may also be the prefix of a
        TF V2 checkpoint (tf_utils.py:66-90 restores by these names), read by
        tf_checkpoint.py; non-float entries (global_step, ...) are skipped."""
        if isinstance(weights, str):
            from .tf_checkpoint import read_checkpoint
            weights = {k: v for k, v in read_checkpoint(weights).items() if v.dtype.kind == "f"}
        for k, v in weights.items():
            self.preloaded[k] = np.asarray(v, np.float32)
            if k in self.vars:
                with torch.no_grad():
                    self.vars[k].copy_(torch.from_numpy(self.preloaded[k]).to(self.device))

    def _init_value(self, name, shape, init):
        if name in self.preloaded:
            v = self.preloaded[name]
            assert tuple(v.shape) == tuple(shape), (name, v.shape, shape)
            return v
        if init[0] == "uniform":
            return self.rng.uniform(-init[1], init[1], size=shape).astype(np.float32)
        if init[0] == "zeros":
            return np.zeros(shape, np.float32)
        if init[0] == "ones":
            return np.ones(shape, np.float32)
        if init[0] == "glorot":
            if len(shape) == 1:
                fan_in = fan_out = shape[0]
            else:
                recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                fan_in, fan_out = shape[-2] * recept, shape[-1] * recept
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return self.rng.uniform(-lim, lim, size=shape).astype(np.float32)
        raise ValueError(init)

    def get(self, name, shape, init=("glorot",)):
        """tf.get_variable with reuse: create on first request, then share."""
        shape = tuple(int(s) for s in shape)
        if name in self.vars:
            assert self.specs[name][1] == shape, (name, self.specs[name][1], shape)
            return self.vars[name]
        size = int(np.prod(shape))
        off = self.used
        if off + size > self.capacity:
            raise RuntimeError("VariableStore capacity exceeded (%d floats)" % self.capacity)
        self.used = off + (size + 3) // 4 * 4     # keep every variable 16-byte aligned
        self.specs[name] = (off, shape)
        self.order.append(name)
        val = torch.from_numpy(np.ascontiguousarray(self._init_value(name, shape, init)).reshape(-1))
        self.flat[off:off + size].copy_(val)
        # `.data` gives the view its own autograd version counter: creating a later
        # variable (an in-place write into `flat`) must not invalidate tensors that
        # an earlier op already saved for backward.
        v = self.flat.data[off:off + size].view(shape).requires_grad_(True)
        v.grad = self.gflat.data[off:off + size].view(shape)
        self.vars[name] = v
        return v

    def flat_params(self):
        return self.flat[:self.used]

    def flat_grads(self):
        return self.gflat[:self.used]

    # -- access -----------------------------------------------------------
    def __getitem__(self, name):
        return self.vars[name]

    def names(self):
        return list(self.order)

    def grad(self, name):
        o, shp = self.specs[name]
        return self.gflat[o:o + int(np.prod(shp))].view(shp)

    def zero_grad(self):
        self.gflat[:self.used].zero_()
        for n, v in self.vars.items():     # re-attach (autograd may have replaced .grad)
            o, shp = self.specs[n]
            v.grad = self.gflat.data[o:o + int(np.prod(shp))].view(shp)

    def save_checkpoint(self, prefix):
        """Writes the variables as a TF V2 checkpoint (`<prefix>.index`, `<prefix>.data-00000-of-00001`) under their
        TF names."""
        from .tf_checkpoint import write_checkpoint
        write_checkpoint(prefix, self.state_dict())

    def state_dict(self):
        return {n: self.vars[n].detach().cpu().numpy().copy() for n in self.order}

    def num_params(self):
        return sum(int(np.prod(s)) for _, s in self.specs.values())


_default = None


def default_store(device="cuda"):
    global _default
    if _default is None:
        _default = VariableStore(device)
    return _default


def reset_default_store():
    global _default
    _default = None
