"""Hyper-parameter containers.

Mirrors the reference's `BaseParams` contract (reference base_params.py:10-28):
every component exposes `class_params()` returning a `Bunch` of defaults and
`get_updated_params(options)` that overrides a default only when the option has
exactly the same Python type (base_params.py:26) -- an `int` given for a `float`
default is silently ignored, as in the reference.

The reference depends on the third-party `bunch` package (not installed here);
`Bunch` below is a minimal attribute/key dual-access dict with the same surface.
"""


class Bunch(dict):
    """dict whose keys are also attributes (surface of `bunch.Bunch`)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError:
            raise AttributeError(key)

    def copy(self):
        return Bunch(self)

    def __deepcopy__(self, memo):
        import copy
        return Bunch((k, copy.deepcopy(v, memo)) for k, v in self.items())


class BaseParams(object):
    """Base class for dealing with parameters (reference base_params.py:10)."""

    @classmethod
    def class_params(cls):
        return Bunch()

    @classmethod
    def add_parse_options(cls, parser=None):
        pass

    @classmethod
    def get_updated_params(cls, options):
        params = cls.class_params()
        for attr in list(params.keys()):
            if attr in options:
                if type(params[attr]) == type(options[attr]):
                    params[attr] = options[attr]
        return params
