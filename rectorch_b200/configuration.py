"""``DataConfig`` (rectorch/configuration.py:26-46): the JSON data configuration the data reader takes
(SURVEY.md section 8f N3).  The reference builds it on ``munch.DefaultMunch(None, json)``; this is a dependency-free
equivalent with the behaviour the reader relies on: attribute and item access to the JSON keys, ``None`` for a
missing key, dict equality, the same ``str`` / ``repr``.  ``ModelConfig`` / ``ConfigManager`` are outside the hot
path and are not provided.
"""
import json
from os.path import exists

__all__ = ['DataConfig']


class _AttrDict(dict):
    """dict with attribute access; a missing key reads as ``None`` (``DefaultMunch(None, mapping)``)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.get(name, None)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        self.pop(name, None)


class DataConfig(_AttrDict):
    """``DataConfig(file_path)``: the JSON data configuration (keys such as ``data_path``, ``proc_path``, ``seed``,
    ``threshold``, ``separator``, ``header``, ``u_min``, ``i_min``, ``heldout``, ``test_prop``, ``topn``)."""

    def __init__(self, file_path):
        with open(file_path, "r") as fh:
            super(DataConfig, self).__init__(json.load(fh))

    def __str__(self):
        return "DataConfig(" + ", ".join(["%s=%s" % (k, self[k]) for k in self]) + ")"

    def __repr__(self):
        return str(self)
