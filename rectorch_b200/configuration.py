"""Data configuration object (rectorch/configuration.py:26-46).

Only ``DataConfig`` is mirrored: it is what ``DataReader`` / ``DatasetManager`` take.  The reference builds it
on ``munch.DefaultMunch(None, json)``; this is a dependency-free equivalent with the same behaviour the data
path relies on: attribute and item access to the JSON keys, ``None`` for a missing key, dict equality.
"""
import json

__all__ = ['DataConfig']


class DataConfig(dict):
    """``DataConfig(file_path)``: the JSON data configuration (keys such as ``proc_path``, ``topn``)."""

    def __init__(self, file_path):
        with open(file_path, "r") as fh:
            super(DataConfig, self).__init__(json.load(fh))

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.get(name, None)

    def __setattr__(self, name, value):
        self[name] = value

    def __str__(self):
        return "DataConfig(" + ", ".join(["%s=%s" % (k, self[k]) for k in self]) + ")"

    def __repr__(self):
        return str(self)
