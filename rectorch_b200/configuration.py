"""Configuration objects (rectorch/configuration.py:11-160): ``DataConfig``, ``ModelConfig``, ``ConfigManager``.

The reference builds them on ``munch.DefaultMunch(None, json)``; this is a dependency-free equivalent with the
behaviour its callers rely on: attribute and item access to the JSON keys, ``None`` for a missing key, dict
equality, the same ``str`` / ``repr`` and the singleton protocol of ``ConfigManager``.
"""
import json
from os.path import exists

__all__ = ['DataConfig', 'ModelConfig', 'ConfigManager']


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


class Singleton(type):
    """Metaclass that lets clients access one unique instance (configuration.py:13-23)."""

    def __init__(cls, name, bases, attrs):
        super().__init__(name, bases, attrs)
        cls._instance = None

    def __call__(cls, *args, **kwargs):
        if cls._instance is None:
            cls._instance = super().__call__(*args, **kwargs)
        return cls._instance


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


class ModelConfig():
    """``ModelConfig(file_path)``: the ``model`` / ``train`` / ``test`` / ``sampler`` sections of the JSON model
    configuration as attribute dictionaries (configuration.py:49-91)."""

    def __init__(self, file_path):
        with open(file_path, "r") as fh:
            json_cfg = json.load(fh)
        self.model = _AttrDict(json_cfg["model"])
        self.train = _AttrDict(json_cfg["train"])
        self.test = _AttrDict(json_cfg["test"])
        self.sampler = _AttrDict(json_cfg["sampler"])

    def __str__(self):
        return "ModelConfig(model={}, train={}, test={}, sampler={}".format(
            self.model, self.train, self.test, self.sampler)

    def __repr__(self):
        return str(self)


class ConfigManager(metaclass=Singleton):
    """Singleton wrapper of both configurations (configuration.py:94-160): ``ConfigManager(data_config_path,
    model_config_path)`` creates it, ``ConfigManager.get()`` returns it or raises if it does not exist yet."""

    @classmethod
    def get(cls):
        if cls._instance:
            return cls._instance
        raise Exception("Singleton object not instantiated!")

    def __init__(self, data_config_path, model_config_path):
        assert exists(data_config_path), "Data config file does not exist."
        assert exists(model_config_path), "Model config file does not exist."
        self.data_config = DataConfig(data_config_path)
        self.model_config = ModelConfig(model_config_path)

    def __str__(self):
        return "ConfigManager(data_config=%s, model_config=%s" % (self.data_config, self.model_config)

    def __repr__(self):
        return str(self)
