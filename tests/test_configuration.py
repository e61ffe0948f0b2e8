"""CPU: rectorch_b200.configuration against the reference's own tests (rectorch/tests/test_configuration.py)."""
import json

import pytest

from rectorch_b200.configuration import ConfigManager, DataConfig, ModelConfig

CFG_D = {"data_path": "path/to/data", "proc_path": "path/to/preproc/folder", "seed": 42, "threshold": 1,
         "separator": ",", "u_min": 5, "i_min": 2, "heldout": 100, "test_prop": 0.2, "topn": 1}
CFG_M = {"model": {"param1": 1, "param2": 2, "param3": 3.0, "param4": "four"}, "train": {"train_param1": 1},
         "test": {"test_param2": 2.0}, "sampler": {"sampler_param3": 3}}


def _write(tmp_path, name, cfg):
    p = str(tmp_path / name)
    json.dump(cfg, open(p, "w"))
    return p


def test_DataConfig(tmp_path):
    data_config = DataConfig(_write(tmp_path, "d.json", CFG_D))
    assert str(data_config) == data_config.__repr__()
    for k in CFG_D:
        assert hasattr(data_config, k)
        assert data_config[k] == CFG_D[k] and getattr(data_config, k) == CFG_D[k]
    assert data_config.header is None                 # DefaultMunch(None, ...): a missing key reads as None
    assert data_config == DataConfig(_write(tmp_path, "d2.json", CFG_D))


def test_ModelConfig(tmp_path):
    model_config = ModelConfig(_write(tmp_path, "m.json", CFG_M))
    assert str(model_config) == model_config.__repr__()
    for sec in ("model", "train", "test", "sampler"):
        assert hasattr(model_config, sec)
        for k in CFG_M[sec]:
            assert getattr(model_config, sec)[k] == CFG_M[sec][k]
            assert getattr(getattr(model_config, sec), k) == CFG_M[sec][k]


def test_ConfigManager(tmp_path):
    ConfigManager._instance = None
    with pytest.raises(Exception):
        ConfigManager.get()
    d, m = _write(tmp_path, "d.json", CFG_D), _write(tmp_path, "m.json", CFG_M)
    man = ConfigManager(d, m)
    assert ConfigManager.get() is man and ConfigManager(d, m) is man
    assert str(man) == repr(man)
    assert man.data_config.proc_path == CFG_D["proc_path"] and man.model_config.model.param4 == "four"
    ConfigManager._instance = None
