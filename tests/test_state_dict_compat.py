"""The product's PointRCNN exposes exactly the reference model's state-dict keys and shapes
(fixture generated from the reference definition by tools/make_statedict_fixture.py), so the
checkpoints the reference publishes (README.md:127-132) load by key."""
import json
import os

import torch

from conftest import load

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pointrcnn_state_dict.json")


def test_state_dict_matches_reference_definition():
    load("config").use_default_yaml("rcnn")
    torch.manual_seed(0)
    model = load("net.point_rcnn").PointRCNN(num_classes=2, use_xyz=True, mode="TEST")
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(GOLD) as f:
        ref = json.load(f)
    assert sorted(mine) == sorted(ref)
    assert mine == ref


def test_cfg_merge_semantics(tmp_path):
    cfgm = load("config")
    cfgm.reset_cfg()
    y = tmp_path / "c.yaml"
    y.write_text("RPN:\n    LOC_XZ_FINE: True\nCLS_MEAN_SIZE: [[1.5, 1.6, 3.9]]\nTEST:\n    RPN_NMS_THRESH: 0.8\n")
    cfgm.cfg_from_file(str(y))
    assert cfgm.cfg.RPN.LOC_XZ_FINE is True and cfgm.cfg.TEST.RPN_NMS_THRESH == 0.8
    assert cfgm.cfg.CLS_MEAN_SIZE.dtype.name == "float32"  # lists coerce to the ndarray dtype (config.py:205-206)
    import pytest
    y.write_text("NOT_A_KEY: 1\n")
    with pytest.raises(KeyError):
        cfgm.cfg_from_file(str(y))
    y.write_text("RPN:\n    NUM_POINTS: 'many'\n")
    with pytest.raises(ValueError):
        cfgm.cfg_from_file(str(y))
    cfgm.cfg_from_list(["RPN.NUM_POINTS", "32768", "TAG", "double"])
    assert cfgm.cfg.RPN.NUM_POINTS == 32768 and cfgm.cfg.TAG == "double"
    cfgm.reset_cfg()
