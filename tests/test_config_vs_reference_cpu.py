"""config.py against the REFERENCE lib/config.py (imported live from /root/reference, build container only): the
defaults, the merge of tools/cfgs/default.yaml, cfg_from_list overrides -- every key, value and dtype equal; and
use_default_yaml() (the tree-less shortcut the bench and tests use) equals default.yaml + eval_mode on every key
that is not training-only."""
import os
import sys

import numpy as np
import pytest

from conftest import load, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import refnet_cpu as rn                      # noqa: E402

pytestmark = pytest.mark.skipif(not rn.available(), reason="reference tree not present")

TRAINING_ONLY = ("TRAIN.", "AUG_", "GT_AUG_", "RCNN.HARD_BG_RATIO")


def _flat(d, pre=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, pre + k + "."))
        else:
            out[pre + k] = v
    return out


def _differences(a, b):
    bad = []
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            bad.append((k, "missing on one side"))
            continue
        x, y = a[k], b[k]
        if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
            if not (isinstance(x, np.ndarray) and isinstance(y, np.ndarray) and x.dtype == y.dtype and np.array_equal(x, y)):
                bad.append((k, x, y))
        elif type(x) is not type(y) or x != y:
            bad.append((k, x, y))
    return bad


def test_config_equals_reference_module():
    cfgm = load("config")
    yaml_file = os.path.join(rn.REF, "tools", "cfgs", "default.yaml")
    overrides = ["RPN.LOC_SCOPE", "4.0", "TEST.RPN_POST_NMS_TOP_N", "50", "RCNN.USE_DEPTH", "False"]
    with rn.reference_imports():
        import lib.config as rc
        ref_default = _flat(rc.cfg)
        rc.cfg_from_file(yaml_file)
        ref_yaml = _flat(rc.cfg)
        rc.cfg_from_list(overrides)
        ref_list = _flat(rc.cfg)
    try:
        cfgm.reset_cfg()
        assert _differences(ref_default, _flat(cfgm.cfg)) == []
        cfgm.cfg_from_file(yaml_file)
        assert _differences(ref_yaml, _flat(cfgm.cfg)) == []
        cfgm.cfg_from_list(overrides)
        assert _differences(ref_list, _flat(cfgm.cfg)) == []
        cfgm.use_default_yaml("rcnn")
        want = dict(ref_yaml, **{"RCNN.ENABLED": True, "RPN.ENABLED": True, "RPN.FIXED": True})
        diff = [d for d in _differences(want, _flat(cfgm.cfg)) if not d[0].startswith(TRAINING_ONLY)]
        assert diff == []
    finally:
        cfgm.use_default_yaml("rcnn")
