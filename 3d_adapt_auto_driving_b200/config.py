"""Global `cfg` for the PointRCNN path -- mirror of pointrcnn/lib/config.py.

Same public surface: cfg, cfg_from_file, cfg_from_list, save_config_to_file, with the
reference's strict merge semantics (unknown key -> KeyError, type mismatch -> ValueError,
ndarray-typed keys coerce lists; config.py:193-220).  Differences that the drop-in has to
absorb: no `easydict` dependency (AttrDict below), yaml.safe_load instead of
the Loader-less yaml.load that PyYAML 6 rejects (config.py:188).

Only the keys are restated, the values below are the reference's defaults (config.py:8-181);
`use_default_yaml()` applies the overrides of tools/cfgs/default.yaml that matter at
inference so that the benchmark and the tests run the published architecture without the
reference tree being present.
"""
import numpy as np


class AttrDict(dict):
    """dict with attribute access, nested dicts converted on assignment (EasyDict subset)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


edict = AttrDict


def _defaults():
    f32 = np.float32
    return {
        "TAG": "default", "CLASSES": "Car", "INCLUDE_SIMILAR_TYPE": False,
        "AUG_DATA": True, "AUG_METHOD_LIST": ["rotation", "scaling", "flip"], "SCALE_MIN_MAX_RANGE": [0.95, 1.05],
        "AUG_METHOD_PROB": [0.5, 0.5, 0.5], "AUG_ROT_RANGE": 18,
        "GT_AUG_ENABLED": False, "GT_EXTRA_NUM": 15, "GT_AUG_RAND_NUM": False, "GT_AUG_APPLY_PROB": 0.75,
        "GT_AUG_HARD_RATIO": 0.6,
        "PC_REDUCE_BY_RANGE": True,
        "PC_AREA_SCOPE": np.array([[-40, 40], [-1, 3], [0, 70.4]]),
        "CLS_MEAN_SIZE": np.array([[1.52, 1.63, 3.88]], dtype=f32),
        "RPN": {
            "ENABLED": True, "FIXED": False, "USE_INTENSITY": True,
            "LOC_XZ_FINE": False, "LOC_SCOPE": 3.0, "LOC_BIN_SIZE": 0.5, "NUM_HEAD_BIN": 12,
            "BACKBONE": "pointnet2_msg", "USE_BN": True, "NUM_POINTS": 16384,
            "SA_CONFIG": {
                "NPOINTS": [4096, 1024, 256, 64],
                "RADIUS": [[0.1, 0.5], [0.5, 1.0], [1.0, 2.0], [2.0, 4.0]],
                "NSAMPLE": [[16, 32], [16, 32], [16, 32], [16, 32]],
                "MLPS": [[[16, 16, 32], [32, 32, 64]], [[64, 64, 128], [64, 96, 128]],
                         [[128, 196, 256], [128, 196, 256]], [[256, 256, 512], [256, 384, 512]]],
            },
            "FP_MLPS": [[128, 128], [256, 256], [512, 512], [512, 512]],
            "CLS_FC": [128], "REG_FC": [128], "DP_RATIO": 0.5,
            "LOSS_CLS": "DiceLoss", "FG_WEIGHT": 15, "FOCAL_ALPHA": [0.25, 0.75], "FOCAL_GAMMA": 2.0,
            "REG_LOSS_WEIGHT": [1.0, 1.0, 1.0, 1.0], "LOSS_WEIGHT": [1.0, 1.0], "NMS_TYPE": "normal",
            "SCORE_THRESH": 0.3,
        },
        "RCNN": {
            "ENABLED": False, "USE_RPN_FEATURES": True, "USE_MASK": True, "MASK_TYPE": "seg", "USE_INTENSITY": False,
            "USE_DEPTH": True, "USE_SEG_SCORE": False, "ROI_SAMPLE_JIT": False, "ROI_FG_AUG_TIMES": 10,
            "REG_AUG_METHOD": "multiple", "POOL_EXTRA_WIDTH": 1.0,
            "LOC_SCOPE": 1.5, "LOC_BIN_SIZE": 0.5, "NUM_HEAD_BIN": 9, "LOC_Y_BY_BIN": False, "LOC_Y_SCOPE": 0.5,
            "LOC_Y_BIN_SIZE": 0.25, "SIZE_RES_ON_ROI": False,
            "USE_BN": False, "DP_RATIO": 0.0, "BACKBONE": "pointnet", "XYZ_UP_LAYER": [128, 128], "NUM_POINTS": 512,
            "SA_CONFIG": {"NPOINTS": [128, 32, -1], "RADIUS": [0.2, 0.4, 100], "NSAMPLE": [64, 64, 64],
                          "MLPS": [[128, 128, 128], [128, 128, 256], [256, 256, 512]]},
            "CLS_FC": [256, 256], "REG_FC": [256, 256],
            "LOSS_CLS": "BinaryCrossEntropy", "FOCAL_ALPHA": [0.25, 0.75], "FOCAL_GAMMA": 2.0,
            "CLS_WEIGHT": np.array([1.0, 1.0, 1.0], dtype=f32), "CLS_FG_THRESH": 0.6, "CLS_BG_THRESH": 0.45,
            "CLS_BG_THRESH_LO": 0.05, "REG_FG_THRESH": 0.55, "FG_RATIO": 0.5, "ROI_PER_IMAGE": 64,
            "HARD_BG_RATIO": 0.6, "SCORE_THRESH": 0.3, "NMS_THRESH": 0.1,
        },
        "TRAIN": {
            "SPLIT": "train", "VAL_SPLIT": "smallval", "LR": 0.002, "LR_CLIP": 0.00001, "LR_DECAY": 0.5,
            "DECAY_STEP_LIST": [50, 100, 150, 200, 250, 300], "LR_WARMUP": False, "WARMUP_MIN": 0.0002,
            "WARMUP_EPOCH": 5, "BN_MOMENTUM": 0.9, "BN_DECAY": 0.5, "BNM_CLIP": 0.01,
            "BN_DECAY_STEP_LIST": [50, 100, 150, 200, 250, 300], "OPTIMIZER": "adam", "WEIGHT_DECAY": 0.0,
            "MOMENTUM": 0.9, "MOMS": [0.95, 0.85], "DIV_FACTOR": 10.0, "PCT_START": 0.4, "GRAD_NORM_CLIP": 1.0,
            "RPN_PRE_NMS_TOP_N": 12000, "RPN_POST_NMS_TOP_N": 2048, "RPN_NMS_THRESH": 0.85,
            "RPN_DISTANCE_BASED_PROPOSE": True,
        },
        "TEST": {"SPLIT": "val", "RPN_PRE_NMS_TOP_N": 9000, "RPN_POST_NMS_TOP_N": 300, "RPN_NMS_THRESH": 0.7,
                 "RPN_DISTANCE_BASED_PROPOSE": True},
    }


__C = AttrDict(_defaults())
cfg = __C

# what tools/cfgs/default.yaml changes relative to the defaults above, restricted to the keys
# read at inference (yaml lines 17-18, 24-52, 85-109, 132-134, 162-166)
_DEFAULT_YAML_INFERENCE = {
    "INCLUDE_SIMILAR_TYPE": True,
    "CLS_MEAN_SIZE": [[1.52563191462, 1.62856739989, 3.88311640418]],
    "RPN": {"USE_INTENSITY": False, "LOC_XZ_FINE": True, "LOSS_CLS": "SigmoidFocalLoss"},
    "RCNN": {"ENABLED": True, "ROI_SAMPLE_JIT": True},
    "TEST": {"RPN_POST_NMS_TOP_N": 100, "RPN_NMS_THRESH": 0.8},
}


def reset_cfg():
    """Back to the reference defaults (tests call this; the reference has a single global)."""
    __C.clear()
    for k, v in _defaults().items():
        __C[k] = v


def use_default_yaml(eval_mode="rcnn"):
    """reset + tools/cfgs/default.yaml (inference keys) + what eval_rcnn.py sets for --eval_mode
    (eval_rcnn.py:877-894: 'rcnn' => RCNN.ENABLED, RPN.ENABLED = RPN.FIXED = True)."""
    reset_cfg()
    _merge_a_into_b(AttrDict(_DEFAULT_YAML_INFERENCE), __C)
    if eval_mode == "rpn":
        __C.RPN.ENABLED, __C.RCNN.ENABLED = True, False
    elif eval_mode == "rcnn":
        __C.RCNN.ENABLED = True
        __C.RPN.ENABLED = __C.RPN.FIXED = True
    return __C


def cfg_from_file(filename):
    """Merge a yaml file into cfg (config.py:184-190)."""
    import yaml
    with open(filename, "r") as f:
        yaml_cfg = AttrDict(yaml.safe_load(f))
    _merge_a_into_b(yaml_cfg, __C)


def _merge_a_into_b(a, b):
    if not isinstance(a, AttrDict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError("{} is not a valid config key".format(k))
        old_type = type(b[k])
        if old_type is not type(v):
            if isinstance(b[k], np.ndarray):
                v = np.array(v, dtype=b[k].dtype)
            else:
                raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(b[k]), type(v), k))
        if isinstance(v, AttrDict):
            try:
                _merge_a_into_b(a[k], b[k])
            except Exception:
                print("Error under config key: {}".format(k))
                raise
        else:
            b[k] = v


def cfg_from_list(cfg_list):
    """`--set KEY VALUE ...` overrides (config.py:223-242)."""
    from ast import literal_eval
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        keys = k.split(".")
        d = __C
        for sub in keys[:-1]:
            assert sub in d
            d = d[sub]
        sub = keys[-1]
        assert sub in d
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        assert type(value) == type(d[sub]), "type {} does not match original type {}".format(type(value), type(d[sub]))
        d[sub] = value


def save_config_to_file(cfg, pre="cfg", logger=None):
    out = logger.info if logger is not None else print
    for key, val in cfg.items():
        if isinstance(val, AttrDict):
            out("\n%s.%s = edict()" % (pre, key))
            save_config_to_file(val, pre=pre + "." + key, logger=logger)
        else:
            out("%s.%s: %s" % (pre, key, val))
