"""CfgNode (yacs-like) + detectron2 defaults for every key the reference reads + @configurable."""
import ast
import copy
import functools
import inspect
import os

import yaml


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass

    def dump(self):
        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v)
                    for k, v in n.items()}
        return yaml.safe_dump(plain(self))

    @staticmethod
    def _load_yaml_with_base(path):
        with open(path) as f:
            cfg = yaml.safe_load(f) or {}
        base = cfg.pop("_BASE_", None)
        if base is not None:
            if not os.path.isabs(base):
                base = os.path.join(os.path.dirname(path), base)
            b = CfgNode._load_yaml_with_base(base)
            CfgNode._merge_dict(cfg, b)
            return b
        return cfg

    @staticmethod
    def _merge_dict(a, b):
        for k, v in a.items():
            if isinstance(v, dict) and isinstance(b.get(k), dict):
                CfgNode._merge_dict(v, b[k])
            else:
                b[k] = v

    @staticmethod
    def _coerce(v, like):
        if isinstance(v, str) and not isinstance(like, str):
            try:
                v = ast.literal_eval(v)
            except Exception:
                pass
        if isinstance(like, tuple) and isinstance(v, list):
            v = tuple(v)
        if isinstance(like, list) and isinstance(v, tuple):
            v = list(v)
        if isinstance(like, float) and isinstance(v, int):
            v = float(v)
        return v

    def _merge(self, other, path=""):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self:
                    self[k] = CfgNode()
                self[k]._merge(v, path + k + ".")
            else:
                if k not in self:
                    raise KeyError(f"Non-existent config key: {path}{k}")
                self[k] = self._coerce(v, self[k])

    def merge_from_file(self, path, allow_unsafe=True):
        self._merge(self._load_yaml_with_base(path))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent config key: {k}")
            node[parts[-1]] = self._coerce(v, node[parts[-1]])


def get_cfg():
    C = CfgNode
    _C = C()
    _C.VERSION = 2
    _C.MODEL = C({
        "LOAD_PROPOSALS": False, "MASK_ON": False, "KEYPOINT_ON": False, "DEVICE": "cuda",
        "META_ARCHITECTURE": "GeneralizedRCNN", "WEIGHTS": "",
        "PIXEL_MEAN": [103.530, 116.280, 123.675], "PIXEL_STD": [1.0, 1.0, 1.0],
        "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
        "FPN": {"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
        "PROPOSAL_GENERATOR": {"NAME": "RPN", "MIN_SIZE": 0},
        "ANCHOR_GENERATOR": {"NAME": "DefaultAnchorGenerator", "SIZES": [[32, 64, 128, 256, 512]],
                             "ASPECT_RATIOS": [[0.5, 1.0, 2.0]], "ANGLES": [[-90, 0, 90]], "OFFSET": 0.0},
        "RPN": {"HEAD_NAME": "StandardRPNHead", "IN_FEATURES": ["res4"], "BOUNDARY_THRESH": -1,
                "IOU_THRESHOLDS": [0.3, 0.7], "IOU_LABELS": [0, -1, 1], "BATCH_SIZE_PER_IMAGE": 256,
                "POSITIVE_FRACTION": 0.5, "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                "BBOX_REG_WEIGHTS": (1.0, 1.0, 1.0, 1.0), "SMOOTH_L1_BETA": 0.0, "LOSS_WEIGHT": 1.0,
                "PRE_NMS_TOPK_TRAIN": 12000, "PRE_NMS_TOPK_TEST": 6000, "POST_NMS_TOPK_TRAIN": 2000,
                "POST_NMS_TOPK_TEST": 1000, "NMS_THRESH": 0.7, "CONV_DIMS": [-1]},
        "ROI_HEADS": {"NAME": "Res5ROIHeads", "NUM_CLASSES": 80, "IN_FEATURES": ["res4"],
                      "IOU_THRESHOLDS": [0.5], "IOU_LABELS": [0, 1], "BATCH_SIZE_PER_IMAGE": 512,
                      "POSITIVE_FRACTION": 0.25, "SCORE_THRESH_TEST": 0.05, "NMS_THRESH_TEST": 0.5,
                      "PROPOSAL_APPEND_GT": True},
        "ROI_BOX_HEAD": {"NAME": "", "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                         "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0), "SMOOTH_L1_BETA": 0.0,
                         "POOLER_RESOLUTION": 14, "POOLER_SAMPLING_RATIO": 0, "POOLER_TYPE": "ROIAlignV2",
                         "NUM_FC": 0, "FC_DIM": 1024, "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "",
                         "CLS_AGNOSTIC_BBOX_REG": False, "TRAIN_ON_PRED_BOXES": False,
                         "USE_FED_LOSS": False, "USE_SIGMOID_CE": False, "FED_LOSS_FREQ_WEIGHT_POWER": 0.5,
                         "FED_LOSS_NUM_CLASSES": 50},
        "RESNETS": {"DEPTH": 50, "OUT_FEATURES": ["res4"], "NUM_GROUPS": 1, "NORM": "FrozenBN",
                    "WIDTH_PER_GROUP": 64, "STRIDE_IN_1X1": True, "RES5_DILATION": 1,
                    "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64},
    })
    _C.INPUT = C({"MIN_SIZE_TRAIN": (800,), "MIN_SIZE_TRAIN_SAMPLING": "choice", "MAX_SIZE_TRAIN": 1333,
                  "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333, "RANDOM_FLIP": "horizontal", "FORMAT": "BGR",
                  "MASK_FORMAT": "polygon", "CROP": {"ENABLED": False, "TYPE": "relative_range", "SIZE": [0.9, 0.9]}})
    _C.DATASETS = C({"TRAIN": (), "TEST": (), "PROPOSAL_FILES_TRAIN": (), "PROPOSAL_FILES_TEST": (),
                     "PRECOMPUTED_PROPOSAL_TOPK_TRAIN": 2000, "PRECOMPUTED_PROPOSAL_TOPK_TEST": 1000})
    _C.DATALOADER = C({"NUM_WORKERS": 4, "ASPECT_RATIO_GROUPING": True, "SAMPLER_TRAIN": "TrainingSampler",
                       "REPEAT_THRESHOLD": 0.0, "FILTER_EMPTY_ANNOTATIONS": True})
    _C.SOLVER = C({"LR_SCHEDULER_NAME": "WarmupMultiStepLR", "MAX_ITER": 40000, "BASE_LR": 0.001,
                   "MOMENTUM": 0.9, "NESTEROV": False, "WEIGHT_DECAY": 0.0001, "WEIGHT_DECAY_NORM": 0.0,
                   "GAMMA": 0.1, "STEPS": (30000,), "WARMUP_FACTOR": 1.0 / 1000, "WARMUP_ITERS": 1000,
                   "WARMUP_METHOD": "linear", "CHECKPOINT_PERIOD": 5000, "IMS_PER_BATCH": 16,
                   "REFERENCE_WORLD_SIZE": 0, "BIAS_LR_FACTOR": 1.0, "WEIGHT_DECAY_BIAS": None,
                   "CLIP_GRADIENTS": {"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0},
                   "AMP": {"ENABLED": False}, "BASE_LR_END": 0.0})
    _C.TEST = C({"EXPECTED_RESULTS": [], "EVAL_PERIOD": 0, "DETECTIONS_PER_IMAGE": 100,
                 "AUG": {"ENABLED": False}, "PRECISE_BN": {"ENABLED": False, "NUM_ITER": 200}})
    _C.OUTPUT_DIR = "./output"
    _C.SEED = -1
    _C.CUDNN_BENCHMARK = False
    _C.VIS_PERIOD = 0
    _C.GLOBAL = C({"HACK": 1.0})
    return _C


def configurable(init_func=None, *, from_config=None):
    """@configurable: cls(cfg, *args, **kw) -> cls(**cls.from_config(cfg, *args, **kw))."""
    if init_func is not None:
        @functools.wraps(init_func)
        def wrapped(self, *args, **kwargs):
            if _called_with_cfg(*args, **kwargs):
                explicit = _get_args_from_config(type(self).from_config, *args, **kwargs)
                init_func(self, **explicit)
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    def wrapper(orig_func):
        @functools.wraps(orig_func)
        def wrapped(*args, **kwargs):
            if _called_with_cfg(*args, **kwargs):
                return orig_func(**_get_args_from_config(from_config, *args, **kwargs))
            return orig_func(*args, **kwargs)
        wrapped.from_config = from_config
        return wrapped
    return wrapper


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.pop("cfg", None), CfgNode)


def _get_args_from_config(from_config_func, *args, **kwargs):
    sig = inspect.signature(from_config_func)
    if any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values()):
        return from_config_func(*args, **kwargs)
    supported = set(sig.parameters.keys())
    extra = {k: kwargs.pop(k) for k in list(kwargs) if k not in supported}
    ret = from_config_func(*args, **kwargs)
    ret.update(extra)
    return ret
