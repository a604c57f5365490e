"""Config surface of the accelerated path: a small yacs-style CfgNode able to load the reference's YAML
files unchanged (incl. `_BASE_` chains and CLI `KEY VALUE` overrides, tools/train_net.py:318-333) and the
key set of cubercnn/config/config.py:4-159 plus the detectron2 defaults the model path reads."""
import ast
import copy
import os

import yaml


class CfgNode(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if type(v) is dict else v

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass

    @staticmethod
    def _read(path):
        with open(path) as f:
            d = yaml.safe_load(f) or {}
        base = d.pop("_BASE_", None)
        if base is None:
            return d
        if not os.path.isabs(base):
            base = os.path.join(os.path.dirname(path), base)
        out = CfgNode._read(base)
        CfgNode._overlay(out, d)
        return out

    @staticmethod
    def _overlay(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                CfgNode._overlay(dst[k], v)
            else:
                dst[k] = v

    @staticmethod
    def _cast(v, like):
        if isinstance(v, str) and not isinstance(like, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
        if isinstance(like, tuple) and isinstance(v, list):
            v = tuple(v)
        elif isinstance(like, list) and isinstance(v, tuple):
            v = list(v)
        elif isinstance(like, float) and isinstance(v, int) and not isinstance(v, bool):
            v = float(v)
        return v

    def _merge(self, src, prefix=""):
        for k, v in src.items():
            if isinstance(v, dict):
                self.setdefault(k, CfgNode())._merge(v, prefix + k + ".")
            elif k not in self:
                raise KeyError(f"Non-existent config key: {prefix}{k}")
            else:
                self[k] = self._cast(v, self[k])

    def merge_from_file(self, path, allow_unsafe=True):
        self._merge(self._read(path))

    def merge_from_list(self, opts):
        if len(opts) % 2:
            raise ValueError("override list must be KEY VALUE pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            *parents, leaf = key.split(".")
            for p in parents:
                node = node[p]
            if leaf not in node:
                raise KeyError(f"Non-existent config key: {key}")
            node[leaf] = self._cast(val, node[leaf])


def get_cfg():
    """detectron2 defaults for every key the model path / YAMLs touch (values: detectron2 v0.6)."""
    N = CfgNode
    c = N()
    c.VERSION = 2
    c.MODEL = N(dict(
        DEVICE="cuda", META_ARCHITECTURE="GeneralizedRCNN", WEIGHTS="", MASK_ON=False, KEYPOINT_ON=False,
        LOAD_PROPOSALS=False, PIXEL_MEAN=[103.530, 116.280, 123.675], PIXEL_STD=[1.0, 1.0, 1.0],
        BACKBONE=dict(NAME="build_resnet_backbone", FREEZE_AT=2),
        FPN=dict(IN_FEATURES=[], OUT_CHANNELS=256, NORM="", FUSE_TYPE="sum"),
        PROPOSAL_GENERATOR=dict(NAME="RPN", MIN_SIZE=0),
        ANCHOR_GENERATOR=dict(NAME="DefaultAnchorGenerator", SIZES=[[32, 64, 128, 256, 512]],
                              ASPECT_RATIOS=[[0.5, 1.0, 2.0]], ANGLES=[[-90, 0, 90]], OFFSET=0.0),
        RPN=dict(HEAD_NAME="StandardRPNHead", IN_FEATURES=["res4"], BOUNDARY_THRESH=-1, IOU_THRESHOLDS=[0.3, 0.7],
                 IOU_LABELS=[0, -1, 1], BATCH_SIZE_PER_IMAGE=256, POSITIVE_FRACTION=0.5,
                 BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0, BBOX_REG_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
                 SMOOTH_L1_BETA=0.0, LOSS_WEIGHT=1.0, PRE_NMS_TOPK_TRAIN=12000, PRE_NMS_TOPK_TEST=6000,
                 POST_NMS_TOPK_TRAIN=2000, POST_NMS_TOPK_TEST=1000, NMS_THRESH=0.7, CONV_DIMS=[-1]),
        ROI_HEADS=dict(NAME="Res5ROIHeads", NUM_CLASSES=80, IN_FEATURES=["res4"], IOU_THRESHOLDS=[0.5],
                       IOU_LABELS=[0, 1], BATCH_SIZE_PER_IMAGE=512, POSITIVE_FRACTION=0.25, SCORE_THRESH_TEST=0.05,
                       NMS_THRESH_TEST=0.5, PROPOSAL_APPEND_GT=True),
        ROI_BOX_HEAD=dict(NAME="", BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0,
                          BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), SMOOTH_L1_BETA=0.0, POOLER_RESOLUTION=14,
                          POOLER_SAMPLING_RATIO=0, POOLER_TYPE="ROIAlignV2", NUM_FC=0, FC_DIM=1024, NUM_CONV=0,
                          CONV_DIM=256, NORM="", CLS_AGNOSTIC_BBOX_REG=False, TRAIN_ON_PRED_BOXES=False,
                          USE_FED_LOSS=False, USE_SIGMOID_CE=False, FED_LOSS_FREQ_WEIGHT_POWER=0.5,
                          FED_LOSS_NUM_CLASSES=50),
        RESNETS=dict(DEPTH=50, OUT_FEATURES=["res4"], NUM_GROUPS=1, NORM="FrozenBN", WIDTH_PER_GROUP=64,
                     STRIDE_IN_1X1=True, RES5_DILATION=1, RES2_OUT_CHANNELS=256, STEM_OUT_CHANNELS=64)))
    c.INPUT = N(dict(MIN_SIZE_TRAIN=(800,), MIN_SIZE_TRAIN_SAMPLING="choice", MAX_SIZE_TRAIN=1333, MIN_SIZE_TEST=800,
                     MAX_SIZE_TEST=1333, RANDOM_FLIP="horizontal", FORMAT="BGR", MASK_FORMAT="polygon"))
    c.DATASETS = N(dict(TRAIN=(), TEST=()))
    c.DATALOADER = N(dict(NUM_WORKERS=4, ASPECT_RATIO_GROUPING=True, SAMPLER_TRAIN="TrainingSampler",
                          REPEAT_THRESHOLD=0.0, FILTER_EMPTY_ANNOTATIONS=True))
    c.SOLVER = N(dict(LR_SCHEDULER_NAME="WarmupMultiStepLR", MAX_ITER=40000, BASE_LR=0.001, MOMENTUM=0.9,
                      NESTEROV=False, WEIGHT_DECAY=0.0001, WEIGHT_DECAY_NORM=0.0, GAMMA=0.1, STEPS=(30000,),
                      WARMUP_FACTOR=0.001, WARMUP_ITERS=1000, WARMUP_METHOD="linear", CHECKPOINT_PERIOD=5000,
                      IMS_PER_BATCH=16, REFERENCE_WORLD_SIZE=0, BIAS_LR_FACTOR=1.0, WEIGHT_DECAY_BIAS=None,
                      CLIP_GRADIENTS=dict(ENABLED=False, CLIP_TYPE="value", CLIP_VALUE=1.0, NORM_TYPE=2.0),
                      AMP=dict(ENABLED=False)))
    c.TEST = N(dict(EXPECTED_RESULTS=[], EVAL_PERIOD=0, DETECTIONS_PER_IMAGE=100))
    c.OUTPUT_DIR = "./output"
    c.SEED = -1
    c.CUDNN_BENCHMARK = False
    c.VIS_PERIOD = 0
    return c


def get_cfg_defaults(cfg):
    """Adds the Cube R-CNN keys (same names and defaults as cubercnn/config/config.py:4-159)."""
    D, L, M, R = cfg.DATASETS, cfg.DATALOADER, cfg.MODEL, cfg.MODEL.RPN
    D.CATEGORY_NAMES, D.IGNORE_NAMES = [], []
    L.BALANCE_DATASETS = False
    D.TRUNCATION_THRES, D.VISIBILITY_THRES, D.MIN_HEIGHT_THRES, D.MAX_DEPTH = 0.99, 0.01, 0.00, 1e8
    D.MODAL_2D_BOXES, D.TRUNC_2D_BOXES = False, True
    R.IGNORE_THRESHOLD, R.OBJECTNESS_UNCERTAINTY = 0.5, "IoUness"
    M.ROI_CUBE_HEAD = CfgNode(dict(
        NAME="CubeHead", POOLER_RESOLUTION=7, POOLER_SAMPLING_RATIO=0, POOLER_TYPE="ROIAlignV2", NUM_CONV=0,
        CONV_DIM=256, NUM_FC=2, FC_DIM=1024, Z_TYPE="direct", POSE_TYPE="6d", INVERSE_Z_WEIGHT=False,
        VIRTUAL_DEPTH=True, VIRTUAL_FOCAL=512.0, DISENTANGLED_LOSS=True, CLUSTER_BINS=1, ALLOCENTRIC_POSE=True,
        CHAMFER_POSE=True, SHARED_FC=True, DIMS_PRIORS_ENABLED=True, DIMS_PRIORS_FUNC="exp", USE_CONFIDENCE=1.0,
        LOSS_W_3D=1.0, LOSS_W_XY=1.0, LOSS_W_Z=1.0, LOSS_W_DIMS=1.0, LOSS_W_POSE=1.0, LOSS_W_JOINT=1.0,
        SCALE_ROI_BOXES=0.0))
    M.USE_BN, M.STABILIZE = True, 0.01
    M.DLA = CfgNode(dict(TYPE="dla34", TRICKS=False))
    M.RESNETS.TORCHVISION = True
    M.WEIGHTS_PRETRAIN = ""
    cfg.SOLVER.TYPE = "sgd"
    cfg.TEST.DETECTIONS_PER_IMAGE = 100
    cfg.TEST.VISIBILITY_THRES, cfg.TEST.TRUNCATION_THRES = 0.5, 0.5
    cfg.INPUT.RANDOM_FLIP = "horizontal"
    return cfg


_CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "configs")


def load_cfg(config_file, opts=()):
    cfg = get_cfg_defaults(get_cfg())
    if not os.path.exists(config_file):
        config_file = os.path.join(_CONFIG_DIR, config_file)
    cfg.merge_from_file(config_file)
    cfg.merge_from_list(list(opts))
    return cfg
