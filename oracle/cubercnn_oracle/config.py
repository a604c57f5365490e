"""Cube R-CNN config keys (restates cubercnn/config/config.py:4-159) on the d2lite CfgNode."""
import os

from detectron2.config import CfgNode as CN
from detectron2.config import get_cfg as _d2_get_cfg

_CUBE_HEAD = dict(
    NAME="CubeHead", POOLER_RESOLUTION=7, POOLER_SAMPLING_RATIO=0, POOLER_TYPE="ROIAlignV2", NUM_CONV=0,
    CONV_DIM=256, NUM_FC=2, FC_DIM=1024, Z_TYPE="direct", POSE_TYPE="6d", INVERSE_Z_WEIGHT=False,
    VIRTUAL_DEPTH=True, VIRTUAL_FOCAL=512.0, DISENTANGLED_LOSS=True, CLUSTER_BINS=1, ALLOCENTRIC_POSE=True,
    CHAMFER_POSE=True, SHARED_FC=True, DIMS_PRIORS_ENABLED=True, DIMS_PRIORS_FUNC="exp", USE_CONFIDENCE=1.0,
    LOSS_W_3D=1.0, LOSS_W_XY=1.0, LOSS_W_Z=1.0, LOSS_W_DIMS=1.0, LOSS_W_POSE=1.0, LOSS_W_JOINT=1.0,
    SCALE_ROI_BOXES=0.0)


def get_cfg_defaults(cfg):
    cfg.DATASETS.CATEGORY_NAMES = []
    cfg.DATASETS.IGNORE_NAMES = []
    cfg.DATALOADER.BALANCE_DATASETS = False
    for k, v in dict(TRUNCATION_THRES=0.99, VISIBILITY_THRES=0.01, MIN_HEIGHT_THRES=0.0, MAX_DEPTH=1e8,
                     MODAL_2D_BOXES=False, TRUNC_2D_BOXES=True).items():
        cfg.DATASETS[k] = v
    cfg.MODEL.RPN.IGNORE_THRESHOLD = 0.5
    cfg.MODEL.RPN.OBJECTNESS_UNCERTAINTY = "IoUness"
    cfg.MODEL.ROI_CUBE_HEAD = CN(dict(_CUBE_HEAD))
    cfg.MODEL.USE_BN = True
    cfg.MODEL.STABILIZE = 0.01
    cfg.MODEL.DLA = CN({"TYPE": "dla34", "TRICKS": False})
    cfg.MODEL.RESNETS.TORCHVISION = True
    cfg.MODEL.WEIGHTS_PRETRAIN = ""
    cfg.SOLVER.TYPE = "sgd"
    cfg.TEST.DETECTIONS_PER_IMAGE = 100
    cfg.TEST.VISIBILITY_THRES = 0.5
    cfg.TEST.TRUNCATION_THRES = 0.5
    cfg.INPUT.RANDOM_FLIP = "horizontal"
    return cfg


def get_cfg():
    return get_cfg_defaults(_d2_get_cfg())


CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "configs")


def load_cfg(config_file, opts=()):
    """config_file: path to a reference-format YAML (the repo ships verbatim-compatible copies under configs/)."""
    cfg = get_cfg()
    if not os.path.isabs(config_file) and not os.path.exists(config_file):
        config_file = os.path.join(CONFIG_DIR, config_file)
    cfg.merge_from_file(config_file)
    cfg.merge_from_list(["MODEL.DEVICE", "cpu", "MODEL.WEIGHTS_PRETRAIN", "none", "VIS_PERIOD", 0] + list(opts))
    return cfg
