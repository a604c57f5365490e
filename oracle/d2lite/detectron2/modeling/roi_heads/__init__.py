from .box_head import ROI_BOX_HEAD_REGISTRY, FastRCNNConvFCHead, build_box_head  # noqa: F401
from .fast_rcnn import FastRCNNOutputLayers, _log_classification_stats  # noqa: F401
from .roi_heads import (ROI_HEADS_REGISTRY, ROIHeads, StandardROIHeads, build_roi_heads,  # noqa: F401
                        select_foreground_proposals)
