from .meta_arch import META_ARCH_REGISTRY, GeneralizedRCNN  # noqa: F401
from .backbone import BACKBONE_REGISTRY, Backbone, FPN  # noqa: F401
from .proposal_generator import PROPOSAL_GENERATOR_REGISTRY, RPN, build_proposal_generator  # noqa: F401
from .roi_heads import ROI_HEADS_REGISTRY, StandardROIHeads, build_roi_heads  # noqa: F401
from .anchor_generator import DefaultAnchorGenerator, build_anchor_generator  # noqa: F401
from .box_regression import Box2BoxTransform  # noqa: F401
from .matcher import Matcher  # noqa: F401
from .poolers import ROIPooler  # noqa: F401
