from .backbone import Backbone  # noqa: F401
from .build import BACKBONE_REGISTRY, build_backbone  # noqa: F401
from .fpn import FPN, LastLevelMaxPool  # noqa: F401
