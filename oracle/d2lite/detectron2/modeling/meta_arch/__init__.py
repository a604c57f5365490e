from .build import META_ARCH_REGISTRY, build_model  # noqa: F401
from .rcnn import GeneralizedRCNN  # noqa: F401
