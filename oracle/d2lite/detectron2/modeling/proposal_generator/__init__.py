from .build import PROPOSAL_GENERATOR_REGISTRY, build_proposal_generator  # noqa: F401
from .rpn import RPN, RPN_HEAD_REGISTRY, StandardRPNHead, build_rpn_head  # noqa: F401
