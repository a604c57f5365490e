"""ORACLE — test infrastructure only: CPU/fp32 plain-PyTorch restatement of the Cube R-CNN model
path, built on oracle/d2lite.  Travels to the GPU box (unlike
/root/reference) and is what tests/, smoke() and bench.py's cpu_baseline / --impl reference use.

Pinned in the build container against the reference's own code (oracle/ref_runner.py) by
tests/golden/make_model_golden.py + tests/test_model_oracle.py: identical state_dict key set, identical
losses / proposals / detections on seeded inputs.

Only the configuration the BASELINE configs use is restated (disentangled chamfer loss, allocentric
6D pose, direct virtual depth, shared-FC CubeHead, IoUness RPN); other switches raise.
"""
import os
import sys

_D2 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "d2lite")
if _D2 not in sys.path:
    sys.path.insert(0, _D2)

from .config import get_cfg, get_cfg_defaults, load_cfg  # noqa: E402,F401
from .model import RCNN3D, build_model  # noqa: E402,F401
