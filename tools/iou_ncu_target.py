import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, boxgen
from omni3d_b200 import box3d
L = float(os.environ.get("L", "1.0")); n = int(os.environ.get("N", "1000"))
a = torch.from_numpy(boxgen.random_boxes(n, L, 0)).cuda(); b = torch.from_numpy(boxgen.random_boxes(n, L, 5)).cuda()
for _ in range(2):
    box3d.iou_box3d(a, b)
torch.cuda.synchronize()
