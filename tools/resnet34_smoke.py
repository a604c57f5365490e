"""ResNet34-FPN Cube R-CNN: two eager train steps of the product path on cuda:0 at a small size (does it run; are the
losses finite).  Numeric parity of this backbone against the oracle is a round-2 test."""
import torch, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omni3d_b200 import cubercnn as pc, synth
from omni3d_b200.train import FlatSGDTrainer
cfg = pc.load_cfg("cubercnn_ResNet34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none", "SOLVER.BASE_LR", 0.0025])
torch.manual_seed(0)
m = pc.build_model(cfg).train()
tr = FlatSGDTrainer(cfg, m, use_graph=False)
items = synth.make_batch(2, 128, 192, num_gt=4, seed=1, image_dtype=torch.uint8)
for i in range(2):
    tr.step(items)
print("R34", tr.status()["losses"])
