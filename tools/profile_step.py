"""One profiled train step (BASELINE configs[1]) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from omni3d_b200 import synth, cubercnn as pc
from omni3d_b200.train import FlatSGDTrainer
B = int(os.environ.get("BATCH", "32")); S = int(os.environ.get("SIZE", "640"))
cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none", "SOLVER.BASE_LR", 0.0025])
torch.manual_seed(0)
model = pc.build_model(cfg).train()
tr = FlatSGDTrainer(cfg, model)
items = synth.make_batch(B, S, S, num_gt=8, seed=0, image_dtype=torch.uint8)
items = [{**it, "image": it["image"].cuda(), "gt": {k: v.cuda() for k, v in it["gt"].items()}} for it in items]
for _ in range(2):
    tr.step(items)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(items)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", tr.status()["total_loss"])
