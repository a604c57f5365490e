"""CUDA-event breakdown of one train step into phases (forward sections, backward, optimizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from omni3d_b200 import synth, cubercnn as pc
from omni3d_b200.cubercnn.model import collate_gt
from omni3d_b200.train import FlatSGDTrainer
B = int(os.environ.get("BATCH", "32")); S = int(os.environ.get("SIZE", "640"))
cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none", "SOLVER.BASE_LR", 0.0025])
torch.manual_seed(0)
model = pc.build_model(cfg).train()
tr = FlatSGDTrainer(cfg, model)
items = synth.make_batch(B, S, S, num_gt=8, seed=0)
items = [{**it, "image": it["image"].cuda(), "gt": {k: v.cuda() for k, v in it["gt"].items()}} for it in items]
for _ in range(3):
    tr.step(items)
ev = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((name, e))
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
mark("start")
tr.flat_g.zero_()
x, sizes = model.preprocess_image(items); mark("preprocess")
feats_bu = model.backbone.bottom_up(x); mark("dla34 fwd")
# FPN part only
from omni3d_b200.nnfunc import ConvBias
bb = model.backbone
results, prev = {}, None
for f, s in reversed(list(zip(bb.in_features, bb.stages))):
    lat, outc = getattr(bb, "fpn_lateral%d" % s), getattr(bb, "fpn_output%d" % s)
    prev = ConvBias.apply(feats_bu[f], lat.weight, lat.bias, prev, 1, 0, False, False)
    results["p%d" % s] = ConvBias.apply(prev, outc.weight, outc.bias, None, 1, 1, False, False)
features = {k: results[k] for k in bb._out_features}; mark("fpn fwd")
gt = collate_gt(items, model.device); mark("collate gt")
pg = model.proposal_generator
feats = [features[f] for f in pg.in_features]
anchors_l = pg.anchor_generator([tuple(f.shape[1:3]) for f in feats], feats[0].device)
logits_l, deltas_l = pg.rpn_head(feats); mark("rpn head convs")
anchors = torch.cat(anchors_l, 0); logits, deltas = torch.cat(logits_l, 1), torch.cat(deltas_l, 1)
labels, idx = pg.label_and_sample_anchors(anchors, gt["boxes"], gt["classes"], gt["present"]); mark("rpn label+sample")
l_rpn = pg.losses(anchors, logits, deltas, labels, idx, gt["boxes"]); mark("rpn losses")
props = pg.predict_proposals(anchors_l, [l.detach() for l in logits_l], [d.detach() for d in deltas_l], sizes); mark("rpn proposals (decode/topk/nms)")
rh = model.roi_heads
ratios = [1.0] * B; Ks = [it["K"] for it in items]
fl = [features[f] for f in rh.in_features]
smp = rh.label_and_sample_proposals(props[0], props[2], gt); mark("roi label+sample")
xr = rh.pool(fl, smp["boxes"], smp["valid"]); mark("roi_align box")
scores, dl = rh.box_branch(xr); mark("box head GEMMs")
losses = rh.box_losses(scores, dl, smp); mark("box losses")
Fc = smp["fcap"]; K = rh.num_classes
fb, fc_, fv = smp["boxes"][:, :Fc], smp["classes"][:, :Fc], smp["valid"][:, :Fc]
fv = fv & (fc_ >= 0) & (fc_ < K)
xc = rh.pool(fl, fb, fv); mark("roi_align cube")
Kb, v2r, _ = rh.per_box_camera(Ks, ratios, [s[0] for s in sizes], Fc, B, x.device)
raw = rh.cube_outputs(xc, fc_.reshape(-1)); mark("cube head GEMMs")
losses.update(rh.cube_losses_fused(raw, fb.reshape(-1, 4), fc_.reshape(-1), fv.reshape(-1), smp["gt_boxes3D"][:, :Fc].reshape(-1, 9),
                                   smp["gt_poses"][:, :Fc].reshape(-1, 3, 3), Kb, v2r)); mark("cube loss")
losses.update(l_rpn)
total = sum(losses.values()); mark("sum")
t_fwd_host = time.perf_counter() - t0
total.backward(); mark("backward (all)")
t_bwd_host = time.perf_counter() - t0
torch.cuda.synchronize()
prev_e = ev[0][1]
tot = 0
for name, e in ev[1:]:
    ms = prev_e.elapsed_time(e); tot += ms
    print("%8.3f ms  %s" % (ms, name)); prev_e = e
print("%8.3f ms  TOTAL (GPU)   host enqueue: fwd %.1f ms, fwd+bwd %.1f ms" % (tot, t_fwd_host * 1e3, t_bwd_host * 1e3))
