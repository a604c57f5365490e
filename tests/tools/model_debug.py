import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from omni3d_b200 import synth, cubercnn as pc
from oracle import cubercnn_oracle as co, model_io
from oracle_capture import run_oracle_train, to_injection
H, W = 128, 192
torch.manual_seed(0); orc = co.build_model(co.load_cfg("cubercnn_DLA34_FPN.yaml"))
torch.manual_seed(0); prod = pc.build_model(pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none"]))
items = synth.make_batch(2, H, W, with_gt=False, seed=7)
prod.train(); orc.train()
with torch.no_grad():
    x, _ = prod.preprocess_image(items)
    xr = orc.preprocess_image(model_io.to_d2_inputs(items)).tensor
    print("preprocess rel", float((x[..., :3].float().cpu().permute(0,3,1,2) - xr).norm() / xr.norm()))
    # layer by layer through the bottom-up
    from omni3d_b200.cubercnn.backbone import conv_bn
    bu, ob = prod.backbone.bottom_up, orc.backbone.bottom_up
    a, b = x, xr
    for name in ("base_layer", "level0", "level1"):
        seq = getattr(bu, name); a = conv_bn(a, seq[0], seq[1]); b = getattr(ob, name)(b)
        print(name, "rel", float((a.float().cpu().permute(0,3,1,2) - b).norm() / b.norm()), tuple(a.shape))
    for i in range(2, 6):
        a = getattr(bu, "level%d" % i)(a); b = getattr(ob, "level%d" % i)(b)
        print("level%d" % i, "rel", float((a.float().cpu().permute(0,3,1,2) - b).norm() / b.norm()), tuple(a.shape))
    feats = prod.backbone(x); ref = orc.backbone(xr)
    for k in ref:
        print(k, "rel", float((feats[k].float().cpu().permute(0,3,1,2) - ref[k]).norm() / ref[k].norm()))
items = synth.make_batch(2, H, W, num_gt=4, seed=1)
ref_losses, _, cap = run_oracle_train(orc, items)
prod.train(); prod.zero_grad()
losses = prod(items, _inject=to_injection(cap, "cuda"))
for k, v in ref_losses.items():
    print(k, float(losses[k]), float(v))
sum(losses.values()).backward()
ref_g = {n: p.grad for n, p in orc.named_parameters() if p.grad is not None}
got_g = {n: p.grad for n, p in prod.named_parameters() if p.grad is not None}
print("grad key diff", sorted(set(ref_g) ^ set(got_g))[:10])
rows = []
for n, g in ref_g.items():
    if n in got_g:
        rows.append((float((got_g[n].float().cpu() - g).norm() / (g.norm() + 1e-12)), float(g.norm()), n))
rows.sort(reverse=True)
for r in rows[:25]: print("%.3f  |g|=%.3e  %s" % r)
print("median rel", sorted(r[0] for r in rows)[len(rows)//2])
