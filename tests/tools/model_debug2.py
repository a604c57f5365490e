import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from omni3d_b200 import synth, cubercnn as pc
from oracle import cubercnn_oracle as co, model_io
from oracle_capture import run_oracle_train, to_injection
torch.manual_seed(0); orc = co.build_model(co.load_cfg("cubercnn_DLA34_FPN.yaml"))
torch.manual_seed(0); prod = pc.build_model(pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none"]))
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))
for (Bn, H, W) in [(2, 128, 192), (4, 256, 320)]:
    print("=== batch", Bn, H, W)
    items = synth.make_batch(Bn, H, W, num_gt=4, seed=1)
    # (1) the oracle's backbone on the GPU under bf16 autocast vs itself in fp32 on the CPU: inherent sensitivity
    ob = copy.deepcopy(orc.backbone).cuda().train()
    orc.train()
    xr = orc.preprocess_image(model_io.to_d2_inputs(items)).tensor
    xr.requires_grad_(False)
    f32 = orc.backbone(xr)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        f16 = ob(xr.cuda())
    for k in f32: print(" torch-autocast-bf16 vs fp32", k, rel(f16[k].float().cpu(), f32[k]))
    # grads of a fixed linear functional of the features
    gs = {k: torch.randn(v.shape, generator=torch.Generator().manual_seed(3)) for k, v in f32.items()}
    orc.backbone.zero_grad(); sum((f32[k] * gs[k]).sum() for k in f32).backward()
    ob.zero_grad(); sum((f16[k].float() * gs[k].cuda()).sum() for k in f16).backward()
    ga = {n: p.grad.clone() for n, p in orc.backbone.named_parameters() if p.grad is not None}
    gb = {n: p.grad for n, p in ob.named_parameters() if p.grad is not None}
    errs = sorted(rel(gb[n].cpu(), ga[n]) for n in ga)
    print(" torch-autocast grads: median rel", errs[len(errs)//2], "p90", errs[int(.9*len(errs))])
    # (2) ours vs fp32
    prod.train(); prod.zero_grad()
    x, _ = prod.preprocess_image(items)
    fm = prod.backbone(x)
    for k in f32: print(" ours vs fp32", k, rel(fm[k].float().cpu().permute(0,3,1,2), f32[k].detach()))
    sum((fm[k].float() * gs[k].permute(0,2,3,1).cuda()).sum() for k in fm).backward()
    gm = {n: p.grad for n, p in prod.backbone.named_parameters() if p.grad is not None}
    errs = sorted((rel(gm[n].float().cpu(), ga[n]), n) for n in ga if n in gm)
    print(" ours grads: median rel", errs[len(errs)//2][0], "p90", errs[int(.9*len(errs))][0], "min", errs[0], "max", errs[-1])
    errs2 = sorted((rel(gm[n].float().cpu(), gb[n].cpu()), n) for n in gb if n in gm)
    print(" ours vs torch-autocast grads: median", errs2[len(errs2)//2][0])
    for e, n in errs[:6] + errs[-6:]: print("   %.3f %s" % (e, n))
