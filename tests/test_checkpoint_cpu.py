"""SURVEY 8f-4 (host logic, no GPU): checkpoint files in the reference's format round-trip through omni3d_b200.checkpoint,
a reference-layout state_dict (the oracle's == the reference's, test_model_oracle.py) loads key for key, and the
BatchNorm fold is algebraically the eval-mode BatchNorm."""
import os

import pytest
import torch

from omni3d_b200 import checkpoint as ck
from omni3d_b200 import cubercnn as pc
from oracle import cubercnn_oracle as co


@pytest.fixture(scope="module")
def models():
    torch.manual_seed(0)
    orc = co.build_model(co.load_cfg("cubercnn_DLA34_FPN.yaml"))
    torch.manual_seed(1)
    prod = pc.build_model(pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.DEVICE", "cpu", "MODEL.WEIGHTS_PRETRAIN", "none"]))
    return prod, orc


def test_reference_layout_checkpoint_loads_key_for_key(models, tmp_path):
    prod, orc = models
    # a detectron2 checkpointer file as tools/train_net.py writes it, saved from a DDP-wrapped model
    path = os.path.join(tmp_path, "model_final.pth")
    torch.save({"model": {"module." + k: v for k, v in orc.state_dict().items()}, "iteration": 1234, "optimizer": {}}, path)
    info = ck.load_checkpoint(prod, path)
    assert info["iteration"] == 1234 and info["missing"] == [] and info["unexpected"] == []
    sd = prod.state_dict()
    for k, v in orc.state_dict().items():
        assert torch.equal(sd[k], v), k
    # shape mismatches raise instead of being skipped silently
    bad = {k: v for k, v in orc.state_dict().items()}
    bad["roi_heads.box_predictor.cls_score.weight"] = torch.zeros(11, 1024)
    with pytest.raises(ValueError):
        ck.load_checkpoint(prod, {"model": bad})


def test_checkpointer_resume_and_periodic_only_one(models, tmp_path):
    prod, _ = models
    d = str(tmp_path)
    cp = ck.Checkpointer(prod, d)
    per = ck.PeriodicCheckpointerOnlyOne(cp, period=5, max_iter=12)
    for it in range(12):
        per.step(it)
    assert sorted(f for f in os.listdir(d) if f.endswith(".pth")) == ["model_final.pth", "model_recent.pth"]
    before = {k: v.clone() for k, v in prod.state_dict().items()}
    with torch.no_grad():
        for p in prod.parameters():
            p.add_(1.0)
    out = ck.Checkpointer(prod, d).resume_or_load("", resume=True)
    assert out["iteration"] == 11
    for k, v in prod.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert ck.Checkpointer(prod, os.path.join(d, "empty")).resume_or_load("", resume=True) == {}


def test_folded_conv_params_equal_eval_batchnorm():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 16, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(16).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 2); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.2, 3)
    x = torch.randn(2, 8, 9, 9)
    w, b = ck.folded_conv_params(conv.weight, bn)
    ref = bn(conv(x))
    got = torch.nn.functional.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
