"""Checkpoint interop with the reference — SURVEY 8f-4.

The reference saves / loads through detectron2's DetectionCheckpointer (tools/train_net.py:128-145,443-445,
demo/demo.py:152-154) and cubercnn/solver/checkpoint.py:5-27 (PeriodicCheckpointerOnlyOne: one `model_recent.pth`
per period + `model_final.pth`).  File format = torch.save({"model": state_dict, "optimizer": ..., "scheduler": ...,
"iteration": i}); the zoo files of MODEL_ZOO.md:9-16 are such files.  The accelerated model keeps the reference's
parameter names and layouts (fp32 OIHW masters), so a reference state_dict loads key for key:

    load_checkpoint(model, path_or_dict)             -> {"iteration": ..., "missing": [...], "unexpected": [...]}
    save_checkpoint(model, path, trainer=None, **extra)
    Checkpointer(model, save_dir, trainer).resume_or_load(path, resume) / .save(name, **extra)   (the calls the scripts make)
    PeriodicCheckpointerOnlyOne(checkpointer, period, max_iter).step(iteration)

and fold_batchnorm(model) folds eval-mode BatchNorm into the preceding convolution for inference (demo.py): the
bf16 conv then carries scale and shift in its epilogue (bias + residual + ReLU) and the separate BN pass disappears.
"""
import os

import torch

# keys the reference model has but that are not parameters of the accelerated path (none today) / legacy prefixes
_STRIP_PREFIXES = ("module.",)            # DistributedDataParallel wrap (train_net.py:451)


def _state_of(obj):
    if isinstance(obj, (str, os.PathLike)):
        obj = torch.load(obj, map_location="cpu", weights_only=False)
    extra = {}
    if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict):
        extra = {k: v for k, v in obj.items() if k != "model"}
        obj = obj["model"]
    sd = {}
    for k, v in obj.items():
        for p in _STRIP_PREFIXES:
            if k.startswith(p):
                k = k[len(p):]
        sd[k] = torch.as_tensor(v) if not isinstance(v, torch.Tensor) else v
    return sd, extra


def load_checkpoint(model, path_or_dict, strict_shapes=True):
    """Load a reference / own checkpoint.  Shape mismatches raise (the reference's checkpointer warns and skips: a
    silently skipped 3D head is worse than an error); missing / unexpected keys are returned like load_state_dict's."""
    sd, extra = _state_of(path_or_dict)
    own = model.state_dict()
    bad = [(k, tuple(v.shape), tuple(own[k].shape)) for k, v in sd.items() if k in own and tuple(v.shape) != tuple(own[k].shape)]
    if bad and strict_shapes:
        raise ValueError("checkpoint tensors with a different shape than the model: %s" % bad[:5])
    for k, _, _ in bad:
        sd.pop(k)
    res = model.load_state_dict(sd, strict=False)
    from . import nnfunc
    nnfunc.invalidate_packed()             # packed bf16 copies of the old weights are stale now
    return {"iteration": extra.get("iteration", -1), "missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys),
            "extra": extra}


def save_checkpoint(model, path, trainer=None, **extra):
    """torch.save({"model": state_dict (contiguous fp32 tensors in the reference's layouts), "trainer": ..., **extra})."""
    sd = {k: v.detach().to("cpu").contiguous().clone() for k, v in model.state_dict().items()}
    data = {"model": sd}
    if trainer is not None:
        data["trainer"] = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in trainer.state_dict().items()}
        data.setdefault("iteration", trainer.iteration - 1)
    data.update(extra)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(data, path)
    return path


class Checkpointer:
    """The DetectionCheckpointer calls tools/train_net.py:128-145 and demo/demo.py:152-154 make."""

    def __init__(self, model, save_dir="", trainer=None, **checkpointables):
        self.model, self.save_dir, self.trainer = model, save_dir, trainer

    def _last(self):
        p = os.path.join(self.save_dir, "last_checkpoint")
        if os.path.exists(p):
            return os.path.join(self.save_dir, open(p).read().strip())
        return None

    def has_checkpoint(self):
        return self._last() is not None

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        info = load_checkpoint(self.model, path)
        tr = info["extra"].get("trainer")
        if self.trainer is not None and tr is not None and (checkpointables is None or "trainer" in checkpointables):
            self.trainer.load_state_dict(tr)
        return {"iteration": info["iteration"], **{k: v for k, v in info.items() if k in ("missing", "unexpected")}}

    def resume_or_load(self, path, resume=True):
        if resume and self.has_checkpoint():
            return self.load(self._last())
        return self.load(path, checkpointables=[])

    def save(self, name, **extra):
        path = os.path.join(self.save_dir, name + ".pth")
        save_checkpoint(self.model, path, self.trainer, **extra)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(os.path.basename(path))
        return path


class PeriodicCheckpointerOnlyOne:
    """cubercnn/solver/checkpoint.py:5-27: a single `<prefix>_recent.pth` every `period` iterations, `<prefix>_final.pth`
    at max_iter - 1."""

    def __init__(self, checkpointer, period, max_iter=None, file_prefix="model"):
        self.checkpointer, self.period, self.max_iter, self.file_prefix = checkpointer, int(period), max_iter, file_prefix

    def step(self, iteration, **kwargs):
        iteration = int(iteration)
        state = {"iteration": iteration, **kwargs}
        if (iteration + 1) % self.period == 0:
            self.checkpointer.save("{}_recent".format(self.file_prefix), **state)
        if self.max_iter is not None and iteration >= self.max_iter - 1:
            self.checkpointer.save("{}_final".format(self.file_prefix), **state)


# ---- BatchNorm folding for inference ---------------------------------------------------------------------------------
def folded_conv_params(conv_weight, bn):
    """eval-mode BatchNorm(conv(x)) == conv'(x) + b' with w' = w * gamma * rstd (per output channel),
    b' = beta - running_mean * gamma * rstd.  fp32 in, fp32 out (the bf16 pack happens on the folded weight)."""
    rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
    scale = bn.weight.float() * rstd
    w = conv_weight.float() * scale.view(-1, 1, 1, 1)
    b = bn.bias.float() - bn.running_mean.float() * scale
    return w, b


def fold_batchnorm(model, enable=True):
    """Switch the backbone's conv+BN pairs to the folded single-kernel path whenever their BatchNorm is in eval mode
    (model.eval() / freeze_bn).  The fold is recomputed from the live parameters / running statistics per parameter
    version, so loading another checkpoint needs no re-fold; training-mode BatchNorm is never folded."""
    from .cubercnn import backbone as B
    B.set_bn_folding(model, enable)
    return model
