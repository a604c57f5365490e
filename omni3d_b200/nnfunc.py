"""torch.autograd.Function wrappers that put the hand-written sm_100a kernels on the autograd tape.

Activations are NHWC bf16 tensors (N,H,W,C); parameters stay fp32 masters in the reference's layouts
(conv weight OIHW, so checkpoints / optimizers see the reference's tensors) and are re-packed to the
kernels' bf16 OHWI layouts once per parameter version.  Forward AND backward run on libc3d.so:
  conv fwd  -> c3d_conv2d_fwd          dgrad -> c3d_conv2d_fwd with flipped/transposed weights
  wgrad     -> c3d_conv2d_wgrad        BN    -> c3d_bn_finalize / c3d_bn_apply / c3d_bn_bwd
"""
import torch

from . import conv as K
from . import kernels as Kx

_pack_cache = {}
_epoch = 0


def invalidate_packed():
    """Call after parameters were updated outside autograd's view (the fused SGD kernel writes through raw
    pointers, so tensor._version does not move)."""
    global _epoch
    _epoch += 1
    _pack_cache.clear()


def _packed(w, kind):
    """bf16 kernel-layout copies of an fp32 OIHW master; cached for nn.Parameters only (per storage/version/
    epoch) — temporaries (padded stem weight, fused RPN predictor weight) are packed on the fly."""
    cacheable = isinstance(w, torch.nn.Parameter)
    key = (w.data_ptr(), kind)
    ver = (w._version, _epoch)
    hit = _pack_cache.get(key) if cacheable else None
    if hit is not None and hit[0] == ver and hit[2] == tuple(w.shape):
        return hit[1]
    with torch.no_grad():
        if kind == "fwd":        # (Cout,Cin,KH,KW) -> (Cout,KH,KW,Cin)
            p = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        elif kind == "dgrad":    # -> (Cin,KH,KW,Cout), taps rotated by 180 degrees
            p = w.flip(2, 3).permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)
        else:
            raise ValueError(kind)
    if cacheable:
        _pack_cache[key] = (ver, p, tuple(w.shape))
    return p


def _dgrad(dy, w, stride, pad, in_hw):
    """dx for y = conv(x, w, stride, pad): stride-1 conv of (zero-stuffed) dy with the rotated weights."""
    KH = w.shape[2]
    wp = _packed(w, "dgrad")
    if stride == 1:
        return K.conv2d_fwd(dy, wp, stride=1, pad=KH - 1 - pad)
    assert stride == 2
    H, W = in_hw
    z = Kx.zero_stuff2(dy, H, W)
    if KH == 1:                      # 1x1 stride 2: pure scatter + 1x1 conv
        return K.conv2d_fwd(z, wp, stride=1, pad=0)
    return K.conv2d_fwd(z, wp, stride=1, pad=KH - 1 - pad)


def _wgrad_to_master(x, dy, w, stride, pad):
    dw = K.conv2d_wgrad(x, dy, w.shape[2], w.shape[3], stride, pad)      # (Cout,KH,KW,Cin) fp32
    return dw.permute(0, 3, 1, 2)


class ConvBNAct(torch.autograd.Function):
    """out = [relu]( BN_train|eval( conv(x, w) ) [+ residual] ) — dla.py:40-68 BasicBlock halves, Root, project."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, running_mean, running_var, residual, stride, pad, relu, training, eps,
                momentum):
        x = x.contiguous()
        wp = _packed(w, "fwd")
        if training:
            y, stats = K.conv2d_fwd(x, wp, stride=stride, pad=pad, want_stats=True)
            count = y.numel() // y.shape[-1]
            mean, rstd = Kx.bn_finalize(stats, count, eps, momentum, running_mean, running_var)
        else:
            y = K.conv2d_fwd(x, wp, stride=stride, pad=pad)
            mean, rstd = running_mean, torch.rsqrt(running_var + eps)
        res = residual.contiguous() if residual is not None else None
        out = Kx.bn_apply(y, mean, rstd, gamma, beta, res, relu)
        ctx.save_for_backward(x, w, gamma, y, mean, rstd, out)
        ctx.cfg = (stride, pad, relu, training, residual is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, gamma, y, mean, rstd, out = ctx.saved_tensors
        stride, pad, relu, training, has_res = ctx.cfg
        dout = dout.contiguous()
        dgamma = torch.zeros_like(gamma)
        dbeta = torch.zeros_like(gamma)
        dy, dres = Kx.bn_bwd(dout, out, y, mean, rstd, gamma, relu, dgamma, dbeta, has_res and ctx.needs_input_grad[6],
                              frozen=not training)
        dx = _dgrad(dy, w, stride, pad, x.shape[1:3]) if ctx.needs_input_grad[0] else None
        dw = _wgrad_to_master(x, dy, w, stride, pad) if ctx.needs_input_grad[1] else None
        return dx, dw, dgamma, dbeta, None, None, dres, None, None, None, None, None, None


class ConvBias(torch.autograd.Function):
    """out = [relu]( conv(x, w) + b [+ up2(addend)] ) — FPN lateral/output convs and the RPN head conv."""

    @staticmethod
    def forward(ctx, x, w, bias, addend, stride, pad, relu, out_fp32):
        x = x.contiguous()
        add = addend.contiguous() if addend is not None else None
        out = K.conv2d_fwd(x, _packed(w, "fwd"), bias, stride, pad, relu=relu, addend=add, up2=add is not None,
                           out_fp32=out_fp32)
        ctx.save_for_backward(x, w, out if relu else None)
        ctx.cfg = (stride, pad, relu, addend is not None, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, out = ctx.saved_tensors
        stride, pad, relu, has_add, has_bias = ctx.cfg
        dzb, dbias = Kx.bias_act_bwd(dout, out, relu, has_bias)
        dadd = Kx.sumpool2(dzb) if has_add and ctx.needs_input_grad[3] else None
        dx = _dgrad(dzb, w, stride, pad, x.shape[1:3]) if ctx.needs_input_grad[0] else None
        dw = _wgrad_to_master(x, dzb, w, stride, pad) if ctx.needs_input_grad[1] else None
        return dx, dw, dbias, dadd, None, None, None, None


class MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return Kx.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return Kx.maxpool2_bwd(x, dy.contiguous())


class ROIAlign(torch.autograd.Function):
    """feats (tuple of NHWC bf16 maps), rois (R,6) [batch, level, x1,y1,x2,y2] -> (R,7,7,C) bf16."""

    @staticmethod
    def forward(ctx, rois, strides, pooled, *feats):
        feats = [f.contiguous() for f in feats]
        ctx.save_for_backward(rois, *feats)
        ctx.cfg = (strides, pooled)
        return Kx.roi_align_fwd(feats, strides, rois, pooled)

    @staticmethod
    def backward(ctx, dout):
        rois, *feats = ctx.saved_tensors
        strides, pooled = ctx.cfg
        grads = Kx.roi_align_bwd(feats, strides, rois, dout.contiguous(), pooled)
        return (None, None, None) + tuple(g.to(torch.bfloat16) for g in grads)
