"""ctypes front-end of the tcgen05 implicit-GEMM convolution entry points (include/c3d.h).

Tensors are torch CUDA tensors used as raw device buffers: activations NHWC bf16 (N,H,W,C),
weights OHWI bf16 (Cout,KH,KW,Cin).  No torch types cross the ABI.
"""
import ctypes

import torch

from . import _lib


class ConvDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("Cout", ctypes.c_int32), ("KH", ctypes.c_int32), ("KW", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("pad", ctypes.c_int32), ("relu", ctypes.c_int32),
                ("out_fp32", ctypes.c_int32), ("add_mode", ctypes.c_int32),
                ("x_pix_stride", ctypes.c_int64), ("y_pix_stride", ctypes.c_int64),
                ("add_pix_stride", ctypes.c_int64), ("y_img_stride", ctypes.c_int64), ("y_h_stride", ctypes.c_int64),
                ("y_w_stride", ctypes.c_int64), ("y_offset", ctypes.c_int64), ("out_h", ctypes.c_int32),
                ("out_w", ctypes.c_int32)]


_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        P = ctypes.POINTER(ConvDesc)
        L.c3d_conv2d_tiles.restype = i32
        L.c3d_conv2d_tiles.argtypes = [P, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.c3d_conv2d_fwd.restype = i32
        L.c3d_conv2d_fwd.argtypes = [P, vp, vp, vp, vp, vp, vp, vp]
        L.c3d_conv2d_wgrad.restype = i32
        L.c3d_conv2d_wgrad.argtypes = [P, vp, vp, vp, vp]
        L.c3d_conv2d_wgrad_ex.restype = i32
        L.c3d_conv2d_wgrad_ex.argtypes = [P, vp, vp, vp, i32, vp]
        L.c3d_pack_conv_weight.restype = i32
        L.c3d_pack_conv_weight.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
        _bound = True
    return L


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def out_hw(H, W, KH, KW, stride, pad):
    return (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1


def make_desc(x, w, stride=1, pad=0, relu=False, out_fp32=False, add_mode=0):
    N, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2, (x.shape, w.shape)
    return ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, pad, int(relu), int(out_fp32), add_mode, 0, 0, 0, 0, 0, 0, 0, 0, 0)


def num_tiles(desc):
    L = _bind()
    t, th, tw = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(L.c3d_conv2d_tiles(ctypes.byref(desc), ctypes.byref(t), ctypes.byref(th), ctypes.byref(tw)))
    return t.value, th.value, tw.value


def conv2d_fwd(x, w, bias=None, stride=1, pad=0, relu=False, addend=None, up2=False, out_fp32=False,
               want_stats=False, out=None, out_place=None, out_hw_override=None):
    """x (N,H,W,Cin) bf16, w (Cout,KH,KW,Cin) bf16 -> y (N,Ho,Wo,Cout) bf16|fp32 [, stats (tiles,2,Cout)]."""
    L = _bind()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    assert w.dtype == torch.bfloat16 and w.is_contiguous()
    add_mode = 0 if addend is None else (2 if up2 else 1)
    d = make_desc(x, w, stride, pad, relu, out_fp32, add_mode)
    Ho, Wo = out_hw(d.H, d.W, d.KH, d.KW, stride, pad)
    if out_hw_override is not None:
        Ho, Wo = out_hw_override
        d.out_h, d.out_w = Ho, Wo
    if out_place is not None:        # (img_stride, h_stride, w_stride, offset) in pixels of `out`
        d.y_img_stride, d.y_h_stride, d.y_w_stride, d.y_offset = out_place
    if out is None:
        out = torch.empty((d.N, Ho, Wo, d.Cout), device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    stats = None
    if want_stats:
        t, _, _ = num_tiles(d)
        stats = torch.empty((t, 2, d.Cout), device=x.device, dtype=torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if addend is not None:
        assert addend.dtype == torch.bfloat16 and addend.is_contiguous()
    _lib.check(L.c3d_conv2d_fwd(ctypes.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(addend), _ptr(out), _ptr(stats),
                                _stream()))
    return (out, stats) if want_stats else out


def pack_conv_weight(w, want_fwd=True, want_dgrad=True):
    """fp32 (Cout,Cin,KH,KW) -> bf16 (Cout,KH,KW,Cin) and bf16 (Cin,KH,KW,Cout) rotated, in ONE launch."""
    L = _bind()
    Cout, Cin, KH, KW = w.shape
    w = w.detach()
    ohwi = (not w.is_contiguous()) and w.permute(0, 2, 3, 1).is_contiguous()     # channels_last master storage
    if not ohwi:
        w = w.contiguous()
    f = torch.empty((Cout, KH, KW, Cin), device=w.device, dtype=torch.bfloat16) if want_fwd else None
    g = torch.empty((Cin, KH, KW, Cout), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    _lib.check(L.c3d_pack_conv_weight(_ptr(w), Cout, Cin, KH, KW, int(ohwi), _ptr(f), _ptr(g), _stream()))
    return f, g


def conv2d_wgrad(x, dy, KH, KW, stride=1, pad=0, dw=None, oihw=False):
    """dW (Cout,KH,KW,Cin) [or (Cout,Cin,KH,KW) when oihw] fp32 (+)= wgrad(x (N,H,W,Cin) bf16, dy bf16)."""
    L = _bind()
    N, H, W, Cin = x.shape
    Cout = dy.shape[3]
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    if dw is None:
        dw = torch.zeros((Cout, Cin, KH, KW) if oihw else (Cout, KH, KW, Cin), device=x.device, dtype=torch.float32)
    d = ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, pad, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    _lib.check(L.c3d_conv2d_wgrad_ex(ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw), int(oihw), _stream()))
    return dw
