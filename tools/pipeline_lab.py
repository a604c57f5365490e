"""Lab: where does the time of a pipeline stage go in the tcgen05 conv kernels?  Runs selected layer shapes with parts of the
pipeline switched off (libc3d_lab.so, built by `make -C omni3d_b200/csrc lab`; the product library has none of this):
  C3D_CONV_DBG / C3D_WGRAD_DBG bits: 1 no MMA, 2 no TMA, 4 no B (weights / x) loads, 8 no A (x / dY) loads, 16 no epilogue.
usage: python tools/pipeline_lab.py fwd|wgrad "0,1,2,..."  (results are garbage numerically; only the times matter)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("C3D_LIB_PATH", os.path.join(ROOT, "omni3d_b200", "libc3d_lab.so"))
import torch
from omni3d_b200 import conv as K

kind = sys.argv[1]
modes = [int(m) for m in sys.argv[2].split(",")]
var = "C3D_CONV_DBG" if kind == "fwd" else "C3D_WGRAD_DBG"
SHAPES = [("fpn 256->256 3x3 @160", 160, 256, 256, 3), ("l4 256->256 3x3 @40", 40, 256, 256, 3), ("l5 512->512 3x3 @20", 20, 512, 512, 3),
          ("l3 128->128 3x3 @80", 80, 128, 128, 3), ("l2 64->64 3x3 @160", 160, 64, 64, 3), ("root 448->128 1x1 @80", 80, 448, 128, 1)]
N = 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print("%s  env: %s" % (kind, {k: v for k, v in os.environ.items() if k.startswith("C3D_") and k != "C3D_LIB_PATH"}))
print("%-24s" % "shape" + "".join("%9s" % ("m%d" % m) for m in modes) + "   (ms; m0 = everything on)")
for name, H, Cin, Cout, k in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * 0.05).bfloat16()
    dy = torch.randn(N, H, H, Cout, device="cuda", generator=g).bfloat16()
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    bias = torch.zeros(Cout, device="cuda")
    row = []
    for m in modes:
        os.environ[var] = str(m)
        if kind == "fwd":
            t = timeit(lambda: K.conv2d_fwd(x, w, bias, stride=1, pad=k // 2))
        else:
            t = timeit(lambda: K.conv2d_wgrad(x, dy, k, k, 1, k // 2, dw=dw))
        row.append(t)
    print("%-24s" % name + "".join("%9.3f" % t for t in row), flush=True)
