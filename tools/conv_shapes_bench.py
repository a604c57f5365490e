"""Per-shape timing of the tcgen05 conv kernels on the layer shapes of DLA34_FPN @ 640^2, batch 32
(SURVEY.md section 8a-2 census): CUDA-event ms, algorithmic TFLOP/s.  Also the target of the ncu --set full
captures under profiles/."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from omni3d_b200 import conv as K

STEM_C = 16 if os.environ.get("C3D_CONV_NO_HALO") else 8     # NHWC8 image for the rolling-halo stem kernel
SHAPES = [  # name, H, W, Cin, Cout, k, stride, pad, real_cin
    ("stem7x7_3(%d)->16@640" % STEM_C, 640, 640, STEM_C, 16, 7, 1, 3, 3),
    ("level0_16->16@640", 640, 640, 16, 16, 3, 1, 1, 16),
    ("level1_16->32s2@640", 640, 640, 16, 32, 3, 2, 1, 16),
    ("l2_64->64@160", 160, 160, 64, 64, 3, 1, 1, 64),
    ("l3_128->128@80", 80, 80, 128, 128, 3, 1, 1, 128),
    ("l4_256->256@40", 40, 40, 256, 256, 3, 1, 1, 256),
    ("l5_512->512@20", 20, 20, 512, 512, 3, 1, 1, 512),
    ("root_448->128_1x1@80", 80, 80, 448, 128, 1, 1, 0, 448),
    ("fpn_out_256->256@160", 160, 160, 256, 256, 3, 1, 1, 256),
    ("fpn_out_256->256@80", 80, 80, 256, 256, 3, 1, 1, 256),
    ("fpn_lat_64->256_1x1@160", 160, 160, 64, 256, 1, 1, 0, 64),
]
N = int(os.environ.get("BATCH", "32"))
ITERS = int(os.environ.get("ITERS", "5"))
only = os.environ.get("ONLY")


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS


out = []
for name, H, W, Cin, Cout, k, s, p, rc in SHAPES:
    if only and only not in name:
        continue
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * 0.05).bfloat16()
    Ho, Wo = K.out_hw(H, W, k, k, s, p)
    dy = torch.randn(N, Ho, Wo, Cout, device="cuda", generator=g).bfloat16()
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    flop = 2.0 * N * Ho * Wo * Cout * k * k * rc
    t_f = timeit(lambda: K.conv2d_fwd(x, w, stride=s, pad=p, want_stats=True))
    t_w = timeit(lambda: K.conv2d_wgrad(x, dy, k, k, s, p, dw=dw))
    rec = {"shape": name, "fwd_ms": t_f, "fwd_tflops": flop / t_f / 1e9, "wgrad_ms": t_w, "wgrad_tflops": flop / t_w / 1e9,
           "gflop": flop / 1e9}
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_shapes.json"), "w"), indent=1)
