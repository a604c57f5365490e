"""box3d_overlap sweep (BASELINE configs[4]): pairs/s on the GPU (CUDA events, device-resident
inputs) for cross N x M = 1k..1M pairs, dense (L=1) and sparse (L=10) regimes."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import boxgen  # noqa: E402
from omni3d_b200 import box3d  # noqa: E402


def time_gpu(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    out = []
    for L, regime in [(1.0, "dense"), (10.0, "sparse")]:
        for n in [32, 100, 320, 1000, 3000]:
            a = torch.from_numpy(boxgen.inject_degenerate(boxgen.random_boxes(n, L, 0), 0.01, 1)[0]).cuda()
            b = torch.from_numpy(boxgen.random_boxes(n, L, 5)).cuda()
            med, best = time_gpu(lambda: box3d.iou_box3d(a, b))
            frac = float((box3d.iou_box3d(a, b)[1] > 0).float().mean())
            rec = {"regime": regime, "N": n, "M": n, "pairs": n * n, "ms_median": med, "ms_min": best,
                   "pairs_per_s": n * n / (med * 1e-3), "overlap_frac": frac,
                   "alg_GBps": (96 * 2 * n + 8 * n * n) / (med * 1e-3) / 1e9}
            out.append(rec); print(json.dumps(rec), flush=True)
    for L, regime in [(1.0, "dense"), (10.0, "sparse")]:
        n = 1_000_000
        a = torch.from_numpy(boxgen.random_boxes(n, L, 0)).cuda()
        b = torch.from_numpy(boxgen.random_boxes(n, L, 5)).cuda()
        med, best = time_gpu(lambda: box3d.iou_box3d_paired(a, b), iters=5)
        rec = {"regime": regime + "_paired", "pairs": n, "ms_median": med, "ms_min": best,
               "pairs_per_s": n / (med * 1e-3), "alg_GBps": 200 * n / (med * 1e-3) / 1e9}
        out.append(rec); print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "iou3d_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
