"""GPU parity tests: libc3d's box3d_overlap / iou_box3d vs the oracle, through the C-ABI.
Bar: bit-exact face counts AND bit-exact vol/iou (same fp32 op order, no FMA contraction)."""
import os

import numpy as np
import pytest
import torch

import boxgen
from oracle import iou3d as oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests/golden/iou3d_wrapper_golden.npz"))


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("N,M,L,seed", [(1, 1, 1.0, 0), (7, 3, 1.0, 1), (33, 65, 1.0, 2), (100, 100, 1.0, 3),
                                         (128, 96, 3.0, 4), (100, 100, 10.0, 5), (257, 31, 2.0, 6)])
def test_cross_bit_exact(N, M, L, seed):
    from omni3d_b200 import box3d
    a, b = boxgen.random_boxes(N, L, seed), boxgen.random_boxes(M, L, seed + 77)
    vol, iou, nf = box3d.iou_box3d(_gpu(a), _gpu(b), with_counts=True)
    vo, uo, nfo, _ = oracle.iou_box3d(a, b, threads=os.cpu_count(), with_counts=True)
    assert np.array_equal(nf.cpu().numpy(), nfo), "face counts differ"
    assert np.array_equal(vol.cpu().numpy(), vo)
    assert np.array_equal(iou.cpu().numpy(), uo)


def test_known_answers_gpu():
    from omni3d_b200 import box3d
    b = GOLD["ka_boxes"]
    vol, iou, nf = box3d.iou_box3d(_gpu(b[:1]), _gpu(b), with_counts=True)
    np.testing.assert_allclose(iou.cpu().numpy()[0], GOLD["ka_expect"], atol=2e-6)
    assert int(nf[0, 5]) == 0


@pytest.mark.parametrize("name", ["dense", "sparse", "mid"])
def test_wrapper_matches_reference_fixture(name, capsys):
    from omni3d_b200 import box3d
    dt, gt = GOLD[f"{name}_dt"], GOLD[f"{name}_gt"]
    iou, bad = box3d.box3d_overlap(_gpu(dt), _gpu(gt), return_bad_counts=True)
    assert np.array_equal(iou.cpu().numpy(), GOLD[f"{name}_iou"])
    assert bad[0] == int((~GOLD[f"{name}_coplanar_ok"]).sum())
    assert bad[1] == int((~GOLD[f"{name}_nonzero_ok"]).sum())
    out = capsys.readouterr().out
    assert "non-coplanar boxes at eval" in out and "zero volume boxes at eval" in out


def test_cpu_tensor_inputs_return_cpu():
    from omni3d_b200 import box3d
    dt, gt = GOLD["mid_dt"], GOLD["mid_gt"]
    iou = box3d.box3d_overlap(torch.from_numpy(dt), torch.from_numpy(gt))
    assert iou.device.type == "cpu" and np.array_equal(iou.numpy(), GOLD["mid_iou"])


def test_empty_and_ragged():
    from omni3d_b200 import box3d
    a = _gpu(boxgen.random_boxes(5, 1.0, 0))
    e = torch.zeros((0, 8, 3), device="cuda")
    assert tuple(box3d.box3d_overlap(a, e).shape) == (5, 0)
    assert tuple(box3d.box3d_overlap(e, a).shape) == (0, 5)
    assert tuple(box3d.iou_box3d(e, e)[1].shape) == (0, 0)
    with pytest.raises(ValueError):
        box3d.box3d_overlap(torch.zeros(3, 8, 2, device="cuda"), a)


def test_paired_bit_exact():
    from omni3d_b200 import box3d
    a, b = boxgen.random_boxes(5000, 1.0, 11), boxgen.random_boxes(5000, 1.0, 12)
    vol, iou, nf = box3d.iou_box3d_paired(_gpu(a), _gpu(b), with_counts=True)
    vo, uo, nfo = oracle.iou_box3d_paired(a, b, threads=os.cpu_count())
    assert np.array_equal(nf.cpu().numpy(), nfo)
    assert np.array_equal(vol.cpu().numpy(), vo) and np.array_equal(iou.cpu().numpy(), uo)


def test_overflow_fallback_path():
    """Pairs whose clipped lists exceed the shared-memory capacity take the global-memory path;
    the 45-degree cube (43 tris/side) is fine, so stress with many near-identical rotated copies."""
    from omni3d_b200 import box3d
    rng = np.random.default_rng(0)
    a = boxgen.random_boxes(400, 0.3, 21)
    b = boxgen.random_boxes(400, 0.3, 22)
    vol, iou, nf = box3d.iou_box3d(_gpu(a), _gpu(b), with_counts=True)
    vo, uo, nfo, ns = oracle.iou_box3d(a, b, threads=os.cpu_count(), with_counts=True)
    assert np.array_equal(nf.cpu().numpy(), nfo)
    assert np.array_equal(iou.cpu().numpy(), uo)
    # report (not assert) whether the overflow path was exercised
    print("pairs with >64 tris on a side:", int((ns.max(-1) > 64).sum()))


def test_full_size_properties_1m_pairs():
    """BASELINE configs[4] top size (1000 x 1000): size-independent properties."""
    from omni3d_b200 import box3d
    a, b = boxgen.random_boxes(1000, 10.0, 0), boxgen.random_boxes(1000, 10.0, 5)
    ga, gb = _gpu(a), _gpu(b)
    vol, iou, nf = box3d.iou_box3d(ga, gb, with_counts=True)
    iou_n = iou.cpu().numpy(); nf_n = nf.cpu().numpy(); vol_n = vol.cpu().numpy()
    assert np.isfinite(iou_n).all() and iou_n.min() >= 0 and iou_n.max() <= 1 + 1e-6
    assert ((nf_n == 0) == (vol_n == 0)).all()
    # disjoint bounding spheres => exactly zero
    c1, c2 = a.mean(1), b.mean(1)
    r1 = np.linalg.norm(a - c1[:, None], axis=2).max(1); r2 = np.linalg.norm(b - c2[:, None], axis=2).max(1)
    far = np.linalg.norm(c1[:, None] - c2[None], axis=2) > (r1[:, None] + r2[None]) * 1.01
    assert (iou_n[far] == 0).all()
    # symmetry up to fp32 rounding for all but a handful of pairs (the reference algorithm's coplanar
    # de-dup is itself asymmetric: the CPU oracle shows the same 4 outliers of 1e6 on this input)
    _, iou_t = box3d.iou_box3d(gb, ga)
    assert ((iou_t.t() - iou).abs() > 5e-5).float().mean().item() < 1e-4
    _, self_iou = box3d.iou_box3d_paired(ga, ga)
    assert (self_iou - 1).abs().max().item() < 1e-5
    # sampled rows against the oracle, bit-exact
    rows = np.arange(0, 1000, 97)
    vo, uo, nfo, _ = oracle.iou_box3d(a[rows], b, threads=os.cpu_count(), with_counts=True)
    assert np.array_equal(nf_n[rows], nfo) and np.array_equal(iou_n[rows], uo)


def test_determinism():
    from omni3d_b200 import box3d
    a, b = _gpu(boxgen.random_boxes(300, 1.0, 1)), _gpu(boxgen.random_boxes(300, 1.0, 2))
    r1 = box3d.iou_box3d(a, b, with_counts=True)
    r2 = box3d.iou_box3d(a, b, with_counts=True)
    for x, y in zip(r1, r2):
        assert torch.equal(x, y)


# ---- all (image, category) groups of Omni3Deval.computeIoU in one segmented launch ---------------------
def _groups(seed, G=40):
    rng = np.random.default_rng(seed)
    dts, gts = [], []
    for g in range(G):
        n, m = int(rng.integers(0, 40)), int(rng.integers(0, 12))
        if g == 3:
            n, m = 0, 5
        if g == 4:
            n, m = 6, 0
        if g == 5:
            n, m = 0, 0
        if g == 6:
            n, m = 100, 30
        L = float(rng.choice([1.0, 3.0, 10.0]))
        a = boxgen.random_boxes(n, L, seed * 1000 + g) if n else np.zeros((0, 8, 3), np.float32)
        if n >= 8:
            a, _ = boxgen.inject_degenerate(a, 0.1, g)
        dts.append(a)
        gts.append(boxgen.random_boxes(m, L, seed * 1000 + g + 500) if m else np.zeros((0, 8, 3), np.float32))
    return dts, gts


@pytest.mark.parametrize("seed", [0, 1])
def test_segmented_overlap_equals_per_group_oracle(seed, capsys):
    from omni3d_b200 import evaluation as ev
    dts, gts = _groups(seed)
    res, bad = ev.box3d_overlap_segmented(dts, gts, return_bad_counts=True)
    nb = np.zeros(2, np.int64)
    for a, b, r in zip(dts, gts, res):
        assert r.shape == (len(a), len(b)) and r.dtype == np.float32
        if len(a) and len(b):
            ref, k = oracle.box3d_overlap(a, b)
            assert np.array_equal(r, ref)
        if len(a):
            c, z = oracle.check_boxes(a)
            nb += [int((~c).sum()), int((~z).sum())]
    assert list(bad) == list(nb) and nb.min() > 0
    out = capsys.readouterr().out
    assert "non-coplanar boxes at eval" in out and "zero volume boxes at eval" in out


def test_compute_ious_3d_matches_reference_call_pattern():
    """== {(imgId, catId): computeIoU(imgId, catId)} of omni3d_evaluation.py:1339-1343 with box3d_overlap = the oracle:
    score-sorted (stable), truncated to maxDets, [] for empty groups."""
    from omni3d_b200 import evaluation as ev
    rng = np.random.default_rng(7)
    dts, gts = {}, {}
    imgs, cats = [11, 12, 13], [0, 1, 2]
    for i in imgs:
        for c in cats:
            n, m = int(rng.integers(0, 9)), int(rng.integers(0, 4))
            if (i, c) == (12, 1):
                n, m = 0, 0
            a = boxgen.random_boxes(n, 2.0, i * 10 + c) if n else []
            b = boxgen.random_boxes(m, 2.0, i * 10 + c + 7) if m else []
            sc = np.round(rng.uniform(0, 1, n), 1)                      # ties: the merge sort keeps input order
            dts[i, c] = [{"score": float(s), "bbox3D": a[k].tolist(), "bbox": [0, 0, 1, 1]} for k, s in enumerate(sc)]
            gts[i, c] = [{"bbox3D": b[k].tolist(), "bbox": [0, 0, 1, 1]} for k in range(m)]
    got = ev.compute_ious_3d(dts, gts, imgs, cats, max_dets=5)
    for i in imgs:
        for c in cats:
            dt, gt = dts[i, c], gts[i, c]
            if not dt and not gt:
                assert got[i, c] == []
                continue
            inds = np.argsort([-d["score"] for d in dt], kind="mergesort")
            dt = [dt[k] for k in inds][:5]
            ious, prox = got[i, c]
            assert prox is None
            if dt and gt:
                ref, _ = oracle.box3d_overlap(np.array([d["bbox3D"] for d in dt], np.float32),
                                              np.array([g["bbox3D"] for g in gt], np.float32))
                assert np.array_equal(ious, ref)
            else:
                assert ious == []
