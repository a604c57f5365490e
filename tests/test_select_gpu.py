"""GPU tests of the selection / sampling kernels (omni3d_b200/csrc/select_ops.cu): sorted top-k segments vs torch.topk,
the fused proposal labelling + sampling vs the ORACLE's pre-sampling labels (bit-exact) and its sampling invariants /
distribution, and the fused anchor sampler."""
import pytest
import torch

from omni3d_b200 import synth

pytestmark = pytest.mark.gpu


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("B,shapes", [(3, [(76800, 2000), (19200, 2000), (4800, 2000), (1200, 1200), (300, 300)]),
                                      (2, [(7500, 7500)]), (4, [(100, 256)]), (2, [(102300, 256), (102300, 256)]),
                                      (1, [(50000, 8192)]), (5, [(17, 1), (33, 5)])])
def test_topk_segments_vs_torch(B, shapes):
    from omni3d_b200 import kernels as Kx
    g = _gen(1)
    vals = [torch.randn(B, n, device="cuda", generator=g) for n, _ in shapes]
    vals[0][0, ::7] = -float("inf")                           # masked candidates
    if shapes[0][0] > 64:
        vals[0][-1, 5:40] = 1.25                              # a run of exact ties
    big = torch.randn(B, sum(n for n, _ in shapes) + 8, device="cuda", generator=g)   # row-strided input
    view = big[:, 3:3 + shapes[0][0]]
    view.copy_(vals[0]); vals[0] = view
    out_v, out_i, cnt = Kx.topk_segments([(v, k) for v, (_, k) in zip(vals, shapes)], want_idx64=True, want_counts=True)
    col = 0
    for s, (v, (n, k)) in enumerate(zip(vals, shapes)):
        kk = min(n, k)
        rv, _ = v.topk(kk, dim=1)
        gv, gi = out_v[:, col:col + k], out_i[:, col:col + k]
        assert torch.equal(gv[:, :kk], rv), (n, k)            # identical value sequence (sorted descending)
        assert torch.equal(torch.gather(v, 1, gi[:, :kk]), gv[:, :kk])
        for b in range(B):                                    # a selection, not a multiset: indices are distinct
            assert gi[b, :kk].unique().numel() == kk
        if k > n:
            assert torch.isinf(gv[:, n:]).all() and (gv[:, n:] < 0).all()
        assert torch.equal(cnt[:, s].long(), torch.isfinite(rv).sum(1))
        # ties come out in ascending index order
        eq = gv[:, :-1] == gv[:, 1:]
        eq[:, kk - 1:] = False
        assert (gi[:, :-1][eq] < gi[:, 1:][eq]).all()
        col += k
    gv32, gi32 = Kx.topk_segments([(vals[0], shapes[0][1])])
    assert gi32.dtype == torch.int32 and torch.equal(gi32.long(), out_i[:, :shapes[0][1]])


def _gt_cuda(items):
    from omni3d_b200.cubercnn.model import collate_gt
    return collate_gt(items, torch.device("cuda"))


def test_label_sample_prelabels_equal_oracle_and_torch_formulation():
    """the kernel's matcher / ignore rule / class labels == the oracle's label_and_sample_proposals right before its
    sampling (bit-exact) == the product's batched torch formulation (ROIHeads3D.match_proposals)."""
    from omni3d_b200 import cubercnn as pc
    from omni3d_b200 import kernels as Kx
    from omni3d_b200.cubercnn.roi_heads import ROIHeads3D
    from oracle import cubercnn_oracle as co
    import test_glue_cpu as T
    pcfg, ocfg = T._cfgs()
    torch.manual_seed(0)
    orc = co.build_model(ocfg)
    RH = pcfg.MODEL.ROI_HEADS
    K, thr, ithr = RH.NUM_CLASSES, RH.IOU_THRESHOLDS[0], pcfg.MODEL.RPN.IGNORE_THRESHOLD
    rh = ROIHeads3D.__new__(ROIHeads3D)
    rh.num_classes, rh.iou_thresh, rh.ignore_thresh = K, thr, ithr
    P = 96
    for seed in (3, 4, 5):
        items = T._label_case(seed)
        boxes, counts = T.synthetic_proposals(items, P, seed)
        cap, sampled, d2 = T.oracle_prelabels(orc, items, boxes, counts)
        gt = _gt_cuda(items)
        out = Kx.label_sample_proposals(boxes.cuda(), counts.cuda(), gt, K, 64, 16, thr, ithr, want_prelabels=True,
                                        want_index=True, rng=Kx.rng_state(torch.device("cuda"), seed=seed))
        midx, miou, cls = out["pre"]
        T.check_prelabels(midx.cpu(), miou.cpu(), cls.cpu(), {k: v.cpu() for k, v in gt.items()}, cap, d2, counts, P)
        # == the torch formulation on the same device
        pvalid = torch.arange(P, device="cuda")[None] < counts.cuda()[:, None]
        allb = torch.cat([boxes.cuda(), gt["boxes"]], 1)
        allv = torch.cat([pvalid, gt["present"] & (gt["classes"] >= 0)], 1)
        m2, i2, c2 = ROIHeads3D.match_proposals(rh, allb, allv, gt)
        assert torch.equal(c2, cls) and torch.equal(i2, miou)
        assert torch.equal(m2[c2 >= 0], midx[c2 >= 0])
        # sampling invariants (roi_heads.py:826-860): quotas, fg first, distinct picks, fields gathered from the match
        S, Fcap = 64, 16
        for i in range(len(items)):
            lab = cls[i]
            nfg_all, nbg_all = int(((lab >= 0) & (lab < K)).sum()), int((lab == K).sum())
            nfg, nbg = min(nfg_all, Fcap), min(nbg_all, S - min(nfg_all, Fcap))
            v = out["valid"][i]
            assert int(v.sum()) == nfg + nbg and v[:nfg + nbg].all()
            sc, ix = out["classes"][i], out["index"][i]
            assert ((sc[:nfg] >= 0) & (sc[:nfg] < K)).all() and (sc[nfg:nfg + nbg] == K).all() and (sc[nfg + nbg:] == -1).all()
            assert ix[:nfg + nbg].unique().numel() == nfg + nbg
            assert torch.equal(sc[:nfg + nbg], lab[ix[:nfg + nbg]])
            assert torch.equal(out["boxes"][i, :nfg + nbg], allb[i][ix[:nfg + nbg]])
            mi = midx[i][ix[:nfg + nbg]]
            assert torch.equal(out["gt_boxes"][i, :nfg + nbg], gt["boxes"][i][mi])
            assert torch.equal(out["gt_boxes3D"][i, :nfg + nbg], gt["boxes3D"][i][mi][:, :9])
            assert torch.equal(out["gt_poses"][i, :nfg + nbg], gt["poses"][i][mi])
            # the oracle (quota 512 / 128 here) sampled from the same candidate sets
            osm = sampled[i].gt_classes
            assert int((osm < K).sum()) == min(nfg_all, 128) and int((osm == K).sum()) == min(nbg_all, 512 - min(nfg_all, 128))
        assert float(out["stats"][0]) == sum(min(int(((cls[i] >= 0) & (cls[i] < K)).sum()), Fcap) for i in range(len(items)))


def test_label_sample_distribution_is_iou_weighted():
    """one foreground slot drawn for 4096 identical images (independent Philox streams): the pick frequencies follow
    (IoU + 1e-4) / sum — the first draw of torch.multinomial(weights) (rpn.py:317-324) — and successive launches differ."""
    from omni3d_b200 import kernels as Kx
    B = 4096
    gtb = torch.tensor([[10.0, 10.0, 110.0, 110.0]], device="cuda")
    props = torch.tensor([[10.0, 10.0, 110.0, 110.0], [10.0, 10.0, 110.0, 90.0], [10.0, 10.0, 110.0, 70.0], [10.0, 10.0, 110.0, 65.0],
                          [200.0, 200.0, 240.0, 240.0], [300.0, 300.0, 340.0, 340.0]], device="cuda")
    gt = {"boxes": gtb[None].repeat(B, 1, 1), "classes": torch.full((B, 1), 7, device="cuda"),
          "present": torch.ones((B, 1), dtype=torch.bool, device="cuda"), "boxes3D": torch.zeros((B, 1, 9), device="cuda"),
          "poses": torch.eye(3, device="cuda").repeat(B, 1, 1, 1)}
    pb = props[None].repeat(B, 1, 1)
    pc = torch.full((B,), 6, dtype=torch.int32, device="cuda")
    rng = Kx.rng_state(torch.device("cuda"), seed=1234)
    out = Kx.label_sample_proposals(pb, pc, gt, 50, 4, 1, 0.5, 0.5, append_gt=False, rng=rng, want_prelabels=True, want_index=True)
    iou = out["pre"][1][0, :4]
    assert (out["pre"][2][0, :4] == 7).all() and (out["pre"][2][0, 4:] == 50).all()
    w = (iou + 1e-4) / (iou + 1e-4).sum()
    pick = out["index"][:, 0]
    freq = torch.bincount(pick, minlength=6)[:4].float() / B
    assert (freq - w).abs().max().item() < 4 * (0.25 / B) ** 0.5 + 0.01, (freq, w)
    # the two background boxes fill the next slots in random order
    assert set(out["index"][:, 1].unique().tolist()) == {4, 5}
    step0 = int(rng[1])
    out2 = Kx.label_sample_proposals(pb, pc, gt, 50, 4, 1, 0.5, 0.5, append_gt=False, rng=rng, want_index=True)
    assert int(rng[1]) == step0 + 1 and not torch.equal(out2["index"][:, 0], pick)


def test_anchor_sample_kernel_semantics_and_distribution():
    from omni3d_b200 import cubercnn as pc
    from omni3d_b200 import kernels as Kx
    from omni3d_b200.cubercnn import rpn as prpn
    items = synth.make_batch(4, 256, 320, num_gt=6, seed=11)
    gt = _gt_cuda(items)
    cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none"])
    strides = [4, 8, 16, 32, 64]
    ag = prpn.AnchorGenerator(cfg.MODEL.ANCHOR_GENERATOR.SIZES, cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS, strides)
    anchors = torch.cat(ag([(64, 80), (32, 40), (16, 20), (8, 10), (4, 5)], torch.device("cuda")))
    valid = gt["present"] & (gt["classes"] >= 0)
    ign = gt["present"] & (gt["classes"] < 0)
    idx, miou, lab, ioa, best_idx = Kx.anchor_match(anchors, gt["boxes"], valid, ign, 0.05)
    n_total, cap = 256, 256
    rng = Kx.rng_state(torch.device("cuda"), seed=5)
    out = Kx.anchor_sample(lab, miou, ioa, best_idx, valid, ign, n_total, cap, 0.5, rng=rng)
    A = anchors.shape[0]
    for b in range(4):
        npos_all, nneg_all = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        npos, nneg = min(npos_all, cap), min(nneg_all, n_total - min(npos_all, cap))
        o = out[b]
        pos, zero = o == 1, o == 0
        best = torch.zeros(A, dtype=torch.bool, device="cuda")
        best[best_idx[b][valid[b]].long()] = True
        best &= lab[b] == 1
        assert (lab[b][pos] == 1).all() and (lab[b][zero] == 0).all()
        assert npos <= int(pos.sum()) <= npos + int(best.sum())
        assert (o[best] == 1).all()
        # sampled negatives: 0, or -1 when inside an ignore region (rule active with > 1 sampled negatives)
        hit = (lab[b] == 0) & (ioa[b] >= 0.5)
        assert int(zero.sum()) <= nneg and int(zero.sum()) >= nneg - int(hit.sum())
        assert not (zero & hit).any() or nneg <= 1 or not bool(ign[b].any())
    out2 = Kx.anchor_sample(lab, miou, ioa, best_idx, valid, ign, n_total, cap, 0.5, rng=rng)
    assert not torch.equal(out, out2)                      # the kernel advanced the Philox step counter
    # == the torch formulation's invariants on the same inputs (counts of each label per image)
    head = prpn.RPNWithIgnore.__new__(prpn.RPNWithIgnore)
    head.iou_thresholds, head.batch_size_per_image, head.positive_fraction = [0.05, 0.05], n_total, 1.0
    head.ignore_thresh, head.generator, head.fused_sampling = 0.5, None, False
    ref, _ = prpn.RPNWithIgnore.label_and_sample_anchors(head, anchors, gt["boxes"], gt["classes"], gt["present"])
    # same quotas: sampled positives (+ forced best anchors) and sampled negatives (minus ignore hits) per image
    for b in range(4):
        npos = min(int((lab[b] == 1).sum()), cap)
        assert npos <= int((ref[b] == 1).sum()) <= npos + int(valid[b].sum())
        assert abs(int((ref[b] >= 0).sum()) - int((out[b] >= 0).sum())) <= int(valid[b].sum()) + int(((lab[b] == 0) & (ioa[b] >= 0.5)).sum())


# ---- batched inference post-processing == the oracle's per-image fast_rcnn_inference_single_image ---------
def _rand_dets(B, P, K, seed, spread, peaked, clustered=False):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(B, P, 1, 2, generator=g) * torch.tensor([300.0, 200.0])
    wh = torch.rand(B, P, K, 2, generator=g) * spread + 8
    c = ctr + (torch.rand(B, P, K, 2, generator=g) - 0.5) * 20
    if clustered:            # every box of a class nearly coincides: the greedy NMS keeps one or two per class
        wh = torch.rand(B, P, K, 2, generator=g) * 6 + 40
        c = torch.tensor([150.0, 100.0]) + (torch.rand(B, P, K, 2, generator=g) - 0.5) * 4
    boxes = torch.cat([c - wh / 2, c + wh / 2], -1)                     # some stick out of the image: clipped
    logits = torch.randn(B, P, K + 1, generator=g) * (3.0 if peaked else 0.3)
    return torch.softmax(logits, -1), boxes


@pytest.mark.parametrize("P,K,peaked,max_cand", [(200, 50, False, 8192), (60, 50, True, 8192), (300, 20, False, 512),
                                                 (1000, 50, False, 8192)])
def test_select_detections_equals_oracle_inference(P, K, peaked, max_cand):
    """identical kept (box, score, class) lists, in order, for every image — including the exact-rounds path (max_cand 512:
    heavy suppression pushes the 100th survivor beyond the top-M candidates) and a proposal with a non-finite box."""
    from omni3d_b200.cubercnn.roi_heads import ROIHeads3D
    from oracle.cubercnn_oracle.model import FastRCNNOutputs
    B = 3
    probs, boxes = _rand_dets(B, P, K, 5, 120.0, peaked, clustered=max_cand < 8192)
    boxes[1, 7, 3, 2] = float("nan")
    counts = torch.tensor([P, P - 5, P], dtype=torch.int32)
    sizes = [(200, 300)] * B
    rh = ROIHeads3D.__new__(ROIHeads3D)
    rh.num_classes, rh.test_topk, rh.test_score_thresh, rh.test_nms_thresh = K, 100, 0.01, 0.5
    rh.nms_trick_max_numel = 4000                                       # the CPU oracle runs torchvision's CPU threshold
    hw = torch.tensor(sizes, dtype=torch.float32)
    det = ROIHeads3D.select_detections(rh, probs.cuda(), boxes.cuda(), counts.cuda(), hw.cuda(), max_candidates=max_cand)
    orc = FastRCNNOutputs.__new__(FastRCNNOutputs)
    orc.test_score_thresh, orc.test_nms_thresh, orc.test_topk_per_image = 0.01, 0.5, 100
    used_rounds = False
    for i in range(B):
        n = int(counts[i])
        res, _ = FastRCNNOutputs._inference_one(orc, boxes[i, :n].reshape(n, -1), probs[i, :n], sizes[i])
        c = det["counts_host"][i]
        assert c == len(res), (i, c, len(res))
        assert torch.equal(det["scores"][i, :c].cpu(), res.scores)
        assert torch.equal(det["classes"][i, :c].cpu(), res.pred_classes)
        assert torch.equal(det["boxes"][i, :c].cpu(), res.pred_boxes.tensor)
        assert torch.equal(det["scores_full"][i, :c].cpu(), res.scores_full)
        # the kept pairs really are (proposal, class) entries of the input
        pi, ci = det["prop"][i, :c].cpu(), det["classes"][i, :c].cpu()
        assert torch.equal(probs[i][pi, ci], res.scores)
        used_rounds |= bool((probs[i, :n, :K] > 0.01).sum() > max_cand and c < 100)
    if max_cand == 512:
        assert used_rounds           # fewer than 100 survivors although more candidates than M: the exact-rounds path ran
