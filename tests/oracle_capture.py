"""Test helper: run the CPU oracle's training forward while capturing the RNG-dependent decisions
(sampled anchor labels, proposals, sampled proposals) so they can be injected into the CUDA path —
the sampling parity protocol of SURVEY.md section 7 ("inject the oracle's sampled indices")."""
import torch

from oracle import model_io


def run_oracle_train(orc, items, seed=123):
    from detectron2.utils.events import EventStorage
    cap = {}
    pg, rh = orc.proposal_generator, orc.roi_heads
    o_label, o_pred, o_samp = pg.label_and_sample_anchors, pg.predict_proposals, rh.label_and_sample_proposals

    def label(anchors, gt_instances):
        lab, boxes = o_label(anchors, gt_instances)
        cap["anchor_labels"] = torch.stack(lab).clone()
        return lab, boxes

    def pred(*a, **k):
        props = o_pred(*a, **k)
        cap["proposals"] = [(p.proposal_boxes.tensor.clone(), p.objectness_logits.clone()) for p in props]
        return props

    def samp(proposals, targets):
        out = o_samp(proposals, targets)
        cap["sampled"] = [{"boxes": p.proposal_boxes.tensor.clone(), "classes": p.gt_classes.clone(),
                           "gt_boxes": p.gt_boxes.tensor.clone(), "gt_boxes3D": p.gt_boxes3D.clone(),
                           "gt_poses": p.gt_poses.clone()} for p in out]
        return out

    pg.label_and_sample_anchors, pg.predict_proposals, rh.label_and_sample_proposals = label, pred, samp
    try:
        orc.train()
        orc.zero_grad()
        torch.manual_seed(seed)
        with EventStorage(0) as st:
            losses = orc(model_io.to_d2_inputs(items))
            sum(losses.values()).backward()
            scalars = st.latest()
    finally:
        pg.label_and_sample_anchors, pg.predict_proposals, rh.label_and_sample_proposals = o_label, o_pred, o_samp
    return losses, scalars, cap


def to_injection(cap, device, post_k=1000, fcap=128):
    B = len(cap["proposals"])
    pb = torch.zeros((B, post_k, 4)); ps = torch.full((B, post_k), -float("inf")); pc = torch.zeros(B, dtype=torch.int32)
    for i, (b, s) in enumerate(cap["proposals"]):
        n = len(b)
        pb[i, :n], ps[i, :n], pc[i] = b, s, n
    S = max(len(s["classes"]) for s in cap["sampled"])
    smp = {"boxes": torch.zeros((B, S, 4)), "valid": torch.zeros((B, S), dtype=torch.bool),
           "classes": torch.full((B, S), -1, dtype=torch.long), "gt_boxes": torch.zeros((B, S, 4)),
           "gt_boxes3D": torch.zeros((B, S, 9)), "gt_poses": torch.eye(3).repeat(B, S, 1, 1)}
    for i, s in enumerate(cap["sampled"]):
        n = len(s["classes"])
        smp["boxes"][i, :n], smp["classes"][i, :n], smp["valid"][i, :n] = s["boxes"], s["classes"], True
        smp["gt_boxes"][i, :n], smp["gt_boxes3D"][i, :n], smp["gt_poses"][i, :n] = s["gt_boxes"], s["gt_boxes3D"][:, :9], s["gt_poses"]
    smp = {k: v.to(device) for k, v in smp.items()}
    smp["fcap"] = fcap
    return {"anchor_labels": cap["anchor_labels"].to(device),
            "proposals": (pb.to(device), ps.to(device), pc.to(device)), "sampled": smp}
