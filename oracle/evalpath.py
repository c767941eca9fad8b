"""oracle.evalpath -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

CPU restatement of the reference's test-time metric, SURVEY §8f rank 4:
    COOCC_Ray.evaluation_semantic   P/coocc/detectors/coocc_ray.py:659-684
    fast_hist                       P/coocc/detectors/coocc_ray.py:726-730
Pinned against those reference lines executed as they are (oracle/refshim.reference_evaluation_semantic ->
tests/golden/reference_eval.npz, tests/test_oracle_eval.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def fast_hist(pred, label, max_label):
    """coocc_ray.py:726-730: hist[label][pred] by one bincount."""
    n = np.bincount(max_label * label.astype(np.int64).ravel() + pred.astype(np.int64).ravel(), minlength=max_label ** 2)
    return n[:max_label ** 2].reshape(max_label, max_label)


def evaluation_semantic(pred, gt, eval_type, visible_mask=None, empty_idx=0):
    """coocc_ray.py:659-684.  pred [1,C,X,Y,Z] logits, gt [1,H,W,D] labels (255 = noise, excluded)."""
    _, H, W, D = gt.shape
    up = F.interpolate(pred, size=[H, W, D], mode="trilinear", align_corners=False)      # :661
    p = torch.argmax(up[0], dim=0).cpu().numpy()                                        # :662
    g = gt[0].cpu().numpy().astype(np.int64)
    keep = g != 255                                                                      # :667
    if eval_type == "SC":                                                                # :669-673
        return fast_hist((p != empty_idx).astype(np.int64)[keep], (g != empty_idx).astype(np.int64)[keep], 2), None
    if eval_type == "SSC":                                                               # :676-684
        hist_occ = None
        if visible_mask is not None:
            m = keep & (visible_mask[0].cpu().numpy() != 0)
            hist_occ = fast_hist(p[m], g[m], 17)
        return fast_hist(p[keep], g[keep], 17), hist_occ
    raise ValueError(eval_type)
