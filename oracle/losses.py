"""oracle.losses -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

CPU restatement (torch-CPU fp32 / int64) of the occupancy head's voxel losses, SURVEY §8f rank 1:
    OccHead.loss_voxel        P/coocc/dense_heads/occ_head.py:267-293
    CE_ssc_loss               P/utils/semkitti.py:139-149
    sem_scal_loss             P/utils/semkitti.py:92-136
    geo_scal_loss             P/utils/semkitti.py:62-89
    lovasz_softmax (+ _flat, lovasz_grad, flatten_probas)
                              P/coocc/dense_heads/lovasz_softmax.py:20-34, 156-229
    class weights             occ_head.py:135-138 with P/utils/nusc_param.py:10-12

Pinned against the unmodified reference functions executed in the build container
(oracle/make_golden.py -> tests/golden/reference_losses.npz, tests/test_oracle_losses.py); the
reference has no tests of its own for them.  Differentiable through torch autograd, so
the GPU kernels' gradients are checked against it as well.
"""
import math

import torch
import torch.nn.functional as F

# P/utils/nusc_param.py:10-12 (class 0 = free)
NUSC_CLASS_FREQUENCIES = [2242961742295, 25985376, 1561108, 28862014, 196106643, 15920504, 2158753, 26539491,
                          4004729, 34838681, 75173306, 2255027978, 50959399, 646022466, 869055679, 1446141335,
                          1724391378]


def class_weights():
    """occ_head.py:137: 1 / log(freq + 0.001), evaluated in float64 and used in the logits' dtype (:289)."""
    return torch.tensor([1.0 / math.log(f + 0.001) for f in NUSC_CLASS_FREQUENCIES], dtype=torch.float64)


def downsample_labels(target, H, empty_idx=0):
    """occ_head.py:269-280.  target [B, H*r, W*r, D*r] integer labels -> int64 [B,H,W,D].

    Each r^3 cell votes; in a cell whose labels do not sum to `empty_idx` every 0 is first replaced by
    a value that occurs nowhere else (:274-276), so zeros never form a majority there; torch.mode
    returns the smallest of the most frequent values (:277); negative winners become 255 (:278)."""
    B, HH, WW, DD = target.shape
    r = target.shape[2] // H          # :271 -- the reference divides the label grid's *second* spatial extent by the
    if r == 1:                        # output's first one (equal for the square grids of every config)
        return target.long()
    W, D = WW // r, DD // r
    cells = target.reshape(B, H, r, W, r, D, r).permute(0, 1, 3, 5, 2, 4, 6).reshape(B, H, W, D, r ** 3).long()
    empty = cells.sum(-1) == empty_idx
    flat = cells.reshape(-1, r ** 3).clone()
    # zeros of the non-empty cells -> unique negatives; numbering them 1..n in row-major order is what
    # the reference's masked assignment does, any injective choice gives the same vote
    z = (flat == 0) & (~empty.reshape(-1, 1))
    flat[z] = -torch.arange(1, int(z.sum()) + 1, device=flat.device)
    srt, _ = torch.sort(flat, dim=-1)
    n = r ** 3
    # run lengths in the sorted row; the first longest run is the smallest most-frequent value
    best_val = srt[:, 0].clone()
    best_cnt = torch.zeros(len(srt), dtype=torch.long, device=srt.device)
    run_val = srt[:, 0].clone()
    run_cnt = torch.zeros(len(srt), dtype=torch.long, device=srt.device)
    for i in range(n):
        same = srt[:, i] == run_val
        run_cnt = torch.where(same, run_cnt + 1, torch.ones_like(run_cnt))
        run_val = srt[:, i]
        better = run_cnt > best_cnt
        best_cnt = torch.where(better, run_cnt, best_cnt)
        best_val = torch.where(better, run_val, best_val)
    out = best_val.reshape(B, H, W, D)
    out[out < 0] = 255
    return out


def _valid(pred, target, ignore_index):
    """[B,C,...] logits, [B,...] labels -> (softmax probabilities [P,C], labels [P]) of the non-ignored voxels."""
    C = pred.shape[1]
    p = F.softmax(pred, dim=1).movedim(1, -1).reshape(-1, C)
    t = target.reshape(-1)
    keep = t != ignore_index
    return p[keep], t[keep]


def ce_ssc_loss(pred, target, weight=None, ignore_index=255):
    """semkitti.py:139-149: class-weighted mean cross-entropy over the non-ignored voxels."""
    return F.cross_entropy(pred, target.long(), weight=weight, ignore_index=ignore_index, reduction="mean")


def _bce_to_one(x):
    """F.binary_cross_entropy(x, ones_like(x)) for a scalar: -max(log x, -100)."""
    return -torch.clamp(torch.log(x), min=-100.0)


def sem_scal_loss(pred, target, ignore_index=255):
    """semkitti.py:92-136: for every class with at least one (non-ignored) voxel, BCE-to-one of its soft
    precision, recall and specificity; mean over those classes."""
    p, t = _valid(pred, target, ignore_index)
    loss, count = 0.0, 0.0
    for c in range(p.shape[1]):
        is_c = (t == c).to(p.dtype)
        n_c = is_c.sum()
        if n_c > 0:                                             # :114
            count += 1.0
            nom = (p[:, c] * is_c).sum()                        # :116
            if p[:, c].sum() > 0:                               # :118-123
                loss = loss + _bce_to_one(nom / p[:, c].sum())
            loss = loss + _bce_to_one(nom / n_c)                # :124-127
            if (1 - is_c).sum() > 0:                            # :128-134
                loss = loss + _bce_to_one(((1 - p[:, c]) * (1 - is_c)).sum() / (1 - is_c).sum())
    return loss / count


def geo_scal_loss(pred, target, ignore_index=255, non_empty_idx=0):
    """semkitti.py:62-89: the same three terms for the binary empty / non-empty split, eps = 1e-5."""
    p, t = _valid(pred, target, ignore_index)
    empty_p = p[:, non_empty_idx]
    nonempty_p = 1 - empty_p
    nonempty_t = (t != non_empty_idx).to(p.dtype)
    eps = 1e-5
    inter = (nonempty_t * nonempty_p).sum()
    precision = inter / (nonempty_p.sum() + eps)
    recall = inter / (nonempty_t.sum() + eps)
    spec = ((1 - nonempty_t) * empty_p).sum() / ((1 - nonempty_t).sum() + eps)
    return _bce_to_one(precision) + _bce_to_one(recall) + _bce_to_one(spec)


def lovasz_grad(fg_sorted):
    """lovasz_softmax.py:20-34."""
    n = len(fg_sorted)
    gts = fg_sorted.sum()
    inter = gts - fg_sorted.float().cumsum(0)
    union = gts + (1 - fg_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    if n > 1:
        jac[1:n] = jac[1:n] - jac[0:-1]
    return jac


def lovasz_softmax(pred, target, ignore=255):
    """lovasz_softmax.py:156-203 with classes='present' on softmax(pred) (occ_head.py:292): per class with
    foreground, sort |fg - p_c| descending and dot it with the (detached) Lovasz gradient; mean."""
    p, t = _valid(pred, target, ignore)
    if p.numel() == 0:
        return p.sum() * 0.0
    losses = []
    for c in range(p.shape[1]):
        fg = (t == c).float()
        if fg.sum() == 0:
            continue
        err = (fg - p[:, c]).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        losses.append(torch.dot(err_sorted, lovasz_grad(fg[perm]).detach()))
    return sum(losses) / len(losses)


def loss_voxel(output_voxels, target_voxels, tag="c_0", weights=(1.0, 1.0, 1.0, 1.0), empty_idx=0, balance=True):
    """occ_head.py:267-293.  output_voxels [B,C,H,W,D] logits, target_voxels [B,H*r,W*r,D*r] labels."""
    tv = downsample_labels(target_voxels, output_voxels.shape[2], empty_idx)
    cw = class_weights().to(output_voxels.dtype) if balance else torch.ones(17, dtype=output_voxels.dtype) / 17
    cw = cw.to(output_voxels.device)
    return {
        "loss_voxel_ce_%s" % tag: weights[0] * ce_ssc_loss(output_voxels, tv, cw, 255),
        "loss_voxel_sem_scal_%s" % tag: weights[1] * sem_scal_loss(output_voxels, tv, 255),
        "loss_voxel_geo_scal_%s" % tag: weights[2] * geo_scal_loss(output_voxels, tv, 255, empty_idx),
        "loss_voxel_lovasz_%s" % tag: weights[3] * lovasz_softmax(output_voxels, tv, 255),
    }
