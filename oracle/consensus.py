"""Oracle: structure-via-consensus segmentation loss (numpy fp64), forward and gradient.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ref tricks/consensus_loss.py:63-168 (`StructureConsensuLossFunction`), the segmentation criterion of the live
training recipe (ref train.py:228-229,258: `seg_criterion(final_seg, msk, msk)` with alpha=10, beta=5, 'idx', 'idx';
SURVEY.md 8f-4).  For every blob id s (sorted unique values of `blobs`, ref :84-86), with M = [blobs == s],
p = softmax(logit, dim=1) (ref :114) and cnt_n = sum_x M[n, x]:

    t[n, c]   = sum_x M p[n, c, x] / cnt_n   ('idx', ref :131-137; 0 where cnt_n == 0)  or  / (H*W)   ('all', ref :138-140)
    y         = target at the first pixel of the blob in (n, h, w) order                              (ref :111-112)
    loss_avg  = mean_n [cnt_n > 0] * -log t[n, y]                                                     (ref :142-148)
    kl        = sum over (n, c, x in blob, p != 0) of t[n, c] * (log t[n, c] - log p[n, c, x])        (ref :153-163)
    loss_dev  = kl / #{(n, c, x) in blob with p != 0}   ('idx', ref :168)   or   kl / (N*H*W)         ('all', ref :166)
    L_s       = alpha * loss_avg + beta * loss_dev;      loss = mean_s L_s                            (ref :170,96)

The gradient is the closed form of what autograd produces for the reference (gradients flow through the blob mean in
BOTH terms: the KL target is not detached, ref :157).
"""
import numpy as np


def _softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def consensus_loss(logit, blobs, target, alpha=10.0, beta=5.0, reduce_pixel="idx", reduce_pixel_kl="idx",
                   want_grad=True, softmax_dtype=np.float64, ids=None):
    """logit (N, C, H, W); blobs, target (N, H, W) or (N, 1, H, W) integer-valued -> (loss, dloss/dlogit or None).

    ``softmax_dtype=np.float32`` reproduces the reference's `p != 0` underflow pattern exactly (the test that needs it
    says so); everything else is fp64.  ``ids`` restricts the blobs that are evaluated (default: every value present,
    ref :84; `ids=[s]` is the reference's per-blob method, ref :98-171)."""
    z = np.asarray(logit, np.float64)
    N, C, H, W = z.shape
    bl = np.asarray(blobs).reshape(N, H, W)
    tg = np.asarray(target).reshape(N, H, W)
    if softmax_dtype == np.float32:
        p = _softmax(np.asarray(logit, np.float32)).astype(np.float64)
    else:
        p = _softmax(z)
    ids = np.unique(bl) if ids is None else np.asarray(ids)
    total = 0.0
    grad_p = np.zeros_like(p)
    for s in ids:
        M = (bl == s)
        lab = tg[M]
        assert np.unique(lab).size == 1, "labels inside a blob must agree (ref :103)"
        y = int(lab[0])
        Mf = M[:, None].astype(np.float64)
        cnt = M.reshape(N, -1).sum(axis=1).astype(np.float64)              # support of the blob per sample
        valid = cnt > 0
        S1 = (p * Mf).sum(axis=(2, 3))                                     # (N, C)
        if reduce_pixel != "all":
            den = np.where(valid, cnt, 1.0)[:, None]
            t = np.where(valid[:, None], S1 / den, 0.0)
        else:
            den = np.full((N, 1), float(H * W))
            t = S1 / den
        with np.errstate(divide="ignore"):
            loss_avg = np.where(valid, -np.log(np.where(valid, t[:, y], 1.0)), 0.0).mean()
        nzm = (p * Mf) != 0                                                # in the blob and not underflowed
        logp = np.where(nzm, np.log(np.where(nzm, p, 1.0)), 0.0)
        nz = nzm.sum(axis=(2, 3)).astype(np.float64)                       # (N, C)
        S2 = logp.sum(axis=(2, 3))
        tlogt = np.where(nz > 0, t * np.log(np.where(nz > 0, t, 1.0)), 0.0)
        kl = (tlogt * nz - t * S2).sum()
        D = nz.sum() if reduce_pixel_kl != "all" else float(N * H * W)
        loss_dev = kl / D
        total += alpha * loss_avg + beta * loss_dev
        if want_grad:
            dLdt = np.zeros((N, C))
            dLdt[valid, y] += alpha / N * (-1.0 / t[valid, y])
            dLdt += np.where(nz > 0, beta / D * ((np.log(np.where(nz > 0, t, 1.0)) + 1.0) * nz - S2), 0.0)
            g = (dLdt / den)[:, :, None, None] * Mf                        # through the blob mean
            g = g - np.where(nzm, beta / D * t[:, :, None, None] / np.where(nzm, p, 1.0), 0.0)   # through log p
            grad_p += g
    loss = total / len(ids)
    if not want_grad:
        return loss, None
    grad_p /= len(ids)
    dz = p * (grad_p - (p * grad_p).sum(axis=1, keepdims=True))            # softmax backward
    return loss, dz
