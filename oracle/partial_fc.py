"""Oracle: PartialFC class-sharded head, all ranks simulated in one process (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ref headers/partial_fc.py:
  :34-36    shard geometry (num_local, class_start, num_sample)
  :77-94    sample(): label remap, positive set, perm/topk/sort, searchsorted
  :106-116  prepare(): all_gather labels, sample, normalize(sub_weight)
  :118-177  forward_backward(): all_gather features, logits, margin, distributed softmax,
            label smoothing (eps 0.1, local shard only), loss, grad, backward, reduce_scatter
plus the margin callable (oracle/margins.py).

Integer parts (shards, remap, index) are exact.  The random draw is an INPUT here
(``perm`` = the tensor ``torch.rand(num_local)`` returned in the ref):
bit-exact sampling is defined relative to the same draw.  Tie rule at the k-th value:
lowest class index wins (torch.topk leaves ties unspecified; tests detect a boundary
tie in the ref vectors and compare as sets there).
"""
import numpy as np

from .margins import l2_normalize, margin_apply, margin_dcos, normalize_bwd

EPSILON = 0.1  # ref partial_fc.py:154


def shard_geometry(num_classes, world_size, rank, sample_rate=1.0):
    """ref partial_fc.py:34-36 -> (num_local, class_start, num_sample)."""
    num_local = num_classes // world_size + int(rank < num_classes % world_size)
    class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    num_sample = int(sample_rate * num_local)
    return num_local, class_start, num_sample


def remap_labels(total_label, class_start, num_local):
    """ref partial_fc.py:79-81.  -> int64 copy: off-shard -> -1, on-shard -= class_start."""
    tl = np.asarray(total_label, np.int64).copy()
    on = (class_start <= tl) & (tl < class_start + num_local)
    tl[~on] = -1
    tl[on] -= class_start
    return tl


def select_index(perm, positive, num_sample):
    """ref partial_fc.py:84-90.  perm float32 (num_local,), positive sorted unique int64.
    -> sorted int64 index of the sampled classes."""
    if num_sample - positive.size >= 0:
        p = np.asarray(perm, np.float32).copy()
        p[positive] = 2.0
        # top-k by value, ties -> lowest index: stable sort on (-value)
        order = np.argsort(-p.astype(np.float64), kind="stable")
        return np.sort(order[:num_sample]).astype(np.int64)
    return positive.astype(np.int64)


def sample(total_label, perm, class_start, num_local, num_sample, sample_rate):
    """ref partial_fc.py:77-94 -> (remapped labels, index or None)."""
    tl = remap_labels(total_label, class_start, num_local)
    if int(sample_rate) == 1:
        return tl, None
    on = tl != -1
    positive = np.unique(tl[on])
    index = select_index(perm, positive, num_sample)
    tl[on] = np.searchsorted(index, tl[on])
    return tl, index


def rank_forward(total_features, sub_weight, tl, kind, s, m, a=0.0, k=0.0):
    """One rank, before any collective: -> dict(cos, logits, wn, rowmax)."""
    x = np.asarray(total_features, np.float64)
    w = np.asarray(sub_weight, np.float64)
    wn = l2_normalize(w)
    cos = x @ wn.T
    logits = margin_apply(cos, tl, kind, s, m, a, k)
    return dict(cos=cos, logits=logits, wn=wn, w=w, rowmax=logits.max(axis=1))


def step(features_per_rank, labels_per_rank, weights_per_rank, num_classes, kind, s, m,
         a=0.0, k=0.0, sample_rate=1.0, perms_per_rank=None):
    """Whole PartialFC.forward_backward over W simulated ranks.

    features_per_rank[r] (B, D), labels_per_rank[r] (B,), weights_per_rank[r] (num_local_r, D)
    -> dict with per-rank lists: x_grad (B, D), w_grad (n_s, D) [grad of sub_weight], index,
       total_label (remapped), and scalars loss; plus rowmax / rowsum / target prob (B_tot,).
    """
    W = len(features_per_rank)
    B = features_per_rank[0].shape[0]
    B_tot = B * W
    X = np.concatenate([np.asarray(f, np.float64) for f in features_per_rank], 0)
    L = np.concatenate([np.asarray(l, np.int64) for l in labels_per_rank], 0)

    fw, tls, idxs = [], [], []
    for r in range(W):
        num_local, class_start, num_sample = shard_geometry(num_classes, W, r, sample_rate)
        assert weights_per_rank[r].shape[0] == num_local
        perm = None if perms_per_rank is None else perms_per_rank[r]
        tl, index = sample(L, perm, class_start, num_local, num_sample, sample_rate)
        sub_w = weights_per_rank[r] if index is None else weights_per_rank[r][index]
        fw.append(rank_forward(X, sub_w, tl, kind, s, m, a, k))
        tls.append(tl)
        idxs.append(index)

    # distributed softmax: all_reduce MAX, all_reduce SUM  (:135-144)
    gmax = np.max(np.stack([f["rowmax"] for f in fw], 0), axis=0)
    exps = [np.exp(f["logits"] - gmax[:, None]) for f in fw]
    gsum = np.sum(np.stack([e.sum(axis=1) for e in exps], 0), axis=0)
    probs = [e / gsum[:, None] for e in exps]

    # loss: all_reduce SUM of the target probability  (:159-163)
    tgt = np.zeros(B_tot)
    for r in range(W):
        rows = np.nonzero(tls[r] != -1)[0]
        tgt[rows] += probs[r][rows, tls[r][rows]]
    loss = -np.mean(np.log(np.maximum(tgt, 1e-30)))

    x_grads_full, w_grads = [], []
    for r in range(W):
        p, tl, f = probs[r], tls[r], fw[r]
        n_s = p.shape[1]
        rows = np.nonzero(tl != -1)[0]
        g = p.copy()
        one_hot = np.full((rows.size, n_s), EPSILON / (n_s - 1))       # (:149-156)
        one_hot[np.arange(rows.size), tl[rows]] = 1.0 - EPSILON
        g[rows] -= one_hot
        g /= B_tot                                                      # (:166-167)
        dcos = g * margin_dcos(f["cos"], tl, kind, s, m, a, k)
        x_grads_full.append(dcos @ f["wn"])
        dwn = dcos.T @ X
        w_grads.append(normalize_bwd(f["w"], f["wn"], dwn))

    # reduce_scatter SUM then * world_size  (:172-175)
    total_dx = np.sum(np.stack(x_grads_full, 0), axis=0)
    x_grad = [total_dx[r * B:(r + 1) * B] * W for r in range(W)]
    return dict(x_grad=x_grad, w_grad=w_grads, index=idxs, total_label=tls, loss=loss,
                rowmax=gmax, rowsum=gsum, target_prob=tgt,
                logits=[f["logits"] for f in fw], dx_full=x_grads_full)


def forward_loss(features_per_rank, labels_per_rank, weights_per_rank, num_classes, kind, s, m, a=0.0, k=0.0, chunk=256):
    """Loss of the full-class step only (ref :132-163), evaluated in row chunks so that BASELINE-size problems
    (B_tot = 1024 x 93,431 classes) fit in a few hundred MB: used by bench.py to check the loss its first step reports
    on N ranks.  -> (loss, rowmax (B_tot,), rowsum (B_tot,))."""
    W = len(features_per_rank)
    X = np.concatenate([np.asarray(f, np.float64) for f in features_per_rank], 0)
    L = np.concatenate([np.asarray(l, np.int64) for l in labels_per_rank], 0)
    B_tot = X.shape[0]
    geo = [shard_geometry(num_classes, W, r) for r in range(W)]
    wns = [l2_normalize(np.asarray(w, np.float64)) for w in weights_per_rank]
    tls = [remap_labels(L, g[1], g[0]) for g in geo]
    gmax, gsum, tgt = np.empty(B_tot), np.empty(B_tot), np.zeros(B_tot)
    for lo in range(0, B_tot, chunk):
        hi = min(lo + chunk, B_tot)
        logits = [margin_apply(X[lo:hi] @ wns[r].T, tls[r][lo:hi], kind, s, m, a, k) for r in range(W)]
        mx = np.max(np.stack([l.max(axis=1) for l in logits], 0), axis=0)                      # all_reduce MAX (:136)
        exps = [np.exp(l - mx[:, None]) for l in logits]
        sm = np.sum(np.stack([e.sum(axis=1) for e in exps], 0), axis=0)                        # all_reduce SUM (:141)
        for r in range(W):
            rows = np.nonzero(tls[r][lo:hi] != -1)[0]
            tgt[lo + rows] += exps[r][rows, tls[r][lo:hi][rows]] / sm[rows]                   # all_reduce SUM (:162)
        gmax[lo:hi], gsum[lo:hi] = mx, sm
    return -np.mean(np.log(np.maximum(tgt, 1e-30))), gmax, gsum
