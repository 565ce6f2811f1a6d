"""Occluded-face verification embedding extraction — the GPU side of ref eval/verification.py:238-305 (``test``)
and eval/qeval_mxnet.py:285-397 (BASELINE config 5 / SURVEY.md 8a-16).

The reference walks a list of two uint8 image sets (originals and horizontally flipped copies), normalises each batch
with ((x / 255) - 0.5) / 0.5, runs ``backbone(img) -> (feature, final_seg)``, sums the two embedding sets and
L2-normalises the sum.  Its last batch is taken as ``data[bb - batch_size: bb]`` (a full batch that overlaps the previous
one) and only the new rows are kept; that indexing is reproduced here.  The 10-fold threshold evaluation that follows
(sklearn KFold / ROC on the host) is outside the hot path (out of scope).

B200 specifics: batches stay uint8 until they are on the device (38.5 MB per 1024 images instead of 154 MB of fp32),
the flipped set is produced on the device when the caller does not supply one, and the model runs the same fused NHWC
kernels as in training (eval-mode BN coefficients, fused FM gate, DAP + argmax mask).
"""
import numpy as np
import torch

__all__ = ["extract_embeddings", "random_block_occlusion", "test"]


def _as_uint8_tensor(data):
    t = torch.from_numpy(data) if isinstance(data, np.ndarray) else data
    if t.dim() != 4:
        raise ValueError("expected an (N, C, 112, 112) image set")
    return t


@torch.no_grad()
def extract_embeddings(data_list, backbone, batch_size, is_gray=False, device=None, return_masks=False):
    """data_list: [images] or [images, flipped images], each (N, 3, H, W) uint8 / float in 0..255 (host or device).
    -> (embeddings (N, D) float64 numpy, L2-normalised sum over the sets, list of per-set raw embeddings[, masks])."""
    device = torch.device(device) if device is not None else next(backbone.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("msml_b200.eval: embedding extraction has only a CUDA (sm_100a) implementation")
    sets = [_as_uint8_tensor(d) for d in data_list]
    if len(sets) == 1:
        sets.append(None)                          # flipped on the device
    if len(sets) != 2:
        raise ValueError("data_list must hold the originals and (optionally) the flipped copies")
    n = sets[0].shape[0]
    if batch_size > n:
        raise ValueError("batch_size %d exceeds the %d images of the set (ref :263 would index out of range)" % (batch_size, n))
    was_training = backbone.training
    backbone.eval()
    embeddings_list, masks = [], []
    try:
        for si, data in enumerate(sets):
            emb = None
            ba = 0
            while ba < n:
                bb = min(ba + batch_size, n)
                count = bb - ba
                src = sets[0] if data is None else data
                _data = src[bb - batch_size: bb].to(device, non_blocking=True)          # ref :263
                if data is None:
                    _data = torch.flip(_data, dims=[3])
                _data = _data.float()
                if is_gray:                                                               # ref :250-254, :269
                    _data = ((0.2989 * _data[:, 0] + 0.5870 * _data[:, 1] + 0.1140 * _data[:, 2]) / 3)[:, None]
                    img = _data / 255
                else:
                    img = ((_data / 255) - 0.5) / 0.5                                     # ref :267
                feature, final_seg = backbone(img)
                if emb is None:
                    emb = torch.zeros((n, feature.shape[1]), dtype=torch.float32, device=device)
                emb[ba:bb] = feature[batch_size - count:].float()                         # ref :281
                if return_masks and si == 0 and final_seg is not None:
                    masks.append((ba, bb, final_seg[batch_size - count:].max(1)[1]))      # ref train.py:357 argmax mask
                ba = bb
            embeddings_list.append(emb.cpu().numpy().astype(np.float64))
    finally:
        backbone.train(was_training)
    total = embeddings_list[0] + embeddings_list[1]                                       # ref :298
    norm = np.maximum(np.linalg.norm(total, axis=1, keepdims=True), 1e-12)                # sklearn.preprocessing.normalize (l2)
    out = (total / norm, embeddings_list)
    if return_masks:
        m = torch.zeros((n,) + tuple(masks[0][2].shape[1:]), dtype=torch.int64, device=device) if masks else None
        for ba, bb, mk in masks:
            m[ba:bb] = mk
        out = out + (m.cpu().numpy() if m is not None else None,)
    return out


def test(data_set, backbone, batch_size, nfolds=10, is_gray=False):
    """Signature of ref eval/verification.py:238.  Returns (acc1, std1, acc2, std2, xnorm, embeddings_list) with the
    accuracy fields computed by the caller-side evaluator when scikit-learn is importable, else NaN: the k-fold ROC
    evaluation is host code outside this package's scope, only the embedding side is accelerated."""
    data_list, issame_list = data_set[0], data_set[1]
    embeddings, embeddings_list = extract_embeddings(data_list, backbone, batch_size, is_gray=is_gray)
    xnorm = float(np.mean([np.linalg.norm(e, axis=1).mean() for e in embeddings_list]))  # ref :286-294
    acc2, std2 = float("nan"), float("nan")
    try:
        from sklearn.model_selection import KFold
        emb1, emb2 = embeddings[0::2], embeddings[1::2]
        dist = np.sum(np.square(emb1 - emb2), 1)
        issame = np.asarray(issame_list, dtype=bool)
        thresholds = np.arange(0, 4, 0.01)
        accs = []
        for train, val in KFold(n_splits=nfolds, shuffle=False).split(np.arange(len(issame))):
            acc_train = [np.mean((dist[train] < t) == issame[train]) for t in thresholds]
            best = thresholds[int(np.argmax(acc_train))]
            accs.append(np.mean((dist[val] < best) == issame[val]))
        acc2, std2 = float(np.mean(accs)), float(np.std(accs))
    except ImportError:
        pass
    return 0.0, 0.0, acc2, std2, xnorm, embeddings_list


def random_block_occlusion(images, lo, hi, generator=None):
    """Device twin of ref datasets/augment/rand_occ.py:36-72 RandomBlock(lo, hi, 'black') for a uint8 batch that already
    lives on the GPU: one black square per image whose AREA is ratio % of the frame, ratio uniform in [lo, hi) (integer
    percent), side = int(sqrt(ratio) * W), position uniform over the placements that keep it inside (ref :45-71).  It draws
    from a torch generator, so it is statistically, not bitwise, the reference's transform; the bit-exact host version is
    msml_b200.datasets.augment.random_block_batch (what bench.py's config-5 workload uses)."""
    n, _, h, w = images.shape
    dev = images.device
    ratio = torch.randint(lo, max(hi, lo + 1), (n,), device=dev, generator=generator).double() * 0.01
    side = torch.floor(torch.sqrt(ratio * w * w)).long()
    left = (torch.rand(n, device=dev, generator=generator) * (w - side + 1)).long().clamp(max=w - 1)
    top = (torch.rand(n, device=dev, generator=generator) * (w - side + 1)).long().clamp(max=h - 1)
    ys = torch.arange(h, device=dev)[None, :, None]
    xs = torch.arange(w, device=dev)[None, None, :]
    inside = ((ys >= top[:, None, None]) & (ys < (top + side)[:, None, None]) &
              (xs >= left[:, None, None]) & (xs < (left + side)[:, None, None]))
    return images * (~inside)[:, None].to(images.dtype)
