"""Oracle: DAP head of the occlusion-segmentation branch + 2-way argmax mask (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ref backbones/osb/unet.py:158-161 (``DAP = PixelShuffle(k) -> AvgPool2d(k)``) and :223.
PixelShuffle(k): y[b, c, k*h+i, k*w+j] = x[b, c*k*k + i*k + j, h, w]
AvgPool2d(k)   : out[b, c, h, w] = mean_{i,j} y[b, c, k*h+i, k*w+j]
=> out[b, c, h, w] = mean over the k*k consecutive input channels [c*k*k, (c+1)*k*k).

Argmax mask: ref train.py:357 / eval/qeval_mxnet.py:347,370 ``final_seg[b].max(0)[1]``
(index of the max over the 2 channels, first index on ties).
"""
import numpy as np


def dap_fwd(x, k=3, dtype=np.float64):
    """x (B, C*k*k, H, W) -> (B, C, H, W)."""
    x = np.asarray(x, dtype)
    B, CK, H, W = x.shape
    kk = k * k
    assert CK % kk == 0
    return x.reshape(B, CK // kk, kk, H, W).sum(axis=2) / kk


def dap_bwd(dout, k=3, dtype=np.float64):
    """dout (B, C, H, W) -> dx (B, C*k*k, H, W)."""
    d = np.asarray(dout, dtype)
    kk = k * k
    return np.repeat(d / kk, kk, axis=1)


def argmax_mask(seg):
    """seg (B, 2, H, W) -> (B, H, W) int64; ties -> 0 (torch.max semantics)."""
    seg = np.asarray(seg)
    return (seg[:, 1] > seg[:, 0]).astype(np.int64)
