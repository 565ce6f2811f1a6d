"""Oracle: margin heads (numpy fp64).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ref headers/margin_losses.py:
  Softmax.forward    :41-68    out = e W^T + b
  AMCosFace.forward  :241-305  cos = norm(e) norm(W)^T ; target: cos - (m - k(theta - a)) ; * s
  AMArcFace.forward  :356-418  theta = acos(cos) ; target: theta + (m - k(theta - a)) ; cos(.) * s
Rows with label == -1 get no margin (:285,:400).  No clamp before acos (as the ref).

``margin_apply`` is the PartialFC margin callable (ref headers/partial_fc.py:132 expects
``margin_softmax(logits, total_label)``; the ref ships none - SURVEY.md F3 - it is exactly
:390-418 / :275-303 with the ``F.linear(normalize, normalize)`` prefix removed).

``margin_dcos`` is d(logit)/d(cos) (the adaptive term carries gradient):
  Arc target: s (1-k) sin((1-k) theta + m + k a) / sin(theta)      Cos target: s (1 - k / sin(theta))
  everything else: s
"""
import numpy as np


def l2_normalize(x, eps=1e-12):
    n = np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x / np.maximum(n, eps)


def margin_apply(cos, label, kind, s, m, a=0.0, k=0.0):
    """cos (B, C) float64, label (B,) int with -1 = no target -> logits (B, C)."""
    cos = np.asarray(cos, np.float64)
    label = np.asarray(label)
    rows = np.nonzero(label != -1)[0]
    cols = label[rows]
    if kind == "arc":
        theta = np.arccos(cos)
        t = theta[rows, cols]
        theta[rows, cols] = t + (m - k * (t - a))
        return np.cos(theta) * s
    if kind == "cos":
        out = cos.copy()
        t = np.arccos(cos[rows, cols])
        out[rows, cols] -= m - k * (t - a)
        return out * s
    raise ValueError("margin kind error")


def margin_dcos(cos, label, kind, s, m, a=0.0, k=0.0):
    """Elementwise d logit / d cos, same shape as cos."""
    cos = np.asarray(cos, np.float64)
    label = np.asarray(label)
    d = np.full_like(cos, s)
    rows = np.nonzero(label != -1)[0]
    cols = label[rows]
    t = np.arccos(cos[rows, cols])
    if kind == "arc":
        d[rows, cols] = s * (1.0 - k) * np.sin((1.0 - k) * t + m + k * a) / np.sin(t)
    elif kind == "cos":
        d[rows, cols] = s * (1.0 - k / np.sin(t))
    else:
        raise ValueError("margin kind error")
    return d


def am_head_fwd(embedding, weight, label, kind, s, m, a=0.0, k=0.0):
    """Full in-model head: AMArcFace / AMCosFace forward -> (logits, cos, en, wn)."""
    e = np.asarray(embedding, np.float64)
    w = np.asarray(weight, np.float64)
    en, wn = l2_normalize(e), l2_normalize(w)
    cos = en @ wn.T
    return margin_apply(cos, label, kind, s, m, a, k), cos, en, wn


def normalize_bwd(x, xn, dxn, eps=1e-12):
    """Backward of row-wise L2 normalise: dx = (dxn - xn * rowsum(xn*dxn)) / max(||x||, eps)."""
    n = np.maximum(np.sqrt((x * x).sum(axis=1, keepdims=True)), eps)
    return (dxn - xn * (xn * dxn).sum(axis=1, keepdims=True)) / n


def am_head_bwd(embedding, weight, label, dlogits, kind, s, m, a=0.0, k=0.0):
    """-> (d_embedding, d_weight) for the in-model head."""
    e = np.asarray(embedding, np.float64)
    w = np.asarray(weight, np.float64)
    _, cos, en, wn = am_head_fwd(e, w, label, kind, s, m, a, k)
    dcos = np.asarray(dlogits, np.float64) * margin_dcos(cos, label, kind, s, m, a, k)
    den = dcos @ wn
    dwn = dcos.T @ en
    return normalize_bwd(e, en, den), normalize_bwd(w, wn, dwn)


def softmax_head_fwd(embedding, weight, bias):
    return np.asarray(embedding, np.float64) @ np.asarray(weight, np.float64).T + np.asarray(bias, np.float64)
