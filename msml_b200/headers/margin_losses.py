"""Margin heads — drop-in for ref headers/margin_losses.py (Softmax :18-68, AMCosFace :203-315,
AMArcFace :318-428), plus the margin callable PartialFC expects (ref headers/partial_fc.py:30,132;
the reference ships none).

The contraction runs on the tcgen05 GEMM of libmsml_b200.so (bf16 operands, fp32 accumulate) and
the margin / scale and its derivative are CUDA kernels (msml_margin_fwd / msml_margin_bwd);
only the two tiny L2-normalisations are left to PyTorch autograd.
"""
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter

from .. import ops

__all__ = ["Softmax", "AMCosFace", "AMArcFace", "MarginSoftmax", "ArcFace", "CosFace"]


class MarginSoftmax:
    """``margin_softmax(logits, label) -> logits`` callable carrying (kind, s, m, a, k).

    PartialFC reads the attributes and fuses the margin into its GEMM epilogue; calling the
    object applies the same margin to a materialised cosine matrix (rows whose label is -1 are
    only scaled), i.e. ref margin_losses.py:390-418 ('arc') / :275-303 ('cos').
    """

    def __init__(self, kind, s=64.0, m=0.5, a=0.0, k=0.0):
        if kind not in ("arc", "cos"):
            raise ValueError("margin kind error")
        self.kind, self.s, self.m, self.a, self.k = kind, float(s), float(m), float(a), float(k)

    def __call__(self, logits, label):
        return ops.margin_logits(logits, label, self.kind, self.s, self.m, self.a, self.k)

    def __repr__(self):
        return "MarginSoftmax(kind=%s, s=%g, m=%g, a=%g, k=%g)" % (self.kind, self.s, self.m, self.a, self.k)


def ArcFace(s=64.0, m=0.5):
    return MarginSoftmax("arc", s, m)


def CosFace(s=64.0, m=0.4):
    return MarginSoftmax("cos", s, m)


class Softmax(nn.Module):
    """Plain FC head: out = e W^T + b  (ref :41-68); device_id must be None (ref :55)."""

    def __init__(self, in_features, out_features, device_id):
        super().__init__()
        self.in_features, self.out_features, self.device_id = in_features, out_features, device_id
        self.weight = Parameter(torch.empty(out_features, in_features))
        self.bias = Parameter(torch.empty(out_features))
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, embedding, label):
        if self.device_id is not None:
            raise ValueError("DataParallel is not implemented yet.")
        return ops.cosine_logits(embedding.float(), self.weight) + self.bias


class _AMHead(nn.Module):
    kind = None

    def __init__(self, in_features, out_features, device_id, s, m, a, k):
        super().__init__()
        self.in_features, self.out_features, self.device_id = in_features, out_features, device_id
        self.s, self.m, self.a, self.k = s, m, a, k
        self.weight = Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, embedding, label):
        if self.device_id is not None:
            raise ValueError("DataParallel is not implemented yet.")
        cos = ops.cosine_logits(F.normalize(embedding.float()), F.normalize(self.weight))
        return ops.margin_logits(cos, label, self.kind, self.s, self.m, self.a, self.k)

    def __repr__(self):
        return "%s(in_features = %d, out_features = %d, s = %s, m = %s, a = %s, k = %s)" % (
            self.__class__.__name__, self.in_features, self.out_features, self.s, self.m, self.a, self.k)


class AMCosFace(_AMHead):
    """s * (cos(theta) - m + k (theta - a)) at the target; k = 0 is CosFace (ref :203-315)."""
    kind = "cos"

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.4, a=1.2, k=0.1):
        super().__init__(in_features, out_features, device_id, s, m, a, k)


class AMArcFace(_AMHead):
    """s * cos(theta + m - k (theta - a)) at the target; k = 0 is ArcFace (ref :318-428)."""
    kind = "arc"

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.5, a=1.2, k=0.1):
        super().__init__(in_features, out_features, device_id, s, m, a, k)
