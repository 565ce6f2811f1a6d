"""Structure-via-consensus segmentation criterion — drop-in for ref tricks/consensus_loss.py:28-178
(`StructureConsensuLossFunction`), the `seg_criterion` of the live training recipe (ref train.py:228-229,258;
SURVEY.md 8f-4).

The reference loops over the blobs in Python and spends ~40 ATen kernels per blob on the (N, C, H, W) map; here the
loss is one pass over the logits + a one-CTA finalize, and its gradient one elementwise pass (csrc/seg_loss.cu,
ops.consensus_loss).  Same constructor arguments, same `forward(logit, blobs, target)`.

Differences from the reference, all on inputs its training loop never produces:
  * blob ids must be integers 0 .. num_blobs-1 (the occlusion mask of ref train.py:258 is {0, 1}: num_blobs=2, the
    default); the reference accepts arbitrary values because it calls torch.unique (a host synchronisation);
  * a blob whose pixels carry different labels makes the loss NaN instead of tripping the reference's assert (ref :103),
    again because an assert needs a host synchronisation;
  * reduce_pixel='all': a sample that lacks a blob gets a zero gradient from it; the reference produces NaN there
    (log of an empty blob mean, ref :142).
"""
import torch
from torch import nn

from .. import ops

__all__ = ["StructureConsensuLossFunction"]


class StructureConsensuLossFunction(nn.Module):
    def __init__(self, consensus_loss_alpha=10.0, consensus_loss_beta=5.0, reduce_pixel='idx', reduce_pixel_kl='idx',
                 num_blobs=2):
        super().__init__()
        self.consensus_loss_alpha = consensus_loss_alpha
        self.consensus_loss_beta = consensus_loss_beta
        self.reduce_pixel = reduce_pixel
        self.reduce_pixel_kl = reduce_pixel_kl
        self.num_blobs = num_blobs

    def structure_via_consensus(self, logit, blobs, target):
        """logit (N, C, H, W) pre-softmax; blobs (N, 1, H, W) or (N, H, W) blob ids; target (N, H, W) labels -> 0-dim loss."""
        return ops.consensus_loss(logit, blobs, target, self.consensus_loss_alpha, self.consensus_loss_beta,
                                  self.reduce_pixel, self.reduce_pixel_kl, self.num_blobs)

    def structure_via_consensus_over_blob(self, idx_blob, target, logit):
        """Loss of ONE blob given as a boolean mask (N, H, W) (ref :98-171): pixels outside the mask belong to no blob."""
        blobs = torch.where(idx_blob, 0, -1)
        return ops.consensus_loss(logit, blobs, target, self.consensus_loss_alpha, self.consensus_loss_beta,
                                  self.reduce_pixel, self.reduce_pixel_kl, 1)

    def forward(self, logit, blobs, target):
        return self.structure_via_consensus(logit, blobs, target)
