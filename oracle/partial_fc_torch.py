"""Oracle / CPU baseline: the reference's single-rank PartialFC step restated op for op in torch-CPU fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `partial_fc.py` (numpy fp64, all ranks simulated) is the CHECKER;
this module is the TIMED CPU PORT of the same step for bench.py's cpu_baseline / --impl reference legs: it issues the
very ATen calls the reference issues (multi-threaded, fp32, the full-matrix `acos_ / cos_ / mul_` passes, the dense
one-hot, autograd for the two backward GEMMs), so its time is what the reference's code costs on the host cores — the
fp64 checker is deliberately simple and 20x slower, and was the wrong thing to time.

Follows ref headers/partial_fc.py (world_size == 1, sample_rate == 1, so the collectives are identities):
  :115      norm_weight = normalize(sub_weight)
  :96-99    logits = linear(total_features, norm_weight)
  :132      margin_softmax(logits, total_label)  == ref headers/margin_losses.py:390-418 (arc) / :275-303 (cos), k = 0
  :135-144  max, exp, sum, div
  :149-167  dense one-hot with local label smoothing eps = 0.1, loss, grad = (p - t) / B_tot
  :169      logits.backward(grad)
and ref train.py:188-191,299-300 for the class-centre SGD (momentum 0.9, weight decay 5e-4).
Pinned by tests/test_oracle_golden.py::test_torch_port_of_the_head_matches_reference (pfc_w1_full, pfc_w1_d512).
"""
import torch
import torch.nn.functional as F

EPSILON = 0.1  # ref partial_fc.py:154


def margin_inplace(logits, label, kind, s, m, a=0.0, k=0.0):
    """ref margin_losses.py:390-418 (arc) / :275-303 (cos), in place on the cosine matrix like the reference: dense
    m_hot, the adaptive term on an advanced-index copy, then full-matrix acos_ / cos_ / mul_ passes."""
    index = torch.where(label != -1)[0]
    m_hot = torch.zeros(index.size(0), logits.size(1), dtype=logits.dtype)
    m_hot.scatter_(1, label[index, None], m)
    m_hot[range(0, index.size(0)), label[index]] -= k * (logits[index, label[index]].acos_() - a)
    if kind == "arc":
        logits.acos_()
        logits[index] += m_hot
        logits.cos_().mul_(s)
    elif kind == "cos":
        logits[index] -= m_hot
        logits.mul_(s)
    else:
        raise ValueError("margin kind error")
    return logits


def head_step(features, label, weight, kind="arc", s=64.0, m=0.5):
    """features (B, D) fp32 L2-normalised, label (B,) int64 in [0, C), weight (C, D) fp32 (requires no grad).
    -> (x_grad (B, D), w_grad (C, D), loss 0-d)."""
    sub_weight = weight.detach().requires_grad_(True)
    norm_weight = F.normalize(sub_weight)                                  # :115
    total_features = features.detach().clone().requires_grad_(True)       # :124-127 (W = 1: the gather is a copy)
    B_tot = total_features.size(0)
    logits = F.linear(total_features, norm_weight)                         # :98
    logits = margin_inplace(logits, label, kind, s, m)                     # :132
    with torch.no_grad():
        max_fc = torch.max(logits, dim=1, keepdim=True)[0]                 # :135
        logits_exp = torch.exp(logits - max_fc)                            # :139-140
        logits_sum_exp = logits_exp.sum(dim=1, keepdims=True)              # :141
        logits_exp.div_(logits_sum_exp)                                    # :144
        grad = logits_exp
        index = torch.where(label != -1)[0]                                # :149
        one_hot = torch.zeros(index.size(0), grad.size(1), dtype=grad.dtype)       # :150-151 (dense, as the reference)
        one_hot.scatter_(1, label[index, None], 1)
        one_hot = (1 - EPSILON) * one_hot                                          # :153-155
        one_hot[torch.where(one_hot == 0)] = EPSILON / (one_hot.shape[1] - 1)
        loss = torch.zeros(B_tot, 1)
        loss[index] = grad[index].gather(1, label[index, None])           # :159-161
        loss_v = loss.clamp_min_(1e-30).log_().mean() * (-1)               # :163
        grad[index] -= one_hot                                             # :166
        grad.div_(B_tot)                                                   # :167
    logits.backward(grad)                                                  # :169
    return total_features.grad, sub_weight.grad, loss_v


def sgd_update(weight, mom, w_grad, lr=0.1, momentum=0.9, weight_decay=5e-4):
    """ref train.py:188-191,299: torch.optim.SGD on the class centres with the supplied momentum buffer, in place."""
    with torch.no_grad():
        d = w_grad.add(weight, alpha=weight_decay)
        mom.mul_(momentum).add_(d)
        weight.add_(mom, alpha=-lr)
