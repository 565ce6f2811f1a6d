"""Oracle: y = prelu(batch_norm(x) [+ res]) and its backward (numpy fp64).  TEST INFRASTRUCTURE ONLY.

Follows the chain the reference builds from stock modules in its residual units:
  ref backbones/frb/iresnet.py:56-67 (bn1 | bn2 -> prelu | bn3 -> += identity),
  ref backbones/fm/fmoperator.py:52-68 (bn -> prelu, bn3 -> += identity -> prelu3)
with torch.nn.BatchNorm2d semantics (training: biased batch variance for the normalisation,
unbiased for running_var, momentum 0.1, eps 1e-5) and torch.nn.PReLU (per-channel slope).
Layout here: x (P, C) with P = N*H*W rows.
"""
import numpy as np


def bn_act_fwd(x, gamma, beta, prelu=None, res=None, training=True, running_mean=None, running_var=None,
               momentum=0.1, eps=1e-5):
    x = np.asarray(x, np.float64)
    P = x.shape[0]
    if training:
        mean, var = x.mean(0), x.var(0)
        new_rm = None if running_mean is None else (1 - momentum) * running_mean + momentum * mean
        new_rv = None if running_var is None else (1 - momentum) * running_var + momentum * var * P / max(P - 1, 1)
    else:
        mean, var, new_rm, new_rv = np.asarray(running_mean, np.float64), np.asarray(running_var, np.float64), running_mean, running_var
    invstd = 1.0 / np.sqrt(var + eps)
    u = (x - mean) * invstd * gamma + beta
    if res is not None:
        u = u + np.asarray(res, np.float64)
    y = u if prelu is None else np.where(u > 0, u, u * prelu)
    return y, dict(mean=mean, invstd=invstd, u=u, running_mean=new_rm, running_var=new_rv)


def bn_act_bwd(dy, x, gamma, beta, prelu=None, res=None, training=True, running_mean=None, running_var=None, eps=1e-5):
    """-> dict(dx, dgamma, dbeta, dprelu, dres)."""
    dy = np.asarray(dy, np.float64)
    x = np.asarray(x, np.float64)
    P = x.shape[0]
    _, st = bn_act_fwd(x, gamma, beta, prelu, res, training, running_mean, running_var, eps=eps)
    xh = (x - st["mean"]) * st["invstd"]
    if prelu is None:
        du, dprelu = dy, None
    else:
        neg = ~(st["u"] > 0)
        du = np.where(neg, dy * prelu, dy)
        dprelu = (dy * st["u"] * neg).sum(0)
    dbeta, dgamma = du.sum(0), (du * xh).sum(0)
    if training:
        dx = gamma * st["invstd"] * (du - dbeta / P - xh * dgamma / P)
    else:
        dx = gamma * st["invstd"] * du
    return dict(dx=dx, dgamma=dgamma, dbeta=dbeta, dprelu=dprelu, dres=du if res is not None else None)
