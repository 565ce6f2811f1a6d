"""Oracle: elementwise tail of the Feature-Masking operator (numpy, fp64 by default).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ref backbones/fm/fmoperator.py:
  :113-117  act   in {tanh, sigmoid}              (``mask_norm``)
  :71-81    arith in {add, sub, div, mul}         (``arith_*``)
  :288      g = act(z)            z = res_block(same_conv(cat(yf, yo)))
  :304      x = arith(identity, g)
  :307-308  x += f_out            (peer branch only)
  :310      x += identity

so   out = arith(yf, act(z)) [+ f_out] + yf          (default: yf * (1 + sigmoid(z)))

Backward (d = dout), with g = act(z), g' = g(1-g) for sigmoid, 1-g^2 for tanh:
  add : dyf = 2d            dg =  d
  sub : dyf = 2d            dg = -d
  mul : dyf = d(g+1)        dg =  d*yf
  div : dyf = d(1/g+1)      dg = -d*yf/g^2
  dz = dg * g'      df_out = d
``dyf`` here is the DIRECT part only (the path through cat->conv is autograd's).
"""
import numpy as np

ACTS = ("tanh", "sigmoid")
ARITHS = ("add", "sub", "div", "mul")


def act_fwd(z, act):
    if act == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if act == "tanh":
        return np.tanh(z)
    raise ValueError("activation type error")


def act_grad_from_g(g, act):
    return g * (1.0 - g) if act == "sigmoid" else 1.0 - g * g


def fm_gate_fwd(yf, z, act="sigmoid", arith="mul", f_out=None, dtype=np.float64):
    yf = np.asarray(yf, dtype)
    g = act_fwd(np.asarray(z, dtype), act)
    if arith == "add":
        x = yf + g
    elif arith == "sub":
        x = yf - g
    elif arith == "div":
        x = yf / g
    elif arith == "mul":
        x = yf * g
    else:
        raise ValueError("arith type error")
    if f_out is not None:
        x = x + np.asarray(f_out, dtype)
    return x + yf


def fm_gate_bwd(dout, yf, z, act="sigmoid", arith="mul", dtype=np.float64):
    """Returns (dyf_direct, dz).  df_out == dout."""
    d = np.asarray(dout, dtype)
    yf = np.asarray(yf, dtype)
    g = act_fwd(np.asarray(z, dtype), act)
    if arith == "add":
        dyf, dg = 2.0 * d, d
    elif arith == "sub":
        dyf, dg = 2.0 * d, -d
    elif arith == "mul":
        dyf, dg = d * (g + 1.0), d * yf
    elif arith == "div":
        dyf, dg = d * (1.0 / g + 1.0), -d * yf / (g * g)
    else:
        raise ValueError("arith type error")
    return dyf, dg * act_grad_from_g(g, act)


# ---------------------------------------------------------------------------
# Extension named by BASELINE.json's north_star ("decoder mask logits resized to
# each feature scale, normalised into gates and multiplied into the feature
# maps").  The reference never takes this branch: its gate is
# C-channel at the feature resolution.  Here the mask is (B, Hm, Wm, Cm) with
# Cm in {1, C}, resized to (H, W) by nearest-neighbour (torch 'nearest' rule:
# src = floor(dst * Hm / H)), broadcast over channels when Cm == 1.
# Layout is NHWC (channels-last physical order), matching the CUDA kernels.
# ---------------------------------------------------------------------------
def nearest_index(out_size, in_size):
    return np.minimum((np.arange(out_size, dtype=np.int64) * in_size) // out_size, in_size - 1)


def fm_mask_fwd(yf, m, act="sigmoid", arith="mul", dtype=np.float64):
    """yf (B,H,W,C), m (B,Hm,Wm,Cm) logits -> out (B,H,W,C)."""
    yf = np.asarray(yf, dtype)
    B, H, W, C = yf.shape
    _, Hm, Wm, Cm = m.shape
    ih, iw = nearest_index(H, Hm), nearest_index(W, Wm)
    z = np.asarray(m, dtype)[:, ih][:, :, iw]           # (B,H,W,Cm)
    z = np.broadcast_to(z, (B, H, W, C)) if Cm == 1 else z
    return fm_gate_fwd(yf, z, act, arith, None, dtype)


def fm_mask_bwd(dout, yf, m, act="sigmoid", arith="mul", dtype=np.float64):
    """Returns (dyf, dm) with dm reduced over broadcast channels and resize fan-out."""
    yf = np.asarray(yf, dtype)
    B, H, W, C = yf.shape
    _, Hm, Wm, Cm = m.shape
    ih, iw = nearest_index(H, Hm), nearest_index(W, Wm)
    z = np.asarray(m, dtype)[:, ih][:, :, iw]
    zb = np.broadcast_to(z, (B, H, W, C)) if Cm == 1 else z
    dyf, dz = fm_gate_bwd(dout, yf, zb, act, arith, dtype)
    if Cm == 1:
        dz = dz.sum(axis=3, keepdims=True)
    dm = np.zeros((B, Hm, Wm, Cm), dtype)
    hh, ww = np.meshgrid(ih, iw, indexing="ij")
    np.add.at(dm, (slice(None), hh, ww), dz)
    return dyf, dm
