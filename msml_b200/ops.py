"""torch.autograd wrappers over the C ABI (include/msml_b200.h).  PyTorch is plumbing here: it
owns device memory and streams and supplies the autograd graph; every op below is one or more
hand-written sm_100a kernels from libmsml_b200.so.  No op has a PyTorch/CPU fallback.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ACT, ARITH, check, dtype_code, load, require_cuda, stream_ptr


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dense_like(ref, t):
    """Return t laid out with exactly ref's strides (ref is dense)."""
    if t.stride() == ref.stride() and t.dtype == ref.dtype:
        return t
    out = torch.empty_like(ref)
    out.copy_(t)
    return out


def _dense(t):
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------
# K-A  mask fusion tail of the FM operator      ref backbones/fm/fmoperator.py:288,304-310
# --------------------------------------------------------------------------------------------
class _FMGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, yf, z, f_out, act, arith):
        require_cuda(yf, z, f_out)
        lib = load()
        yf_d = _dense(yf)
        z_d = _dense_like(yf_d, z)
        fo_d = _dense_like(yf_d, f_out) if f_out is not None else None
        out = torch.empty_like(yf_d)
        check(lib.msml_fm_gate_fwd(_ptr(yf_d), _ptr(z_d), _ptr(fo_d), _ptr(out), yf_d.numel(),
                                   dtype_code(yf_d.dtype), ACT[act], ARITH[arith], stream_ptr()))
        ctx.save_for_backward(yf_d, z_d)
        ctx.cfg = (act, arith, f_out is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        yf, z = ctx.saved_tensors
        act, arith, has_fout = ctx.cfg
        lib = load()
        d = _dense_like(yf, dout)
        dyf = torch.empty_like(yf)
        dz = torch.empty_like(yf)
        check(lib.msml_fm_gate_bwd(_ptr(d), _ptr(yf), _ptr(z), _ptr(dyf), _ptr(dz), yf.numel(),
                                   dtype_code(yf.dtype), ACT[act], ARITH[arith], stream_ptr()))
        return dyf, dz, (d if has_fout else None), None, None


def fm_gate(yf, z, act="sigmoid", arith="mul", f_out=None):
    """out = arith(yf, act(z)) [+ f_out] + yf, fused forward and backward."""
    if act not in ACT:
        raise ValueError("activation type error")
    if arith not in ARITH:
        raise ValueError("arith type error")
    if yf.shape != z.shape:
        raise ValueError("fm_gate: yf %s and z %s must have the same shape" % (tuple(yf.shape), tuple(z.shape)))
    if z.dtype != yf.dtype:
        z = z.to(yf.dtype)
    return _FMGate.apply(yf, z, f_out, act, arith)


def fm_gate_multi_fwd(yfs, zs, act="sigmoid", arith="mul"):
    """One launch over several scales (inference / microbench).  Returns the list of outputs."""
    lib = load()
    require_cuda(*yfs, *zs)
    n = len(yfs)
    outs = [torch.empty_like(y) for y in yfs]
    arr = ctypes.c_void_p * n
    sizes = (ctypes.c_int64 * n)(*[y.numel() for y in yfs])
    check(lib.msml_fm_gate_fwd_multi(n, arr(*[y.data_ptr() for y in yfs]), arr(*[z.data_ptr() for z in zs]),
                                     arr(*[o.data_ptr() for o in outs]), sizes, dtype_code(yfs[0].dtype),
                                     ACT[act], ARITH[arith], stream_ptr()))
    return outs


def fm_gate_multi_bwd(douts, yfs, zs, act="sigmoid", arith="mul"):
    lib = load()
    require_cuda(*douts, *yfs, *zs)
    n = len(yfs)
    dyfs = [torch.empty_like(y) for y in yfs]
    dzs = [torch.empty_like(y) for y in yfs]
    arr = ctypes.c_void_p * n
    sizes = (ctypes.c_int64 * n)(*[y.numel() for y in yfs])
    check(lib.msml_fm_gate_bwd_multi(n, arr(*[d.data_ptr() for d in douts]), arr(*[y.data_ptr() for y in yfs]),
                                     arr(*[z.data_ptr() for z in zs]), arr(*[o.data_ptr() for o in dyfs]),
                                     arr(*[o.data_ptr() for o in dzs]), sizes, dtype_code(yfs[0].dtype),
                                     ACT[act], ARITH[arith], stream_ptr()))
    return dyfs, dzs


# --------------------------------------------------------------------------------------------
# K-C  input assembly of the FM operator: cat(yf, yo) + zero pad     ref backbones/fm/fmoperator.py:277-279
# --------------------------------------------------------------------------------------------
def _fm_cat_launch(yf_d, yo_d, multiple):
    B, C, H, W = yf_d.shape
    Co = yo_d.shape[1]
    Ct = -(-(C + Co) // multiple) * multiple
    cat = torch.empty((B, Ct, H, W), dtype=yf_d.dtype, device=yf_d.device, memory_format=torch.channels_last)
    check(load().msml_fm_cat_fwd(_ptr(yf_d), _ptr(yo_d), _ptr(cat), B * H * W, C, Co, Ct, dtype_code(yf_d.dtype), stream_ptr()))
    return cat, Ct - C - Co


class _FMCat(torch.autograd.Function):
    """(cat(yf, yo, zeros), yf): yf has two consumers inside the FM operator, this concat and the fused tail.  The second
    output is yf itself for the tail; the backward kernel then adds the tail's gradient while it gathers the yf columns of
    the concat's gradient, instead of autograd adding a strided slice in one more pass (same idea as bn_act_fork)."""

    @staticmethod
    def forward(ctx, yf, yo, multiple):
        yf_d = yf.contiguous(memory_format=torch.channels_last)
        yo_d = yo.to(yf.dtype).contiguous(memory_format=torch.channels_last)
        cat, _ = _fm_cat_launch(yf_d, yo_d, multiple)
        ctx.cfg = (tuple(yf.shape), yo.shape[1], cat.shape[1], yo.dtype)
        ctx.set_materialize_grads(False)
        return cat, yf_d.view_as(yf_d)

    @staticmethod
    def backward(ctx, dcat, dtail):
        (B, C, H, W), Co, Ct, yo_dtype = ctx.cfg
        if dcat is None:
            return dtail, None, None
        d = dcat.contiguous(memory_format=torch.channels_last)
        dyf = torch.empty((B, C, H, W), dtype=d.dtype, device=d.device, memory_format=torch.channels_last)
        dadd = _dense_like(dyf, dtail) if dtail is not None else None
        dyo = (torch.empty((B, Co, H, W), dtype=d.dtype, device=d.device, memory_format=torch.channels_last)
               if ctx.needs_input_grad[1] else None)
        check(load().msml_fm_cat_bwd(_ptr(d), _ptr(dadd), _ptr(dyf), _ptr(dyo), B * H * W, C, Co, Ct, dtype_code(d.dtype),
                                     stream_ptr()))
        return dyf, (dyo.to(yo_dtype) if dyo is not None else None), None


def fm_cat(yf, yo, multiple=8):
    """-> (x, pad, yf_tail): x = cat((yf, yo), dim=1) with `pad` zero channels appended up to a multiple of ``multiple``
    (cuDNN's bf16 tensor-core kernels need C % 8 == 0), one kernel each way; yf_tail is yf for the operator's fused tail
    (see _FMCat).  Like every operator of this package it has only the CUDA implementation: CPU tensors and feature
    channel counts that are not whole 16-byte vectors raise (every iResNet stage has C % 8 == 0)."""
    if yf.dim() != 4 or yo.dim() != 4 or yf.shape[0] != yo.shape[0] or yf.shape[2:] != yo.shape[2:]:
        raise ValueError("fm_cat: yf %s and yo %s must agree in batch and spatial size" % (tuple(yf.shape), tuple(yo.shape)))
    require_cuda(yf, yo)
    vn = 4 if yf.dtype == torch.float32 else 8
    if yf.shape[1] % vn or multiple % vn:
        raise ValueError("fm_cat: feature channels (%d) and the padding multiple (%d) must be multiples of %d (16-byte vectors)"
                         % (yf.shape[1], multiple, vn))
    C, Co = yf.shape[1], yo.shape[1]
    pad = (-(C + Co)) % multiple
    if torch.is_grad_enabled() and (yf.requires_grad or yo.requires_grad):
        x, yf_tail = _FMCat.apply(yf, yo, multiple)
        return x, pad, yf_tail
    yf_d = yf.contiguous(memory_format=torch.channels_last)
    x, pad = _fm_cat_launch(yf_d, yo.to(yf.dtype).contiguous(memory_format=torch.channels_last), multiple)
    return x, pad, yf_d


# Extension (north_star): low-resolution / single-channel mask, NHWC.
class _FMMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, yf, m, act, arith):
        require_cuda(yf, m)
        lib = load()
        B, C, H, W = yf.shape
        _, Cm, Hm, Wm = m.shape
        yf_d = yf.contiguous(memory_format=torch.channels_last)
        m_d = m.to(yf.dtype).contiguous(memory_format=torch.channels_last)
        out = torch.empty_like(yf_d)
        check(lib.msml_fm_mask_fwd(_ptr(yf_d), _ptr(m_d), _ptr(out), B, H, W, C, Hm, Wm, Cm,
                                   dtype_code(yf_d.dtype), ACT[act], ARITH[arith], stream_ptr()))
        ctx.save_for_backward(yf_d, m_d)
        ctx.cfg = (act, arith)
        return out

    @staticmethod
    def backward(ctx, dout):
        yf, m = ctx.saved_tensors
        act, arith = ctx.cfg
        lib = load()
        B, C, H, W = yf.shape
        _, Cm, Hm, Wm = m.shape
        d = _dense_like(yf, dout)
        dyf = torch.empty_like(yf)
        dm = torch.zeros(m.shape, dtype=torch.float32, device=m.device).contiguous(memory_format=torch.channels_last)
        check(lib.msml_fm_mask_bwd(_ptr(d), _ptr(yf), _ptr(m), _ptr(dyf), _ptr(dm), B, H, W, C, Hm, Wm, Cm,
                                   dtype_code(yf.dtype), ACT[act], ARITH[arith], stream_ptr()))
        return dyf, dm.to(m.dtype), None, None


def fm_mask(yf, m, act="sigmoid", arith="mul"):
    """yf (B,C,H,W), m (B,Cm,Hm,Wm) mask logits with Cm in {1, C}: resize(nearest) + gate + fuse."""
    if act not in ACT:
        raise ValueError("activation type error")
    if arith not in ARITH:
        raise ValueError("arith type error")
    return _FMMask.apply(yf, m, act, arith)


# --------------------------------------------------------------------------------------------
# K-P  peer-guided branch of the FM operator: masked products + MSE      ref backbones/fm/fmoperator.py:293-302
# --------------------------------------------------------------------------------------------
PEER_MODE = {"given": 0, "invert": 1}


class _FMPeerMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, yf, yt, mode, act):
        require_cuda(src, yf, yt)
        lib = load()
        yf_d = _dense(yf)
        src_d = _dense_like(yf_d, src)
        yt_d = _dense_like(yf_d, yt) if yt is not None else None
        pf = torch.empty_like(yf_d)
        pt = torch.empty_like(yf_d) if yt is not None else None
        check(lib.msml_fm_peer_mul_fwd(_ptr(src_d), _ptr(yf_d), _ptr(yt_d), _ptr(pf), _ptr(pt), yf_d.numel(), dtype_code(yf_d.dtype),
                                       PEER_MODE[mode], ACT[act], stream_ptr()))
        ctx.save_for_backward(src_d, yf_d, yt_d)
        ctx.cfg = (mode, act)
        ctx.set_materialize_grads(False)
        if yt is None:
            return pf
        return pf, pt

    @staticmethod
    def backward(ctx, dpf, dpt=None):
        src, yf, yt = ctx.saved_tensors
        mode, act = ctx.cfg
        if dpf is None and dpt is None:
            return None, None, None, None, None
        lib = load()
        if dpf is None:                             # only the teacher product was used downstream
            dpf = torch.zeros_like(yf)
        if yt is not None and dpt is None:
            dpt = torch.zeros_like(yf)
        dpf_d = _dense_like(yf, dpf)
        dpt_d = _dense_like(yf, dpt) if yt is not None else None
        dsrc, dyf = torch.empty_like(yf), torch.empty_like(yf)
        check(lib.msml_fm_peer_mul_bwd(_ptr(dpf_d), _ptr(dpt_d), _ptr(src), _ptr(yf), _ptr(yt), _ptr(dsrc), _ptr(dyf), yf.numel(),
                                       dtype_code(yf.dtype), PEER_MODE[mode], ACT[act], stream_ptr()))
        return dsrc, dyf, None, None, None


def fm_peer_mul(src, yf, yt=None, mode="given", act="sigmoid"):
    """(m_bar * yf, m_bar * yt) in one pass (ref fmoperator.py:295-299).  mode "given": ``src`` is m_bar (the output of
    conv_m); mode "invert": ``src`` is the PRE-activation z and m_bar = 1 - act(z) (ref :160-166 with mask_trans 'invert'),
    formed in registers.  ``yt`` (the frozen teacher's feature map) gets no gradient; without it only the first product is
    returned."""
    if mode not in PEER_MODE:
        raise ValueError("mask_trans type error")
    if act not in ACT:
        raise ValueError("activation type error")
    if src.shape != yf.shape or (yt is not None and yt.shape != yf.shape):
        raise ValueError("fm_peer_mul: m_bar, yf and yt must have one shape")
    return _FMPeerMul.apply(src, yf, yt.detach() if yt is not None else None, mode, act)


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        require_cuda(a, b)
        lib = load()
        a_d = _dense(a)
        b_d = _dense_like(a_d, b)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        ws_bytes = lib.msml_mse_workspace()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
        check(lib.msml_mse_fwd(_ptr(a_d), _ptr(b_d), a_d.numel(), dtype_code(a_d.dtype), _ptr(out), _ptr(ws), ws_bytes, stream_ptr()))
        ctx.save_for_backward(a_d, b_d)
        ctx.need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        lib = load()
        g = gout.to(torch.float32).contiguous()
        da = torch.empty_like(a)
        db = torch.empty_like(a) if ctx.need[1] else None
        check(lib.msml_mse_bwd(_ptr(a), _ptr(b), _ptr(g), _ptr(da), _ptr(db), a.numel(), dtype_code(a.dtype), stream_ptr()))
        return (da if ctx.need[0] else None), db


def mse_loss(a, b):
    """torch.nn.MSELoss()(a, b) (ref fmoperator.py:300, mean reduction) as one reduction pass: fp32 accumulation from the
    storage dtype, fp32 scalar result, deterministic."""
    if a.shape != b.shape:
        raise ValueError("mse_loss: shapes differ")
    if a.numel() == 0:
        raise ValueError("mse_loss of an empty tensor")
    return _MSE.apply(a, b)


# --------------------------------------------------------------------------------------------
# K-N  fused BatchNorm (+ residual) (+ PReLU), NHWC            ref backbones/frb/iresnet.py:56-67,
#                                                               backbones/fm/fmoperator.py:52-68
# --------------------------------------------------------------------------------------------
class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, prelu, res, running_mean, running_var, nbt, training, momentum, eps, fork=False,
                emit_next=False, pre_ws=None):
        require_cuda(x, gamma, beta, prelu, res)
        lib = load()
        B, C, H, W = x.shape
        P = B * H * W
        x_d = x.contiguous(memory_format=torch.channels_last)
        res_d = _dense_like(x_d, res) if res is not None else None
        y = torch.empty_like(x_d)
        stats = torch.empty((2, C), dtype=torch.float32, device=x.device)
        ws_bytes = lib.msml_bn_workspace(P, C)
        # chained statistics (msml_bn_fwd_ex): pre_ws is this op's workspace, already holding the slab statistics of x
        # (left there by the op that wrote x); next_ws receives those of y for the BatchNorm that consumes it
        ws = pre_ws if pre_ws is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        next_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if emit_next else None
        if emit_next or pre_ws is not None:
            check(lib.msml_bn_fwd_ex(_ptr(x_d), _ptr(res_d), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(prelu), _ptr(running_mean),
                                     _ptr(running_var), _ptr(nbt) if training else None, _ptr(stats[0]), _ptr(stats[1]), P, C,
                                     dtype_code(x_d.dtype), int(training), float(momentum), float(eps), _ptr(ws), ws_bytes,
                                     _ptr(next_ws), ws_bytes if emit_next else 0, int(pre_ws is not None), stream_ptr()))
        else:
            check(lib.msml_bn_fwd(_ptr(x_d), _ptr(res_d), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(prelu), _ptr(running_mean),
                                  _ptr(running_var), _ptr(nbt) if training else None, _ptr(stats[0]), _ptr(stats[1]), P, C,
                                  dtype_code(x_d.dtype), int(training), float(momentum), float(eps), _ptr(ws), ws_bytes, stream_ptr()))
        ctx.save_for_backward(x_d, res_d if prelu is not None else None, gamma, beta, prelu, stats)
        ctx.cfg = (training, res is not None)
        ctx.params = (gamma, beta, prelu)          # the Parameter objects themselves (for the direct-gradient path)
        ctx.fork = fork
        ctx.set_materialize_grads(False)           # an unused output arrives as None in backward, not as a zero tensor
        if fork:                                   # second output: x itself, for the skip branch (see bn_act_fork)
            return y, x_d.view_as(x_d)
        if emit_next:                              # second output: the next BatchNorm's workspace (no gradient)
            ctx.mark_non_differentiable(next_ws)
            return y, next_ws
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, res, gamma, beta, prelu, stats = ctx.saved_tensors
        training, has_res = ctx.cfg
        if not ctx.fork:                           # a second output that is not the skip branch (next_ws) has no gradient
            dskip = None
        lib = load()
        B, C, H, W = x.shape
        P = B * H * W
        if dy is None:                             # only the skip branch carried a gradient (or nothing did)
            return (dskip,) + (None,) * 13
        dy_d = _dense_like(x, dy)
        dadd = _dense_like(x, dskip) if dskip is not None else None
        dx = torch.empty_like(x)
        both = has_res and prelu is not None
        dres = torch.empty_like(x) if both else None
        ws_bytes = lib.msml_bn_workspace(P, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        # Direct-gradient path (engine.TrainStep): the parameters' .grad are views of one flat fp32 buffer and the
        # kernel adds dgamma / dbeta / dprelu into them itself: what AccumulateGrad would do in one more kernel each.
        params = [p for p in ctx.params if p is not None]
        direct = all(_direct_grad(p) for p in params)
        if direct:
            g_ptrs = [_ptr(p.grad) if p is not None else None for p in ctx.params]
            grads = None
        else:
            grads = torch.empty((3, C), dtype=torch.float32, device=x.device)
            g_ptrs = [_ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2])]
        # a residual without PReLU passes its gradient straight through (dres = dy): no extra stream
        check(lib.msml_bn_bwd(_ptr(dy_d), _ptr(x), _ptr(res) if both else None, _ptr(gamma), _ptr(beta),
                              _ptr(prelu), _ptr(stats[0]), _ptr(stats[1]), _ptr(dx), _ptr(dres), _ptr(dadd), g_ptrs[0], g_ptrs[1],
                              g_ptrs[2], P, C, dtype_code(x.dtype), int(training), int(direct), _ptr(ws), ws_bytes, stream_ptr()))
        d_res = dres if both else (dy_d if has_res else None)
        if direct:
            return (dx, None, None, None, d_res) + (None,) * 9
        dprelu = grads[2] if prelu is not None else None
        return (dx, grads[0], grads[1], dprelu, d_res) + (None,) * 9


def _direct_grad(p):
    """True when the engine has marked ``p`` as living in its flat gradient buffer (see engine.TrainStep)."""
    return (getattr(p, "_msml_direct_grad", False) and p.grad is not None and p.grad.dtype == torch.float32
            and _is_dense(p.grad) and p.grad.data_ptr() % 16 == 0)


_CHAIN_ATTR = "_msml_bn_next_ws"


def _chain_enabled():
    """Chained BN statistics (msml_bn_fwd_ex) need the split launches; MSML_BN_CHAIN=0 turns them off (A/B)."""
    return os.environ.get("MSML_BN_CHAIN", "1") != "0" and os.environ.get("MSML_BN_FUSED", "0") in ("", "0")


def _take_chained_ws(x, bn):
    """The workspace a producer (bn_act(..., emit_next_stats=True)) attached to ``x``: the slab statistics of exactly this
    tensor, valid for one training-mode BatchNorm over it.  Consumed (detached from ``x``) here."""
    ws = getattr(x, _CHAIN_ATTR, None)
    if ws is None:
        return None
    delattr(x, _CHAIN_ATTR)
    if not bn.training or x._version != ws[1]:
        return None                                # eval mode, or x was written to since: fall back to reading x
    return ws[0]


def bn_act_fork(x, bn, prelu=None):
    """(prelu(bn(x)), x): for a tensor with TWO consumers, the normalisation and a skip connection (ref iresnet.py:56-67
    `identity = x; out = self.bn1(x)`).  Use the second output for the skip branch: the backward kernel then adds the
    skip gradient while it writes dx, instead of autograd summing the two contributions in one more pass."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return bn_act(x, bn, prelu), x
    _check_bn_args(x, bn, prelu)
    a = prelu.weight if prelu is not None else None
    return _BNAct.apply(x, bn.weight, bn.bias, a, None, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                        bn.training, bn.momentum, bn.eps, True, False, _take_chained_ws(x, bn))


def _check_bn_args(x, bn, prelu):
    if x.dim() != 4:
        raise ValueError("bn_act expects a 4-D (B, C, H, W) tensor")
    if bn.momentum is None or not bn.track_running_stats or not bn.affine:
        raise RuntimeError("bn_act supports affine BatchNorm2d with running statistics and a fixed momentum")
    if prelu is not None and prelu.weight.numel() != x.shape[1]:
        raise ValueError("bn_act: PReLU must have one slope per channel")


def bn_act(x, bn, prelu=None, res=None, emit_next_stats=False):
    """y = prelu(bn(x) [+ res]) with ``bn`` an nn.BatchNorm2d and ``prelu`` an nn.PReLU (or None): statistics,
    normalisation, residual add and activation in two fused passes (forward) / two (backward).

    emit_next_stats: y is the input of ANOTHER training-mode BatchNorm next (the end of one residual unit feeding the next
    unit's bn1, ref iresnet.py:56-67): the apply pass also leaves the batch statistics of y in that op's workspace, which
    travels with y; the next bn_act / bn_act_fork over y picks it up and skips its own statistics pass."""
    _check_bn_args(x, bn, prelu)
    a = prelu.weight if prelu is not None else None
    if res is not None and res.dtype != x.dtype:
        res = res.to(x.dtype)
    emit = bool(emit_next_stats) and bn.training and _chain_enabled()
    out = _BNAct.apply(x, bn.weight, bn.bias, a, res, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                       bn.training, bn.momentum, bn.eps, False, emit, _take_chained_ws(x, bn))
    if emit:
        y, next_ws = out
        setattr(y, _CHAIN_ATTR, (next_ws, y._version))
        return y
    return out


# --------------------------------------------------------------------------------------------
# K-B  DAP + argmax mask                        ref backbones/osb/unet.py:158-161,223; train.py:357
# --------------------------------------------------------------------------------------------
class _DAP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, want_mask):
        require_cuda(x)
        lib = load()
        B, CK, H, W = x.shape
        kk = k * k
        if CK % kk:
            raise ValueError("DAP: channels %d not divisible by k*k=%d" % (CK, kk))
        G = CK // kk
        x_d = _dense(x)
        cl = x_d.is_contiguous(memory_format=torch.channels_last) and not x_d.is_contiguous()
        y = torch.empty((B, G, H, W), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        mask = torch.empty((B, H, W), dtype=torch.int64, device=x.device) if want_mask else None
        check(lib.msml_dap_fwd(_ptr(x_d), _ptr(y), _ptr(mask), B, G, kk, H, W, int(cl), dtype_code(x.dtype), stream_ptr()))
        ctx.cfg = (kk, G, cl)
        if want_mask:
            ctx.mark_non_differentiable(mask)
            return y, mask
        return y

    @staticmethod
    def backward(ctx, dy, *_):
        kk, G, cl = ctx.cfg
        lib = load()
        B, _, H, W = dy.shape
        dy_d = dy.contiguous(memory_format=torch.channels_last) if cl else dy.contiguous()
        dx = torch.empty((B, G * kk, H, W), dtype=dy.dtype, device=dy.device,
                         memory_format=torch.channels_last if cl else torch.contiguous_format)
        check(lib.msml_dap_bwd(_ptr(dy_d), _ptr(dx), B, G, kk, H, W, int(cl), dtype_code(dy.dtype), stream_ptr()))
        return dx, None, None


def dap(x, k=3):
    """PixelShuffle(k)+AvgPool2d(k) == mean over k*k channel groups; (B, G*k*k, H, W) -> (B, G, H, W)."""
    return _DAP.apply(x, k, False)


def dap_with_mask(x, k=3):
    """-> (seg (B,G,H,W), argmax mask (B,H,W) int64 with first-index tie rule), one kernel."""
    return _DAP.apply(x, k, True)


# --------------------------------------------------------------------------------------------
# K-S  structure-via-consensus segmentation criterion          ref tricks/consensus_loss.py:63-178
# --------------------------------------------------------------------------------------------
class _Consensus(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, blobs, target, alpha, beta, pixel_all, kl_all, K):
        lib = load()
        N, C, H, W = logit.shape
        z = _dense(logit)
        cl = z.is_contiguous(memory_format=torch.channels_last) and not z.is_contiguous()
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        coef = torch.empty(2 * K * N * C, dtype=torch.float32, device=z.device)
        ws_bytes = lib.msml_consensus_workspace(N, C, H * W, K)
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=z.device)
        check(lib.msml_consensus_fwd(_ptr(z), _ptr(blobs), _ptr(target), N, C, H * W, K, int(cl), dtype_code(z.dtype),
                                     float(alpha), float(beta), int(pixel_all), int(kl_all), _ptr(loss), _ptr(coef), _ptr(ws),
                                     ws_bytes, stream_ptr()))
        ctx.save_for_backward(z, blobs, coef)
        ctx.cfg = (K, cl)
        return loss

    @staticmethod
    def backward(ctx, gout):
        z, blobs, coef = ctx.saved_tensors
        K, cl = ctx.cfg
        N, C, H, W = z.shape
        g = gout.to(torch.float32).contiguous()
        dz = torch.empty_like(z)
        check(load().msml_consensus_bwd(_ptr(z), _ptr(blobs), _ptr(coef), _ptr(g), _ptr(dz), N, C, H * W, K, int(cl),
                                        dtype_code(z.dtype), stream_ptr()))
        return dz, None, None, None, None, None, None, None


def consensus_loss(logit, blobs, target, alpha=10.0, beta=5.0, reduce_pixel="idx", reduce_pixel_kl="idx", num_blobs=2):
    """Structure-via-consensus loss of ref tricks/consensus_loss.py: logit (N, C, H, W), 2 <= C <= 4; blobs (N, 1, H, W)
    or (N, H, W) with integer ids in [0, num_blobs) (-1: pixel of no blob); target (N, H, W) labels -> 0-dim fp32 loss."""
    require_cuda(logit, blobs, target)
    if logit.dim() != 4:
        raise ValueError("consensus_loss expects logits of shape (N, C, H, W)")
    N, C, H, W = logit.shape
    if blobs.numel() != N * H * W or target.numel() != N * H * W:
        raise ValueError("consensus_loss: blobs %s / target %s do not match logits %s" % (tuple(blobs.shape), tuple(target.shape),
                                                                                         tuple(logit.shape)))
    b = blobs.reshape(N, H * W).to(torch.int64).contiguous()
    t = target.reshape(N, H * W).to(torch.int64).contiguous()
    return _Consensus.apply(logit, b, t, alpha, beta, reduce_pixel == "all", reduce_pixel_kl == "all", int(num_blobs))


# --------------------------------------------------------------------------------------------
# tcgen05 GEMM + in-model margin heads          ref headers/margin_losses.py:275-303,390-418
# --------------------------------------------------------------------------------------------
def _pad_k(t):
    """bf16 (rows, k) -> contiguous with k padded to a multiple of 8 (TMA pitch rule)."""
    t = t.to(torch.bfloat16)
    k = t.shape[1]
    if k % 8:
        t = torch.nn.functional.pad(t, (0, 8 - k % 8))
    return t.contiguous(), k


def gemm_tn(a, b):
    """fp32 (M, N) = a (M, K) @ b (N, K)^T on tcgen05 tensor cores (operands rounded to bf16)."""
    require_cuda(a, b)
    lib = load()
    a_p, k = _pad_k(a)
    b_p, k2 = _pad_k(b)
    if k != k2:
        raise ValueError("gemm_tn: K mismatch %d vs %d" % (k, k2))
    M, N = a.shape[0], b.shape[0]
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    check(lib.msml_gemm_bf16_tn(_ptr(a_p), a_p.shape[1], _ptr(b_p), b_p.shape[1], _ptr(c), N, M, N, k, stream_ptr()))
    return c


def gemm(a, b, a_mn=False, b_mn=False, block_n=256, pair=False):
    """fp32 (M, N) = op(a) @ op(b)^T; a_mn / b_mn: the operand is stored (K, M) / (K, N) row-major and read in
    place through MN-major UMMA descriptors (no transposed copy).  pair=True runs the CTA-pair mainloop
    (tcgen05.mma.cta_group::2, 256 x block_n tiles, block_n in {128, 256})."""
    require_cuda(a, b)
    lib = load()
    a = a.to(torch.bfloat16).contiguous()
    b = b.to(torch.bfloat16).contiguous()
    (K, M) = a.shape if a_mn else a.shape[::-1]
    (K2, N) = b.shape if b_mn else b.shape[::-1]
    if K != K2:
        raise ValueError("gemm: K mismatch %d vs %d" % (K, K2))
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    fn = lib.msml_gemm_bf16_pair if pair else lib.msml_gemm_bf16
    check(fn(_ptr(a), a.shape[1], int(a_mn), _ptr(b), b.shape[1], int(b_mn), _ptr(c), N, M, N, K, block_n, stream_ptr()))
    return c


class _CosineGemm(torch.autograd.Function):
    """cos = en @ wn^T with both gradients, all three contractions on tcgen05."""

    @staticmethod
    def forward(ctx, en, wn):
        ctx.save_for_backward(en, wn)
        return gemm_tn(en, wn)

    @staticmethod
    def backward(ctx, dcos):
        en, wn = ctx.saved_tensors
        pad = (-dcos.shape[1]) % 8            # TMA pitch rule: row pitch multiple of 16 bytes
        d16 = torch.nn.functional.pad(dcos, (0, pad)).to(torch.bfloat16) if pad else dcos.to(torch.bfloat16)
        wn16 = wn.to(torch.bfloat16)
        en16 = en.to(torch.bfloat16)
        if pad:
            wn16 = torch.nn.functional.pad(wn16, (0, 0, 0, pad))
        d_en = gemm(d16, wn16, a_mn=False, b_mn=True)                      # dcos (B,C) x Wn (C,D) read in place
        d_wn = gemm(d16, en16, a_mn=True, b_mn=True)[:wn.shape[0]]        # dcos^T (C,B) x En (B,D), both in place
        return d_en.to(en.dtype), d_wn.to(wn.dtype)


class _Margin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cos, label, kind, s, m, a, k):
        require_cuda(cos, label)
        lib = load()
        cos = cos.contiguous()
        label = label.contiguous().to(torch.int64)
        mp = _lib.margin_params(kind, s, m, a, k)
        ctx.save_for_backward(cos, label)
        ctx.mp = mp
        out = cos.clone()
        B, C = out.shape
        check(lib.msml_margin_fwd(_ptr(out), _ptr(label), B, C, C, ctypes.byref(mp), stream_ptr()))
        return out

    @staticmethod
    def backward(ctx, dl):
        cos, label = ctx.saved_tensors
        lib = load()
        g = dl.to(torch.float32).contiguous().clone()
        B, C = g.shape
        check(lib.msml_margin_bwd(_ptr(g), _ptr(cos), _ptr(label), B, C, C, ctypes.byref(ctx.mp), stream_ptr()))
        return g, None, None, None, None, None, None


def cosine_logits(en, wn):
    return _CosineGemm.apply(en, wn)


def margin_logits(cos, label, kind, s, m, a=0.0, k=0.0):
    """In-place-free margin + scale on a cosine matrix (rows with label -1 get scale only)."""
    return _Margin.apply(cos.float(), label, kind, s, m, a, k)


# --------------------------------------------------------------------------------------------
# Convolution plumbing (cuDNN does the math): bf16 shadow weights + direct weight-gradient accumulation.
# Under autocast every convolution casts its fp32 weight to bf16 (one kernel), casts the bf16 weight gradient
# back (another) and AccumulateGrad adds it into .grad (a third).  engine.TrainStep keeps ONE bf16 shadow per
# weight, refreshed for all weights by a single multi-tensor copy per step; the shadow enters the graph through
# _ShadowWeight, whose backward adds the bf16 gradient straight into the fp32 flat-gradient view.
# Without an engine (no ``_msml_shadow`` attribute) these helpers are exactly ``module(x)``.
# --------------------------------------------------------------------------------------------
class _ShadowWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w):
        ctx.param = w
        return w._msml_shadow.view_as(w._msml_shadow)

    @staticmethod
    def backward(ctx, g):
        p = ctx.param
        if _direct_grad(p):
            if g.dtype == torch.bfloat16 and g.stride() == p.grad.stride() and _is_dense(g):
                _PENDING_WGRADS.append((p.grad, g))     # added by ONE multi-tensor kernel in flush_weight_grads()
            else:
                p.grad.add_(g)                          # e.g. the slice of a zero-padded weight's gradient
            return None
        return g.to(p.dtype)


_PENDING_WGRADS = []


def _is_dense(t):
    """True when t's elements occupy one gap-free block of memory (any permutation of dims)."""
    expect = 1
    for size, stride in sorted(zip(t.shape, t.stride()), key=lambda q: q[1]):
        if size == 1:
            continue
        if stride != expect:
            return False
        expect *= size
    return True


def discard_pending_weight_grads():
    """Drop queued weight gradients (a backward pass that raised half-way would otherwise leak them into the next step)."""
    _PENDING_WGRADS.clear()
    join_weight_grad_stream()


def flush_weight_grads():
    """Add every bf16 weight gradient queued by _ShadowWeight.backward into its fp32 flat-gradient view: one launch
    (msml_accum_bf16_multi) for all of them.  engine.TrainStep calls this right after the backward pass."""
    join_weight_grad_stream()
    if not _PENDING_WGRADS:
        return
    lib = load()
    n = len(_PENDING_WGRADS)
    dst = (ctypes.c_void_p * n)(*[d.data_ptr() for d, _ in _PENDING_WGRADS])
    src = (ctypes.c_void_p * n)(*[g.data_ptr() for _, g in _PENDING_WGRADS])
    cnt = (ctypes.c_int64 * n)(*[g.numel() for _, g in _PENDING_WGRADS])
    try:
        check(lib.msml_accum_bf16_multi(n, dst, src, cnt, stream_ptr()))
    finally:
        _PENDING_WGRADS.clear()


def _weight_of(m):
    w = m.weight
    sh = getattr(w, "_msml_shadow", None)
    # training steps only: the engine refreshes the shadows at the start of each step, so they would be one update
    # behind for an evaluation pass run between two steps
    if (sh is not None and m.training and torch.is_grad_enabled() and w.requires_grad and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == sh.dtype):
        return _ShadowWeight.apply(w)
    return w


def _shadow_active(m):
    w = m.weight
    sh = getattr(w, "_msml_shadow", None)
    return (sh is not None and m.training and torch.is_grad_enabled() and w.requires_grad and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == sh.dtype)


# Weight gradients on a side stream.  In the backward pass only the data gradient of a convolution is on the critical
# path (the next layer's backward needs it); the weight gradient is needed by the optimizer alone.  The BN / activation
# kernels between two convolutions are latency-bound and leave most SMs idle, so the wgrad kernels run on a second stream
# and fill that time; they are joined before the gradients are consumed (flush_weight_grads).  Inside the captured
# CUDA graph this becomes a fork / join of kernel nodes.
_WGRAD_STREAMS = {}
_WGRAD_KEEPALIVE = []          # (dy, x) of wgrads still running on the side stream
_WGRAD_SIDE = {"enabled": False, "dirty": False}


def set_wgrad_side_stream(enabled):
    _WGRAD_SIDE["enabled"] = bool(enabled)


def side_stream_enabled():
    return _WGRAD_SIDE["enabled"]


def run_on_side_stream(fn, *args):
    """Run ``fn(*args)`` on the side stream (forked from the current stream); returns (result, join) where ``join()`` makes
    the current stream wait for it.  Used to run the occlusion-segmentation branch concurrently with the first stage of the
    recognition branch, which does not need its outputs until the first FM operator."""
    dev = torch.cuda.current_device()
    main = torch.cuda.current_stream(dev)
    side = _wgrad_stream(torch.device("cuda", dev))
    side.wait_stream(main)
    with torch.cuda.stream(side):
        out = fn(*args)

    def join():
        torch.cuda.current_stream(dev).wait_stream(side)
    return out, join


def _wgrad_stream(device):
    st = _WGRAD_STREAMS.get(device)
    if st is None:
        st = torch.cuda.Stream(device)
        _WGRAD_STREAMS[device] = st
    return st


def join_weight_grad_stream():
    """Make the current stream wait for every weight gradient launched on the side stream."""
    if _WGRAD_SIDE["dirty"]:
        for dev, st in _WGRAD_STREAMS.items():
            torch.cuda.current_stream(dev).wait_stream(st)
        _WGRAD_SIDE["dirty"] = False
    _WGRAD_KEEPALIVE.clear()


class _ConvShadow(torch.autograd.Function):
    """conv2d(x, bf16 shadow of w) whose backward computes dgrad on the current stream and wgrad on the side stream."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, dilation, groups):
        sh = w._msml_shadow
        ctx.save_for_backward(x, sh)
        ctx.param = w
        ctx.cfg = (stride, padding, dilation, groups)
        return torch.nn.functional.conv2d(x, sh, None, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, dy):
        x, sh = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.cfg
        p = ctx.param
        cb = torch.ops.aten.convolution_backward
        out_pad = [0] * len(stride)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = cb(dy, x, sh, None, stride, padding, dilation, False, out_pad, groups, [True, False, False])[0]
        if _WGRAD_SIDE["enabled"] and _direct_grad(p):
            main = torch.cuda.current_stream(dy.device)
            side = _wgrad_stream(dy.device)
            side.wait_stream(main)                       # dy (and x) are complete on the main stream
            with torch.cuda.stream(side):
                dw = cb(dy, x, sh, None, stride, padding, dilation, False, out_pad, groups, [False, True, False])[1]
            _WGRAD_KEEPALIVE.append((dy, x))             # their memory must not be reused before the join
            _WGRAD_SIDE["dirty"] = True
        else:
            dw = cb(dy, x, sh, None, stride, padding, dilation, False, out_pad, groups, [False, True, False])[1]
        if _direct_grad(p):
            if dw.dtype == torch.bfloat16 and dw.stride() == p.grad.stride() and _is_dense(dw):
                _PENDING_WGRADS.append((p.grad, dw))
            else:
                join_weight_grad_stream()
                p.grad.add_(dw)
            return dx, None, None, None, None, None
        return dx, dw.to(p.dtype), None, None, None, None


def conv2d(x, m):
    """``m(x)`` for an nn.Conv2d, through the bf16 shadow weight when the engine installed one."""
    if not _shadow_active(m):
        return m(x)
    if m.bias is None and x.is_cuda and x.dtype == m.weight._msml_shadow.dtype and m.padding_mode == "zeros":
        return _ConvShadow.apply(x, m.weight, tuple(m.stride), tuple(m.padding), tuple(m.dilation), m.groups)
    return torch.nn.functional.conv2d(x, _ShadowWeight.apply(m.weight), m.bias, m.stride, m.padding, m.dilation, m.groups)


def conv_transpose2d(x, m):
    w = _weight_of(m)
    if w is m.weight:
        return m(x)
    return torch.nn.functional.conv_transpose2d(x, w, m.bias, m.stride, m.padding, m.output_padding, m.groups, m.dilation)


def linear(x, m):
    w = _weight_of(m)
    if w is m.weight:
        return m(x)
    return torch.nn.functional.linear(x, w, m.bias)


_ZERO_PAD = {}


def cat_channels_padded(parts, multiple=8):
    """cat(parts, dim=1) with zero channels appended up to a multiple of ``multiple`` (cuDNN's bf16 tensor-core
    kernels need C % 8 == 0; otherwise it pads the activation itself, in a separate kernel, in forward, dgrad and
    wgrad).  Returns (tensor, number of padding channels)."""
    c = sum(t.shape[1] for t in parts)
    pad = (-c) % multiple
    if pad and parts[0].is_cuda:
        B, _, H, W = parts[0].shape
        key = (B, pad, H, W, parts[0].dtype, parts[0].device)
        z = _ZERO_PAD.get(key)
        if z is None:
            z = torch.zeros((B, pad, H, W), dtype=parts[0].dtype, device=parts[0].device).contiguous(memory_format=torch.channels_last)
            _ZERO_PAD[key] = z
        return torch.cat(list(parts) + [z], dim=1), pad
    return torch.cat(list(parts), dim=1), 0


def padded_params(m, pad_in=0, pad_out=0, transposed=False):
    """(weight, bias) of a conv module with zero input / output channels appended (weight through the bf16 shadow when
    there is one).  ConvTranspose2d stores its weight as (in, out, kh, kw)."""
    w = _weight_of(m)
    if pad_in or pad_out:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_out, 0, pad_in) if transposed else (0, 0, 0, 0, 0, pad_in, 0, pad_out))
    b = m.bias
    if b is not None and pad_out:
        b = torch.nn.functional.pad(b, (0, pad_out))
    return w, b


def conv2d_padded_in(x, m, pad):
    """conv2d whose input carries ``pad`` extra zero channels: the weight gets matching zero input channels."""
    if not pad:
        return conv2d(x, m)
    w = torch.nn.functional.pad(_weight_of(m), (0, 0, 0, 0, 0, pad))
    return torch.nn.functional.conv2d(x, w, m.bias, m.stride, m.padding, m.dilation, m.groups)


# --------------------------------------------------------------------------------------------
# Backward-progress markers: identity in forward; in backward they tell the engine that every parameter used
# AFTER this point of the forward pass has its gradient complete, so that its slice of the flat gradient buffer
# can be all-reduced while the rest of the backward pass still runs (engine.TrainStep, world_size > 1).
# --------------------------------------------------------------------------------------------
_MARKER_CALLBACK = None


def set_grad_marker_callback(fn):
    global _MARKER_CALLBACK
    _MARKER_CALLBACK = fn


class _GradMarker(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if _MARKER_CALLBACK is not None:
            _MARKER_CALLBACK(ctx.tag)
        return g, None


def grad_marker(x, tag):
    if _MARKER_CALLBACK is None or not x.requires_grad or not torch.is_grad_enabled():
        return x
    y = _GradMarker.apply(x, tag)
    ws = getattr(x, _CHAIN_ATTR, None)          # same values, same version counter: chained BN statistics stay valid
    if ws is not None:
        delattr(x, _CHAIN_ATTR)
        setattr(y, _CHAIN_ATTR, (ws[0], y._version))
    return y


def launch_count():
    return load().msml_launch_count()


def launch_count_reset():
    load().msml_launch_count_reset()
