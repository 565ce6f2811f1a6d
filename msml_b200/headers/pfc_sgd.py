"""Fused optimizer for the PartialFC head.

The reference trains the class centres with a stock optimizer over ``module_partial_fc.parameters()`` and lets
PartialFC swap the sampled rows in and out of it every step (ref train.py:188-191,299-300; headers/partial_fc.py:93-94
gather, :112-114 optimizer-state surgery, :101-104 scatter back):

    opt_pfc = torch.optim.SGD([{'params': module_partial_fc.parameters()}], lr=..., momentum=0.9, weight_decay=5e-4)
    ...  opt_pfc.step(); module_partial_fc.update()

``PartialFCSGD(module_partial_fc, lr=..., momentum=0.9, weight_decay=5e-4)`` is a drop-in for that optimizer: same
hyper-parameters, same ``param_groups`` (LR schedulers work), same ``step()`` / ``zero_grad()``.  Its step is ONE kernel
(msml_pfc_sgd_update, csrc/pfc_sgd_kernels.cuh) that applies momentum SGD with weight decay to the sampled rows
in place in the shard (``weight`` / ``weight_mom``), so ``module_partial_fc.update()`` has nothing left to scatter and
becomes a no-op FOR THAT STEP: ``step()`` raises a flag on the module that ``update()`` consumes, so a module that is later
stepped by a stock optimizer again scatters its gathered rows back as the reference does.  Arithmetic is
torch.optim.SGD's with an existing momentum buffer (PartialFC always supplies one); momentum == 0 ignores dampening and
leaves the buffer untouched, as torch does.  ``capturable_lr`` in the defaults tells ``engine.TrainStep`` to turn the
learning rate into a device tensor at capture time (the kernel reads it from memory), so LR schedules survive graph replays.
"""
import ctypes

import torch

from .._lib import check, load, stream_ptr

__all__ = ["PartialFCSGD"]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class PartialFCSGD(torch.optim.Optimizer):
    def __init__(self, module, lr, momentum=0.9, dampening=0.0, weight_decay=0.0, nesterov=False, emit_normalized=False,
                 fuse_projection=False):
        if not isinstance(lr, torch.Tensor) and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        for attr in ("weight", "weight_mom", "sub_weight", "num_local", "sample_rate"):
            if not hasattr(module, attr):
                raise TypeError("PartialFCSGD drives a PartialFC module (missing attribute %r)" % attr)
        load()
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                        capturable_lr=True)
        super().__init__([{"params": list(module.parameters())}], defaults)
        self.module = module
        # emit_normalized (sample_rate 1 only): the update kernel also writes the NEXT step's unit-norm bf16 centres and
        # 1/||w|| (it holds the updated row in registers anyway), so PartialFC skips its msml_wnorm_cast pass (ref
        # partial_fc.py:115).  Opt-in because it assumes nobody else writes module.weight between two steps; call
        # module.invalidate_normalized() after loading / editing the class centres by hand.
        self.emit_normalized = bool(emit_normalized) and int(module.sample_rate) == 1
        # fuse_projection: PartialFC.forward_backward(label, features, THIS optimizer) then runs the head in raw mode —
        # sub_weight.grad holds dWn = dcos^T X, the gradient with respect to the NORMALISED centres — and step() applies the
        # backward of normalize(sub_weight) (ref partial_fc.py:115) to the row it holds in registers before the update
        # (msml_pfc_sgd_update_raw).  The dcos GEMM loses its <Wn, dWn> column reduction and the dW GEMM its Wn stream.
        # Opt-in: the gradient is only final inside this optimizer, so do not read or clip sub_weight.grad yourself.
        self.fuse_projection = bool(fuse_projection)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        m = self.module
        dw = m.sub_weight.grad
        if dw is None:
            return loss
        g = self.param_groups[-1]
        if dw.dtype != torch.float32 or not dw.is_contiguous():
            dw = dw.to(torch.float32).contiguous()
        n_s, D = dw.shape
        index = None if int(m.sample_rate) == 1 else m.index
        if index is not None and index.numel() != n_s:
            raise RuntimeError("PartialFCSGD: gradient has %d rows but the sample index has %d" % (n_s, index.numel()))
        lr = g["lr"]
        lr_dev = lr if isinstance(lr, torch.Tensor) and lr.is_cuda else None        # captured steps read it from memory
        if lr_dev is not None and lr_dev.dtype != torch.float32:
            raise RuntimeError("PartialFCSGD: a tensor lr must be float32")
        wn = inv = None
        if self.emit_normalized:
            wn = m._buf("wn", (n_s, D), torch.bfloat16)
            inv = m._buf("inv_norm", (n_s,), torch.float32)
        raw = bool(getattr(m, "_grad_is_raw", False))
        if raw and not self.fuse_projection:
            raise RuntimeError("PartialFCSGD: the module produced a raw (unprojected) gradient for another optimizer")
        fn = load().msml_pfc_sgd_update_raw if raw else load().msml_pfc_sgd_update
        check(fn(_ptr(m.weight), _ptr(m.weight_mom), _ptr(dw), _ptr(index), n_s, m.num_local, D,
                 _ptr(lr_dev), 0.0 if lr_dev is not None else float(lr), float(g["momentum"]),
                 float(g["weight_decay"]), float(g["dampening"]), int(bool(g["nesterov"])), _ptr(wn), _ptr(inv), stream_ptr()))
        m._wn_fresh = self.emit_normalized
        m._fused_step_done = True           # this step's rows are already in the shard: the next update() has nothing to scatter
        return loss
