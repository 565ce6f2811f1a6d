"""FlatSGD: the backbone optimizer of the training step as ONE kernel.

The reference steps ``torch.optim.SGD(backbone.parameters(), lr, momentum=0.9, weight_decay=5e-4)`` after
``clip_grad_norm_(backbone.parameters(), 5)`` (ref train.py:186-191, 297-299).  engine.TrainStep already keeps every
backbone gradient in one flat fp32 buffer; FlatSGD lays the parameters and the momentum buffers out the same way
(``bind_flat``: each ``p.data`` / ``state[p]['momentum_buffer']`` becomes a view at the gradient's offset) so that the
whole update is one streaming pass of ``msml_sgd_flat`` (csrc/sgd_flat_kernels.cuh), which also writes the bf16 copy
of the new weights that the next step's autocast convolutions read.  On a B200 that pass replaces torch's
multi-tensor fused SGD over ~300 tensors plus the multi-tensor fp32 -> bf16 weight copy.

It IS a torch.optim.SGD (param_groups, LR schedulers, state_dict with per-parameter ``momentum_buffer``), restricted
to what the kernel implements: one parameter group, dampening 0, no ``maximize``.  ``grad_scale`` follows the protocol
of torch's fused optimizers: a 0-dim fp32 device tensor the gradients are divided by inside the kernel (the engine
passes clip coefficient x world size).  There is no CPU / unbound fallback: ``step()`` raises until ``bind_flat`` ran.
"""
import torch

from ._lib import check, load, require_cuda, stream_ptr

__all__ = ["FlatSGD"]


def _pad4(k):
    return (k + 3) // 4 * 4


class FlatSGD(torch.optim.SGD):
    def __init__(self, params, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if dampening != 0.0:
            raise ValueError("FlatSGD implements dampening == 0 (ref train.py:186-191 uses the default)")
        super().__init__(params, lr=lr, momentum=momentum, dampening=0.0, weight_decay=weight_decay, nesterov=nesterov,
                         foreach=False, fused=False)
        if len(self.param_groups) != 1:
            raise ValueError("FlatSGD takes one parameter group (uniform lr / momentum / weight decay, as the reference's backbone optimizer)")
        self.param_groups[0]["capturable_lr"] = True        # engine.TrainStep turns lr into a device scalar before capture
        self.grad_scale = None
        self._w = self._m = self._g = self._shadow = self._lr_buf = None
        self._offset, self._bound = {}, []

    # ------------------------------------------------------------------ layout
    @torch.no_grad()
    def bind_flat(self, params, flat_grad):
        """``params`` in buffer order; ``p.grad`` must already be the view of ``flat_grad`` (fp32, 16-byte aligned) at the
        running offset, every parameter padded to a multiple of 4 elements.  Moves the parameters into a flat buffer with the
        same offsets (``p.data`` is re-pointed; values, shapes and strides are kept) and creates the momentum views."""
        require_cuda(flat_grad)
        if flat_grad.dtype != torch.float32 or flat_grad.dim() != 1 or flat_grad.data_ptr() % 16:
            raise ValueError("flat_grad must be a 1-D, 16-byte aligned fp32 buffer")
        own = {id(p) for p in self.param_groups[0]["params"]}
        n = flat_grad.numel()
        if n != sum(_pad4(p.numel()) for p in params):
            raise ValueError("flat_grad does not match the parameters (each padded to a multiple of 4 elements)")
        momentum = self.param_groups[0]["momentum"]
        self._w = torch.zeros(n, dtype=torch.float32, device=flat_grad.device)
        self._m = torch.zeros(n, dtype=torch.float32, device=flat_grad.device) if momentum != 0 else None
        self._g, self._shadow = flat_grad, None
        self._lr_buf = torch.zeros((), dtype=torch.float32, device=flat_grad.device)
        self._offset, self._bound, off = {}, list(params), 0
        for p in params:
            if id(p) not in own:
                raise ValueError("bind_flat: a parameter that this optimizer does not own")
            if p.dtype != torch.float32 or p.device != flat_grad.device:
                raise ValueError("bind_flat: fp32 parameters on the gradient buffer's device only")
            if p.grad is None or p.grad.data_ptr() != flat_grad.data_ptr() + 4 * off or p.grad.stride() != p.stride():
                raise ValueError("bind_flat: p.grad is not the view of flat_grad at offset %d" % off)
            view = torch.as_strided(self._w, p.size(), p.stride(), off)
            view.copy_(p)
            p.data = view
            if self._m is not None:
                mview = torch.as_strided(self._m, p.size(), p.stride(), off)
                old = self.state[p].get("momentum_buffer") if p in self.state else None
                if old is not None:
                    mview.copy_(old)
                self.state[p]["momentum_buffer"] = mview
            self._offset[id(p)] = off
            off += _pad4(p.numel())

    def is_bound(self):
        return self._w is not None

    def shadow_view(self, p):
        """bf16 tensor with p's shape and strides inside the flat shadow buffer that every ``step()`` rewrites from the new
        weights (None for a parameter that is not bound).  Call ``refresh_shadows()`` once after taking the views, and
        whenever weights were changed by anything but ``step()``."""
        off = self._offset.get(id(p))
        if off is None:
            return None
        if self._shadow is None:
            self._shadow = torch.zeros(self._w.numel(), dtype=torch.bfloat16, device=self._w.device)
        return torch.as_strided(self._shadow, p.size(), p.stride(), off)

    @torch.no_grad()
    def refresh_shadows(self):
        if self._shadow is not None:
            self._shadow.copy_(self._w)

    # ------------------------------------------------------------------ torch.optim.Optimizer surface
    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FlatSGD.step() takes no closure")
        if self._w is None:
            raise RuntimeError("FlatSGD.step(): not bound to a flat gradient buffer (engine.TrainStep binds it; see bind_flat)")
        g = self.param_groups[0]
        if g.get("maximize") or g["dampening"] != 0:
            raise RuntimeError("FlatSGD implements dampening == 0 and maximize == False")
        lr = g["lr"]
        if isinstance(lr, torch.Tensor):
            if lr.device != self._w.device or lr.dtype != torch.float32:
                raise RuntimeError("FlatSGD: a tensor lr must be a fp32 scalar on the parameters' device")
        else:
            lr = self._lr_buf.fill_(float(lr))             # a python lr is baked into a captured graph (engine makes it a tensor)
        scale = self.grad_scale
        if scale is not None and (scale.device != self._w.device or scale.dtype != torch.float32 or scale.numel() != 1):
            raise RuntimeError("FlatSGD.grad_scale must be a fp32 scalar on the parameters' device")
        check(load().msml_sgd_flat(self._w.data_ptr(), self._m.data_ptr() if self._m is not None else None, self._g.data_ptr(),
                                   self._shadow.data_ptr() if self._shadow is not None else None, self._w.numel(), lr.data_ptr(),
                                   scale.data_ptr() if scale is not None else None, float(g["momentum"]), float(g["weight_decay"]),
                                   int(bool(g["nesterov"])), stream_ptr()))
        return None

    def zero_grad(self, set_to_none=True):
        """The gradients are views of the flat buffer and must stay: zero it instead of dropping them."""
        if self._g is None:
            return super().zero_grad(set_to_none)
        self._g.zero_()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.param_groups[0].setdefault("capturable_lr", True)
        if self._m is None:
            return
        for p in self._bound:                              # torch deep-copies the loaded state: move it back into the views
            mview = torch.as_strided(self._m, p.size(), p.stride(), self._offset[id(p)])
            buf = self.state[p].get("momentum_buffer") if p in self.state else None
            if buf is not None and buf.data_ptr() != mview.data_ptr():
                mview.copy_(buf)
            elif buf is None:
                mview.zero_()
            self.state[p]["momentum_buffer"] = mview
