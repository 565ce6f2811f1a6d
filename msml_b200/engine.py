"""Training-step engine for the MSML + PartialFC hot path on B200.

The reference's step (train.py:283-300, PartialFC variant) is a few thousand small kernels; on a
B200 that is launch-bound from Python.  ``TrainStep`` keeps the reference's semantics

    features, _ = backbone(img)
    x_grad, loss = pfc.forward_backward(label, F.normalize(features), opt_pfc)
    features.backward(x_grad); clip_grad_norm_(backbone, 5); opt_backbone.step(); opt_pfc.step(); pfc.update()

and runs it the B200 way:
  * gradients of all used backbone parameters live in ONE flat fp32 buffer ordered by backward completion; one NCCL
    all-reduce per stage starts while the rest of the backward pass still runs; the clip coefficient and the 1/W of the
    mean ride in the fused SGD's grad_scale;
  * conv / linear weights have bf16 shadows refreshed by one multi-tensor copy; their bf16 gradients are added into the
    flat buffer by one launch; BN / PReLU gradients are added by the BN backward kernels themselves;
  * with ``FlatSGD`` (flat_sgd.py) as the backbone optimizer the parameters and momentum buffers are flat too and the
    whole backbone update, clip scale included, is ONE kernel that also writes next step's bf16 shadow weights;
  * two streams: convolution weight gradients (off the critical path) and the occlusion-segmentation branch (not needed
    before the first FM operator) run on a side stream and fill the SMs that the latency-bound BN kernels leave idle;
  * after three eager warm-up steps the whole step — cuDNN convolutions, this library's kernels, NCCL collectives, the
    SGD updates, the stream forks and joins — is captured into a CUDA graph and replayed with no Python or launch
    overhead; the next batch's host-to-device copy runs under the current step (``prefetch``).

Limits of the captured mode (documented, checked): with a sampled PartialFC (sample_rate < 1) the gathered batch must
not exceed num_sample (the positives-outnumber-the-sample branch has a data-dependent size); the sampling itself —
torch.rand through the graph-registered generator, radix select, gathers into fixed-capacity buffers — is captured.  With ``fused=True`` optimizers the learning rates are turned into device tensors
at capture time, so torch LR schedulers keep working across replays; with other optimizers, and for momentum /
weight decay, the values are baked in (call ``recapture()`` after changing them).  ``use_graph=False`` runs the identical step eagerly.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .flat_sgd import FlatSGD  # noqa: F401  (re-exported: the engine's backbone optimizer)


class TrainStep:
    def __init__(self, backbone, pfc, opt_backbone, opt_pfc, batch_shape, world_size=1, max_norm=5.0,
                 use_graph=True, device=None, wgrad_side_stream=True):
        self.backbone, self.pfc = backbone, pfc
        self.opt_backbone, self.opt_pfc = opt_backbone, opt_pfc
        self.world_size, self.max_norm, self.use_graph = world_size, max_norm, use_graph
        self.device = device or next(backbone.parameters()).device
        # NHWC everywhere: channels-last weights spare cuDNN a weight-layout conversion per convolution, forward and wgrad
        self.backbone.to(memory_format=torch.channels_last)
        self.static_img = torch.zeros(batch_shape, device=self.device).contiguous(memory_format=torch.channels_last)
        self.static_label = torch.zeros(batch_shape[0], dtype=torch.int64, device=self.device)
        self.static_loss = None
        self.flat = None
        self.graph = None
        self._used = None
        self._shadow_src, self._shadow_dst = [], []
        self._emitted, self._emitted_versions = [], None      # weights whose bf16 shadows the optimizer kernel writes (FlatSGD)
        self._copy_stream, self._has_staged = None, False
        # convolution weight gradients run on a second stream and fill the idle SMs under the latency-bound BN kernels
        self.wgrad_side_stream = wgrad_side_stream
        if use_graph and int(pfc.sample_rate) != 1 and batch_shape[0] * world_size > pfc.num_sample:
            raise ValueError("captured TrainStep with a sampled PartialFC needs batch * world_size <= num_sample (the branch of "
                             "ref partial_fc.py:89-90 has a data-dependent size); use use_graph=False")

    # ------------------------------------------------------------------ one eager step
    def _discover_used_params(self):
        """One throw-away forward/backward to find which parameters receive gradients (the OSB
        branch does not when there is no segmentation loss: its outputs are detached, ref
        unet.py:227-230) and to build the flat gradient buffer over exactly those."""
        bn_state = {k: v.clone() for k, v in self.backbone.state_dict().items() if "running_" in k or "num_batches" in k}
        feat, _ = self.backbone(self.static_img.normal_())
        feat.sum().backward()
        self.backbone.load_state_dict(bn_state, strict=False)      # the probe must not touch the BN statistics
        used = [p for p in self.backbone.parameters() if p.grad is not None]
        used, self._buckets = self._order_for_overlap(used)
        pad4 = lambda k: (k + 3) // 4 * 4     # every gradient starts on a 16-byte boundary (vector accesses in the kernels)
        n = sum(pad4(p.numel()) for p in used)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        off = 0
        for p in used:      # same strides as the parameter (conv weights are channels-last): no layout conversion per step
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)
            p._msml_direct_grad = True          # fused kernels may add their parameter gradients into p.grad themselves
            off += pad4(p.numel())
        self._used = used
        if hasattr(self.opt_backbone, "bind_flat"):
            # FlatSGD: parameters and momentum buffers move into flat buffers with the gradients' offsets, and the whole
            # optimizer step (+ next step's bf16 shadow weights) becomes one kernel
            self.opt_backbone.bind_flat(used, self.flat)
        # element ranges of the flat buffer per bucket (bucket i is complete when the backward pass crosses marker i)
        self._bucket_ranges, lo = {}, 0
        for tag, cnt in self._buckets:
            hi = lo + sum(pad4(p.numel()) for p in used[self._bucket_start[tag]:self._bucket_start[tag] + cnt])
            self._bucket_ranges[tag] = (lo, hi)
            lo = hi
        self.static_img.zero_()
        self._install_shadows()

    def _order_for_overlap(self, used):
        """Order the parameters so that the ones whose gradients finish first in the backward pass come first in the
        flat buffer: [stage 3 (layer4, fm_ops.3, bn2, fc, features) | stage 2 | stage 1 | stage 0 | rest (stem, OSB)].
        Returns (ordered list, [(tag, count)])."""
        names = {id(p): n for n, p in self.backbone.named_parameters()}

        def stage_of(name):
            for i in (3, 2, 1, 0):
                if name.startswith("frb.layer%d." % (i + 1)) or name.startswith("frb.fm_ops.%d." % i):
                    return i
            if name.startswith(("frb.bn2.", "frb.fc.", "frb.features.")):
                return 3
            return -1                                   # stem and anything outside the FRB stages: reduced last
        groups = {3: [], 2: [], 1: [], 0: [], -1: []}
        for p in used:
            groups[stage_of(names.get(id(p), ""))].append(p)
        ordered, buckets, self._bucket_start = [], [], {}
        for tag in (3, 2, 1, 0, -1):
            self._bucket_start[tag] = len(ordered)
            buckets.append((tag, len(groups[tag])))
            ordered += groups[tag]
        return ordered, buckets

    def _on_marker(self, tag):
        """Backward crossed marker `tag`: the gradients of bucket `tag` are complete (after the queued bf16 weight
        gradients are flushed) -> start its all-reduce; NCCL's stream runs it under the rest of the backward pass."""
        if self.world_size <= 1 or tag in self._reduced:
            return
        ops.flush_weight_grads()
        lo, hi = self._bucket_ranges[tag]
        if hi > lo:
            self._works.append(dist.all_reduce(self.flat[lo:hi], async_op=True))
        self._reduced.add(tag)

    def share_state_from(self, other):
        """Make this TrainStep operate on `other`'s flat gradient buffer, buckets and shadow weights (bench.py runs an
        eager twin of the captured step to time individual launches)."""
        for k in ("flat", "_used", "_buckets", "_bucket_ranges", "_bucket_start", "_shadow_src", "_shadow_dst", "_emitted",
                  "_emitted_versions"):
            setattr(self, k, getattr(other, k))

    def _install_shadows(self):
        """bf16 shadow of every conv / linear weight that runs under autocast (ops._ShadowWeight): refreshed by ONE
        multi-tensor copy per step instead of one cast kernel per layer, and gradients flow into the flat buffer."""
        self._shadow_src, self._shadow_dst = [], []
        self._emitted, self._emitted_versions = [], None
        if not getattr(self.backbone, "fp16", False):
            return
        flat_opt = self.opt_backbone if hasattr(self.opt_backbone, "shadow_view") else None
        for m in self.backbone.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
                w = m.weight
                emitted = flat_opt.shadow_view(w) if flat_opt is not None else None
                if emitted is not None:                 # rewritten by every FlatSGD.step(): no per-step copy
                    w._msml_shadow = emitted
                    self._emitted.append(w)
                    continue
                w._msml_shadow = torch.empty_like(w, dtype=torch.bfloat16)      # preserves strides (channels-last weights)
                self._shadow_src.append(w.detach())
                self._shadow_dst.append(w._msml_shadow)
        self.refresh_shadows()

    def refresh_shadows(self):
        """Recompute the optimizer-emitted bf16 shadow weights from the fp32 weights.  Done automatically when a weight
        was written in place through the Parameter (``load_state_dict``, ``p.copy_``: its version counter moves); call it
        yourself after writes that bypass the counter (``p.data`` / ``p.detach()`` aliases, ``dist.broadcast(p.data)``)."""
        if self._emitted:
            self.opt_backbone.refresh_shadows()
            self._emitted_versions = [w._version for w in self._emitted]

    def _sync_shadows(self):
        if self._emitted and self._emitted_versions != [w._version for w in self._emitted]:
            self.refresh_shadows()

    def _step(self, img, label):
        ops.discard_pending_weight_grads()          # nothing may survive from an aborted earlier step
        self.flat.zero_()
        if self._shadow_dst:
            torch._foreach_copy_(self._shadow_dst, self._shadow_src)
        self._works, self._reduced = [], set()
        ops.set_grad_marker_callback(self._on_marker if self.world_size > 1 else None)   # markers are placed in forward
        ops.set_wgrad_side_stream(self.wgrad_side_stream)
        try:
            return self._step_body(img, label)
        finally:
            ops.set_grad_marker_callback(None)
            ops.set_wgrad_side_stream(False)

    def _step_body(self, img, label):
        feat, _seg = self.backbone(img)
        featn = F.normalize(feat)
        x_grad, loss = self.pfc.forward_backward(label, featn, self.opt_pfc)
        featn.backward(x_grad)                      # accumulates into the views of self.flat
        ops.flush_weight_grads()                    # the queued bf16 weight gradients, one multi-tensor launch
        if self.world_size > 1:
            # buckets 3..0 were started by the markers while the backward pass was still running (NCCL over NVLink on its
            # own stream); what is left is the small tail (stem, stage 0 if its marker did not fire)
            for tag, _cnt in self._buckets:
                if tag not in self._reduced:
                    lo, hi = self._bucket_ranges[tag]
                    if hi > lo:
                        self._works.append(dist.all_reduce(self.flat[lo:hi], async_op=True))
            for w in self._works:
                w.wait()
        fused_clip = False
        if self.max_norm is not None:               # == clip_grad_norm_(used params, max_norm)
            norm = torch.linalg.vector_norm(self.flat)
            if self.world_size > 1:
                # flat still holds the SUM over ranks: fold the 1/world_size of the mean into the same scale
                norm = norm / self.world_size
            if self._fused_sgd_takes_scale():
                # torch's fused SGD divides every gradient by `grad_scale` inside its own kernel (the GradScaler hook):
                # clipping (and the 1/world_size of the all-reduce mean) costs no extra pass over the 48 M gradients
                inv = torch.clamp((norm + 1e-6) / self.max_norm, min=1.0)
                self.opt_backbone.grad_scale = (inv * self.world_size if self.world_size > 1 else inv).reshape(())
                fused_clip = True
            else:
                coef = torch.clamp(self.max_norm / (norm + 1e-6), max=1.0)
                self.flat.mul_(coef / self.world_size if self.world_size > 1 else coef)
        elif self.world_size > 1:
            self.flat.div_(self.world_size)
        self.opt_backbone.step()
        if fused_clip:
            self.opt_backbone.grad_scale = None
        self.opt_pfc.step()
        self.pfc.update()
        self.pfc.sub_weight.grad = None
        return loss

    def _fused_sgd_takes_scale(self):
        opt = self.opt_backbone
        if hasattr(opt, "bind_flat"):               # FlatSGD divides by grad_scale inside its kernel, like torch's fused SGD
            return True
        return (isinstance(opt, torch.optim.SGD) and all(g.get("fused") for g in opt.param_groups)
                and all(not g.get("maximize") for g in opt.param_groups))

    # ------------------------------------------------------------------ capture / replay
    def _prepare(self):
        if self._used is None:
            self._discover_used_params()

    def _snapshot(self):
        snap = [t.detach().clone() for t in list(self.backbone.parameters()) + list(self.backbone.buffers())]
        snap += [self.pfc.weight.clone(), self.pfc.weight_mom.clone()]
        return snap

    def _restore(self, snap, had_momentum):
        with torch.no_grad():
            live = list(self.backbone.parameters()) + list(self.backbone.buffers()) + [self.pfc.weight, self.pfc.weight_mom]
            for t, s in zip(live, snap):
                t.copy_(s)                          # in place: the captured graph keeps pointing at the live tensors
            if not had_momentum:                    # momentum buffers created by the warm-up start from zero again
                for st in self.opt_backbone.state.values():
                    if st.get("momentum_buffer") is not None:
                        st["momentum_buffer"].zero_()

    def _make_lr_capturable(self):
        """Learning rates become 0-dim device tensors: the captured optimizer kernels read them from memory, and torch's LR
        schedulers update a tensor lr in place (``param_group["lr"].fill_``), so a schedule that steps every iteration
        (ref train.py:193-197, 238) works across replays without recapturing."""
        for opt in (self.opt_backbone, self.opt_pfc):
            if opt is None:
                continue
            for g in opt.param_groups:
                if not (g.get("fused") or g.get("capturable_lr")):
                    continue        # only torch's fused kernels and headers.PartialFCSGD take a tensor lr without a host read
                if not isinstance(g["lr"], torch.Tensor):
                    g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=self.device)
                elif g["lr"].device != self.device:
                    g["lr"] = g["lr"].to(self.device)

    def recapture(self, preserve_state=True):
        """Three eager warm-up steps on noise (lazy state: momentum buffers, cuDNN autotuning, workspaces), then the
        capture.  The warm-up and the capture pass train on noise, so model / optimizer state is snapshotted before
        and restored afterwards unless preserve_state is False."""
        self._prepare()
        self._sync_shadows()
        self.graph = None
        self._make_lr_capturable()
        had_momentum = any(st.get("momentum_buffer") is not None for st in self.opt_backbone.state.values())
        snap = self._snapshot() if preserve_state else None
        mom_snap = ([st["momentum_buffer"].clone() for st in self.opt_backbone.state.values() if st.get("momentum_buffer") is not None]
                    if preserve_state and had_momentum else None)
        self.static_img.normal_()
        self.static_label.random_(0, self.pfc.num_classes)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step(self.static_img, self.static_label)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.static_loss = self._step(self.static_img, self.static_label)
        self.graph = g
        if preserve_state:
            self._restore(snap, had_momentum)
            self.refresh_shadows()                  # the restore rewrote the weights behind the optimizer-emitted shadows
            if getattr(self.opt_pfc, "emit_normalized", False):
                # the captured step expects the previous step's normalised centres (headers.PartialFCSGD emits them):
                # the restore just rewrote the centres, so produce them once, eagerly, for the first replay
                self.pfc._normalize_weight(force=True)
                self.pfc._wn_fresh = False
            if mom_snap is not None:
                with torch.no_grad():
                    bufs = [st["momentum_buffer"] for st in self.opt_backbone.state.values() if st.get("momentum_buffer") is not None]
                    for b, s in zip(bufs, mom_snap):
                        b.copy_(s)
        torch.cuda.synchronize(self.device)

    def prefetch(self, img, label):
        """Start the host -> device copy of the NEXT step's inputs (pinned host tensors) on a copy stream, so that it runs
        under the current step; the following ``step()`` call (no arguments) consumes them.  What a data loader's
        prefetcher does (ref utils: DataLoaderX / CUDAPrefetcher in datasets/dataloaderx.py)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staged_img = torch.empty_like(self.static_img)
            self._staged_label = torch.empty_like(self.static_label)
            self._staged_free = torch.cuda.Event()
            self._staged_ready = torch.cuda.Event()
            self._staged_free.record(torch.cuda.current_stream(self.device))
        self._copy_stream.wait_event(self._staged_free)          # the previous staged batch has been consumed
        with torch.cuda.stream(self._copy_stream):
            self._staged_img.copy_(img, non_blocking=True)
            self._staged_label.copy_(label, non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._has_staged = True

    def __call__(self, img=None, label=None):
        """img (B,3,H,W) and label (B,) may live on the host (pinned) or the device; with no arguments the batch staged by
        ``prefetch`` is used."""
        self._prepare()
        self._sync_shadows()
        if img is None:
            if not self._has_staged:
                raise RuntimeError("TrainStep(): no inputs given and nothing staged by prefetch()")
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._staged_ready)
            img, label = self._staged_img, self._staged_label
            self._has_staged = False
            staged = True
        else:
            staged = False
        if not self.use_graph:
            loss = self._step(img.to(self.device, non_blocking=True).contiguous(memory_format=torch.channels_last),
                              label.to(self.device, non_blocking=True))
            if staged:
                self._staged_free.record(torch.cuda.current_stream(self.device))
            return loss
        if self.graph is None:
            self.recapture()
        self.static_img.copy_(img, non_blocking=True)
        self.static_label.copy_(label, non_blocking=True)
        if staged:
            self._staged_free.record(torch.cuda.current_stream(self.device))
        self.graph.replay()
        return self.static_loss


@torch.no_grad()
def broadcast_parameters(module, src=0):
    """ref train.py:133-134: every rank starts from rank 0's weights."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
