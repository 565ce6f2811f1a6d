"""PartialFC — drop-in for ref headers/partial_fc.py (class-sharded FC, sampled negatives,
distributed softmax cross-entropy with the reference's local label smoothing).

Same constructor, attributes and methods as the reference (``forward_backward(label, features,
optimizer) -> (x_grad, loss_v)``, ``update()``, ``save_params()``, ``sample()``, ``prepare()``),
but the work is done by libmsml_b200.so:

  sample           msml_pfc_remap / mark_positive / select (radix select + ordered compaction) /
                   searchsorted / gather_rows              (ref :77-94; torch.rand stays in torch so
                   the generator is consumed exactly as in the reference)
  prepare          msml_wnorm_cast: fp32 master rows -> unit-norm bf16 + 1/||w||            (ref :115)
  forward_backward msml_head_fwd (tcgen05 GEMM, margin/scale/online-softmax epilogue; the fp32 logits are
                   never materialised — the backward keeps one bf16 (B_tot x n_s) dcos matrix in the
                   workspace) -> ONE all-gather of per-row (max, sum, target) instead of
                   the reference's three all-reduces (:136,141,162) -> msml_head_merge_stats ->
                   msml_head_bwd (recompute + dX + dW with the normalise-backward epilogue)
                   -> reduce-scatter of dX (:172-175)

Class shards exchange data through an ncclComm_t owned by the library (headers/_comm.NativeComm: three collectives
per step, enqueued from C between the kernels); with world_size == 1 no process group is needed.  ``margin_softmax`` must carry (kind, s, m, a, k): a
``msml_b200.headers.MarginSoftmax`` (ArcFace()/CosFace()) or an AMArcFace/AMCosFace module.
"""
import contextlib
import ctypes
import logging
import os

import torch
from torch.nn import Module
from torch.nn.parameter import Parameter

from .. import _lib
from .._lib import check, load, stream_ptr
from ._comm import NativeComm, TorchDistComm

EPSILON = 0.1  # label smoothing baked into the gradient (ref :154)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class PartialFC(Module):
    @torch.no_grad()
    def __init__(self, rank, local_rank, world_size, batch_size, resume,
                 margin_softmax, num_classes, sample_rate=1.0, embedding_size=512, prefix="./", comm=None):
        super().__init__()
        self.num_classes = num_classes
        self.rank = rank
        self.local_rank = local_rank
        self.device = torch.device("cuda:{}".format(local_rank))
        self.world_size = world_size
        self.batch_size = batch_size
        self.margin_softmax = margin_softmax
        self.sample_rate = sample_rate
        self.embedding_size = embedding_size
        self.prefix = prefix
        # shard geometry, ref :34-36
        self.num_local = num_classes // world_size + int(rank < num_classes % world_size)
        self.class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
        self.num_sample = int(self.sample_rate * self.num_local)

        for attr in ("kind", "s", "m", "a", "k"):
            if not hasattr(margin_softmax, attr):
                raise TypeError("margin_softmax must expose (kind, s, m, a, k) — use msml_b200.headers.ArcFace()/"
                                "CosFace()/MarginSoftmax or an AMArcFace/AMCosFace module; the margin is fused "
                                "into the tcgen05 epilogue and there is no unfused fallback")
        self._margin = _lib.margin_params(margin_softmax.kind, margin_softmax.s, margin_softmax.m,
                                          margin_softmax.a, margin_softmax.k)
        load()  # fail loudly right here if the CUDA library is missing
        # default: the library-owned communicator (three collectives per step enqueued from C); pass a TorchDistComm /
        # LockstepComm to keep the collectives in Python
        with torch.cuda.device(self.device):
            self.comm = comm if comm is not None else NativeComm(world_size, rank, device=self.device)

        self.weight_name = os.path.join(self.prefix, "rank:{}_softmax_weight.pt".format(self.rank))
        self.weight_mom_name = os.path.join(self.prefix, "rank:{}_softmax_weight_mom.pt".format(self.rank))

        def fresh():
            return torch.normal(0, 0.01, (self.num_local, self.embedding_size), device=self.device)

        if resume:
            try:
                self.weight = torch.load(self.weight_name).to(self.device)
                logging.info("softmax weight resume successfully!")
            except (FileNotFoundError, KeyError, IndexError):
                self.weight = fresh()
                logging.info("softmax weight resume fail!")
            try:
                self.weight_mom = torch.load(self.weight_mom_name).to(self.device)
                logging.info("softmax weight mom resume successfully!")
            except (FileNotFoundError, KeyError, IndexError):
                self.weight_mom = torch.zeros_like(self.weight)
                logging.info("softmax weight mom resume fail!")
        else:
            self.weight = fresh()
            self.weight_mom = torch.zeros_like(self.weight)
            logging.info("softmax weight init successfully!")
            logging.info("softmax weight mom init successfully!")
        self.stream = torch.cuda.Stream(local_rank)

        self.index = None
        if int(self.sample_rate) == 1:
            self.update = lambda: 0
            self.sub_weight = Parameter(self.weight)
            self.sub_weight_mom = self.weight_mom
        else:
            self.sub_weight = Parameter(torch.empty((0, 0), device=self.device))

        self.last_grad = None
        self.last_loss = None
        self._ws = {}

    # ------------------------------------------------------------------ helpers
    def _buf(self, name, shape, dtype):
        """Persistent scratch (torch-allocated so the caching allocator sees it): one buffer per name, grown to the
        largest size seen (n_s changes from step to step when the positives outnumber num_sample, ref :89-90) and
        handed out as a view of the requested shape."""
        numel = 1
        for d in shape:
            numel *= int(d)
        t = self._ws.get(name)
        if t is None or t.dtype != dtype or t.numel() < numel:
            t = torch.empty((max(numel, 1),), dtype=dtype, device=self.device)
            self._ws[name] = t
        return t[:numel].view(tuple(shape))

    def save_params(self):
        torch.save(self.weight.data, self.weight_name)
        torch.save(self.weight_mom, self.weight_mom_name)

    # ------------------------------------------------------------------ sampling, ref :77-94
    def _select(self, perm, num_sample, capacity):
        lib = load()
        ws_bytes = lib.msml_pfc_select_workspace(self.num_local)
        ws = self._buf("select_ws", (ws_bytes,), torch.uint8)
        index = self._buf("index", (capacity,), torch.int64)
        n_index = self._buf("n_index", (1,), torch.int64)
        check(lib.msml_pfc_select(_ptr(perm), self.num_local, num_sample, _ptr(index), _ptr(n_index),
                                  _ptr(ws), ws_bytes, stream_ptr()))
        return index, n_index

    @torch.no_grad()
    def sample(self, total_label, remapped=False):
        """ref :77-94 (remapped=True: total_label already holds shard-local labels, msml_head_gather did :79-81).  All per-step tensors (index, gathered rows, their momentum) are views of persistent scratch: a
        sampled step allocates nothing once warm (the caching allocator otherwise recycles ~0.6 GB per step at 1M classes
        through cudaMalloc / cudaFree stalls of tens of ms), and the step can be captured into a CUDA graph."""
        lib = load()
        n = total_label.numel()
        if not remapped:
            check(lib.msml_pfc_remap(_ptr(total_label), n, self.class_start, self.num_local, stream_ptr()))
        if int(self.sample_rate) != 1:
            index = None
            if n > self.num_sample:
                # the positives may outnumber num_sample (ref :89-90): count them first, without
                # touching the generator, exactly as the reference does
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("PartialFC: a captured step needs batch_size * world_size <= num_sample (%d > %d): the "
                                       "positives-outnumber-the-sample branch (ref :89-90) has a data-dependent size" % (n, self.num_sample))
                probe = self._buf("probe", (self.num_local,), torch.float32).zero_()
                check(lib.msml_pfc_mark_positive(_ptr(probe), _ptr(total_label), n, self.num_local, stream_ptr()))
                pos_index, n_pos = self._select(probe, 0, max(n, self.num_sample, 1))
                n_pos_host = int(n_pos.item())
                if n_pos_host > self.num_sample:
                    index, n_index = pos_index[:n_pos_host], n_pos
            if index is None:
                perm = torch.rand(size=[self.num_local], device=self.device)
                check(lib.msml_pfc_mark_positive(_ptr(perm), _ptr(total_label), n, self.num_local, stream_ptr()))
                index, n_index = self._select(perm, self.num_sample, max(n, self.num_sample, 1))
                index = index[:self.num_sample]
            self.index = index
            check(lib.msml_pfc_searchsorted(_ptr(total_label), n, _ptr(index), _ptr(n_index), stream_ptr()))
            rows = index.numel()
            sub_w = self._buf("sub_w", (rows, self.embedding_size), torch.float32)
            sub_m = self._buf("sub_m", (rows, self.embedding_size), torch.float32)
            check(lib.msml_gather_rows_f32(_ptr(self.weight), _ptr(index), _ptr(sub_w), rows, self.embedding_size, stream_ptr()))
            check(lib.msml_gather_rows_f32(_ptr(self.weight_mom), _ptr(index), _ptr(sub_m), rows, self.embedding_size, stream_ptr()))
            self.sub_weight = Parameter(sub_w)
            self.sub_weight_mom = sub_m

    def forward(self, total_features, norm_weight):
        """logits = total_features @ norm_weight^T on the tcgen05 GEMM (ref :96-99).  Kept for API
        parity; forward_backward never materialises the logits."""
        from .. import ops
        torch.cuda.current_stream().wait_stream(self.stream)
        return ops.gemm_tn(total_features, norm_weight)

    @torch.no_grad()
    def update(self):
        """Scatter the sampled rows back into the shard (ref :101-104)."""
        if getattr(self, "_fused_step_done", False):
            # headers.PartialFCSGD.step() just updated the shard rows in place: the gathered copies are stale, not newer.
            # The flag covers that one step only — a stock optimizer stepping this module later scatters as usual.
            self._fused_step_done = False
            return
        lib = load()
        rows = self.index.numel()
        check(lib.msml_scatter_rows_f32(_ptr(self.weight_mom), _ptr(self.index), _ptr(self.sub_weight_mom), rows,
                                        self.embedding_size, stream_ptr()))
        check(lib.msml_scatter_rows_f32(_ptr(self.weight), _ptr(self.index), _ptr(self.sub_weight.data), rows,
                                        self.embedding_size, stream_ptr()))

    def invalidate_normalized(self):
        """The class centres were changed from outside (checkpoint load, manual edit): recompute the normalised bf16 copy
        at the next step even if a PartialFCSGD(emit_normalized=True) had already produced it."""
        self._wn_fresh = False

    def _normalize_weight(self, force=False):
        """ref :115 -> (wn bf16 (n_s, D), inv_norm fp32 (n_s))."""
        lib = load()
        n_s, D = self.sub_weight.shape
        wn = self._buf("wn", (n_s, D), torch.bfloat16)
        inv = self._buf("inv_norm", (n_s,), torch.float32)
        if getattr(self, "_wn_fresh", False) and not force and int(self.sample_rate) == 1:
            self._wn_fresh = False      # produced by the previous step's PartialFCSGD update (one-shot: see emit_normalized)
            return wn, inv
        check(lib.msml_wnorm_cast(_ptr(self.sub_weight.data), _ptr(wn), None, 0, _ptr(inv), n_s, D, stream_ptr()))
        return wn, inv

    def prepare(self, label, optimizer):
        # side stream as in the reference (:107) — except under CUDA-graph capture, where the whole
        # step is one stream-ordered graph anyway
        capturing = torch.cuda.is_current_stream_capturing()
        ctx = contextlib.nullcontext() if capturing else torch.cuda.stream(self.stream)
        with ctx:
            total_label = torch.zeros(size=[self.batch_size * self.world_size], device=self.device, dtype=torch.long)
            self.comm.all_gather(total_label, label)
            self.sample(total_label)
            if optimizer is not None:
                # optimizer surgery, ref :112-114
                optimizer.state.pop(optimizer.param_groups[-1]['params'][0], None)
                optimizer.param_groups[-1]['params'][0] = self.sub_weight
                optimizer.state[self.sub_weight]['momentum_buffer'] = self.sub_weight_mom
            norm_weight = self._normalize_weight()
            if not capturing:
                total_label.record_stream(torch.cuda.current_stream(self.device))
            return total_label, norm_weight

    def _forward_backward_native(self, label, features, optimizer):
        """The product path: two C calls, everything on the current stream (a valid subsumption of the reference's side
        stream, :107 / :97): msml_head_gather [pack -> ncclAllGather -> unpack + remap], sampling, optimizer surgery and
        msml_wnorm_cast as in prepare(), then msml_head_step [fwd GEMM -> ncclAllGather(row stats) -> merge + loss ->
        backward GEMMs -> ncclReduceScatter(dX) -> x world_size]."""
        lib = load()
        W, B, D = self.world_size, self.batch_size, self.embedding_size
        B_tot = B * W
        h = self.comm.handle
        with torch.no_grad():
            feat = features.detach().to(torch.float32).contiguous()
            lab = label.detach().to(torch.int64).contiguous()
            if feat.shape != (B, D) or lab.shape != (B,):
                raise ValueError("PartialFC: expected features %s and label %s, got %s and %s" % ((B, D), (B,), tuple(feat.shape), tuple(lab.shape)))
            x = self._buf("x", (B_tot, D), torch.bfloat16)
            total_label = self._buf("total_label", (B_tot,), torch.int64)
            gbytes = lib.msml_head_gather_workspace(B, W, D)
            gws = self._buf("gather_ws", (gbytes,), torch.uint8)
            check(lib.msml_head_gather(h, _ptr(feat), _ptr(lab), B, D, self.class_start, self.num_local, _ptr(x), _ptr(total_label),
                                       _ptr(gws), gbytes, stream_ptr()))
            self.sample(total_label, remapped=True)
            if optimizer is not None:
                # optimizer surgery, ref :112-114
                optimizer.state.pop(optimizer.param_groups[-1]['params'][0], None)
                optimizer.param_groups[-1]['params'][0] = self.sub_weight
                optimizer.state[self.sub_weight]['momentum_buffer'] = self.sub_weight_mom
            wn, inv_norm = self._normalize_weight()
            n_s = wn.shape[0]
            sbytes = lib.msml_head_step_workspace(B, W, n_s, D)
            sws = self._buf("head_ws", (sbytes,), torch.uint8)
            x_grad = torch.empty((B, D), dtype=torch.float32, device=self.device)
            loss_v = torch.empty((), dtype=torch.float32, device=self.device)
            dw = (torch.empty((n_s, D), dtype=torch.float32, device=self.device) if int(self.sample_rate) == 1
                  else self._buf("dw", (n_s, D), torch.float32))
            # raw mode: the optimizer handed in is a PartialFCSGD(fuse_projection=True) driving THIS module — it applies the
            # normalise backward itself, so the GEMM epilogues skip it and dw is dWn = dcos^T X
            raw = bool(getattr(optimizer, "fuse_projection", False)) and getattr(optimizer, "module", None) is self
            if raw:
                check(lib.msml_head_step_raw(h, _ptr(x), _ptr(wn), _ptr(total_label), B, n_s, D, ctypes.byref(self._margin),
                                             _ptr(x_grad), _ptr(dw), _ptr(loss_v), _ptr(sws), sbytes, stream_ptr()))
            else:
                check(lib.msml_head_step(h, _ptr(x), _ptr(wn), _ptr(inv_norm), _ptr(total_label), B, n_s, D, ctypes.byref(self._margin),
                                         _ptr(x_grad), _ptr(dw), _ptr(loss_v), _ptr(sws), sbytes, stream_ptr()))
            self._grad_is_raw = raw
            self.sub_weight.grad = dw
        self.last_loss = loss_v
        return x_grad, loss_v

    def forward_backward(self, label, features, optimizer):
        """features (B, D) are assumed L2-normalised by the caller (ref :119)."""
        lib = load()
        _lib.require_cuda(label, features)
        if getattr(self.comm, "native", False):
            return self._forward_backward_native(label, features, optimizer)
        W, B, D = self.world_size, self.batch_size, self.embedding_size
        B_tot = B * W
        main = torch.cuda.current_stream(self.device)
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.stream.wait_stream(main)        # label / weights produced on the main stream
        total_label, (wn, inv_norm) = self.prepare(label, optimizer)

        with torch.no_grad():
            # all-gather the embeddings in bf16 (half the bytes of the reference's fp32 gather, :126)
            feat = features.detach().to(torch.float32).contiguous()
            x_local = self._buf("x_local", (B, D), torch.bfloat16)
            check(lib.msml_cast_bf16(_ptr(feat), _ptr(x_local), None, 0, B, D, stream_ptr()))
            x = self._buf("x", (B_tot, D), torch.bfloat16)
            self.comm.all_gather(x, x_local)
            if not capturing:
                main.wait_stream(self.stream)    # ref :97

            n_s = wn.shape[0]
            ws_bytes = lib.msml_head_workspace(B_tot, n_s, D)
            ws = self._buf("head_ws", (ws_bytes,), torch.uint8)
            stats = self._buf("stats", (3, B_tot), torch.float32)
            mp = ctypes.byref(self._margin)
            check(lib.msml_head_fwd(_ptr(x), _ptr(wn), _ptr(total_label), B_tot, n_s, D, mp, _ptr(stats),
                                    _ptr(ws), ws_bytes, stream_ptr()))
            if W > 1:
                gathered = self._buf("gathered", (W * 3, B_tot), torch.float32)
                self.comm.all_gather(gathered, stats)
            else:
                gathered = stats
            gstats = self._buf("gstats", (2, B_tot), torch.float32)
            loss_v = torch.empty((), dtype=torch.float32, device=self.device)
            check(lib.msml_head_merge_stats(_ptr(gathered), W, B_tot, _ptr(gstats), _ptr(loss_v), stream_ptr()))

            dx_full = self._buf("dx_full", (B_tot, D), torch.float32)
            # sample_rate 1: a fresh tensor per step as autograd would produce; sampled: persistent scratch (see sample())
            dw = (torch.empty((n_s, D), dtype=torch.float32, device=self.device) if int(self.sample_rate) == 1
                  else self._buf("dw", (n_s, D), torch.float32))
            check(lib.msml_head_bwd(_ptr(x), _ptr(wn), _ptr(inv_norm), _ptr(total_label), B_tot, n_s, D, mp,
                                    _ptr(gstats), _ptr(dx_full), _ptr(dw), _ptr(ws), ws_bytes, stream_ptr()))
            self._grad_is_raw = False
            self.sub_weight.grad = dw

            # feature gradient reduce-scatter, then * world_size (ref :172-175)
            x_grad = torch.empty((B, D), dtype=torch.float32, device=self.device)
            self.comm.reduce_scatter(x_grad, dx_full)
            if W > 1:
                x_grad = x_grad * W
        self.last_loss = loss_v
        return x_grad, loss_v
