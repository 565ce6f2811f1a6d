"""Collective plumbing for the class-sharded head.

``NativeComm`` is the product path: an ``ncclComm_t`` owned by libmsml_b200.so (``msml_nccl_init``), so that one step is
two C calls — ``msml_head_gather`` and ``msml_head_step`` — that enqueue kernels and THREE NCCL collectives back to back on
one stream (ref headers/partial_fc.py issues six from Python, :110,126,136,141,162,174).  torch.distributed is only used once,
to broadcast the 128-byte NCCL unique id.  With world_size == 1 there is no communicator and the same two calls run
without collectives.

``TorchDistComm`` keeps the collectives in torch.distributed (four per step; NCCL on the B200 box, gloo in the CPU tests
of the host logic).  ``LockstepComm`` runs W ranks as W host threads of ONE process on one GPU — used by the single-GPU
parity tests to exercise the sharded schedule without ever making one kernel wait for another (only host threads wait).
"""
import ctypes
import threading

import torch
import torch.distributed as dist


class NativeComm:
    """Library-owned NCCL communicator for PartialFC (``handle`` is None for world_size == 1)."""

    native = True

    def __init__(self, world_size, rank, device=None, group=None):
        from .._lib import check, load
        self.world_size, self.rank, self.group = world_size, rank, group
        self.handle = None
        if world_size == 1:
            return
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NativeComm(world_size=%d): torch.distributed must be initialised (it carries the 128-byte "
                               "NCCL unique id from rank 0 to the other ranks)" % world_size)
        lib = load()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            check(lib.msml_nccl_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        on_gpu = dist.get_backend(group) == "nccl"
        t = uid.to(device) if on_gpu else uid
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        out = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.msml_nccl_init(raw, rank, world_size, ctypes.byref(out)))
        self.handle = out

    def close(self):
        if self.handle is not None:
            from .._lib import check, load
            h, self.handle = self.handle, None
            check(load().msml_nccl_destroy(h))

    # generic collectives for callers that want them (not used by PartialFC's native path)
    def all_gather(self, out, inp):
        TorchDistComm.all_gather(self, out, inp)

    def reduce_scatter(self, out, inp):
        TorchDistComm.reduce_scatter(self, out, inp)


class TorchDistComm:
    def __init__(self, world_size, rank, group=None):
        self.world_size, self.rank, self.group = world_size, rank, group

    def all_gather(self, out, inp):
        """out (W*n, ...) <- concat over ranks of inp (n, ...)."""
        if self.world_size == 1:
            out.copy_(inp.reshape(out.shape))
            return
        dist.all_gather_into_tensor(out, inp.contiguous(), group=self.group)

    def reduce_scatter(self, out, inp):
        """out (n, ...) <- this rank's slice of the sum over ranks of inp (W*n, ...)."""
        if self.world_size == 1:
            out.copy_(inp.reshape(out.shape))
            return
        if dist.get_backend(self.group) == "gloo":   # gloo has no reduce_scatter
            full = inp.clone()
            dist.all_reduce(full, group=self.group)
            n = out.shape[0]
            out.copy_(full[self.rank * n:(self.rank + 1) * n])
        else:
            dist.reduce_scatter_tensor(out, inp.contiguous(), group=self.group)


class LockstepComm:
    """W ranks = W threads of one process.  Create one shared instance with ``LockstepComm.create(W)``
    and hand ``.view(rank)`` to each rank's PartialFC."""

    class _Shared:
        def __init__(self, world_size):
            self.world_size = world_size
            self.barrier = threading.Barrier(world_size)
            self.slots = [None] * world_size

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world_size = shared, rank, shared.world_size

    @classmethod
    def create(cls, world_size):
        shared = cls._Shared(world_size)
        return [cls(shared, r) for r in range(world_size)]

    def _exchange(self, t):
        torch.cuda.current_stream().synchronize()
        self.shared.slots[self.rank] = t
        self.shared.barrier.wait()
        got = list(self.shared.slots)
        self.shared.barrier.wait()
        return got

    def all_gather(self, out, inp):
        parts = self._exchange(inp.contiguous())
        out.copy_(torch.cat([p.reshape((-1,) + tuple(out.shape[1:])) for p in parts], 0).reshape(out.shape))
        torch.cuda.current_stream().synchronize()
        self.shared.barrier.wait()

    def reduce_scatter(self, out, inp):
        parts = self._exchange(inp.contiguous())
        n = out.shape[0]
        acc = parts[0][self.rank * n:(self.rank + 1) * n].clone()
        for p in parts[1:]:
            acc += p[self.rank * n:(self.rank + 1) * n]
        out.copy_(acc)
        torch.cuda.current_stream().synchronize()
        self.shared.barrier.wait()
