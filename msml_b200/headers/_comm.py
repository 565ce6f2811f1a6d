"""Collective plumbing for the class-sharded head.

``TorchDistComm`` is the product path: torch.distributed (NCCL over NVLink / NVSwitch on the B200
box; gloo in the CPU tests of the host logic).  ``LockstepComm`` runs W ranks as W host threads
of ONE process on one GPU — used by the single-GPU parity tests to exercise the sharded
schedule without ever making one kernel wait for another (only host threads wait).
"""
import threading

import torch
import torch.distributed as dist


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
