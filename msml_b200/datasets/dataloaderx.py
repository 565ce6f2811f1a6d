"""Prefetching data loader — drop-in for ref datasets/dataloaderx.py (`DataLoaderX(local_rank, **DataLoader kwargs)`,
ref train.py:97-105): a background thread drains the worker iterator and the next batch is copied to the GPU on a
side stream while the current step runs (the data path).

Same interface (constructor, iteration protocol, `BackgroundGenerator`), three things done differently because a
B200 step is ~15 ms and the loader must never be what the step waits for:
  * a small device-side queue (`device_prefetch` batches, default 2) instead of a single staged batch; each staged batch
    carries its own CUDA event, so the consumer waits for that batch's copies only, not for the whole side stream, and every tensor handed to the caller is
    `record_stream`-ed on the consumer stream: the reference allocates on the side stream and consumes on the main
    one without telling the caching allocator, so a batch can be recycled while a kernel still reads it;
  * batches are pinned before the copy when the DataLoader did not pin them (an unpinned `non_blocking` copy is
    synchronous and would serialise with the step), and 4-D float tensors can be delivered channels-last, the layout
    the whole backbone runs in (`channels_last=True`);
  * an exception in the worker thread is re-raised in the consumer instead of leaving it blocked on an empty queue.
`msml_b200.engine.TrainStep.prefetch` is the captured-graph counterpart (it copies into the graph's static inputs).
"""
import queue
import threading

import torch
from torch.utils.data import DataLoader

__all__ = ["BackgroundGenerator", "DataLoaderX"]

_END = object()


class BackgroundGenerator:
    """Iterator over `generator`, drained ahead of the consumer by a daemon thread through a bounded queue (ref :12-38;
    same constructor and iterator protocol, the thread is owned rather than inherited)."""

    def __init__(self, generator, local_rank, max_prefetch=6):
        self._source = generator
        self._device = local_rank
        self._items = queue.Queue(max_prefetch)
        self._done = False
        self._worker = threading.Thread(target=self._drain, name="msml-loader", daemon=True)
        self._worker.start()

    def _drain(self):
        try:
            if self._device is not None and torch.cuda.is_available():
                torch.cuda.set_device(self._device)         # worker-side CUDA calls (pinning) bind to this rank's GPU
            for element in self._source:
                self._items.put(element)
        except BaseException as exc:                        # noqa: BLE001 — re-raised in the consumer
            self._items.put(exc)
        finally:
            self._items.put(_END)

    def __iter__(self):
        return self

    def __next__(self):
        if self._done:
            raise StopIteration
        element = self._items.get()
        if element is _END:
            self._done = True
            raise StopIteration
        if isinstance(element, BaseException):
            raise element
        return element

    next = __next__


class DataLoaderX(DataLoader):
    """`DataLoaderX(local_rank, **DataLoader kwargs)`; iterating yields lists of device tensors.  `device_prefetch`
    batches (default 2) are kept in flight on the side stream, each with its own completion event."""

    def __init__(self, local_rank, channels_last=False, max_prefetch=6, device_prefetch=2, **kwargs):
        super().__init__(**kwargs)
        self.local_rank = local_rank
        self.channels_last = channels_last
        self.max_prefetch = max_prefetch
        self.device_prefetch = max(1, int(device_prefetch))
        self.device = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(local_rank)
        self._host_iter = None
        self._staged = []                   # [(list of device tensors, event recorded after their copies)]

    def _upload(self, t):
        if not isinstance(t, torch.Tensor):
            return t
        if not t.is_pinned():
            t = t.pin_memory()
        fmt = torch.channels_last if (self.channels_last and t.dim() == 4 and t.is_floating_point()) else torch.preserve_format
        return t.to(device=self.device, non_blocking=True, memory_format=fmt)

    def _fill(self):
        while self._host_iter is not None and len(self._staged) < self.device_prefetch:
            host_batch = next(self._host_iter, None)
            if host_batch is None:
                self._host_iter = None
                break
            with torch.cuda.stream(self.stream):
                on_device = [self._upload(t) for t in host_batch]
                copied = torch.cuda.Event()
                copied.record(self.stream)
            self._staged.append((on_device, copied))

    def preload(self):
        """Kept for callers of the reference's method name: top the device-side queue up."""
        self._fill()

    def __iter__(self):
        self._host_iter = BackgroundGenerator(super().__iter__(), self.local_rank, self.max_prefetch)
        self._staged = []
        self._fill()
        return self

    def __next__(self):
        if not self._staged:
            raise StopIteration
        on_device, copied = self._staged.pop(0)
        consumer = torch.cuda.current_stream(self.device)
        consumer.wait_event(copied)         # only this batch's copies, not everything queued on the side stream
        for t in on_device:
            if isinstance(t, torch.Tensor):
                t.record_stream(consumer)   # allocated on the side stream, read on the consumer's
        self._fill()
        return on_device
