"""Prefetching data loader — drop-in for ref datasets/dataloaderx.py (`DataLoaderX(local_rank, **DataLoader kwargs)`,
ref train.py:97-105): a background thread drains the worker iterator and the next batch is copied to the GPU on a
side stream while the current step runs (the data path).

Same interface (constructor, iteration protocol, `BackgroundGenerator`), three things done differently because a
B200 step is ~15 ms and the loader must never be what the step waits for:
  * the copy is ordered by CUDA events instead of a whole-stream wait, and every tensor handed to the caller is
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


class BackgroundGenerator(threading.Thread):
    """Iterator over `generator` filled by a daemon thread through a bounded queue (ref :12-38)."""

    def __init__(self, generator, local_rank, max_prefetch=6):
        super().__init__(daemon=True)
        self.queue = queue.Queue(max_prefetch)
        self.generator = generator
        self.local_rank = local_rank
        self.start()

    def run(self):
        try:
            if self.local_rank is not None and torch.cuda.is_available():
                torch.cuda.set_device(self.local_rank)
            for item in self.generator:
                self.queue.put(item)
            self.queue.put(_END)
        except BaseException as e:              # noqa: BLE001 — handed to the consumer, which re-raises it
            self.queue.put(e)

    def next(self):
        item = self.queue.get()
        if item is _END:
            self.queue.put(_END)                # a finished generator keeps raising StopIteration
            raise StopIteration
        if isinstance(item, BaseException):
            self.queue.put(_END)
            raise item
        return item

    __next__ = next

    def __iter__(self):
        return self


class DataLoaderX(DataLoader):
    def __init__(self, local_rank, channels_last=False, max_prefetch=6, **kwargs):
        super().__init__(**kwargs)
        self.local_rank = local_rank
        self.channels_last = channels_last
        self.max_prefetch = max_prefetch
        self.device = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(local_rank)
        self.batch = None
        self._ready = None

    def __iter__(self):
        self.iter = BackgroundGenerator(super().__iter__(), self.local_rank, self.max_prefetch)
        self.preload()
        return self

    def _to_device(self, t):
        if not isinstance(t, torch.Tensor):
            return t
        if not t.is_pinned():
            t = t.pin_memory()
        fmt = torch.channels_last if (self.channels_last and t.dim() == 4 and t.is_floating_point()) else torch.preserve_format
        return t.to(device=self.device, non_blocking=True, memory_format=fmt)

    def preload(self):
        self.batch = next(self.iter, None)
        if self.batch is None:
            return None
        with torch.cuda.stream(self.stream):
            self.batch = [self._to_device(t) for t in self.batch]
            self._ready = torch.cuda.Event()
            self._ready.record(self.stream)

    def __next__(self):
        batch = self.batch
        if batch is None:
            raise StopIteration
        consumer = torch.cuda.current_stream(self.device)
        consumer.wait_event(self._ready)        # only this batch's copy, not everything queued on the side stream
        for t in batch:
            if isinstance(t, torch.Tensor):
                t.record_stream(consumer)       # allocated on the side stream, read on the consumer's
        self.preload()
        return batch
