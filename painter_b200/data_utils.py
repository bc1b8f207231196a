"""Host -> device feeding for the training loop (the reference moves each batch with
`samples.to(device, non_blocking=True)` at the top of the step, engine_train.py:52-56).

`DevicePrefetcher` keeps that contract - every batch is copied from pinned host memory exactly once - but issues
the copy of batch i+1 on a side stream while step i computes, so the ~115 MB/step of images, targets, masks and
valid maps ride under the kernels instead of in front of them.
"""
import torch


class DevicePrefetcher:
    """Iterate over an iterable of tuples of pinned host tensors; yield tuples of device tensors.

    The copy of the next batch is enqueued on `copy_stream` right after the current batch is handed out; the
    consumer's stream waits on the copy's event before first use, and the tensors are `record_stream`-ed so the
    caching allocator does not recycle them early."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.copy_stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)
        self._preload()
        return dev
