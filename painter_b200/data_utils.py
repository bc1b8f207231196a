"""Host -> device feeding for the training loop (the reference moves each batch with
`samples.to(device, non_blocking=True)` at the top of the step, engine_train.py:52-56).

`DevicePrefetcher` keeps that contract - every batch is copied from pinned host memory exactly once - but issues
the copy of batch i+1 on a side stream while step i computes, so the ~115 MB/step of images, targets, masks and
valid maps ride under the kernels instead of in front of them.
"""
import torch


class DevicePrefetcher:
    """Iterate over an iterable of tuples of pinned host tensors; yield tuples of device tensors.

    Two persistent sets of device buffers alternate (no allocator traffic in the loop: per-step allocations on a
    side stream cannot be recycled until their `record_stream` events have drained, which showed up as intermittent
    40 ms stalls).  The copy of batch i+1 into set (i+1) % 2 is ordered after everything the consumer stream had
    enqueued when batch i was handed out - i.e. after the last kernel that read that set (step i-1) - and the consumer
    waits on the copy's event before first use.  A batch whose shapes/dtypes differ from the buffers gets fresh ones.
    """

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.k = 0
        self._next = None
        self._preload()

    def _fits(self, buf, host):
        return buf is not None and len(buf) == len(host) and all(
            b.shape == t.shape and b.dtype == t.dtype for b, t in zip(buf, host))

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self.k
        self.k ^= 1
        if not self._fits(self.bufs[k], host):
            self.bufs[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
        # the set was last read by kernels already enqueued on the consumer stream: order the copy after them
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for b, t in zip(self.bufs[k], host):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next = (self.bufs[k], ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._preload()
        return dev
