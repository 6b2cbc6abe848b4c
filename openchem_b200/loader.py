"""Host -> device feeding for the GraphCNN path: uploads the NEXT batch from pinned host memory on a side stream while
the current step computes (the reference's `cast_inputs` does a synchronous `.to('cuda')` of 4N^2 + 4N*F0 bytes per
molecule every iteration, openchem/models/Graph2Label.py:41-57). Input format is the reference's: dense fp32
`adj_matrix (B,N,N)`, `node_feature_matrix (B,N,F0)`, `labels (B,T)` — or the tensors of a bit-packed
`openchem_b200.compact.CompactGraphBatch` (int32 words; dtypes are preserved), which cuts the upload 32x."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Iterate over (x, adj, labels) host batches; yields device tensors whose upload overlapped the previous step.

    Two device buffer sets are used alternately; a yielded batch stays valid until the next-but-one `next()`.
    """

    def __init__(self, host_batches, device):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.bytes_per_batch = 0
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self.k
        cur = torch.cuda.current_stream(self.device)
        # the slot being overwritten was consumed two steps ago on the compute stream
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self.slots[k] is None or any(s.shape != h.shape or s.dtype != h.dtype for s, h in zip(self.slots[k], host)):
                self.slots[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            for dst, src in zip(self.slots[k], host):
                dst.copy_(src, non_blocking=True)
            self.events[k].record(self.stream)
        self.bytes_per_batch = sum(h.numel() * h.element_size() for h in host)
        self._next = k
        self.k ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        k = self._next
        torch.cuda.current_stream(self.device).wait_event(self.events[k])
        batch = self.slots[k]
        self._preload()
        return batch


class BackgroundIterator:
    """Runs an iterator (e.g. sampler + `CompactGraphDataset.collate`, or a DataLoader) in a daemon thread, `depth` items ahead,
    so the host-side batch assembly of step t+1.. overlaps the Python / launch work of step t instead of sitting between two
    launches. Exceptions raised by the producer are re-raised at the consumer."""

    _END = object()

    def __init__(self, iterable, depth=2):
        import queue
        import threading
        self.q = queue.Queue(maxsize=max(int(depth), 1))
        self.exc = None

        def run():
            try:
                for item in iterable:
                    self.q.put(item)
            except BaseException as e:  # noqa: BLE001 - handed to the consumer
                self.exc = e
            self.q.put(self._END)

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()

    def __iter__(self):
        return self

    def __next__(self):
        item = self.q.get()
        if item is self._END:
            if self.exc is not None:
                raise self.exc
            raise StopIteration
        return item
