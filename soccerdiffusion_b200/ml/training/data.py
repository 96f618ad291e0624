"""Host->device input pipeline of the training loop (reference: ml/training/train.py:193,
``batch = {k: v.to(device, non_blocking=True) ...}``).

``DevicePrefetcher`` issues the pinned-host -> HBM copies of batch i+1 on a dedicated copy stream while the kernels of
step i run on the compute stream (the reference's copy is on the compute stream and serialises with it: 1.5 GB of fp32
images per step at bs=256).  Double-buffered; the compute stream waits on the copy's event before using a batch, and a
buffer is recycled only after the step that consumed it has been enqueued.
"""
from __future__ import annotations

import torch


class DevicePrefetcher:
    def __init__(self, batches, device, depth: int = 2):
        """``batches``: iterable of dicts of (ideally pinned) host tensors."""
        self.it = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.depth = depth
        self.queue = []
        for _ in range(depth):
            self._enqueue()

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        with torch.cuda.stream(self.copy_stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((dev, ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        dev, ev = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev.values():
            t.record_stream(cur)   # the caching allocator must not recycle these while the step uses them
        self._enqueue()
        return dev
