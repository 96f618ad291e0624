"""Host->device input pipeline of the training loop (reference: ml/training/train.py:193,
``batch = {k: v.to(device, non_blocking=True) ...}``).

``DevicePrefetcher`` copies batch i+1 from pinned host memory into one of ``depth`` preallocated device buffer sets
on a dedicated copy stream while the kernels of step i run on the compute stream (the reference's copy is on the
compute stream and serialises with it: 1.5 GB of fp32 images per step at bs=256).  No allocation per batch; a
buffer set is overwritten only after the step that consumed it has finished (event recorded on the compute stream
when the next batch is requested).
"""
from __future__ import annotations

import torch


def _chunked_copy(dst: torch.Tensor, src: torch.Tensor, chunk_bytes: int) -> None:
    """Host->device copy issued as several DMA transfers of at most ``chunk_bytes``: a single 1.5 GB transfer occupies
    its copy engine for ~28 ms, and the memsets / device-to-device copies inside the replayed training-step graph that
    land on the same engine wait behind it; short transfers let them interleave."""
    n = src.numel() * src.element_size()
    if chunk_bytes <= 0 or n <= chunk_bytes or not (src.is_contiguous() and dst.is_contiguous()):
        dst.copy_(src, non_blocking=True)
        return
    d, s_ = dst.view(-1), src.view(-1)
    step = max(1, chunk_bytes // src.element_size())
    for i in range(0, s_.numel(), step):
        d[i:i + step].copy_(s_[i:i + step], non_blocking=True)


class DevicePrefetcher:
    def __init__(self, batches, device, depth: int = 2, chunk_bytes: int = 32 << 20):
        """``batches``: iterable of dicts of (ideally pinned) host tensors, all of the same shapes."""
        self.chunk_bytes = chunk_bytes
        self.it = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.depth = depth
        self.slots = []        # [dict of device tensors, ready_event, free_event]
        self.queue = []        # indices of slots holding a copied batch, in order
        self.next_slot = 0
        self.in_use = None     # slot handed out by the previous __next__
        for _ in range(depth):
            self._enqueue()

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        if len(self.slots) < self.depth:
            dev = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            self.slots.append([dev, torch.cuda.Event(), None])
            i = len(self.slots) - 1
        else:
            i = self.next_slot
        self.next_slot = (i + 1) % self.depth
        dev, ready, free = self.slots[i]
        with torch.cuda.stream(self.copy_stream):
            if free is not None:
                self.copy_stream.wait_event(free)      # the step that read this buffer set has finished
            for k, v in host.items():
                _chunked_copy(dev[k], v, self.chunk_bytes)
            ready.record(self.copy_stream)
        self.queue.append(i)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.in_use is not None:
            ev = torch.cuda.Event()
            ev.record(cur)                              # everything enqueued so far (the previous step) reads in_use
            self.slots[self.in_use][2] = ev
            self.in_use = None
            self._enqueue()
        if not self.queue:
            raise StopIteration
        i = self.queue.pop(0)
        dev, ready, _ = self.slots[i]
        cur.wait_event(ready)
        self.in_use = i
        return dev


# ---- NUMA placement of the feeding process ---------------------------------------------------------------------------------
def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int, sysfs: str = "/sys") -> int | None:
    """NUMA node the GPU's PCIe root is attached to (from sysfs), or None when the platform does not say."""
    import os

    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus", "pci", "devices", bdf, "numa_node")) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


def bind_to_numa_node(node: int | None, sysfs: str = "/sys") -> dict:
    """Pins the calling process (one rank per GPU) to the CPUs of ``node`` and prefers that node for its allocations, so the
    pinned staging buffers it creates afterwards sit next to the GPU's PCIe root (with 8 ranks feeding 8 GPUs from one
    socket's memory, the far socket's ranks copy across the inter-socket link: measured H2D 52 -> 25 GB/s).
    No-op (and says so) when the node is unknown or the platform refuses."""
    import ctypes
    import os

    info = dict(node=node, cpus=None, mempolicy=False)
    if node is None:
        return info
    try:
        with open(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")) as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except (OSError, ValueError):
        return info
    try:   # set_mempolicy(MPOL_PREFERRED, {node}) — x86_64 syscall 238; best effort
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        info["mempolicy"] = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(16 * 64)) == 0
    except (OSError, AttributeError, ValueError):
        pass
    return info
