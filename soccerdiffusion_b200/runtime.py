"""Process-wide run-time switches of the hot path: arithmetic mode and dropout RNG state."""
from __future__ import annotations

import os

import torch

from . import ops

_PRECISIONS = {"fp32": ops.PREC_FP32, "bf16": ops.PREC_BF16}
_precision = _PRECISIONS[os.environ.get("SD_B200_PRECISION", "fp32")]
_seed_base = None
_calls = 0

# torch default of nn.Transformer{En,De}coderLayer, never overridden by the reference
# (encoder/base.py:30-37, decoder.py:26-33)
DROPOUT_P = 0.1


def set_precision(name: str):
    """"fp32": true-fp32 FFMA kernels (<=1e-4 of the reference); "bf16": bf16 tcgen05 GEMMs with fp32
    accumulation, fp32 LayerNorm/softmax/residual stream (<=2e-2)."""
    global _precision
    if name not in _PRECISIONS:
        raise ValueError(f"Invalid precision: {name}")
    _precision = _PRECISIONS[name]


def get_precision() -> int:
    return _precision


def precision_name() -> str:
    return "bf16" if _precision == ops.PREC_BF16 else "fp32"


def manual_seed(seed: int):
    """Seeds the counter-based dropout RNG of the fused kernels."""
    global _seed_base, _calls
    _seed_base = int(seed) & 0xFFFFFFFFFFFF
    _calls = 0


def next_seed() -> int:
    global _seed_base, _calls
    if _seed_base is None:
        _seed_base = torch.initial_seed() & 0xFFFFFFFFFFFF
    _calls += 1
    return (_seed_base * 1000003 + _calls) & 0x7FFFFFFFFFFFFFFF


_fused_trunk = os.environ.get("SD_B200_FUSED_TRUNK", "1") == "1"


def set_fused_trunk(on: bool):
    """bf16 mode: run the trunk's BatchNorm/ReLU/residual/max-pool layers on libsd_b200 kernels (default) or
    through torch's own kernels."""
    global _fused_trunk
    _fused_trunk = bool(on)


def fused_trunk() -> bool:
    return _fused_trunk


def set_dropout(p: float):
    """Overrides the train-mode dropout probability (parity tests run with 0.0; the reference's value is 0.1)."""
    global DROPOUT_P
    if not 0.0 <= p < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), but got {p}")
    DROPOUT_P = float(p)


_fused_layers = os.environ.get("SD_B200_FUSED_LAYERS", "1") == "1"


def set_fused_layers(on: bool):
    """bf16 mode, d_model = 128: run each transformer layer as ONE layer-fused tcgen05 kernel (weights by TMA, residual
    stream in registers) instead of one kernel per GEMM / attention / LayerNorm.  Default on."""
    global _fused_layers
    _fused_layers = bool(on)


def fused_layers() -> bool:
    return _fused_layers


# model.sample(sampler="auto") in bf16 precision mode: batches of at least this many trajectories run the DDIM loop on the
# layer-fused tensor-core kernels (ml/model/model.py:_sample_tc); smaller ones on the fp32 persistent kernels
_tc_sampler_min_batch = int(os.environ.get("SD_B200_TC_SAMPLER_MIN_BATCH", "8"))


def set_tc_sampler_min_batch(n: int):
    global _tc_sampler_min_batch
    _tc_sampler_min_batch = max(1, int(n))


def tc_sampler_min_batch() -> int:
    return _tc_sampler_min_batch


# programmatic dependent launch inside the tensor-core sampler's kernel chain (functional.ddim_sample_tc)
_pdl = os.environ.get("SD_B200_PDL", "1") == "1"


def set_pdl(on: bool):
    global _pdl
    _pdl = bool(on)


def pdl_enabled() -> bool:
    return _pdl


_concurrent_encoders = os.environ.get("SD_B200_CONCURRENT_ENCODERS", "1") == "1"


def set_concurrent_encoders(on: bool):
    """The context encoders are independent until the denoiser: by default the sequence encoders run on side CUDA
    streams beside the image trunk (forward, and — because autograd replays each node on its forward stream —
    backward), joined before the denoiser.  Off = one stream, the reference's call order."""
    global _concurrent_encoders
    _concurrent_encoders = bool(on)


def concurrent_encoders() -> bool:
    return _concurrent_encoders


_side_streams: dict = {}
_SIDE_STREAM_PRIORITY = int(os.environ.get("SD_B200_SIDE_STREAM_PRIORITY", "0"))


def side_streams(device, n: int):
    """Process-wide pool of side CUDA streams per device (created lazily, reused by every model instance)."""
    lst = _side_streams.setdefault((device.type, device.index), [])
    while len(lst) < n:
        # (a higher stream priority for these many-small-kernel streams was A/B-tested next to the trunk's GPU-filling
        # kernels: no measurable difference, so the default priority is kept)
        lst.append(torch.cuda.Stream(device=device, priority=_SIDE_STREAM_PRIORITY))
    return lst[:n]


# ---- weight gradients of the trunk's convolutions on a side stream ---------------------------------------------------------
# In the backward pass the weight gradient of a convolution is a leaf: nothing downstream waits for it until the optimizer
# (or the gradient all-reduce) reads it.  With direct gradient accumulation (below) the trunk launches them on ONE side stream:
# tensor-bound weight-gradient GEMMs then run beside the HBM-bound BatchNorm backward kernels of the main chain.
_wgrad_overlap = os.environ.get("SD_B200_WGRAD_OVERLAP", "1") == "1"
_wgrad_streams: dict = {}
_wgrad_pending: dict = {}


def set_wgrad_overlap(on: bool):
    global _wgrad_overlap
    _wgrad_overlap = bool(on)


def wgrad_overlap() -> bool:
    return _wgrad_overlap and _direct_grads


def wgrad_stream(device):
    key = (device.type, device.index)
    st = _wgrad_streams.get(key)
    if st is None:
        st = _wgrad_streams[key] = torch.cuda.Stream(device=device)
    _wgrad_pending[key] = True
    return st


def join_wgrad_stream(waiter=None, clear: bool = True):
    """``waiter`` (default: the current stream) waits for every weight gradient launched on the side stream so far."""
    for key, pending in list(_wgrad_pending.items()):
        if pending:
            (waiter or torch.cuda.current_stream()).wait_stream(_wgrad_streams[key])
            if clear:
                _wgrad_pending[key] = False


_direct_grads = False


def set_direct_grads(on: bool):
    """Opt-in: backward passes accumulate parameter gradients straight into existing ``p.grad`` buffers (the flat views
    of ``FusedAdamW``) instead of returning them to autograd; see functional._zero_grads."""
    global _direct_grads
    _direct_grads = bool(on)


def direct_grads() -> bool:
    return _direct_grads


_seed_counter = None


def device_seed_counter(device=None):
    """Creates (once) and registers the device-resident dropout seed offset; ``counter.add_(1)`` inside a captured
    training step gives every graph replay fresh dropout masks."""
    global _seed_counter
    if _seed_counter is None:
        _seed_counter = torch.zeros(1, dtype=torch.int64, device=device or "cuda")
        ops.set_dropout_seed_offset(_seed_counter)
    return _seed_counter


def make_cfg(training: bool, dropout_p: float | None = None):
    from .functional import RunCfg

    p = (DROPOUT_P if dropout_p is None else dropout_p) if training else 0.0
    return RunCfg(precision=_precision, p=p, seed=next_seed() if p > 0.0 else 0, stream_base=0)


# ---- content generations (what the plan caches of ml/model/model.py key on) --------------------------------------------
_gen = 0
_weights_gen = 0


def stamp(t):
    """Marks ``t`` as freshly produced by this package (a new process-wide generation number): caches keyed on
    tensor identity treat it as new content even when the allocator reuses its address."""
    global _gen
    _gen += 1
    try:
        t._sd_gen = _gen
    except AttributeError:
        pass
    return t


def bump_weights_generation():
    """Called by everything that rewrites parameters through raw pointers (FusedAdamW steps, graph replays)."""
    global _weights_gen
    _weights_gen += 1


def weights_generation() -> int:
    return _weights_gen


# ---- "this parameter's gradient was written in place" notifications (functional._zero_grads -> FusedAdamW) ---------------
import weakref

_grad_listeners: list = []


def register_grad_listener(bound_method):
    _grad_listeners.append(weakref.WeakMethod(bound_method))


def mark_grad_written(p):
    alive = False
    for ref in _grad_listeners:
        fn = ref()
        if fn is not None:
            alive = True
            fn(p)
    if not alive and _grad_listeners:
        _grad_listeners.clear()


# ---- "the gradients of this part of the model are complete" notifications (backward pass -> overlapped all-reduce) ----------
_grad_ready_cb = None


def set_grad_ready_callback(fn):
    """``fn(tag)`` is called from inside the backward pass when every parameter gradient produced AFTER the tagged point
    of the forward pass (i.e. earlier in the backward pass) is complete.  Tags: "trunk.layer3_onward"."""
    global _grad_ready_cb
    _grad_ready_cb = fn


def grad_ready_enabled() -> bool:
    return _grad_ready_cb is not None


def grad_ready(tag: str):
    if _grad_ready_cb is not None:
        _grad_ready_cb(tag)
