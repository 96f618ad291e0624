#!/usr/bin/env python
"""bench.py — the hot path's headline metric (BASELINE.json): train samples/s, plus p50 DDIM latency at bs=1.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]
                    [--batch B] [--workload full|inscope|denoiser]

A "step" is one pass of the reference's training loop body (ml/training/train.py:193-240) over one synthetic
batch of default.yaml shapes: normalise -> t, eps -> add_noise -> model fwd -> mse -> bwd -> AdamW -> OneCycleLR,
dropout p=0.1 live as in the reference.  N>1: one rank per GPU (torchrun), each rank its own bs=256 slice
(weak scaling), one NCCL all-reduce of the flat gradient per step.

Prints ONE JSON line (rank 0).  `value`: samples/s with inputs resident in HBM; `e2e`: the same step fed from
pinned HOST buffers (H2D inside the timed region) with the loss read back; `roofline`: the dominant kernel
class of libsd_b200 timed with CUDA events; `cpu_baseline`: the oracle port of the reference's CPU path on a
bounded sample; `ddim`: bs=1 30-step trajectory latency (persistent-kernel sampler).

--impl reference: the reference's own CPU PyTorch path (restated in oracle/, the live reference cannot travel to
the GPU box) timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/s (default.yaml denoiser training step, bs=256/GPU)"
UNIT = "samples/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="sd_clocks_", suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's training-loop body in the oracle restatement (torch CPU fp32, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import model_ref, synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = synth.DEFAULT_HP
    bs = args.cpu_batch
    tmpl = _state_template(hp)
    sd = synth.synth_state_dict(tmpl, 0)
    names = [n for n in tmpl if tmpl[n].is_floating_point() and "running_" not in n and n not in ("mean", "std")]
    sd = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    opt = torch.optim.AdamW([sd[n] for n in names], lr=hp["lr"])
    batch = synth.synth_batch(hp, bs, 0)
    times = []
    for it in range(args.warmup + args.steps):
        noise = torch.randn(bs, hp["trajectory_prediction_length"], hp["num_joints"])
        t = torch.randint(0, 1000, (bs,))
        t0 = time.perf_counter()
        opt.zero_grad()
        # dropout p=0.1 is live in the reference's loop; the oracle applies supplied masks only, so the CPU arm
        # runs without dropout (slightly LESS work than the reference: the baseline is not under-reported)
        loss, _ = model_ref.training_loss(batch, noise, t, sd, hp, masks=None, train_bn=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = bs * len(times) / total
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"default.yaml full training step on host CPU, bounded sample bs={bs}",
                            global_batch=bs),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{len(times)} full training steps at bs={bs} (oracle/model_ref.py + torch AdamW)"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def _state_template(hp):
    """name -> empty tensor with the shape the REAL reference's state_dict has for default.yaml (recorded from the
    live reference by oracle/gen_golden.py into tests/golden/manifest.json) — no product code on the CPU arm."""
    import torch

    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as fh:
        c = json.load(fh)["cases"]["default"]
    out = {}
    for n, s in zip(c["state_dict_names"], c["state_dict_shapes"]):
        out[n] = torch.empty(s, dtype=torch.int64 if n.endswith("num_batches_tracked") else torch.float32)
    return out


def cpu_baseline_sample(hp, bs, max_seconds=25.0):
    import torch

    from oracle import model_ref, synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tmpl = _state_template(hp)
    sd = synth.synth_state_dict(tmpl, 0)
    names = [n for n in tmpl if tmpl[n].is_floating_point() and "running_" not in n and n not in ("mean", "std")]
    sd = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    opt = torch.optim.AdamW([sd[n] for n in names], lr=1e-4)
    batch = synth.synth_batch(hp, bs, 0)
    noise = synth.synth_noise("eps", hp, bs, 0)
    t = synth.synth_timesteps(bs, 0)
    n, t_total = 0, 0.0
    for it in range(4):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _ = model_ref.training_loss(batch, noise, t, sd, hp, masks=None, train_bn=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= 1:
            n += 1
            t_total += dt
        if t_total > max_seconds:
            break
    return dict(value=bs * n / t_total, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} full training steps at bs={bs} after 1 warm-up (oracle/model_ref.py, torch CPU fp32 + AdamW)")



# ----------------------------------------------------------------------------------------------------
RESNET18_CONVS = (   # (cin, cout, kernel, stride, input H=W) of layer1..layer4 at 224x224 (conv1 is libsd_b200's own stem kernel)
    [(64, 64, 3, 1, 56)] * 4
    + [(64, 128, 3, 2, 56), (64, 128, 1, 2, 56)] + [(128, 128, 3, 1, 28)] * 3
    + [(128, 256, 3, 2, 28), (128, 256, 1, 2, 28)] + [(256, 256, 3, 1, 14)] * 3
    + [(256, 512, 3, 2, 14), (256, 512, 1, 2, 14)] + [(512, 512, 3, 1, 7)] * 3)


def conv_library_probe(dev, frames: int, resolution: int):
    """Device time of the trunk's LIBRARY convolutions alone (cuDNN: layer1-4 of ResNet18, bf16 channels_last exactly as
    encoder/trunk.py issues them).  Only the passes that ARE library calls are timed: forward, data gradient and weight gradient of
    every convolution EXCEPT the weight gradient of the four 3x3 64->64 convolutions of layer1 and the data gradient of the three
    1x1 stride-2 downsample convolutions, which run on libsd_b200's own kernels (sd_conv3x3_wgrad_c64_bf16,
    sd_conv1x1s2_dgrad_bf16).  Every distinct shape is run back to back between two CUDA events after a warm-up; each call is
    0.05-1 ms of device work, so the queue never runs dry and the events see device time only — no profiler.
    -> (ms per training step, algorithmic flops per step, number of library launches per step)."""
    import collections

    import torch

    scale = resolution / 224.0
    total_ms, total_flops, launches = 0.0, 0.0, 0
    reps = 3
    for (cin, cout, k, st, hin), mult in collections.Counter(RESNET18_CONVS).items():
        hin = int(round(hin * scale))
        cl = torch.channels_last
        x = torch.randn(frames, cin, hin, hin, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        w = torch.randn(cout, cin, k, k, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
        args = ([st, st], [k // 2, k // 2], [1, 1], False, [0, 0], 1)
        y = torch.ops.aten.convolution(x, w, None, *args)
        gy = torch.randn_like(y)
        own_wgrad = (cin, cout, k, st) == (64, 64, 3, 1)
        own_dgrad = k == 1 and st == 2
        passes = [lambda: torch.ops.aten.convolution(x, w, None, *args)]
        if not own_dgrad:
            passes.append(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, *args, [True, False, False]))
        if not own_wgrad:
            passes.append(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, *args, [False, True, False]))
        for _ in range(2):
            for f in passes:
                f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for f in passes:
                f()
        e1.record()
        torch.cuda.synchronize()
        total_ms += mult * e0.elapsed_time(e1) / reps
        total_flops += mult * len(passes) * 2.0 * frames * y.shape[2] * y.shape[3] * cout * cin * k * k
        launches += mult * len(passes)
        del x, w, y, gy
        torch.cuda.empty_cache()
    return total_ms, total_flops, launches


def time_training_leg(hp, bs, dev, precision, workload="full", steps=3, warm=3):
    """ms per step of one more training workload (fresh model + optimizer, captured step when possible) — the extra legs."""
    import gc

    import torch

    import soccerdiffusion_b200 as sd
    from soccerdiffusion_b200 import config
    from soccerdiffusion_b200.dataset.pytorch import Normalizer
    from soccerdiffusion_b200.ml.training import FusedAdamW, GraphedTrainStep, train_step
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    prev = sd.runtime.precision_name()
    sd.set_precision(precision)
    try:
        torch.manual_seed(0)
        model = config.build_model(hp).to(dev).train()
        opt = FusedAdamW(model.parameters(), lr=hp["lr"])
        sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
        norm = Normalizer(model.mean, model.std)
        host = config.synthetic_batch(hp, bs, None, seed=5)
        if workload != "full":
            host.pop("image_data", None)
        if workload == "inscope":
            host["image_tokens"] = torch.randn(bs, hp["image_context_length"], hp["hidden_dim"])
        batch = {k: v.to(dev) for k, v in host.items()}
        pre = workload == "denoiser"
        launch = "graph"
        try:
            g = GraphedTrainStep(model, opt, sch, batch, decoder_pretraining=pre, warmup_steps=2)
            step = lambda: g(batch)
        except Exception as e:   # noqa: BLE001 — an extra leg never costs the headline line
            sys.stderr.write(f"[bench] extra leg {workload}/{precision}: graph capture failed ({type(e).__name__}: {e}); eager\n")
            torch.cuda.synchronize()
            launch = "eager"
            step = lambda: train_step(model, opt, sch, norm, batch, decoder_pretraining=pre)
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return dict(ms_per_step=round(ms, 4), samples_per_s=round(bs * 1e3 / ms, 1), batch=bs, launch=launch, precision_mode=precision)
    finally:
        sd.set_precision(prev)
        step = g = model = opt = None
        gc.collect()
        torch.cuda.empty_cache()


def torch_eager_b200_leg(hp, bs, dev, steps=3, warm=2):
    """BASELINE.md §4 "also reported": the reference's module graph built from the stock torch.nn / torchvision classes it
    consists of (oracle/torch_eager_ref.py), trained eagerly on this B200 — fp32 with PyTorch's default flags, and bf16
    autocast + channels_last: the 'before' of a drop-in user."""
    import gc

    import torch

    from oracle import torch_eager_ref as R
    from soccerdiffusion_b200 import config

    out = {}
    batch = config.synthetic_batch(hp, bs, dev, seed=9)
    acp = R.alphas_cumprod().to(dev)
    mean = torch.full((hp["num_joints"],), 3.14159, device=dev)
    std = torch.full((hp["num_joints"],), 1.8138, device=dev)
    for name, autocast in (("fp32_eager", False), ("bf16_autocast_channels_last", True)):
        torch.manual_seed(0)
        model = R.StockModel(hp).to(dev).train()
        b = dict(batch)
        if autocast:
            model = model.to(memory_format=torch.channels_last)
        opt = torch.optim.AdamW(model.parameters(), lr=hp["lr"])
        for _ in range(warm):
            R.train_step(model, opt, b, acp, mean, std, autocast)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            R.train_step(model, opt, b, acp, mean, std, autocast)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = dict(ms_per_step=round(ms, 3), samples_per_s=round(bs * 1e3 / ms, 1))
        del model, opt
        gc.collect()
        torch.cuda.empty_cache()
    out["batch"] = bs
    out["what"] = ("stock torch.nn.Transformer{En,De}coder + torchvision resnet18 (the reference's own building blocks, "
                   "oracle/torch_eager_ref.py: 12,782,772 parameters as the reference), eager AdamW step, dropout 0.1, cuDNN")
    return out


def cpu_latency_legs(hp, reps=7):
    """encode_input_data at bs=1 and the 30-step DDIM sampler at bs=1 on the host cores (oracle port)."""
    import torch

    from oracle import model_ref, synth

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synth_state_dict(_state_template(hp), 0)
    batch = synth.synth_batch(hp, 1, 0)
    x_T = synth.synth_noise("x_T", hp, 1, 0)
    out = {}
    with torch.no_grad():
        ts = []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            ctx = model_ref.encode_input_data(batch, sd, hp)
            ts.append(time.perf_counter() - t0)
        out["encode_input_data_bs1_p50_ms"] = round(1e3 * sorted(ts[1:])[len(ts[1:]) // 2], 2)
        ts = []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            model_ref.sample_ddim(ctx, x_T, sd, hp, 30)
            ts.append(time.perf_counter() - t0)
        out["ddim30_sampler_bs1_p50_ms"] = round(1e3 * sorted(ts[1:])[len(ts[1:]) // 2], 2)
    out["cores"] = os.cpu_count() or 1
    out["kind"] = "port"
    return out


# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import soccerdiffusion_b200 as sd
    from soccerdiffusion_b200 import config, ops
    from soccerdiffusion_b200.dataset.pytorch import Normalizer
    from soccerdiffusion_b200.ml.training import FusedAdamW, broadcast_parameters, train_step
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sd.set_precision(args.precision)
    peaks = load_peaks()

    hp = dict(config.SCALED if args.config == "scaled" else config.DEFAULT)
    bs = args.batch
    strong = args.global_batch > 0
    if strong:
        if args.global_batch % world != 0:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} GPUs")
        bs = args.global_batch // world
    torch.manual_seed(0)
    model = config.build_model(hp).to(dev)
    model.mean.fill_(3.14159)
    model.std.fill_(1.8138)
    if world > 1:
        broadcast_parameters(model)
    model.train()
    sd.manual_seed(1234 + rank)
    opt = FusedAdamW(model.parameters(), lr=hp["lr"])
    total_steps = 4 * (args.steps + args.warmup) + 64
    lrs = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=hp["lr"], total_steps=total_steps)
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.config["num_train_timesteps"] = hp["train_denoising_timesteps"]
    norm = Normalizer(model.mean, model.std)

    from soccerdiffusion_b200.ml.training import bind_to_numa_node, gpu_numa_node

    numa = bind_to_numa_node(gpu_numa_node(local))   # BEFORE the pinned staging buffers are allocated
    use_u8 = args.workload == "full" and args.input == "uint8"
    host = config.synthetic_batch(hp, bs, None, seed=rank, pin=True, uint8_images=use_u8)
    if args.workload != "full":
        host.pop("image_data")
    batch = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    if args.workload == "inscope":
        # image tokens precomputed: the trunk (library cuDNN, SURVEY.md §8 a8') is outside the timed step; the model takes
        # them through its ``image_tokens`` input (the per-frame embedding interface of FrameEmbeddingCache)
        host["image_tokens"] = torch.randn(bs, hp["image_context_length"], hp["hidden_dim"]).pin_memory()
        batch["image_tokens"] = host["image_tokens"].to(dev)
        h2d = sum(v.numel() * v.element_size() for v in host.values())

    graphed = None
    if args.graph and not args.ncu_range:
        from soccerdiffusion_b200.ml.training import GraphedTrainStep

        try:
            graphed = GraphedTrainStep(model, opt, sch, batch, lr_scheduler=lrs, data_parallel=world > 1,
                                       decoder_pretraining=args.workload == "denoiser", warmup_steps=3)
        except Exception as e:   # capture not possible on this stack: run the step kernel by kernel
            sys.stderr.write(f"[bench] CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); eager launches\n")
            graphed = None
            torch.cuda.synchronize()

    def step(b):
        if graphed is not None:
            return graphed(b)
        return train_step(model, opt, sch, norm, b, lr_scheduler=lrs, data_parallel=world > 1,
                          decoder_pretraining=args.workload == "denoiser")

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(batch)
    sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ops.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.ncu_range:
        torch.cuda.profiler.start()   # ncu --profile-from-start off: only the timed region is captured
    e0.record()
    for _ in range(args.steps):
        step(batch)
    e1.record()
    sync()
    if args.ncu_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = ops.launches()
    clk = clocks.stop() if rank == 0 else None
    if args.ncu_range:
        if rank == 0:
            emit(dict(note="ncu range run: numbers under a profiler are not bench values", steps=args.steps,
                      ms_per_step=ms / args.steps, gpu_launches=launches))
        return

    # ---- e2e: pinned host buffers -> device EVERY step (copy stream, overlapped with the previous step's kernels),
    #      loss read back every step ------------------------------------------------------------------------
    from soccerdiffusion_b200.ml.training import DevicePrefetcher

    step({k: v.to(dev, non_blocking=True) for k, v in host.items()}).item()
    sync()
    ms_e2e = _run_e2e(step, host, dev, warm=2, timed=args.steps)

    # ---- e2e in the OTHER input format.  Default input: RAW uint8 frames — the reference's host
    #      preprocessing (ToDtype + Normalize, dataset/pytorch.py:198-215) runs on the device inside the stem's packing
    #      kernel and the per-step host->device copy is 4x smaller; alternative: float32 frames preprocessed on the host
    ms_e2e_u8, h2d_u8 = None, None
    used_graph = graphed is not None
    if args.workload == "full" and args.alt_input_leg:
        host8 = config.synthetic_batch(hp, bs, None, seed=rank, pin=True, uint8_images=not use_u8)
        h2d_u8 = sum(v.numel() * v.element_size() for v in host8.values())
        step8 = None
        if used_graph:
            graphed = None          # the fp32-frame graph (and its private memory pool) is no longer needed
            import gc

            gc.collect()
            torch.cuda.empty_cache()
            try:
                step8 = GraphedTrainStep(model, opt, sch, {k: v.to(dev) for k, v in host8.items()}, lr_scheduler=lrs,
                                         data_parallel=world > 1, warmup_steps=2)
            except Exception as e:
                sys.stderr.write(f"[bench] alternative-input leg: graph capture failed ({type(e).__name__}: {e}); eager launches\n")
                torch.cuda.synchronize()
        if step8 is None:
            step8 = lambda b: train_step(model, opt, sch, norm, b, lr_scheduler=lrs, data_parallel=world > 1)
        step8({k: v.to(dev, non_blocking=True) for k, v in host8.items()}).item()
        sync()
        ms_e2e_u8 = _run_e2e(step8, host8, dev, warm=2, timed=args.steps)
        del step8

    # pinned host -> device copy bandwidth of this box (explains the gap between `value` and `e2e`: at bs=256 the fp32 frames
    # are 1.55 GB per step)
    probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_dev = torch.empty_like(probe, device=dev)
    probe_dev.copy_(probe, non_blocking=True)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(4):
        probe_dev.copy_(probe, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbps = 4 * probe.numel() / (h0.elapsed_time(h1) * 1e6)
    del probe, probe_dev

    if world > 1:
        tt = torch.tensor([ms, ms_e2e, ms_e2e_u8 or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, mu8 = tt.tolist()
        ms_e2e_u8 = mu8 if ms_e2e_u8 is not None else None

    # ---- per-kernel-class device time (CUDA events around every libsd_b200 GEMM/attention launch) ----
    roofline = None
    roofline_own = None
    kernel_classes = None
    # every rank runs the extra steps (they contain the gradient all-reduce); only rank 0 records events
    nprof = 2
    if rank == 0:
        ops.profile_begin()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    def eager_step(b):
        return train_step(model, opt, sch, norm, b, lr_scheduler=lrs, data_parallel=world > 1,
                          decoder_pretraining=args.workload == "denoiser")

    for _ in range(nprof):
        eager_step(batch)
    pe1.record()
    torch.cuda.synchronize()
    if rank == 0:
        prof = ops.profile_end()
        step_ms = ms / args.steps   # the timed, graph-replayed step: what every share below is a fraction of

        def fmt(v):
            per_step = v["ms"] / nprof
            avg = v["ms"] / max(v["launches"], 1)
            return dict(launches_per_step=v["launches"] // nprof, ms_per_step=round(per_step, 4),
                        tflops=round(v["flops"] / v["ms"] / 1e9, 2) if v["ms"] > 0 else 0.0,
                        gbs=round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else 0.0,
                        # launches shorter than ~50 us are timed with their host launch gap inside the event pair (eager
                        # profile step): no share is claimed for them
                        share_of_step=round(per_step / step_ms, 4) if avg >= 0.05 else None)

        kernel_classes = {k: fmt(v) for k, v in prof.items() if " " not in k}
        kernel_classes["_timing"] = ("CUDA events around every libsd_b200 launch in 2 extra eager steps; share_of_step = class ms / "
                                     "graph-replayed ms_per_step; null where the average launch is < 50 us (host gaps inside the events)")
        shapes = sorted(((k, v) for k, v in prof.items() if " " in k), key=lambda kv: -kv[1]["ms"])[:10]
        kernel_classes["top_shapes"] = {k: fmt(v) for k, v in shapes}
        # the LIBRARY class of the step: the trunk's cuDNN convolutions, timed alone in conv-only CUDA graphs
        conv_ms = conv_flops = None
        if args.workload == "full" and args.precision == "bf16":
            try:
                conv_ms, conv_flops, conv_launches = conv_library_probe(dev, bs * hp["image_context_length"], hp["image_resolution"])
                kernel_classes["cudnn_conv_layer1_4"] = dict(
                    library=True, ms_per_step=round(conv_ms, 3), tflops=round(conv_flops / conv_ms / 1e9, 1),
                    share_of_step=round(conv_ms / step_ms, 4),
                    launches_per_step=conv_launches,
                    note="the LIBRARY passes of the 19 convolutions of ResNet18 layer1-4: fprop + dgrad + wgrad except layer1's four weight "
                         "gradients and the three downsample data gradients (libsd_b200 kernels: conv3x3_wgrad_c64, conv1x1s2_dgrad); each "
                         "distinct shape run alone, back to back, between CUDA events")
            except Exception as e:   # noqa: BLE001
                sys.stderr.write(f"[bench] conv library probe failed: {type(e).__name__}: {e}\n")

        def block(name, v):
            if "gemm" in name or "attention" in name or "fused" in name or "wgrad" in name:
                ach = v["flops"] / v["ms"] / 1e9
                return dict(kernel=name, bound="tensor", achieved=ach, peak=peaks["tf_sust"], unit="TFLOP/s", frac=ach / peaks["tf_sust"],
                            traffic=None, peak_source=peaks["source"] + " (sustained bf16)", launches_per_step=v["launches"] // nprof,
                            avg_launch_ms=v["ms"] / v["launches"], algorithmic_flops_per_launch=v["flops"] / v["launches"],
                            share_of_step=round(v["ms"] / nprof / step_ms, 4), library=False)
            ach = v["bytes"] / v["ms"] / 1e6
            return dict(kernel=name, bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"], traffic=None,
                        peak_source=peaks["source"] + " (copy bandwidth)", launches_per_step=v["launches"] // nprof,
                        avg_launch_ms=v["ms"] / v["launches"], algorithmic_bytes_per_launch=v["bytes"] / v["launches"],
                        share_of_step=round(v["ms"] / nprof / step_ms, 4), library=False)

        top = max((k for k in prof if " " not in k), key=lambda k: prof[k]["ms"])
        roofline_own = block(top, prof[top])
        if conv_ms is not None and conv_ms > prof[top]["ms"] / nprof:
            # the dominant kernel class of the step is a LIBRARY class: report it as such (no kernel credit is claimed for it);
            # `roofline_own` is the largest class of this repo's own kernels
            ach = conv_flops / conv_ms / 1e9
            roofline = dict(kernel="cudnn_conv_layer1_4 (library passes of the layer1-4 convolutions)", library=True, bound="tensor", achieved=ach,
                            peak=peaks["tf_sust"], unit="TFLOP/s", frac=ach / peaks["tf_sust"], traffic=None,
                            peak_source=peaks["source"] + " (sustained bf16)", launches_per_step=conv_launches,
                            avg_launch_ms=conv_ms / conv_launches, algorithmic_flops_per_launch=conv_flops / conv_launches,
                            share_of_step=round(conv_ms / step_ms, 4),
                            how="each distinct convolution shape run alone, back to back, between CUDA events (no profiler)")
        else:
            roofline = roofline_own
    if roofline_own is not None and args.workload == "full" and args.config == "default" and bs == 256:
        # DRAM bytes per launch of the dominant own kernel class from the committed ncu capture of this very command
        # (profiles/r02_dram_traffic.json, written by tools/ncu_summary.py traffic); null when no capture covers it
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_dram_traffic.json")) as fh:
                tr = json.load(fh)
            if roofline_own["kernel"] in tr:
                roofline_own["traffic"] = tr[roofline_own["kernel"]]["dram_bytes_per_launch"]
                roofline_own["traffic_source"] = tr.get("_source")
        except (OSError, ValueError, KeyError):
            pass
    if world > 1:
        dist.barrier()

    # ---- bs=1 DDIM trajectory latency (BASELINE.json configs[2]) ------------------------------------
    ddim = None
    if rank == 0 and not args.no_ddim:
        ddim = ddim_latency(model, hp, dev, args.precision)

    # ---- batched inference, sharded by trajectory over ALL ranks (BASELINE.json configs[4]; no collective) ----
    batched = None
    if not args.no_ddim:
        hpn = dict(hp, _name=args.config)
        batched = batched_inference(model, hpn, dev, args.precision, rank, world, scaled_too=args.config != "scaled")

    distill = None
    if rank == 0 and world == 1 and args.workload == "full" and not args.no_ddim:
        try:
            distill = distill_throughput(hp, dev, min(bs, 64), args.precision)
        except Exception as e:   # an extra leg must never cost the headline line
            sys.stderr.write(f"[bench] distillation leg failed: {type(e).__name__}: {e}\n")

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import synth

            cpu = cpu_baseline_sample(synth.DEFAULT_HP if args.workload == "full" else synth.DEFAULT_HP, args.cpu_batch)
        # ---- extra legs (each a few steps; none may cost the headline line) ----------------------------------------------
        extra = None
        if world == 1 and args.extra and args.workload == "full" and args.config == "default":
            extra = {}
            graphed = None
            import gc

            gc.collect()
            torch.cuda.empty_cache()

            def leg(name, fn):
                try:
                    extra[name] = fn()
                except Exception as e:   # noqa: BLE001
                    extra[name] = dict(error=f"{type(e).__name__}: {e}"[:160])
                    sys.stderr.write(f"[bench] extra leg {name} failed: {type(e).__name__}: {e}\n")
                    torch.cuda.synchronize()

            dhp = dict(config.DEFAULT)
            leg("inscope_step", lambda: time_training_leg(dhp, bs, dev, args.precision, "inscope"))
            leg("denoiser_only_step", lambda: time_training_leg(dict(dhp, hidden_dim=256), bs, dev, args.precision, "denoiser"))
            leg("denoiser_only_step_d128", lambda: time_training_leg(dhp, bs, dev, args.precision, "denoiser"))
            leg("fp32_full_step", lambda: time_training_leg(dhp, bs, dev, "fp32", "full", steps=2, warm=1))
            leg("scaled_config_step_bs128", lambda: time_training_leg(dict(config.SCALED), 128, dev, args.precision, "full"))
            leg("torch_eager_b200", lambda: torch_eager_b200_leg(dhp, bs, dev))
            if not args.no_cpu_baseline:
                from oracle import synth as _synth

                leg("cpu_latency", lambda: cpu_latency_legs(_synth.DEFAULT_HP))
            extra["_note"] = ("inscope_step: image tokens precomputed (trunk outside the step); denoiser_only_step: train.py:221-224 at "
                              "d=256 (decoder_only.yaml); fp32_full_step: the 1e-4 mode; scaled_config_step_bs128: BASELINE.json "
                              "configs[4] architecture; torch_eager_b200: the reference's stock-PyTorch module graph on this GPU")
        gb = bs * world
        fmt_in = {True: "raw uint8 frames; ToDtype(scale)+Normalize fused into the stem packing kernel on the device",
                  False: "float32 frames preprocessed on the host (the reference dataset's output)"}
        alt = (dict(value=gb * args.steps / (ms_e2e_u8 / 1e3), unit=UNIT, h2d_bytes_per_step=h2d_u8, d2h_bytes_per_step=4,
                    input=fmt_in[not use_u8]) if ms_e2e_u8 else None)
        line = dict(
            metric=METRIC, value=gb * args.steps / (ms / 1e3), unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None,
            dtype="f32" if args.precision == "fp32" else "bf16", data="synthetic", impl="ours",
            config=dict(architecture=args.config, workload={"full": "default.yaml full training step incl. ResNet18 trunk (layer1-4 convolutions = cuDNN library calls except "
                                                  "layer1's weight gradients and the downsample data gradients)",
                                  "inscope": "default.yaml training step, image tokens precomputed (trunk outside the step)",
                                  "denoiser": "denoiser-only training step (train.py:221-224)"}[args.workload],
                        global_batch=gb, per_gpu_batch=bs, parallelism=f"dp{world}", dropout_p=0.1,
                        l2="inputs %.2f GB/step per GPU > 126 MB L2; no explicit flush" % (h2d / 1e9),
                        precision_mode=args.precision, input=fmt_in[use_u8] if args.workload == "full" else "no frames",
                        numa=numa,
                        launch="one CUDA graph replay per step" if used_graph else "kernel by kernel"),
            e2e=dict(value=gb * args.steps / (ms_e2e / 1e3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                     loss_readback="every step, pinned async copy, read on the host one step behind the launch front",
                     pipeline="steady state: feeder two batches ahead, K copies issued and K steps executed between the events",
                     h2d_gbps_measured=round(h2d_gbps, 1), input=fmt_in[use_u8] if args.workload == "full" else "no frames"),
            gpu_launches=launches, clocks=clk, kernel_classes=kernel_classes, ddim=ddim, batched_inference=batched, distill=distill,
            **({"e2e_float32_frames" if use_u8 else "e2e_uint8": alt}),
            roofline=roofline, roofline_own=roofline_own, cpu_baseline=cpu, extra=extra)
        emit(line)
    if world > 1:
        # Orderly teardown: the captured step graphs hold NCCL work on the communicator, so they are destroyed FIRST (with the
        # device idle), then the process group.  A watchdog turns a teardown that does not return into a plain exit instead of
        # a hung rank (the line above is already printed).
        import gc

        graphed = None
        step = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(45.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


def _run_e2e(step, host_batch, dev, warm: int, timed: int) -> float:
    """End-to-end loop in steady state; returns the device time (ms) of ``timed`` consecutive iterations.

    One DevicePrefetcher feeds ``warm + timed + 2`` batches from pinned host memory: it runs two batches ahead, so between
    the two timing events exactly ``timed`` host->device copies are issued (for the batches two iterations later) while
    ``timed`` steps execute — copies overlap compute, which is the design being measured; the ``warm`` untimed iterations
    fill that pipeline (with a cold pipeline the first step would wait ~28 ms for its 1.5 GB batch, which a 5-step
    measurement would amortise as +5 ms/step).  Every step's loss is copied device->host (pinned, asynchronous) and read
    on the host one step behind the launch front, inside the timed region, so the device never idles waiting for the
    next launch (a training loop that logs its loss does exactly this)."""
    import torch

    from soccerdiffusion_b200.ml.training import DevicePrefetcher

    feeder = iter(DevicePrefetcher((host_batch for _ in range(warm + timed + 2)), dev))
    bufs = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    pending, losses = None, []
    for i in range(warm + timed):
        if i == warm:
            e0.record(cur)           # in stream order: after the last warm step, before the first timed one
        loss = step(next(feeder))
        buf = bufs[i % 2]
        buf.copy_(loss.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        if pending is not None:
            pending[0].synchronize()
            losses.append(float(pending[1][0]))
        pending = (ev, buf)
    e1.record(cur)
    pending[0].synchronize()
    losses.append(float(pending[1][0]))
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def ddim_latency(model, hp, dev, precision, reps=200):
    import torch

    import soccerdiffusion_b200 as sd
    from soccerdiffusion_b200 import config
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    # the persistent sampler kernels are true-fp32 in either mode; the context encoders (tick legs) run in the
    # bench's precision mode
    model.eval()
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.set_timesteps(30)
    batch = config.synthetic_batch(hp, 1, dev, seed=7)
    x_T = torch.randn(1, hp["trajectory_prediction_length"], hp["num_joints"], device=dev)
    out = {}
    from soccerdiffusion_b200.ml.inference import TrajectorySampler

    from soccerdiffusion_b200.ml.inference import FrameEmbeddingCache

    graphed = TrajectorySampler(model, sch, 30, use_cuda_graph=True)
    # (f)-2: one NEW camera frame per tick is embedded and cached; the tick itself runs on the cached frame tokens
    cache = FrameEmbeddingCache(model, use_cuda_graph=True)
    cache.push(batch["image_data"][0])
    cached_batch = {k: v for k, v in batch.items() if k != "image_data"}
    new_frame = batch["image_data"][0, -1]

    def tick_cached():
        cache.push(new_frame)
        return graphed({**cached_batch, "image_tokens": cache.tokens()}, x_T)

    with torch.no_grad():
        ctx = model.encode_input_data(batch)
        for name, fn in (("sampler", lambda: model.sample(ctx, x_T, sch, denormalize=True)),
                         ("sampler_cta", lambda: model.sample(ctx, x_T, sch, denormalize=True, sampler="cta")),
                         ("tick", lambda: model.sample(model.encode_input_data(batch), x_T, sch, denormalize=True)),
                         ("tick_graph", lambda: graphed(batch, x_T)),
                         ("tick_cached_frames", tick_cached)):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps if name == "sampler" else 30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            out[name + "_p50_ms"] = ts[len(ts) // 2]
            out[name + "_p99_ms"] = ts[min(len(ts) - 1, int(len(ts) * 0.99))]
            if name == "sampler":
                out["sampler_kernel"] = getattr(model, "last_sampler", "?")   # the kernel the bs=1 leg actually launched
    # batched inference (BASELINE.json configs[4]: 512 independent trajectories on 8 GPUs = 64 per GPU, no collective)
    with torch.no_grad():
        ctx64 = [torch.randn(64, sum(c.shape[1] for c in ctx), hp["hidden_dim"], device=dev)]
        x64 = torch.randn(64, hp["trajectory_prediction_length"], hp["num_joints"], device=dev)
        for _ in range(3):
            model.sample(ctx64, x64, sch, denormalize=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model.sample(ctx64, x64, sch, denormalize=True)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        out["batched_bs64_p50_ms"] = ts[len(ts) // 2]
        out["batched_bs64_trajectories_per_s"] = 64e3 / ts[len(ts) // 2]
        out["batched_bs64_kernel"] = getattr(model, "last_sampler", "?")
    out["steps"] = 30
    out["algorithmic_gflop_per_trajectory"] = 2.97
    out["note"] = ("sampler = x_T -> x_0 with the context given (one persistent-kernel launch; 16-CTA cluster kernel, "
                   "sampler_cta = single-CTA kernel); tick = encode_input_data (10x224^2 frames) + sampler, launched "
                   "kernel by kernel; tick_graph = the same tick replayed from one captured CUDA graph; tick_cached_frames = embed "
                   "ONE new frame (trunk on 1 frame, its own graph replay) + graph-replayed tick on the cached frame tokens")
    out["encoder_precision_mode"] = precision
    model.train()
    return out


def batched_inference(model, hp, dev, precision, rank, world, per_gpu=(64, 512), scaled_too=True):
    """Batched inference sharded by INDEPENDENT trajectories (BASELINE.json configs[4]): every rank samples
    its own `per_gpu` trajectories with the 30-step DDIM loop, no collective on the data path; timed on the device per rank
    between barriers, MAX over ranks; trajectories/s is the whole-job aggregate.  "auto" = the tensor-core sampler in bf16
    mode (layer-fused tcgen05 kernels, context K | V projected once, the loop replayed from one CUDA graph); "cta" = the fp32
    persistent one-CTA-per-trajectory kernel, for comparison."""
    import torch
    import torch.distributed as dist

    from soccerdiffusion_b200 import config
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(m, h, B, kind, reps):
        d = h["hidden_dim"]
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        lens = [h["action_context_length"], h["imu_context_length"], h["joint_state_context_length"], h["image_context_length"], 1]
        ctx = [torch.randn(B, n, d, device=dev, generator=g) for n in lens]
        x = torch.randn(B, h["trajectory_prediction_length"], h["num_joints"], device=dev, generator=g)
        sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
        sch.set_timesteps(30)
        for _ in range(3):
            m.sample(ctx, x, sch, denormalize=True, sampler=kind)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m.sample(ctx, x, sch, denormalize=True, sampler=kind)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        return dict(ms_per_batch=round(ms, 3), trajectories_per_s=round(world * B * 1e3 / ms, 1), global_trajectories=world * B,
                    kernel=getattr(m, "last_sampler", "?"))

    out = dict(n_gpus=world, sharding="independent trajectories per rank, no data-path collective", ddim_steps=30,
               precision_mode=precision, timing="CUDA events per rank between barriers, max over ranks")
    model.eval()
    with torch.no_grad():
        for B in per_gpu:
            for kind in (("auto", "cta") if B <= 64 else ("auto",)):
                try:
                    out[f"{hp.get('_name', 'arch')}_per_gpu_{B}_{kind}"] = run(model, hp, B, kind, 10 if B <= 64 else 5)
                except Exception as e:   # noqa: BLE001
                    out[f"per_gpu_{B}_{kind}"] = dict(error=f"{type(e).__name__}: {e}"[:160])
                    torch.cuda.synchronize()
        if scaled_too:
            try:
                shp = dict(config.SCALED)
                torch.manual_seed(3)
                sm = config.build_model(shp).to(dev).eval()
                out["scaled_per_gpu_64_auto"] = run(sm, shp, 64, "auto", 5)
                out["scaled_per_gpu_64_cta"] = run(sm, shp, 64, "cta", 3)
                del sm
            except Exception as e:   # noqa: BLE001
                out["scaled_per_gpu_64_auto"] = dict(error=f"{type(e).__name__}: {e}"[:160])
                torch.cuda.synchronize()
    model.train()
    return out


def distill_throughput(hp, dev, bs, precision, steps=3):
    """distill.py:160-205 as one workload: the teacher encodes the batch and samples a 30-step DDIM
    trajectory per sample under no_grad (persistent sampler kernels), the student is trained for one step on the
    teacher's context (forward_with_context at t = 0, MSE, backward, FusedAdamW)."""
    import torch

    from soccerdiffusion_b200 import config
    from soccerdiffusion_b200.ml.training import FusedAdamW, distill_step
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    torch.manual_seed(1)
    teacher = config.build_model(hp).to(dev).eval()
    student = config.build_model(hp).to(dev).train()
    opt = FusedAdamW(student.parameters(), lr=hp["lr"])
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    batch = config.synthetic_batch(hp, bs, dev, seed=3)
    for _ in range(2):
        distill_step(teacher, student, opt, sch, batch, 30)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        distill_step(teacher, student, opt, sch, batch, 30)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del teacher, student, opt
    return dict(ms_per_step=ms, samples_per_s=bs * 1e3 / ms, batch=bs, teacher_steps=30, precision_mode=precision,
                note="teacher: eval-mode encode + 30-step DDIM per sample (no_grad); student: one training step on the "
                     "teacher's context; launched kernel by kernel")


_REAL_STDOUT = None


def _capture_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SD_B200_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="STRONG scaling: total batch over all GPUs (per-GPU batch = global / N; BASELINE.json configs[3] asks 2048); "
                         "0 = weak scaling with --batch per GPU")
    ap.add_argument("--cpu-batch", type=int, default=8, help="bounded CPU sample batch")
    ap.add_argument("--workload", default="full", choices=["full", "inscope", "denoiser"])
    ap.add_argument("--config", default="default", choices=["default", "scaled"],
                    help="default.yaml, or BASELINE.json configs[4]: 2x depth, 20 frames, T=20 (use a smaller --batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddim", action="store_true")
    ap.add_argument("--input", default="uint8", choices=["uint8", "float32"],
                    help="frame format of the step's input: raw uint8 frames (normalisation fused into the stem's packing kernel) "
                         "or float32 frames preprocessed on the host (the reference dataset's output)")
    ap.add_argument("--no-alt-input-leg", "--no-uint8-leg", dest="alt_input_leg", action="store_false",
                    help="skip the extra end-to-end leg fed with the other frame format")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="skip the extra legs (in-scope / denoiser / fp32 / scaled steps, stock-PyTorch-on-B200, CPU latency legs)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch the training step kernel by kernel instead of replaying one captured CUDA graph")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop and stop after it (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
