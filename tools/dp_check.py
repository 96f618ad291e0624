"""Multi-GPU check of the data-parallel step (run under torchrun, one rank per GPU):
  1. the bucketed, overlapped gradient all-reduce gives the same flat gradient as one plain all-reduce;
  2. after graph-replayed data-parallel steps on different per-rank data, all ranks hold identical parameters.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import soccerdiffusion_b200 as sd  # noqa: E402
from soccerdiffusion_b200 import config, runtime  # noqa: E402
from soccerdiffusion_b200.functional import mse_loss  # noqa: E402
from soccerdiffusion_b200.ml.training import BucketedAllReduce, FusedAdamW, GraphedTrainStep, broadcast_parameters  # noqa: E402
from soccerdiffusion_b200.ml.training.step import q_sample  # noqa: E402
from soccerdiffusion_b200.schedulers import DDIMScheduler  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd.set_precision("bf16")
    runtime.set_dropout(0.0)
    hp = dict(config.DEFAULT, image_resolution=64)
    torch.manual_seed(0)
    model = config.build_model(hp).to(dev).train()
    broadcast_parameters(model)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    batch = config.synthetic_batch(hp, 8, dev, seed=100 + rank)
    noise = torch.randn(8, 10, 20, device=dev)
    t = torch.randint(0, 1000, (8,), device=dev)

    def backward(reducer):
        opt.zero_grad()
        if reducer is not None:
            reducer.begin()
        pred = model(batch, q_sample(sch, model, batch["joint_command"], noise, t), t)
        loss = mse_loss(pred, noise)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return opt.flat_gradients()[0].clone()

    red = BucketedAllReduce(model, opt)
    assert red.split is not None and 0 < red.split < opt.flat_gradients()[0].numel()
    g_bucketed = backward(red)
    assert red.early_done, "the trunk milestone did not fire"
    g_local = backward(None)
    dist.all_reduce(g_local, op=dist.ReduceOp.SUM)
    err = float((g_bucketed - g_local).norm() / g_local.norm())
    assert err < 1e-3, err          # identical up to the summation order of the backward's atomics
    # graph-replayed data-parallel steps
    g = GraphedTrainStep(model, opt, sch, batch, data_parallel=True, warmup_steps=2)
    for i in range(3):
        loss = g(config.synthetic_batch(hp, 8, dev, seed=200 + 10 * i + rank))
    torch.cuda.synchronize()
    flat = opt.flat_parameters()[0]
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "replicas diverged"
    assert torch.isfinite(loss).item()
    if rank == 0:
        print(f"dp_check ok: world={world} bucket split at {red.split}/{flat.numel()} elements, bucketed-vs-plain gradient error {err:.2e}, "
              f"replicas identical after 3 graph-replayed steps")
    del g
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
