"""Host-side logic (CPU): scheduler tables vs the oracle, module/state_dict contract vs the reference's
names, optimizer bookkeeping guards, and the world_size-2 gradient exchange over gloo."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.ddim import DDIMOracle
from soccerdiffusion_b200.schedulers import DDIMScheduler


def test_scheduler_tables_match_oracle():
    s = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    o = DDIMOracle(1000)
    assert np.array_equal(s.betas.numpy(), o.betas)
    assert np.array_equal(s.alphas_cumprod.numpy(), o.alphas_cumprod)
    for n in (30, 10, 1, 1000):
        s.set_timesteps(n)
        o.set_timesteps(n)
        assert s.timesteps.tolist() == o.timesteps.tolist()
        ts, coef = s.schedule_tables()
        ref = np.asarray([o.coefficients(t) for t in ts], dtype=np.float32)
        # the square roots differ by <= 1 ulp: torch's CPU sqrt (what upstream diffusers executes) is not
        # correctly rounded for every input, numpy's (the oracle) is
        assert np.allclose(coef, ref, rtol=1.5e-7, atol=0.0)


def test_scheduler_config_item_assignment():
    s = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    s.config["num_train_timesteps"] = 1000  # train.py:186
    assert s.config["num_train_timesteps"] == 1000 and s.config.num_train_timesteps == 1000
    assert s.config["clip_sample"] is False
    with pytest.raises(ValueError):
        s.set_timesteps(1001)
    with pytest.raises(ValueError):
        DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False).coefficients(10)
    with pytest.raises(NotImplementedError):
        DDIMScheduler(beta_schedule="nope", clip_sample=False)


def _build(hp):
    from soccerdiffusion_b200.ml.model.encoder.image import ImageEncoderType, SequenceEncoderType
    from soccerdiffusion_b200.ml.model.encoder.imu import IMUEncoder
    from soccerdiffusion_b200.ml.model.model import End2EndDiffusionTransformer

    return End2EndDiffusionTransformer(
        num_joints=hp["num_joints"], hidden_dim=hp["hidden_dim"], use_action_history=hp["use_action_history"],
        num_action_history_encoder_layers=hp["num_action_history_encoder_layers"],
        max_action_context_length=hp["action_context_length"], encoder_patch_size=hp["encoder_patch_size"],
        use_imu=hp["use_imu"],
        imu_orientation_embedding_method=IMUEncoder.OrientationEmbeddingMethod(hp["imu_orientation_embedding_method"]),
        num_imu_encoder_layers=hp["num_imu_encoder_layers"], imu_context_length=hp["imu_context_length"],
        use_joint_states=hp["use_joint_states"], joint_state_encoder_layers=hp["joint_state_encoder_layers"],
        joint_state_context_length=hp["joint_state_context_length"], use_images=hp["use_images"],
        image_encoder_type=ImageEncoderType(hp["image_encoder_type"]),
        image_sequence_encoder_type=SequenceEncoderType(hp["image_sequence_encoder_type"]),
        num_image_sequence_encoder_layers=hp["num_image_sequence_encoder_layers"],
        image_context_length=hp["image_context_length"], image_use_final_avgpool=hp.get("image_use_final_avgpool", True),
        image_resolution=hp.get("image_resolution", 480), use_gamestate=hp["use_gamestate"],
        num_decoder_layers=hp["num_decoder_layers"], trajectory_prediction_length=hp["trajectory_prediction_length"])


@pytest.mark.parametrize("case,hp", [("tiny", synth.TINY_HP), ("patch", synth.PATCH_HP), ("default", synth.DEFAULT_HP)])
def test_state_dict_contract_equals_reference(manifest, case, hp):
    """Same names, shapes and order as the real reference's state_dict (checkpoints load unchanged)."""
    c = manifest["cases"][case]
    model = _build(hp)
    sd = model.state_dict()
    assert list(sd.keys()) == c["state_dict_names"]
    assert [list(v.shape) for v in sd.values()] == c["state_dict_shapes"]
    assert sum(p.numel() for p in model.parameters()) == c["n_params"]
    # strict load of a reference-format checkpoint, buffers updated in place (ros.py:144-145)
    mean_before = model.mean
    model.load_state_dict(synth.synth_state_dict(sd, 3))
    assert model.mean is mean_before and float(model.mean[0]) > 2.0


def test_invalid_enums_raise_value_error():
    from soccerdiffusion_b200.ml.model.encoder.image import ImageEncoderType, SequenceEncoderType
    from soccerdiffusion_b200.ml.model.encoder.imu import IMUEncoder

    with pytest.raises(ValueError):
        ImageEncoderType("vgg")
    with pytest.raises(ValueError):
        SequenceEncoderType("lstm")
    with pytest.raises(ValueError):
        IMUEncoder.OrientationEmbeddingMethod("euler")


def test_model_refuses_cpu_inputs():
    from soccerdiffusion_b200 import _lib

    model = _build(synth.PATCH_HP)
    batch = synth.synth_batch(synth.PATCH_HP, 1, 0)
    with pytest.raises(_lib.SdError):
        model(batch, torch.zeros(1, 10, 22), torch.zeros(1, dtype=torch.long))


def test_fused_adamw_refuses_cpu_parameters():
    from soccerdiffusion_b200 import _lib
    from soccerdiffusion_b200.ml.training import FusedAdamW

    with pytest.raises(_lib.SdError):
        FusedAdamW([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist

    from soccerdiffusion_b200.ml.training import allreduce_gradients, broadcast_parameters

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        lin = torch.nn.Linear(8, 4)
        broadcast_parameters(lin)
        w = lin.weight.detach().clone()
        # every rank: gradient of its own slice of a global batch of 6 rows
        xs = torch.arange(48, dtype=torch.float32).reshape(6, 8) / 10
        x = xs[rank * 3:(rank + 1) * 3]
        lin(x).pow(2).mean().backward()
        allreduce_gradients([lin.weight.grad, lin.bias.grad])
        out.put((rank, w.numpy(), lin.weight.grad.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_exchange_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0, g0), (_, w1, g1) = res
    assert np.array_equal(w0, w1)          # broadcast made the replicas identical
    assert np.array_equal(g0, g1)          # all ranks hold the same averaged gradient
    # equals the single-process gradient of the mean over the two half-batch losses
    lin = torch.nn.Linear(8, 4)
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(w0))
    xs = torch.arange(48, dtype=torch.float32).reshape(6, 8) / 10
    torch.manual_seed(0)
    ref = torch.nn.Linear(8, 4)  # rank 0's init (seed 0) incl. bias
    loss = 0.5 * (ref(xs[:3]).pow(2).mean() + ref(xs[3:]).pow(2).mean())
    loss.backward()
    assert np.allclose(g0, ref.weight.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_uint8_frame_preprocessing_equals_torchvision_pipeline():
    """normalize_u8 (the generic device-side path for raw uint8 frames; pure torch ops, so it runs on
    the CPU too) reproduces the reference's host preprocessing — v2.ToDtype(float32, scale=True) followed by
    v2.Normalize(ImageNet mean/std) (dataset/pytorch.py:198-204, ml/inference/ros.py:190-196) — bit for bit, and passes
    float frames through unchanged."""
    import torch
    from torchvision.transforms import v2

    from soccerdiffusion_b200.ml.model.encoder.trunk import normalize_u8

    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (2, 3, 3, 32, 48), generator=g, dtype=torch.uint8)
    pre = v2.Compose([v2.ToDtype(torch.float32, scale=True), v2.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    ref = torch.stack([pre(img) for img in u8.flatten(0, 1)]).view(2, 3, 3, 32, 48)
    assert torch.equal(normalize_u8(u8.flatten(0, 1)).view_as(ref), ref)
    assert normalize_u8(ref) is ref


def test_prefetcher_chunked_copy_is_a_plain_copy():
    """DevicePrefetcher issues large host->device copies as several DMA transfers; whatever the chunking, the destination
    ends up equal to the source (odd sizes, chunk larger than the tensor, non-contiguous source falls back to one copy)."""
    import torch

    from soccerdiffusion_b200.ml.training.data import _chunked_copy

    g = torch.Generator().manual_seed(1)
    for shape, chunk in (((7, 13, 5), 64), ((1000,), 4096), ((33, 3), 1 << 20), ((5, 4), 0)):
        src = torch.randn(*shape, generator=g)
        dst = torch.zeros_like(src)
        _chunked_copy(dst, src, chunk)
        assert torch.equal(dst, src)
    src = torch.randn(6, 8, generator=g).t()          # non-contiguous
    dst = torch.zeros(8, 6)
    _chunked_copy(dst, src, 16)
    assert torch.equal(dst, src)
    u8 = torch.randint(0, 256, (3, 1001), generator=g, dtype=torch.uint8)
    d8 = torch.zeros_like(u8)
    _chunked_copy(d8, u8, 100)
    assert torch.equal(d8, u8)


@pytest.mark.parametrize("rnd,capture", [("r01", "dram_r01_full_step_bf16_s22.csv"), ("r02", "dram_r02_full_step_bf16.csv")])
def test_committed_dram_traffic_matches_its_ncu_capture(tmp_path, rnd, capture):
    """profiles/rNN_dram_traffic.json (what bench.py reports as roofline.traffic) is exactly what tools/ncu_summary.py derives
    from the committed ncu CSV, and the measured DRAM bytes of the HBM-bound kernel classes equal their algorithmic bytes
    (no wasted re-reads): BatchNorm backward at bs=256 moves 10.25-14.25 B per element of its tensors."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv = os.path.join(root, "profiles", capture)
    committed = json.load(open(os.path.join(root, "profiles", f"{rnd}_dram_traffic.json")))
    out = tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "traffic", csv, str(out), "test"],
                   check=True, capture_output=True)
    fresh = json.load(open(out))
    for k, v in fresh.items():
        if k != "_source":
            assert committed[k]["launches"] == v["launches"]
            assert abs(committed[k]["dram_bytes_per_launch"] - v["dram_bytes_per_launch"]) <= 1e-6 * v["dram_bytes_per_launch"]
    assert fresh["bn_bwd"]["launches"] == 19 and fresh["bn_apply"]["launches"] == 19
    # 19 BatchNorm layers of ResNet18 after the stem at 2560 frames: 11 activations of 56x56x64-equivalent size in total
    elems = 2560 * (4 * 56 * 56 * 64 + 5 * 28 * 28 * 128 + 5 * 14 * 14 * 256 + 5 * 7 * 7 * 512)
    per_elem = fresh["bn_bwd"]["dram_bytes_per_launch"] * 19 / elems
    assert 10.0 < per_elem < 14.5, per_elem
    per_elem_apply = fresh["bn_apply"]["dram_bytes_per_launch"] * 19 / elems
    assert 4.0 < per_elem_apply < 6.5, per_elem_apply
