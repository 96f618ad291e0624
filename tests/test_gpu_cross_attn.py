"""GPU parity of the layer-fused cross-attention block (csrc/cross_attn.cu; bf16 mode: rel-L2 <= 2e-2 against an fp32
restatement of the reference block, torch/nn/modules/transformer.py:1137-1139 (norm_first _mha_block) as instantiated by
ml/model/decoder.py:25-35), of the all-layer K/V projection GEMM and of its data-gradient GEMM."""
import math

import pytest
import torch

from util_gpu import rel

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2
D, H, DH = 128, 4, 32


@pytest.fixture(scope="module")
def ops():
    from soccerdiffusion_b200 import _lib, ops

    _lib.load()
    return ops


def bf(t):
    return t.to(torch.bfloat16).float()


def make_block(gen):
    r = lambda *s: torch.randn(*s, generator=gen)
    return dict(in_w=r(3 * D, D) / math.sqrt(D), in_b=r(3 * D) * 0.1, out_w=r(D, D) / math.sqrt(D), out_b=r(D) * 0.1,
                n_w=1 + 0.1 * r(D), n_b=0.1 * r(D))


def ca_ref(x, kv, P, B, T, M, mask_attn=None, mask_out=None):
    """fp32 restatement on the values the kernel sees (bf16 weights and K | V); returns y and the saved tensors."""
    F = torch.nn.functional
    xn = F.layer_norm(x, (D,), P["n_w"], P["n_b"], 1e-5)
    q = bf(xn) @ bf(P["in_w"][:D]).T + P["in_b"][:D]
    k, v = kv[:, :D].view(B, M, H, DH).transpose(1, 2), kv[:, D:].view(B, M, H, DH).transpose(1, 2)
    qh = bf(q).view(B, T, H, DH).transpose(1, 2)
    p = torch.softmax(qh @ k.transpose(-1, -2) / math.sqrt(DH), dim=-1)
    if mask_attn is not None:
        p = p * mask_attn
    attn = (p @ v).transpose(1, 2).reshape(B * T, D)
    o = bf(attn) @ bf(P["out_w"]).T + P["out_b"]
    if mask_out is not None:
        o = o * mask_out
    return x + o, dict(xn=xn, q=q, attn=attn)


def pack(ops, P, r0, rows_total):
    wp = torch.zeros(rows_total, D, device="cuda", dtype=torch.bfloat16)
    ops.pack_weights_bf16([(P["in_w"], 3 * D, r0), (P["out_w"], D, r0 + 3 * D)], wp, D)
    return wp


@pytest.mark.parametrize("rows,L", [(312 * 3, 1), (1000, 4), (128, 2), (77, 3)])
def test_kv_proj_all_layers(ops, rows, L):
    gen = torch.Generator().manual_seed(rows + L)
    mem = torch.randn(rows, D, generator=gen).cuda()
    stride = 640
    W = (torch.randn(L * stride + 64, D, generator=gen) / math.sqrt(D)).cuda()
    biases = [(torch.randn(256, generator=gen) * 0.1).cuda() for _ in range(L)]
    wp = W.to(torch.bfloat16)
    mem_bf = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    if (rows * D) % 8 == 0:
        ops.cast_bf16(mem, mem_bf)
        assert torch.equal(mem_bf, mem.to(torch.bfloat16))
    else:
        mem_bf.copy_(mem)
    kv = torch.full((rows, 256 * L + 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.kv_proj_bf16(mem_bf, wp, 64, stride, biases, kv)
    torch.cuda.synchronize()
    for l in range(L):
        want = mem_bf.float() @ wp[64 + l * stride: 64 + l * stride + 256].float().T + biases[l]
        got = kv[:, 256 * l: 256 * l + 256].float()
        assert rel(got, want) < 6e-3, (l, rel(got, want))
    assert torch.isnan(kv[:, 256 * L:].float()).all()   # columns beyond the layers are untouched


@pytest.mark.parametrize("rows,L", [(312 * 2, 1), (1000, 4), (50, 2)])
def test_kv_dgrad_all_layers(ops, rows, L):
    gen = torch.Generator().manual_seed(rows * 7 + L)
    dkv = torch.randn(rows, 256 * L, generator=gen).cuda().to(torch.bfloat16)
    stride = 1280
    wp = (torch.randn(L * stride, D, generator=gen) / math.sqrt(D)).cuda().to(torch.bfloat16)
    Wall = torch.cat([wp[640 + l * stride: 640 + l * stride + 256] for l in range(L)]).float()
    want = dkv.float() @ Wall
    dmem = torch.full((rows, D), float("nan"), device="cuda")
    ops.kv_dgrad_bf16(dkv, wp, 640, stride, L, dmem, False)
    assert rel(dmem, want) < 1e-3, rel(dmem, want)
    base = torch.randn(rows, D, generator=gen).cuda()
    dmem2 = base.clone()
    ops.kv_dgrad_bf16(dkv, wp, 640, stride, L, dmem2, True)
    assert rel(dmem2, base + want) < 1e-3


def run_fwd(ops, x, kv_all, col0, P, wp, r0, B, T, M, drop=None, save=True):
    y = torch.full_like(x, float("nan"))
    saves = None
    if save:
        b16 = lambda: torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)
        saves = (b16(), b16(), b16(), torch.zeros(B * T, 2, device="cuda"), torch.zeros(B, H, T, device="cuda"))
    ops.ca_block_fwd(x, y, B, T, M, wp, r0, r0 + 3 * D, kv_all, col0, P["in_b"], P["out_b"], P["n_w"], P["n_b"], saves=saves,
                     dropout=drop)
    torch.cuda.synchronize()
    return y, saves


@pytest.mark.parametrize("B,T,M", [(3, 10, 312), (2, 16, 384), (5, 1, 1), (4, 7, 128), (2, 10, 129), (1, 12, 200), (9, 10, 40)])
def test_ca_block_fwd_matches_fp32_restatement(ops, B, T, M):
    assert ops.ca_block_supported(D, H, T, M)
    gen = torch.Generator().manual_seed(B * 1000 + T * 10 + M)
    P = {k: v.cuda() for k, v in make_block(gen).items()}
    x = torch.randn(B * T, D, generator=gen).cuda()
    kv_all = torch.randn(B * M, 3 * 256, generator=gen).cuda().to(torch.bfloat16)   # this layer = column block 1 of 3
    col0 = 256
    r0 = 128
    wp = pack(ops, P, r0, r0 + 4 * D + 64)
    want, sv = ca_ref(x, kv_all[:, col0: col0 + 256].float(), P, B, T, M)
    y, (xn, q, attn, stats, lse) = run_fwd(ops, x, kv_all, col0, P, wp, r0, B, T, M)
    assert rel(y, want) < 5e-3, rel(y, want)
    assert rel(xn.float(), sv["xn"]) < 5e-3
    assert rel(q.float(), sv["q"]) < 6e-3
    assert rel(attn.float(), sv["attn"]) < 1e-2
    mean = x.mean(-1)
    assert rel(stats[:, 0], mean) < 1e-4 or float(mean.abs().max()) < 1e-3
    assert rel(stats[:, 1], torch.rsqrt(x.var(-1, unbiased=False) + 1e-5)) < 1e-4
    # log-sum-exp in the log2 domain
    k = kv_all[:, col0: col0 + D].float().view(B, M, H, DH).transpose(1, 2)
    s = bf(sv["q"]).view(B, T, H, DH).transpose(1, 2) @ k.transpose(-1, -2) / math.sqrt(DH)
    assert rel(lse, torch.logsumexp(s, -1) * math.log2(math.e)) < 2e-3
    # in place
    x2 = x.clone()
    ops.ca_block_fwd(x2, x2, B, T, M, wp, r0, r0 + 3 * D, kv_all, col0, P["in_b"], P["out_b"], P["n_w"], P["n_b"])
    assert torch.equal(x2, y)


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_ca_block_fwd_dropout_masks(ops, p):
    B, T, M = 4, 10, 312
    gen = torch.Generator().manual_seed(5)
    P = {k: v.cuda() for k, v in make_block(gen).items()}
    x = torch.randn(B * T, D, generator=gen).cuda()
    kv = torch.randn(B * M, 256, generator=gen).cuda().to(torch.bfloat16)
    wp = pack(ops, P, 0, 4 * D)
    seed, sid = 99, 7
    m_attn = ops.dropout_mask(B * H * T * M, p, seed, sid, "cuda").view(B, H, T, M)
    m_out = ops.dropout_mask(B * T * D, p, seed, sid + 1, "cuda").view(B * T, D)
    want, _ = ca_ref(x, kv.float(), P, B, T, M, m_attn, m_out)
    y, _ = run_fwd(ops, x, kv, 0, P, wp, 0, B, T, M, drop=(p, seed, sid))
    assert rel(y, want) < 6e-3, rel(y, want)


@pytest.mark.parametrize("B,T,M,p", [(3, 10, 312, 0.0), (2, 16, 384, 0.0), (4, 7, 128, 0.0), (2, 10, 129, 0.1), (5, 10, 312, 0.1),
                                     (3, 3, 20, 0.0), (3, 20, 322, 0.0), (2, 33, 200, 0.1), (2, 64, 384, 0.0)])
def test_ca_block_bwd_matches_autograd_of_restatement(ops, B, T, M, p):
    gen = torch.Generator().manual_seed(B * 100 + T + M)
    P = {k: v.cuda() for k, v in make_block(gen).items()}
    x = torch.randn(B * T, D, generator=gen).cuda()
    kv_all = (torch.randn(B * M, 512, generator=gen)).cuda().to(torch.bfloat16)
    col0 = 256
    wp = pack(ops, P, 0, 4 * D)
    dy = torch.randn(B * T, D, generator=gen).cuda()
    seed, sid = 4242, 10
    drop = (p, seed, sid) if p > 0 else None
    m_attn = ops.dropout_mask(B * H * T * M, p, seed, sid, "cuda").view(B, H, T, M) if p > 0 else None
    m_out = ops.dropout_mask(B * T * D, p, seed, sid + 1, "cuda").view(B * T, D) if p > 0 else None
    # autograd of the restatement (fp32 on bf16-rounded weights / K | V)
    xr = x.clone().requires_grad_(True)
    kvr = kv_all[:, col0: col0 + 256].float().requires_grad_(True)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    F = torch.nn.functional
    xn = F.layer_norm(xr, (D,), Pr["n_w"], Pr["n_b"], 1e-5)
    q = xn @ bf(P["in_w"][:D]).T + Pr["in_b"][:D]
    q.retain_grad()
    k, v = kvr[:, :D].view(B, M, H, DH).transpose(1, 2), kvr[:, D:].view(B, M, H, DH).transpose(1, 2)
    pr = torch.softmax(q.view(B, T, H, DH).transpose(1, 2) @ k.transpose(-1, -2) / math.sqrt(DH), dim=-1)
    if m_attn is not None:
        pr = pr * m_attn
    attn = (pr @ v).transpose(1, 2).reshape(B * T, D)
    o = attn @ bf(P["out_w"]).T + Pr["out_b"]
    if m_out is not None:
        o = o * m_out
    (xr + o).backward(dy)

    y, (xn_s, q_s, attn_s, stats, lse) = run_fwd(ops, x, kv_all, col0, P, wp, 0, B, T, M, drop=drop)
    dx = torch.full_like(x, float("nan"))
    b16 = lambda: torch.full((B * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    g1, dq = b16(), b16()
    dkv = torch.full((B * M, 512), float("nan"), device="cuda", dtype=torch.bfloat16)
    g_n_w, g_n_b = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.ca_block_bwd(dy, dx, x, q_s, attn_s, stats, lse, B, T, M, wp, 0, 3 * D, kv_all, col0, P["n_w"], g1, dq, dkv, g_n_w, g_n_b,
                     dropout=drop)
    torch.cuda.synchronize()
    assert rel(dx, xr.grad) < TOL_BF16, rel(dx, xr.grad)
    assert rel(dq.float(), q.grad) < TOL_BF16, rel(dq.float(), q.grad)
    assert rel(g1.float(), dy * m_out if m_out is not None else dy) < 5e-3
    assert rel(dkv[:, col0: col0 + 256].float(), kvr.grad) < TOL_BF16, rel(dkv[:, col0: col0 + 256].float(), kvr.grad)
    assert torch.isnan(dkv[:, :col0].float()).all()   # other layers' columns untouched
    assert rel(g_n_w, Pr["n_w"].grad) < TOL_BF16, rel(g_n_w, Pr["n_w"].grad)
    assert rel(g_n_b, Pr["n_b"].grad) < TOL_BF16, rel(g_n_b, Pr["n_b"].grad)
    # in place (dx aliases dy)
    dy2 = dy.clone()
    ops.ca_block_bwd(dy2, dy2, x, q_s, attn_s, stats, lse, B, T, M, wp, 0, 3 * D, kv_all, col0, P["n_w"], None, None, dkv, None, None,
                     dropout=drop)
    assert torch.equal(dy2, dx)


# ---- the tensor-core batched sampler (model.sample(sampler="tc")) ----------------------------------------------------------
def test_tc_sampler_matches_reference_golden_default(manifest):
    """bf16 mode, default.yaml: the final trajectory and the per-step predicted noise of the tensor-core DDIM loop against
    the REAL reference's golden trace (north_star: bf16 mode within 2e-2)."""
    import soccerdiffusion_b200 as sd
    from conftest import load_golden
    from oracle import synth
    from soccerdiffusion_b200.schedulers import DDIMScheduler
    from util_gpu import synth_model

    c = manifest["cases"]["default"]
    hp, B, seed, steps = synth.DEFAULT_HP, c["batch_size"], c["seed"], c["ddim_steps"]
    g = load_golden("default")
    model, sdict = synth_model(hp, seed)
    model.eval()
    ctx = [torch.from_numpy(g[k]).cuda() for k in sorted(k for k in g.files if k.startswith("ctx"))]
    x_T = synth.synth_noise("x_T", hp, B, seed).cuda()
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.config["num_train_timesteps"] = 1000
    sch.set_timesteps(steps)
    sd.set_precision("bf16")
    try:
        assert model.tc_sampler_supported(sum(t.shape[1] for t in ctx), x_T.shape[1])
        x0, trace = model.sample(ctx, x_T, sch, return_trace=True, sampler="tc")
        assert model.last_sampler == "tc"
        assert rel(x0, g["ddim_x0"]) < TOL_BF16, rel(x0, g["ddim_x0"])
        assert rel(trace, g["ddim_eps_trace"]) < TOL_BF16, rel(trace, g["ddim_eps_trace"])
        # replay of the captured loop on new inputs == a fresh kernel-by-kernel run on them
        x_T2 = synth.synth_noise("x_T", hp, B, seed + 1).cuda()
        ctx2 = [t + 0.05 * torch.randn_like(t) for t in ctx]
        a = model.sample(ctx2, x_T2, sch, sampler="tc")
        b = model._sample_tc(ctx2, x_T2, sch, False, False, use_graph=False)
        assert torch.equal(a, b)
        assert not torch.equal(a, x0)
        # denormalised output (ros.py:313)
        xd = model.sample(ctx, x_T, sch, denormalize=True, sampler="tc")
        assert rel(xd, g["ddim_x0"] * sdict["std"].numpy() + sdict["mean"].numpy()) < TOL_BF16
        # "auto" picks the tensor-core path in bf16 mode for batches, the fp32 persistent kernels for one trajectory
        from soccerdiffusion_b200 import runtime

        n = runtime.tc_sampler_min_batch()
        rep = lambda t: t.repeat((n + B - 1) // B, 1, 1)[:n].contiguous()
        model.sample([rep(t) for t in ctx], rep(x_T), sch)
        assert model.last_sampler == "tc"
        model.sample([t[:1] for t in ctx], x_T[:1], sch)
        assert model.last_sampler in ("cluster", "cta")
    finally:
        sd.set_precision("fp32")


def test_tc_sampler_batch_70_matches_fp32_kernels():
    """70 trajectories (6 self-attention tiles, the last one ragged) against the fp32 persistent sampler on the same inputs."""
    import soccerdiffusion_b200 as sd
    from oracle import synth
    from soccerdiffusion_b200.schedulers import DDIMScheduler
    from util_gpu import synth_model

    hp = dict(synth.DEFAULT_HP)
    B, d = 70, hp["hidden_dim"]
    model, _ = synth_model(hp, 3)
    model.eval()
    gen = torch.Generator().manual_seed(11)
    ctx = [torch.randn(B, n, d, generator=gen).cuda() for n in (100, 100, 100, 10, 1)]
    x_T = torch.randn(B, hp["trajectory_prediction_length"], hp["num_joints"], generator=gen).cuda()
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.set_timesteps(10)
    want = model.sample(ctx, x_T, sch, sampler="cta")
    sd.set_precision("bf16")
    try:
        got = model.sample(ctx, x_T, sch, sampler="tc")
    finally:
        sd.set_precision("fp32")
    assert rel(got, want) < TOL_BF16, rel(got, want)


@pytest.mark.parametrize("B,T,M,p", [(3, 20, 322, 0.0), (2, 33, 384, 0.0), (2, 64, 100, 0.0), (3, 17, 312, 0.1)])
def test_ca_block_fwd_inference_row_groups(ops, B, T, M, p):
    """T > 16 (the scaled-up config, T = 20): the query rows run in groups of <= 16, one CTA each."""
    assert ops.ca_block_fwd_supported(D, H, T, M) and ops.ca_block_supported(D, H, T, M)
    gen = torch.Generator().manual_seed(B * 1000 + T * 10 + M)
    P = {k: v.cuda() for k, v in make_block(gen).items()}
    x = torch.randn(B * T, D, generator=gen).cuda()
    kv = torch.randn(B * M, 256, generator=gen).cuda().to(torch.bfloat16)
    wp = pack(ops, P, 0, 4 * D)
    seed, sid = 5, 20
    m_attn = ops.dropout_mask(B * H * T * M, p, seed, sid, "cuda").view(B, H, T, M) if p > 0 else None
    m_out = ops.dropout_mask(B * T * D, p, seed, sid + 1, "cuda").view(B * T, D) if p > 0 else None
    want, _ = ca_ref(x, kv.float(), P, B, T, M, m_attn, m_out)
    y, _ = run_fwd(ops, x, kv, 0, P, wp, 0, B, T, M, drop=(p, seed, sid) if p > 0 else None, save=False)
    assert rel(y, want) < 6e-3, rel(y, want)
    y2, _ = run_fwd(ops, x, kv, 0, P, wp, 0, B, T, M, drop=(p, seed, sid) if p > 0 else None, save=True)
    assert torch.equal(y2, y)


def test_tc_sampler_scaled_config_matches_fp32_kernels():
    """BASELINE.json configs[4] architecture (8 decoder layers, T = 20, 322 memory tokens): tensor-core sampler vs the fp32
    persistent kernels on the same inputs."""
    import soccerdiffusion_b200 as sd
    from soccerdiffusion_b200 import config
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    hp = dict(config.SCALED)
    B, d = 9, hp["hidden_dim"]
    torch.manual_seed(5)
    model = config.build_model(hp).cuda().eval()
    gen = torch.Generator().manual_seed(12)
    lens = (hp["action_context_length"], hp["imu_context_length"], hp["joint_state_context_length"], hp["image_context_length"], 1)
    ctx = [torch.randn(B, n, d, generator=gen).cuda() for n in lens]
    x_T = torch.randn(B, hp["trajectory_prediction_length"], hp["num_joints"], generator=gen).cuda()
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.set_timesteps(30)
    want = model.sample(ctx, x_T, sch, sampler="cta")
    sd.set_precision("bf16")
    try:
        assert model.tc_sampler_supported(sum(lens), x_T.shape[1])
        got = model.sample(ctx, x_T, sch)
        assert model.last_sampler == "tc"
    finally:
        sd.set_precision("fp32")
    assert rel(got, want) < TOL_BF16, rel(got, want)


@pytest.mark.parametrize("B,T,J,last", [(3, 10, 20, False), (5, 7, 32, False), (2, 10, 20, True), (70, 10, 20, False)])
def test_ddim_glue_matches_separate_ops(ops, B, T, J, last):
    """fc_out + eta=0 DDIM update + next embedding (+ PE) + step-token K | V broadcast in one launch == the fp32 formulas."""
    gen = torch.Generator().manual_seed(B + T + J)
    r = lambda *s: torch.randn(*s, generator=gen).cuda()
    rows, L, Mm = B * T, 2, 9
    h, x = r(rows, D), r(rows, J)
    fc_w, fc_b, emb_w, emb_b, pe = r(J, D) / math.sqrt(D), 0.1 * r(J), r(D, J) / math.sqrt(J), 0.1 * r(D), 0.1 * r(T, D)
    coef = (0.6, 0.8, 0.9, 0.43588989)
    eps_want = h @ fc_w.T + fc_b
    x0 = (x - coef[0] * eps_want) / coef[1]
    xn_want = coef[2] * x0 + coef[3] * eps_want
    h_want = xn_want @ emb_w.T + emb_b + pe.repeat(B, 1)
    kv = torch.zeros(B * Mm, 256 * L, device="cuda", dtype=torch.bfloat16)
    src = r(2, 256 * L).to(torch.bfloat16)
    x_next, eps = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    hh = h.clone()
    ops.ddim_glue(hh, fc_w, fc_b, x, x_next, eps, coef, emb=None if last else (emb_w, emb_b, pe, T),
                  kv_bcast=None if last else (kv, Mm, Mm - 1, B, src.data_ptr() + 2 * 256 * L, 256 * L))
    torch.cuda.synchronize()
    assert rel(eps, eps_want) < 1e-5 and rel(x_next, xn_want) < 1e-5
    if last:
        assert torch.equal(hh, h) and not kv.any()
    else:
        assert rel(hh, h_want) < 1e-5
        got = kv.view(B, Mm, -1)
        assert torch.equal(got[:, Mm - 1], src[1].expand(B, -1)) and not got[:, :Mm - 1].any()
