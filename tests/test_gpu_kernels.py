"""GPU parity of the individual C-ABI entry points against the oracle (fp32 mode: rel-L2 <= 1e-4, the
tolerance BASELINE.json's north_star states; observed errors are ~1e-6)."""
import math

import numpy as np
import pytest
import torch

from oracle import model_ref
from oracle.ddim import DDIMOracle
from util_gpu import rel

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from soccerdiffusion_b200 import _lib, ops

    _lib.load()
    return ops


def cuda(a):
    return torch.as_tensor(a).cuda()


# ---------------------------------------------------------------------------------------------------
def test_step_token_int_and_float(ops):
    d = 128
    token = torch.randn(1, d // 2)
    freqs = model_ref.step_token_freqs(d)
    for t in (torch.tensor([0, 1, 33, 500, 957, 999]), torch.tensor([0.0, 12.5, 999.0])):
        want = model_ref.step_token(t, token, d)
        out = torch.empty(t.numel(), 1, d, device="cuda")
        ops.step_token(t.cuda(), freqs.cuda(), token.cuda(), out, d, t.numel(), d)
        # sin/cos arguments reach 999 rad: compare absolutely (values are in [-1,1])
        assert (out.cpu() - want).abs().max() < 2e-6


def test_q_sample_and_ddim_step(ops):
    rng = np.random.default_rng(0)
    B, T, J = 5, 10, 20
    jc = rng.uniform(0, 2 * np.pi, (B, T, J)).astype(np.float32)
    mean = rng.uniform(2.5, 3.5, J).astype(np.float32)
    std = rng.uniform(0.8, 1.8, J).astype(np.float32)
    eps = rng.standard_normal((B, T, J)).astype(np.float32)
    t = np.array([0, 1, 500, 957, 999])
    o = DDIMOracle(1000)
    x0 = ((jc - mean) / std).astype(np.float32)
    want = o.add_noise(x0, eps, t)
    xt = torch.empty(B, T, J, device="cuda")
    x0o = torch.empty(B, T, J, device="cuda")
    ops.q_sample(cuda(jc), cuda(mean), cuda(std), cuda(eps), cuda(t), cuda(o.alphas_cumprod), x0o, xt)
    assert rel(xt, want) < 1e-6 and rel(x0o, x0) < 1e-6
    o.set_timesteps(30)
    for ts in (957, 500 - 500 % 33, 0):
        r = o.step(eps, ts, want)
        prev = torch.empty(B, T, J, device="cuda")
        p0 = torch.empty(B, T, J, device="cuda")
        ops.ddim_step(cuda(want), cuda(eps), prev, p0, tuple(float(c) for c in o.coefficients(ts)))
        assert rel(prev, r.prev_sample) < 1e-6 and rel(p0, r.pred_original_sample) < 1e-6
    # last step returns x0_hat exactly
    assert torch.equal(prev, p0)


def test_mse_affine_gather(ops):
    a, b = torch.randn(7, 10, 20), torch.randn(7, 10, 20)
    out = torch.empty((), device="cuda")
    ops.mse_fwd(a.cuda(), b.cuda(), out)
    assert abs(out.item() - torch.nn.functional.mse_loss(a.double(), b.double()).item()) < 1e-6
    g = torch.empty(7, 10, 20, device="cuda")
    ops.mse_bwd(a.cuda(), b.cuda(), None, g)
    assert rel(g, 2 * (a - b) / a.numel()) < 1e-6
    mean, std = torch.rand(20) + 2, torch.rand(20) + 0.5
    y = torch.empty(7, 10, 20, device="cuda")
    ops.affine_joints(a.cuda(), mean.cuda(), std.cuda(), y, 0)
    assert rel(y, (a - mean) / std) < 1e-6
    ops.affine_joints(a.cuda(), mean.cuda(), std.cuda(), y, 1)
    assert rel(y, a * std + mean) < 1e-6
    table = torch.randn(4, 128)
    idx = torch.tensor([3, 0, 0, 2, 1])
    o = torch.empty(5, 1, 128, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.gather_rows(table.cuda(), idx.cuda(), o, 128, err)
    assert torch.equal(o.cpu()[:, 0], table[idx]) and err.item() == 0
    dt = torch.zeros(4, 128, device="cuda")
    do = torch.randn(5, 128)
    dod = do.cuda()
    ops.scatter_add_rows(dod.data_ptr(), 128, idx.cuda(), dt)
    want = torch.zeros(4, 128).index_add_(0, idx, do)
    assert rel(dt, want) < 1e-6


def test_adamw_matches_torch(ops):
    torch.manual_seed(0)
    n = 1003  # exercises the non-multiple-of-4 tail
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-2)
    npad = (n + 3) // 4 * 4
    p = torch.zeros(npad, device="cuda"); p[:n] = p0.cuda()
    m = torch.zeros(npad, device="cuda"); v = torch.zeros(npad, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n)
        ref.grad = g.clone()
        opt.step()
        gp = torch.zeros(npad, device="cuda"); gp[:n] = g.cuda()
        ops.adamw_step(p[:n], gp[:n], m[:n], v[:n], 1e-2, 0.9, 0.999, 1e-8, 1e-2, step)
        assert rel(p[:n], ref.data) < 2e-6


@pytest.mark.parametrize("M,N,K", [(1, 128, 128), (37, 20, 20), (130, 384, 128), (1000, 128, 20), (257, 129, 65)])
def test_gemm_forward_epilogues(ops, M, N, K):
    torch.manual_seed(M + N + K)
    A, W, bias = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    gamma, beta = torch.rand(K) + 0.5, torch.randn(K) * 0.1
    res, pe = torch.randn(M, N), torch.randn(7, N)
    Ad, Wd = A.cuda(), W.cuda()
    # plain + bias
    C = torch.empty(M, N, device="cuda")
    ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, bias=bias.cuda())
    assert rel(C, A.double() @ W.double().T + bias.double()) < 2e-6
    # LayerNorm-on-load + bias + GELU + residual, saving the pre-activation
    if K in (128,):
        mean, rstd = ops.ln_stats(Ad, K)
        pre = torch.empty(M, N, device="cuda")
        ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, ln=(mean, rstd, gamma.cuda(), beta.cuda()),
                 bias=bias.cuda(), pre_out=pre, ldp=N, act=ops.ACT_GELU, residual=res.cuda(), ldr=N)
        xn = torch.nn.functional.layer_norm(A.double(), (K,), gamma.double(), beta.double(), 1e-5)
        z = xn @ W.double().T + bias.double()
        assert rel(pre, z) < 2e-6
        assert rel(C, torch.nn.functional.gelu(z) + res.double()) < 2e-6
    # positional-encoding epilogue (period 7)
    ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, bias=bias.cuda(), pe=pe.cuda(), pe_period=7)
    want = A.double() @ W.double().T + bias.double() + pe.double()[torch.arange(M) % 7]
    assert rel(C, want) < 2e-6


@pytest.mark.parametrize("M,N,K", [(37, 20, 128), (1000, 128, 384), (25600, 128, 128)])
def test_gemm_dgrad_wgrad(ops, M, N, K):
    """dX = dY W (MK x KN) and dW = dY^T X (KM x KN, split-K with atomics), accumulate semantics."""
    torch.manual_seed(1)
    dY, W, X = torch.randn(M, N), torch.randn(N, K), torch.randn(M, K)
    dX = torch.empty(M, K, device="cuda")
    ops.gemm(dY.cuda(), N, ops.MK, W.cuda(), K, ops.KN, dX, K, M, K, N)
    assert rel(dX, dY.double() @ W.double()) < 2e-6
    dW = torch.ones(N, K, device="cuda")
    ops.gemm(dY.cuda(), N, ops.KM, X.cuda(), K, ops.KN, dW, K, N, K, M, accumulate=True)
    assert rel(dW, 1.0 + dY.double().T @ X.double()) < 5e-6
    if K == 128:
        gamma, beta = torch.rand(K) + 0.5, torch.randn(K) * 0.1
        mean, rstd = ops.ln_stats(X.cuda(), K)
        dW.zero_()
        ops.gemm(dY.cuda(), N, ops.KM, X.cuda(), K, ops.KN, dW, K, N, K, M, ln=(mean, rstd, gamma.cuda(), beta.cuda()),
                 accumulate=True)
        xn = torch.nn.functional.layer_norm(X.double(), (K,), gamma.double(), beta.double(), 1e-5)
        assert rel(dW, dY.double().T @ xn) < 5e-6


def test_layernorm_backward(ops):
    torch.manual_seed(2)
    M, d = 333, 128
    x = torch.randn(M, d, dtype=torch.double, requires_grad=True)
    gamma = (torch.rand(d, dtype=torch.double) + 0.5).requires_grad_(True)
    beta = torch.zeros(d, dtype=torch.double, requires_grad=True)
    g = torch.randn(M, d, dtype=torch.double)
    res = torch.randn(M, d, dtype=torch.double)
    y = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    (y * g).sum().backward()
    xd = x.detach().float().cuda()
    mean, rstd = ops.ln_stats(xd, d)
    dx = torch.empty(M, d, device="cuda")
    dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda")
    ops.ln_bwd(g.float().cuda(), xd, mean, rstd, gamma.detach().float().cuda(), res.float().cuda(), dx, dg, db, M, d)
    assert rel(dx, x.grad + res) < 2e-6
    assert rel(dg, gamma.grad) < 5e-6 and rel(db, beta.grad) < 5e-6


@pytest.mark.parametrize("B,H,T,M,dh", [(3, 4, 10, 312, 32), (2, 4, 100, 100, 32), (2, 8, 10, 10, 16),
                                         (1, 4, 20, 322, 64), (2, 4, 1, 11, 8), (1, 4, 12, 33, 128),
                                         (2, 8, 2, 2, 4)])
def test_attention_fwd_bwd(ops, B, H, T, M, dh):
    torch.manual_seed(B * 1000 + T + M)
    d = H * dh
    q = torch.randn(B, T, d, dtype=torch.double, requires_grad=True)
    k = torch.randn(B, M, d, dtype=torch.double, requires_grad=True)
    v = torch.randn(B, M, d, dtype=torch.double, requires_grad=True)
    go = torch.randn(B, T, d, dtype=torch.double)
    qh = q.view(B, T, H, dh).transpose(1, 2); kh = k.view(B, M, H, dh).transpose(1, 2); vh = v.view(B, M, H, dh).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1)
    o = (p @ vh).transpose(1, 2).reshape(B, T, d)
    (o * go).sum().backward()
    qd, kd, vd = (t.detach().float().cuda().contiguous() for t in (q, k, v))
    od = torch.empty(B, T, d, device="cuda")
    lse = torch.empty(B, H, T, device="cuda")
    ops.attention_fwd(qd.data_ptr(), d, kd.data_ptr(), d, vd.data_ptr(), d, od.data_ptr(), d, lse.data_ptr(), B, H, T, M, dh)
    assert rel(od, o) < 2e-6
    dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
    god = go.float().cuda().contiguous()
    ops.attention_bwd(qd.data_ptr(), d, kd.data_ptr(), d, vd.data_ptr(), d, od.data_ptr(), d, god.data_ptr(), d,
                      lse.data_ptr(), dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d, B, H, T, M, dh)
    assert rel(dq, q.grad) < 5e-6 and rel(dk, k.grad) < 5e-6 and rel(dv, v.grad) < 5e-6


def test_dropout_mask_statistics_and_reapply(ops):
    n, p = 1 << 20, 0.1
    m = ops.dropout_mask(n, p, 1234, 5, "cuda")
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.9) < 3e-3
    assert torch.all((m == 0) | ((m - 1 / 0.9).abs() < 1e-6))
    x = torch.randn(n, device="cuda")
    assert torch.equal(ops.dropout_apply(x, p, 1234, 5), x * m)
    assert not torch.equal(ops.dropout_mask(n, p, 1234, 6, "cuda"), m)


def test_bad_arguments_are_reported_not_crashes(ops):
    from soccerdiffusion_b200 import _lib

    x = torch.zeros(4, 100, device="cuda")
    with pytest.raises(_lib.SdError):
        ops.ln_stats(x, 100)  # unsupported width
    with pytest.raises(_lib.SdError):
        ops.attention_fwd(x.data_ptr(), 100, x.data_ptr(), 100, x.data_ptr(), 100, x.data_ptr(), 100, None, 1, 4, 1, 4, 25)
    with pytest.raises(_lib.SdError):
        ops.gemm(x, 100, ops.MK, x, 100, ops.NK, x, 4, 4, 4, 100, precision=7)


# ---------------------------------------------------------------------------------------------------
# bf16 tensor-core (tcgen05) GEMM: exact-arithmetic check against a float64 product of the bf16-ROUNDED
# operands (fp32 accumulation in TMEM => ~1e-6), plus the loose 2e-2 check against the unrounded product.
def _bf(x):
    return x.to(torch.bfloat16).double()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (1, 128, 128), (37, 20, 20), (130, 384, 128), (1000, 128, 20),
                                   (257, 144, 65), (300, 256, 256), (513, 32, 1568), (25600, 384, 128),
                                   (37, 22, 22), (203, 22, 128)])   # J = 22 joints: scalar epilogue path
def test_tc_gemm_forward_epilogues(ops, M, N, K):
    torch.manual_seed(M + N + K)
    A, W, bias = torch.randn(M, K), torch.randn(N, K) / math.sqrt(K), torch.randn(N)
    res, pe = torch.randn(M, N), torch.randn(7, N)
    Ad, Wd = A.cuda(), W.cuda()
    C = torch.empty(M, N, device="cuda")
    ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, precision=ops.PREC_BF16, bias=bias.cuda())
    exact = _bf(A) @ _bf(W).T + bias.double()
    assert rel(C, exact) < 5e-6
    assert rel(C, A.double() @ W.double().T + bias.double()) < 2e-2
    if K in (128, 256):
        gamma, beta = torch.rand(K) + 0.5, torch.randn(K) * 0.1
        mean, rstd = ops.ln_stats(Ad, K)
        pre = torch.empty(M, N, device="cuda")
        ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, precision=ops.PREC_BF16,
                 ln=(mean, rstd, gamma.cuda(), beta.cuda()), bias=bias.cuda(), pre_out=pre, ldp=N, act=ops.ACT_GELU,
                 residual=res.cuda(), ldr=N)
        xn = torch.nn.functional.layer_norm(A, (K,), gamma, beta, 1e-5)
        z = _bf(xn) @ _bf(W).T + bias.double()
        assert rel(pre, z) < 2e-3          # LN rounding differs by fp32 ulps before the bf16 rounding
        assert rel(C, torch.nn.functional.gelu(z) + res.double()) < 2e-3
    ops.gemm(Ad, K, ops.MK, Wd, K, ops.NK, C, N, M, N, K, precision=ops.PREC_BF16, bias=bias.cuda(), pe=pe.cuda(),
             pe_period=7)
    assert rel(C, exact + pe.double()[torch.arange(M) % 7]) < 5e-6


@pytest.mark.parametrize("M,N,K", [(37, 20, 128), (1000, 128, 384), (25600, 128, 128), (3120, 256, 128), (203, 128, 22)])
def test_tc_gemm_dgrad_wgrad(ops, M, N, K):
    torch.manual_seed(1)
    dY, W, X = torch.randn(M, N), torch.randn(N, K), torch.randn(M, K)
    dX = torch.empty(M, K, device="cuda")
    ops.gemm(dY.cuda(), N, ops.MK, W.cuda(), K, ops.KN, dX, K, M, K, N, precision=ops.PREC_BF16)
    assert rel(dX, _bf(dY) @ _bf(W)) < 5e-6
    dW = torch.ones(N, K, device="cuda")
    ops.gemm(dY.cuda(), N, ops.KM, X.cuda(), K, ops.KN, dW, K, N, K, M, precision=ops.PREC_BF16, accumulate=True)
    assert rel(dW, 1.0 + _bf(dY).T @ _bf(X)) < 2e-5
    dW2 = torch.empty(N, K, device="cuda")
    ops.gemm(dY.cuda(), N, ops.KM, X.cuda(), K, ops.KN, dW2, K, N, K, M, precision=ops.PREC_BF16)
    assert rel(dW2, _bf(dY).T @ _bf(X)) < 2e-5
    if K == 128:
        gamma, beta = torch.rand(K) + 0.5, torch.randn(K) * 0.1
        mean, rstd = ops.ln_stats(X.cuda(), K)
        dW.zero_()
        ops.gemm(dY.cuda(), N, ops.KM, X.cuda(), K, ops.KN, dW, K, N, K, M, precision=ops.PREC_BF16,
                 ln=(mean, rstd, gamma.cuda(), beta.cuda()), accumulate=True)
        xn = torch.nn.functional.layer_norm(X, (K,), gamma, beta, 1e-5)
        assert rel(dW, _bf(dY).T @ _bf(xn)) < 2e-3


@pytest.mark.parametrize("B,H,T,M,dh", [(3, 4, 100, 100, 32), (2, 8, 10, 10, 16), (2, 4, 10, 10, 32), (1, 2, 128, 128, 64),
                                         (2, 4, 7, 33, 32), (1, 4, 1, 5, 32)])
def test_tc_attention_fwd_bwd(ops, B, H, T, M, dh):
    """tcgen05 attention (bf16 operands): forward within bf16 rounding of the fp64 result, backward likewise; the
    dropout masks are the exported ones."""
    torch.manual_seed(B * 100 + T + M)
    d = H * dh
    p, seed, sid = 0.1, 4242, 3
    q = torch.randn(B, T, d, dtype=torch.double, requires_grad=True)
    k = torch.randn(B, M, d, dtype=torch.double, requires_grad=True)
    v = torch.randn(B, M, d, dtype=torch.double, requires_grad=True)
    go = torch.randn(B, T, d, dtype=torch.double)
    mask = ops.dropout_mask(B * H * T * M, p, seed, sid, "cuda").view(B, H, T, M).cpu().double()
    qh = q.view(B, T, H, dh).transpose(1, 2); kh = k.view(B, M, H, dh).transpose(1, 2); vh = v.view(B, M, H, dh).transpose(1, 2)
    pr = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1) * mask
    o = (pr @ vh).transpose(1, 2).reshape(B, T, d)
    (o * go).sum().backward()
    qd, kd, vd = (t_.detach().float().cuda().contiguous() for t_ in (q, k, v))
    od = torch.empty(B, T, d, device="cuda")
    lse = torch.empty(B, H, T, device="cuda")
    ops.attention_fwd(qd.data_ptr(), d, kd.data_ptr(), d, vd.data_ptr(), d, od.data_ptr(), d, lse.data_ptr(), B, H, T, M, dh,
                      (p, seed, sid), precision=ops.PREC_BF16)
    assert rel(od, o) < 1.5e-2
    dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
    god = go.float().cuda().contiguous()
    ops.attention_bwd(qd.data_ptr(), d, kd.data_ptr(), d, vd.data_ptr(), d, od.data_ptr(), d, god.data_ptr(), d,
                      lse.data_ptr(), dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d, B, H, T, M, dh, (p, seed, sid),
                      precision=ops.PREC_BF16)
    assert rel(dq, q.grad) < 2.5e-2 and rel(dk, k.grad) < 2.5e-2 and rel(dv, v.grad) < 2.5e-2
