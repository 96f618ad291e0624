"""Pins oracle/ddim.py: known-answer values of diffusers 0.31.0's cosine DDIM schedule,
torch.cumprod equivalence, and the algebraic properties of the eta=0 step."""
import numpy as np
import pytest
import torch

from oracle.ddim import DDIMOracle, betas_for_alpha_bar


def test_known_answers():
    o = DDIMOracle(1000)
    assert float(o.betas[0]) == pytest.approx(4.128422369831242e-05, rel=1e-7)
    assert float(o.betas[999]) == pytest.approx(0.999, rel=1e-7)
    ka = {0: 0.9999586939811707, 33: 0.9958771467208862, 500: 0.4922850430011749, 924: 0.013599730096757412,
          957: 0.004278233740478754, 999: 2.4287349909002387e-09}
    for t, v in ka.items():
        assert float(o.alphas_cumprod[t]) == np.float32(v), (t, float(o.alphas_cumprod[t]), v)


def test_cumprod_matches_torch_bitwise():
    o = DDIMOracle(1000)
    betas = torch.tensor([float(b) for b in betas_for_alpha_bar(1000).astype(np.float64)], dtype=torch.float32)
    acp = torch.cumprod(1.0 - betas, dim=0).numpy()
    assert np.array_equal(acp, o.alphas_cumprod)


def test_set_timesteps_leading():
    o = DDIMOracle(1000)
    o.set_timesteps(30)
    assert o.timesteps.tolist() == list(range(957, -1, -33))
    o.set_timesteps(10)
    assert o.timesteps.tolist() == list(range(900, -1, -100))
    with pytest.raises(ValueError):
        o.set_timesteps(1001)


def test_step_requires_set_timesteps():
    with pytest.raises(ValueError):
        DDIMOracle(1000).step(np.zeros((1, 2, 3), np.float32), 10, np.zeros((1, 2, 3), np.float32))


def test_last_step_returns_pred_x0_exactly():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 10, 20)).astype(np.float32)
    e = rng.standard_normal((2, 10, 20)).astype(np.float32)
    o = DDIMOracle(1000)
    o.set_timesteps(30)
    out = o.step(e, 0, x)
    assert np.array_equal(out.prev_sample, out.pred_original_sample)


def test_add_noise_then_step_recovers_x0():
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((4, 10, 20)).astype(np.float32)
    e = rng.standard_normal((4, 10, 20)).astype(np.float32)
    o = DDIMOracle(1000)
    o.set_timesteps(30)
    for t in (33, 500, 924):
        xt = o.add_noise(x0, e, np.full((4,), t))
        rec = o.step(e, t, xt).pred_original_sample
        scale = 1.0 / np.sqrt(float(o.alphas_cumprod[t]))
        assert np.max(np.abs(rec - x0)) < 4e-6 * scale * 10


def test_step_is_linear_in_sample_and_eps():
    rng = np.random.default_rng(2)
    o = DDIMOracle(1000)
    o.set_timesteps(10)
    x, e = (rng.standard_normal((1, 10, 20)).astype(np.float32) for _ in range(2))
    a = o.step(e, 500, x).prev_sample.astype(np.float64)
    b = o.step(2 * e, 500, 2 * x).prev_sample.astype(np.float64)
    assert np.allclose(b, 2 * a, rtol=1e-6, atol=1e-6)
