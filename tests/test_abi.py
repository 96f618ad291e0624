"""The C-ABI library loads and exports every symbol include/sd_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sd_b200.h")
LIB = os.path.join(ROOT, "soccerdiffusion_b200", "libsd_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from soccerdiffusion_b200 import build

        build.build()
    return ctypes.CDLL(LIB)


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 25
    for must in ("sd_gemm", "sd_attention_fwd", "sd_attention_bwd", "sd_ddim_step", "sd_plan_sample", "sd_adamw_step"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header():
    from soccerdiffusion_b200 import _lib

    bound = set(_lib.SIGNATURES) | {"sd_error_string"}
    assert set(declared_symbols()) <= bound, set(declared_symbols()) - bound


def test_abi_version_and_error_strings(lib):
    assert lib.sd_abi_version() == 1
    lib.sd_error_string.restype = ctypes.c_char_p
    assert b"argument" in lib.sd_error_string(-1).lower() or lib.sd_error_string(-1)
    assert lib.sd_error_string(0)


def test_product_path_refuses_cpu_tensors():
    import torch

    from soccerdiffusion_b200 import _lib
    from soccerdiffusion_b200.schedulers import DDIMScheduler

    s = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    s.set_timesteps(30)
    with pytest.raises(_lib.SdError):
        s.step(torch.zeros(1, 10, 20), 957, torch.zeros(1, 10, 20))
    with pytest.raises(_lib.SdError):
        s.add_noise(torch.zeros(1, 10, 20), torch.zeros(1, 10, 20), torch.zeros(1, dtype=torch.long))


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "soccerdiffusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
