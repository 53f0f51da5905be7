"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol that include/feos_b200.h declares; the Python layer fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "feos_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(feos_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from feos_torch_b200 import _lib

    return _lib


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/feos_b200.h but not exported"


def test_python_binding_covers_header(built):
    assert sorted(built.SIGNATURES) == declared_symbols()


def test_version_and_error_text(built):
    lib = built.load()
    assert lib.feos_b200_version().decode().startswith("0.")
    assert lib.feos_b200_last_error().decode() == ""


def test_argument_errors_do_not_cross_abi(built):
    lib = built.load()
    # n < 0 is rejected before any CUDA call; the text is retrievable and nothing is thrown
    rc = lib.feos_dev_pure_vle_fwd(None, None, -1, None, None, None, None, None, None, None)
    assert rc == -1
    assert b"invalid argument" in lib.feos_b200_last_error()
    with pytest.raises(built.FeosB200Error):
        built.call("feos_dev_pure_state", None, None, None, 5, None, None, None, None)
    # the state-order entry points: empty batch = success without a launch, missing buffers / bad sizes = -1
    assert lib.feos_dev_pure_order(None, None, 0, None, None) == 0
    assert lib.feos_dev_mix_point_order(None, None, None, 0, 0, None, None) == 0
    assert lib.feos_dev_gc_point_order(None, 1, None, None, 0, 0, None, None) == 0
    assert lib.feos_dev_pure_order(None, None, 8, None, None) == -1
    assert lib.feos_dev_mix_point_order(None, None, None, -3, 0, None, None) == -1
    assert lib.feos_dev_gc_point_order(None, 0, None, None, 8, 0, None, None) == -1
    assert b"invalid argument" in lib.feos_b200_last_error()


def test_public_names():
    import feos_torch_b200 as ft

    for name in ("PcSaft", "PcSaftPure"):
        assert hasattr(ft, name)


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product path must raise, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from feos_torch_b200 import PcSaftPure

    p = torch.tensor([[1.5, 3.5, 250.0, 0, 0.03, 1500.0, 1, 1]], dtype=torch.float64)
    with pytest.raises(Exception):
        PcSaftPure(p).vapor_pressure(torch.tensor([300.0], dtype=torch.float64))
