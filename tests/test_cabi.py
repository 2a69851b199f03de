"""CPU-side checks of the drop-in boundary: libgpde_b200.so builds / loads here (nvcc cross-compiles, no GPU
needed), exports every symbol include/gpde_b200.h declares, the ctypes table covers exactly those symbols, and the
host mirrors refuse to run without a CUDA device (there is no CPU fallback).  No compute entry point is called."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpde_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpde_[A-Za-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    import gpde_b200  # noqa: F401
    from gpde_b200 import _lib
    return _lib


def test_header_declares_the_documented_entry_points():
    names = _declared_symbols()
    for must in ("gpde_rom_plan_create", "gpde_rom_forward_f64", "gpde_rom_forward_f32", "gpde_rom_adjoint_f64",
                 "gpde_rom_adjoint_f32", "gpde_rom_stiffness_f64", "gpde_prolong_apply_f64", "gpde_prolong_apply_T_f64",
                 "gpde_vo_plan_create", "gpde_vo_residual_f64", "gpde_vo_residual_f32", "gpde_vo_residual_T_f64",
                 "gpde_vo_residual_T_f32", "gpde_vo_workspace_bytes", "gpde_vo_plan_kernel_path", "gpde_last_error"):
        assert must in names
    # plain C boundary: no torch / C++ types in the header
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code and 'extern "C"' in code


def test_library_exports_every_declared_symbol(lib):
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), "libgpde_b200.so does not export %s" % name


def test_ctypes_table_matches_the_header(lib):
    assert sorted(lib.SIGNATURES) == _declared_symbols()
    loaded = lib.load()
    assert loaded.gpde_version() >= 100
    msg = loaded.gpde_last_error()
    assert msg is None or isinstance(msg, bytes)


def test_no_cpu_fallback(lib):
    from gpde_b200 import VirtualObservables as VO
    from gpde_b200.ROM import ROM
    from gpde_b200.physics import setup_physics
    ph = setup_physics(2, 2, 1)
    with pytest.raises(lib.GpdeLibraryError):
        VO.VoPlan(ph['fom'], torch.device("cpu"))
    rom = ROM.FromPhysics(ph['rom'], dtype=torch.double, device=torch.device("cpu"))
    X = torch.ones(3, rom.Vc_dim, dtype=torch.double)
    F = torch.zeros(3, rom.V_dim, dtype=torch.double)
    with pytest.raises(lib.GpdeLibraryError):
        rom(X, F)
