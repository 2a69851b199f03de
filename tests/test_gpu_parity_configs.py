"""Parity of the CUDA path with the ORACLE at the sizes bench.py times (BASELINE.json configs 2, 3 and a shard of 4).

The oracle side is independent of the product exporter: oracle/fem_p1.py builds the meshes, oracle/vo_ref.CsrAssembler the
fine systems (K_ff, f_eff) and oracle/vo_ref.vo_residual / vo_residual_transposed / oracle/rom_ref.rom_fwd_adjoint restate the
reference arithmetic (VirtualObservables.py:61-69, 662-663; ROM.py:59-100 + autograd).  The CUDA side is the kernel the bench
selects for that configuration (asserted through ``kernel_path``).  Tolerances: relative 1e-10 with FP64 I/O, 1e-5 with FP32
I/O (oracle evaluated in FP64 on the FP32-rounded inputs) -- the tiers BASELINE.json's north_star states.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


_ORACLE_MESH = {}


def oracle_fine(nx_rom, refines):
    """Oracle-side fine mesh, Dirichlet map, pixel map and CSR assembler (cached per mesh)."""
    from oracle import fem_p1, vo_ref
    key = (nx_rom, refines)
    if key not in _ORACLE_MESH:
        nf = nx_rom * 2 ** refines
        cf, cells_f = fem_p1.unit_square_mesh(nf, nf)
        bc_f, _, free_f = fem_p1.dirichlet_left_right(cf, "ND")
        pix = fem_p1.pixel_of_cell(cf, cells_f, nf, nf)
        _ORACLE_MESH[key] = dict(asm=vo_ref.CsrAssembler(cf, cells_f, bc_f, free_f), pix=pix, coords=cf, free=free_f, bc=bc_f)
    return _ORACLE_MESH[key]


def sampled(B, k=8):
    idx = sorted(set([0, 1, B - 1, B // 2] + list(np.random.RandomState(7).randint(0, B, size=k))))
    return idx


# (workload, batch used in the test, expected kernel path for FP64, ptype)
CASES = [
    ("cfg2", 4096, 2, "ND"),      # BASELINE config 2 in full: 64x64, m = 25, B = 4096
    ("cfg2", 4096, 2, "NDP"),     # same mesh, per-sample Dirichlet data
    ("cfg3", 1024, 3, "ND"),      # BASELINE config 3's mesh and m = 256 (a slice of its batch: the workload generator
                                  # needs 2 GB for all 16384 fields; every sample is independent)
    ("cfg4", 16384, 2, "ND"),     # the shard one of 8 GPUs owns of BASELINE config 4 (131072 / 8)
]


@pytest.mark.parametrize("name,B,path,ptype", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_vo_residual_and_transpose_match_oracle_at_bench_sizes(name, B, path, ptype, dtype, dev):
    from oracle import vo_ref
    from gpde_b200.VirtualObservables import VoPlan
    from gpde_b200.workloads import Workload, CONFIGS
    w = Workload(name, B=B, seed=3, ptype=ptype)
    cfg = CONFIGS[name]
    O = oracle_fine(cfg["nx"], cfg["refines"])
    plan = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
    assert plan.kernel_path(w.m, dtype) == path      # FP32 I/O takes the same structured-grid kernels
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    a = torch.tensor(w.log_image, dtype=dtype, device=dev)
    y = torch.tensor(w.y, dtype=dtype, device=dev)
    g = torch.tensor(w.g_fom, dtype=dtype, device=dev)
    V = torch.tensor(w.V, dtype=dtype, device=dev)
    s = torch.tensor(np.random.RandomState(11).standard_normal((B, w.m)), dtype=dtype, device=dev)
    r = plan.residual(a, y, g, V).double().cpu().numpy()
    q = plan.residual_T(a, V, s).double().cpu().numpy()
    # conductivity input (a_is_log = 0) must give the same numbers as the log input
    r_lin = plan.residual(torch.exp(a.double()).to(dtype), y, g, V, a_is_log=False).double().cpu().numpy()
    # oracle in FP64 on the inputs as the kernel sees them (rounded to the I/O type)
    a64, y64, g64 = a.double().cpu().numpy(), y.double().cpu().numpy(), g.double().cpu().numpy()
    V64, s64 = V.double().cpu().numpy(), s.double().cpu().numpy()
    assert np.array_equal(w.physics["fom"].free_dofs, O["free"])
    worst_r = worst_q = worst_l = 0.0
    for b in sampled(B):
        K, f = O["asm"].assemble(np.exp(a64[b][O["pix"]]), g64[b])
        r0 = vo_ref.vo_residual(K, f, V64, y64[b])
        q0 = vo_ref.vo_residual_transposed(K, V64, s64[b])
        worst_r = max(worst_r, rel_err(r[b], r0))
        worst_q = max(worst_q, rel_err(q[b], q0))
        # the conductivity-input call saw exp() rounded to the I/O type: compare it with the oracle on those values
        if dtype == torch.float64:
            worst_l = max(worst_l, rel_err(r_lin[b], r0))
    assert worst_r < tol, worst_r
    assert worst_q < tol, worst_q
    assert worst_l < tol, worst_l


@pytest.mark.parametrize("name,B", [("cfg2", 4096), ("cfg3", 1024), ("cfg4", 16384)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_rom_forward_adjoint_match_oracle_at_bench_sizes(name, B, dtype, dev):
    """u, dL/dlogX, dL/dF of the whole batch against the reference's own torch ops through autograd."""
    from oracle import fem_p1, rom_ref
    from gpde_b200.ROM import ROM
    from gpde_b200.workloads import Workload, CONFIGS
    w = Workload(name, B=B, seed=5)
    cfg = CONFIGS[name]
    cc, cells_c = fem_p1.unit_square_mesh(cfg["nx"], cfg["nx"])
    bc_c, _, _ = fem_p1.dirichlet_left_right(cc, "ND")
    M = torch.tensor(fem_p1.rom_element_tensor(cc, cells_c))
    rom = ROM.FromPhysics(w.physics["rom"], dtype=dtype, device=dev)
    logX = torch.tensor(w.logX, dtype=dtype, device=dev, requires_grad=True)
    F = torch.tensor(w.F, dtype=dtype, device=dev, requires_grad=True)
    gbar = torch.tensor(w.gbar_u, dtype=dtype, device=dev)
    u = rom.solve_log(logX, F)
    u.backward(gbar)
    n_or = B if cfg["nx"] == 4 else min(B, 2048)     # 81 x 81 batched LU on the CPU: keep the oracle to seconds
    u0, gX0, gF0 = rom_ref.rom_fwd_adjoint(M, torch.tensor(bc_c), logX.detach().double().cpu()[:n_or],
                                           F.detach().double().cpu()[:n_or], gbar.double().cpu()[:n_or])
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel_err(u.detach().cpu()[:n_or], u0) < tol
    assert rel_err(logX.grad.cpu()[:n_or], gX0) < tol
    assert rel_err(F.grad.cpu()[:n_or], gF0) < tol
