"""GPU parity of the coarse-grained model (csrc/rom.cu through the C ABI and the ROM /
ReducedOrderModelOperator mirrors) against the oracle and the reference-generated golden vectors.

Tolerances (BASELINE.json north_star): relative <= 1e-10 in FP64, <= 1e-5 in FP32, measured as
max|a-b| / max|b| per tensor."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


class _Phys(object):
    def __init__(self, bc, free):
        self.constrained_dofs, self.free_dofs = np.asarray(bc), np.asarray(free)


def _rom_from_golden(g, dtype, dev):
    from gpde_b200.ROM import ROM
    M = torch.tensor(g['const_M'], dtype=dtype, device=dev)
    return ROM(_Phys(g['const_bc_dofs_rom'], g['const_free_dofs_rom']), M, dtype, dev)


@pytest.mark.parametrize("name", ["rom_4x4_ndp", "rom_8x8_nd"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_rom_against_reference_vectors(name, dtype, dev):
    from oracle import rom_ref
    g = load_golden(name)
    rom = _rom_from_golden(g, dtype, dev)
    logX = torch.tensor(g['in_logX'], dtype=dtype, device=dev)
    F = torch.tensor(g['in_F'], dtype=dtype, device=dev)
    gbar = torch.tensor(g['in_gbar_u'], dtype=dtype, device=dev)
    if dtype == torch.float64:
        want = dict(u=g['out_u'], gx=g['out_grad_x'], gF=g['out_grad_F'], gl=g['out_grad_logX'])
    else:   # the oracle on the float32-rounded inputs
        M, bc = torch.tensor(g['const_M']), torch.tensor(g['const_bc_dofs_rom'])
        lx, Fd, gb = logX.double().cpu(), F.double().cpu(), gbar.double().cpu()
        u, gl, gF = rom_ref.rom_fwd_adjoint(M, bc, lx, Fd, gb)
        want = dict(u=u.numpy(), gl=gl.numpy(), gF=gF.numpy(), gx=None)
    tol = TOL[dtype]

    # ROM.__call__ on conductivities (reference surface)
    x = (torch.exp(logX.double()) + 1e-8).to(dtype).requires_grad_(True)
    Fr = F.clone().requires_grad_(True)
    u = rom(x, Fr)
    assert u.shape == (logX.shape[0], rom.V_dim)
    if dtype == torch.float64:
        assert rel_err(u.detach().cpu(), want['u']) < tol
        u.backward(gbar)
        assert rel_err(x.grad.cpu(), want['gx']) < tol
        assert rel_err(Fr.grad.cpu(), want['gF']) < tol

    # fused exp(.)+1e-8 path used by the operator
    lX = logX.clone().requires_grad_(True)
    Fr2 = F.clone().requires_grad_(True)
    u2 = rom.solve_log(lX, Fr2)
    u2.backward(gbar)
    assert rel_err(u2.detach().cpu(), want['u']) < tol
    assert rel_err(lX.grad.cpu(), want['gl']) < tol
    assert rel_err(Fr2.grad.cpu(), want['gF']) < tol


@pytest.mark.parametrize("name", ["rom_4x4_ndp", "rom_8x8_nd"])
def test_get_stiffness_and_return_stiffness(name, dev):
    g = load_golden(name)
    rom = _rom_from_golden(g, torch.float64, dev)
    x = torch.exp(torch.tensor(g['in_logX'], device=dev)) + 1e-8
    F = torch.tensor(g['in_F'], device=dev)
    u, K = rom(x, F, ReturnStiffness=True)
    assert K.shape == (rom.V_dim, rom.V_dim, x.shape[0])
    assert rel_err(K.cpu(), g['out_K']) < 1e-13
    K0 = rom.GetStiffness(x, DirichletBC=False).cpu().numpy()
    want = np.einsum('ije,be->ijb', g['const_M'], x.cpu().numpy())
    assert rel_err(K0, want) < 1e-13


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_operator_against_reference_vectors(dtype, dev):
    from gpde_b200.components import ReducedOrderModelOperator
    g = load_golden("rom_4x4_ndp")
    rom = _rom_from_golden(g, dtype, dev)
    W = torch.tensor(g['const_W'], dtype=dtype, device=dev)
    op = ReducedOrderModelOperator(rom, W, dtype=dtype, device=dev)
    eff = torch.tensor(g['in_logX'], dtype=dtype, device=dev, requires_grad=True)
    F = torch.tensor(g['in_F'], dtype=dtype, device=dev)
    mu, ls = op.forward(eff, F)
    assert tuple(ls.shape) == tuple(g['out_logsigmas_shape']) and ls.requires_grad
    (mu * torch.tensor(g['out_gbar_y'], dtype=dtype, device=dev)).sum().backward()
    tol = 1e-10 if dtype == torch.float64 else 2e-5   # fp32: inputs AND W u rounded
    assert rel_err(mu.detach().cpu(), g['out_mu_y']) < tol
    assert rel_err(eff.grad.cpu(), g['out_grad_effprop']) < (1e-10 if dtype == torch.float64 else 1e-4)
    assert op.dim_out == W.shape[0] and op.dim_in == 32
    s = op.propagate_samples(eff.detach(), F)
    assert s.shape == mu.shape


@pytest.mark.parametrize("nx,ny,B", [(4, 4, 257), (8, 8, 67), (6, 3, 33), (12, 12, 9), (3, 9, 5), (2, 2, 1)])
def test_rom_fresh_inputs_against_oracle(nx, ny, B, dev):
    """Ragged batches, non-square meshes, the largest admissible coarse mesh (288 cells)."""
    from oracle import fem_p1, rom_ref
    from gpde_b200.ROM import ROM
    rng = np.random.RandomState(nx * 100 + ny)
    c, cells = fem_p1.unit_square_mesh(nx, ny)
    M = fem_p1.rom_element_tensor(c, cells)
    coef = rng.uniform(-.5, .5, size=(B, 4))
    bc, _, free = fem_p1.dirichlet_left_right(c, 'ND')
    g = np.stack([fem_p1.dirichlet_left_right(c, 'NDP', coef[b])[1] for b in range(B)])
    F = fem_p1.full_F_with_applied_bc(len(c), bc, g) + 0.0
    F[:, free] = rng.normal(size=(B, len(free))) * 0.1          # non-zero load as well
    logX = rng.normal(0.4, 0.8, size=(B, len(cells)))
    gbar = rng.normal(size=(B, len(c)))
    u0, gX0, gF0 = rom_ref.rom_fwd_adjoint_closed_form(M, bc, free, logX, F, gbar)

    rom = ROM(_Phys(bc, free), torch.tensor(M, device=dev), torch.float64, dev)
    lX = torch.tensor(logX, device=dev, requires_grad=True)
    Ft = torch.tensor(F, device=dev, requires_grad=True)
    u = rom.solve_log(lX, Ft)
    u.backward(torch.tensor(gbar, device=dev))
    assert rel_err(u.detach().cpu(), u0) < 1e-10
    assert rel_err(lX.grad.cpu(), gX0) < 1e-10
    assert rel_err(Ft.grad.cpu(), gF0) < 1e-10


def test_adjoint_without_stash_recomputes_the_factor(dev):
    from gpde_b200 import ROM as rom_mod
    g = load_golden("rom_8x8_nd")
    rom = _rom_from_golden(g, torch.float64, dev)
    plan = rom._get_plan()
    X = torch.tensor(g['in_logX'], device=dev)
    F = torch.tensor(g['in_F'], device=dev)
    gb = torch.tensor(g['in_gbar_u'], device=dev)
    u, factor = rom_mod._launch_forward(plan, X, F, True, want_factor=True)
    a = rom_mod._launch_adjoint(plan, X, u, factor, gb, True)
    b = rom_mod._launch_adjoint(plan, X, u, None, gb, True)
    # with the stash: the windowed thread-per-sample kernels; without: the cooperative kernel re-factorises -- same numbers
    # to rounding (different elimination order inside a pivot)
    assert rel_err(a[0].cpu(), b[0].cpu()) < 1e-12 and rel_err(a[1].cpu(), b[1].cpu()) < 1e-12
    assert plan.half_bandwidth == 7 and plan.n_free == 63 and plan.factor_doubles == 63 * 8


def test_error_behaviour(dev):
    g = load_golden("rom_4x4_ndp")
    rom = _rom_from_golden(g, torch.float64, dev)
    F = torch.tensor(g['in_F'][:2], device=dev)
    X = torch.ones(2, 32, dtype=torch.float64, device=dev)
    X[1, 5] = 1e-13
    with pytest.raises(ValueError):          # bottleneck/ROM.py:74-76
        rom(X, F)
    rom(torch.ones(2, 32, dtype=torch.float64, device=dev), F)   # flag was cleared
    with pytest.raises(AttributeError):      # ROM.py:71 dereferences F before the None check
        rom(X, None)
    rom.deferred_checks = True
    rom(X, F)                                # no sync, no raise ...
    with pytest.raises(ValueError):
        rom.check()                          # ... until asked
    from gpde_b200 import _lib
    from gpde_b200.ROM import ROM
    cpu_rom = ROM(_Phys(g['const_bc_dofs_rom'], g['const_free_dofs_rom']), torch.tensor(g['const_M']),
                  torch.float64, torch.device("cpu"))
    with pytest.raises(_lib.GpdeLibraryError):   # no CPU fallback
        cpu_rom(torch.ones(2, 32, dtype=torch.float64), torch.tensor(g['in_F'][:2]))


def test_gradcheck_fp64(dev):
    g = load_golden("rom_4x4_ndp")
    rom = _rom_from_golden(g, torch.float64, dev)
    lX = torch.tensor(g['in_logX'][:3], device=dev, requires_grad=True)
    F = torch.tensor(g['in_F'][:3], device=dev, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: rom.solve_log(a, b), (lX, F), eps=1e-6, atol=1e-7, rtol=1e-5,
                                    nondet_tol=0.0)
    x = (torch.exp(lX.detach()) + 1e-8).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a: rom(a, F.detach()), (x,), eps=1e-6, atol=1e-7, rtol=1e-5)


def test_full_size_properties_config2(dev):
    """Batch 4096 (BASELINE config 2): closed-form answer, linearity in F, sample independence."""
    from gpde_b200.physics import setup_physics
    from gpde_b200.ROM import ROM
    ph = setup_physics(4, 4, 0, "ND")
    rom = ROM.FromPhysics(ph['rom'], dtype=torch.float64, device=dev)
    B = 4096
    gen = torch.Generator(device="cpu").manual_seed(0)
    F1 = torch.zeros(B, 25, dtype=torch.float64)
    F1[:, ph['rom'].constrained_dofs] = torch.tensor(ph['rom'].mesh.dirichlet_values("ND"))
    # uniform medium, a different constant per sample: u = x-coordinate for every sample
    Xu = torch.exp(torch.randn(B, 1, generator=gen, dtype=torch.float64)).expand(B, 32).contiguous()
    u = rom(Xu.to(dev), F1.to(dev)).cpu()
    assert (u - torch.tensor(ph['rom'].mesh.coords[:, 0])[None]).abs().max() < 1e-13
    # linearity in F and independence of the samples from their position in the batch
    X = torch.exp(0.4 + 0.8 * torch.randn(B, 32, generator=gen, dtype=torch.float64)).to(dev)
    Fa = torch.randn(B, 25, generator=gen, dtype=torch.float64).to(dev)
    Fb = torch.randn(B, 25, generator=gen, dtype=torch.float64).to(dev)
    ua, ub, uab = rom(X, Fa), rom(X, Fb), rom(X, 2.0 * Fa - 3.0 * Fb)
    assert rel_err((2.0 * ua - 3.0 * ub).cpu(), uab.cpu()) < 1e-12
    perm = torch.randperm(B, generator=gen).to(dev)
    assert torch.equal(rom(X[perm], Fa[perm]), ua[perm])      # bitwise


def test_empty_batch(dev):
    g = load_golden("rom_4x4_ndp")
    rom = _rom_from_golden(g, torch.float64, dev)
    u = rom(torch.ones(0, 32, dtype=torch.float64, device=dev), torch.zeros(0, 25, dtype=torch.float64, device=dev))
    assert u.shape == (0, 25)


def test_large_batch_config4_shard(dev):
    """One GPU's share of BASELINE config 4 at 8 GPUs (131072 / 8 samples... and the whole 131072 batch): forward +
    adjoint on the full batch equal the same samples solved in a small batch (samples are independent, bitwise)."""
    from gpde_b200.components import ReducedOrderModelOperator
    from gpde_b200.workloads import Workload
    w = Workload("cfg2", B=64, seed=4)
    op = ReducedOrderModelOperator.FromPhysics(w.physics, dtype=torch.float64, device=dev)
    B = 131072
    gen = torch.Generator().manual_seed(0)
    logX = (0.4 + 0.8 * torch.randn(B, w.E, generator=gen, dtype=torch.float64)).to(dev).requires_grad_(True)
    F = torch.tensor(w.F[0], device=dev).expand(B, -1).contiguous()
    gbar = torch.randn(B, w.n, generator=gen, dtype=torch.float64).to(dev)
    u = op.rom.solve_log(logX, F)
    u.backward(gbar)
    idx = torch.tensor([0, 1, 4095, 4096, 65535, 100000, B - 1], device=dev)
    lx = logX.detach()[idx].clone().requires_grad_(True)
    us = op.rom.solve_log(lx, F[idx])
    us.backward(gbar[idx])
    assert torch.equal(us.detach(), u.detach()[idx]) and torch.equal(lx.grad, logX.grad[idx])
    assert torch.isfinite(u).all() and torch.isfinite(logX.grad).all()
    # residual of the solve itself: A(x) u = F on the free rows, through GetStiffness on a slice
    K = op.rom.GetStiffness(torch.exp(lx.detach()) + 1e-8, DirichletBC=True)          # [n,n,7]
    res = torch.einsum('ijb,bj->bi', K, us.detach()) - F[idx]
    assert res.abs().max() < 1e-12


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ptype", ["ND", "NDP"])
def test_thread_per_sample_kernels_match_cooperative_kernels(dtype, ptype, dev, monkeypatch):
    """4x4 coarse mesh: the thread-per-sample kernels (rom_tps.cuh, the default for this shape) and the cooperative
    kernels (GPDE_ROM_PATH=coop, read once at plan creation) on the same inputs: u, dL/dlogX, dL/dF, with and without
    a factor argument, conductivity and log-conductivity input, ragged batch (B % 128 != 0)."""
    from gpde_b200 import ROM as rom_mod
    from gpde_b200.ROM import ROM
    from gpde_b200.workloads import Workload
    w = Workload("cfg1", B=333, seed=9, ptype=ptype)
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    X = torch.tensor(w.logX, dtype=dtype, device=dev)
    F = torch.tensor(w.F, dtype=dtype, device=dev)
    gbar = torch.tensor(w.gbar_u, dtype=dtype, device=dev)
    out = {}
    for path in ("tps", "coop"):
        if path == "coop":
            monkeypatch.setenv("GPDE_ROM_PATH", "coop")
        else:
            monkeypatch.delenv("GPDE_ROM_PATH", raising=False)
        rom = ROM.FromPhysics(w.physics['rom'], dtype=dtype, device=dev)
        plan = rom._get_plan()
        assert plan.lanes == (1 if path == "tps" else 8) and plan.half_bandwidth == 3 and plan.n_free == 15
        assert plan.factor_doubles == (0 if path == "tps" else 60)    # no factor stash on the thread-per-sample path
        res = []
        for x_is_log, Xin in ((True, X), (False, torch.exp(X))):
            u, factor = rom_mod._launch_forward(plan, Xin, F, x_is_log, want_factor=True, info=rom._info_word(dev))
            gX, gF = rom_mod._launch_adjoint(plan, Xin, u, factor, gbar, x_is_log, want_gradF=True)
            gX2, _ = rom_mod._launch_adjoint(plan, Xin, u, None, gbar, x_is_log, want_gradF=False)   # factor recomputed
            res += [u, gX, gF, gX2]
        rom.check()
        out[path] = res
    monkeypatch.delenv("GPDE_ROM_PATH", raising=False)
    for a, b in zip(out["tps"], out["coop"]):
        assert rel_err(a.cpu(), b.cpu()) < tol
    # error flag: a non-positive conductivity raises like the reference (ROM.py:74-76) on the default path too
    rom = ROM.FromPhysics(w.physics['rom'], dtype=dtype, device=dev)
    bad = torch.exp(X).clone()
    bad[200, 3] = 0.0
    with pytest.raises(ValueError):
        rom(bad, F)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ptype,B", [("ND", 300), ("NDP", 129), ("NDP", 1)])
def test_windowed_thread_per_sample_kernels_match_cooperative_kernels(dtype, ptype, B, dev, monkeypatch):
    """8x8 coarse mesh (BASELINE config 3): the windowed thread-per-sample kernels (rom_tpw.cuh: band streamed through a
    register window, sample-interleaved factor stash) against the cooperative kernels (GPDE_ROM_PATH=coop) on the same
    inputs: u, dL/dlogX, dL/dF, conductivity and log-conductivity input, ragged batches (B % 128 != 0, B < 128), the
    autograd surface, and the error flag of a non-positive conductivity."""
    from gpde_b200 import ROM as rom_mod
    from gpde_b200.ROM import ROM
    from gpde_b200.workloads import Workload
    w = Workload("cfg3", B=B, seed=4, ptype=ptype)
    tol = 1e-11 if dtype == torch.float64 else 2e-6
    X = torch.tensor(w.logX, dtype=dtype, device=dev)
    F = torch.tensor(w.F, dtype=dtype, device=dev)
    gbar = torch.tensor(w.gbar_u, dtype=dtype, device=dev)
    out = {}
    for path in ("tpw", "coop"):
        if path == "coop":
            monkeypatch.setenv("GPDE_ROM_PATH", "coop")
        else:
            monkeypatch.delenv("GPDE_ROM_PATH", raising=False)
        rom = ROM.FromPhysics(w.physics['rom'], dtype=dtype, device=dev)
        plan = rom._get_plan()
        assert plan.lanes == (2 if path == "tpw" else 32) and plan.half_bandwidth == 7 and plan.n_free == 63
        assert plan.factor_required == (path == "tpw")
        res = []
        for x_is_log, Xin in ((True, X), (False, torch.exp(X))):
            u, factor = rom_mod._launch_forward(plan, Xin, F, x_is_log, want_factor=True, info=rom._info_word(dev))
            if path == "tpw":
                assert factor.numel() == 63 * 8 * 128 * ((B + 127) // 128)
            gX, gF = rom_mod._launch_adjoint(plan, Xin, u, factor, gbar, x_is_log, want_gradF=True)
            gX2, _ = rom_mod._launch_adjoint(plan, Xin, u, factor, gbar, x_is_log, want_gradF=False)
            u_nograd, _ = rom_mod._launch_forward(plan, Xin, F, x_is_log, want_factor=False, info=rom._info_word(dev))
            res += [u, gX, gF, gX2, u_nograd]
        rom.check()
        out[path] = res
    monkeypatch.delenv("GPDE_ROM_PATH", raising=False)
    for k, (a, b) in enumerate(zip(out["tpw"], out["coop"])):
        assert rel_err(a.cpu(), b.cpu()) < tol, k
    rom = ROM.FromPhysics(w.physics['rom'], dtype=dtype, device=dev)
    lx = X.clone().requires_grad_(True)
    Fg = F.clone().requires_grad_(True)
    uu = rom.solve_log(lx, Fg)
    uu.backward(gbar)
    assert rel_err(uu.detach().cpu(), out["coop"][0].cpu()) < tol and rel_err(lx.grad.cpu(), out["coop"][1].cpu()) < tol
    assert rel_err(Fg.grad.cpu(), out["coop"][2].cpu()) < tol
    bad = torch.exp(X).clone()
    bad[B // 2, 77] = 0.0
    with pytest.raises(ValueError):
        rom(bad, F)
