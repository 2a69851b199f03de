"""Batched fine-mesh label solves (gpde_b200/fom_solve.py + csrc/fom_cg.cu) and the DataLoader assembly mirror
(gpde_b200/data.py) against the oracle: the reference's route is one sparse direct solve per sample
(physics/LinearElliptic.py:120-133 through utils/data.py:96-99), restated with the oracle's assembler + scipy spsolve."""
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


def _oracle_solve(nf, a_cell, g):
    import scipy.sparse.linalg as spla
    from oracle import fem_p1
    cf, cells_f = fem_p1.unit_square_mesh(nf, nf)
    bc, _, free = fem_p1.dirichlet_left_right(cf, "ND")
    K, f = fem_p1.assemble_system_free(cf, cells_f, a_cell, bc, g, free)
    return spla.spsolve(K.tocsc(), f)


@pytest.mark.parametrize("nx,refines,B,ptype", [(4, 3, 37, "NDP"), (4, 4, 130, "ND"), (2, 2, 5, "NDP")])
def test_batched_cg_matches_sparse_direct_solves(nx, refines, B, ptype, dev):
    from gpde_b200 import fem, fom_solve
    from gpde_b200.physics import setup_physics, BoundaryConditionEnsemble
    ph = setup_physics(nx, nx, refines, ptype)
    fom = ph['fom']
    nf = nx * 2 ** refines
    rng = np.random.RandomState(nx + B)
    img = fem.sample_log_field(nf, nf, 0.4, 0.8, 0.1, B, rng)
    bce = BoundaryConditionEnsemble(ph, B, ptype, rng=rng)
    gvals = bce.constrained_dofs_values('fom')
    a = torch.exp(torch.tensor(img.reshape(B, -1), device=dev))
    y, info = fom_solve.solve_batched(fom, a, torch.tensor(gvals, device=dev), tol=1e-13, return_info=True)
    assert info["converged"] and info["max_rel_residual"] <= 1e-13
    pix = fom.mesh.pixel_of_cell()
    for b in sorted(set([0, B - 1, B // 2, 3 % B])):
        y0 = _oracle_solve(nf, np.exp(img[b].reshape(-1)[pix]), gvals[b])
        assert rel_err(y[b].cpu(), y0) < 1e-9, b
    # per-cell input plan gives the same labels; warm start from the solution stops at once
    a_cell = a[:, torch.as_tensor(pix, device=dev)].contiguous()
    y2 = fom_solve.solve_batched(fom, a_cell, torch.tensor(gvals, device=dev), tol=1e-13)
    assert rel_err(y2.cpu(), y.cpu()) < 1e-9
    y3, info3 = fom_solve.solve_batched(fom, a, torch.tensor(gvals, device=dev), tol=1e-10, x0=y, return_info=True, check_every=1)
    assert info3["iterations"] <= 1 and rel_err(y3.cpu(), y.cpu()) < 1e-12
    with pytest.raises(ValueError):
        fom_solve.solve_batched(fom, -a, torch.tensor(gvals, device=dev))


def test_dataloader_assembly_on_device_matches_host_route(dev):
    """DataLoader.assemble (utils/data.py:72-119): X_DG through the pixel map, Y by the batched device solver == Y by the
    reference's serial host solves, F_ROM_BC with the Dirichlet values; guards of the reference's constructor."""
    from gpde_b200 import fem
    from gpde_b200.data import DataLoader
    from gpde_b200.physics import setup_physics, BoundaryConditionEnsemble
    ph = setup_physics(4, 4, 3, "NDP")
    N = 12
    rng = np.random.RandomState(0)

    class Sampler(object):
        def sample(self):
            return fem.sample_log_field(32, 32, 0.4, 0.8, 0.15, 1, rng)[0]

    dl = DataLoader.FromSampler(Sampler(), N)
    assert dl.X.shape == (N, 32, 32) and dl.N == N and len(dl) == N
    with pytest.raises(RuntimeError):
        dl.Y
    bce = BoundaryConditionEnsemble(ph, N, "NDP", rng=rng)
    dl.assemble(ph, bce, device=dev)
    host = DataLoader(dl.X.clone())
    host.assemble(ph, bce)
    assert torch.equal(dl.X_DG, host.X_DG) and dl.X_DG.shape == (N, 2048)
    assert rel_err(dl.Y, host.Y) < 1e-9 and dl.Y.dtype == torch.double and dl.Y.device.type == "cpu"
    assert torch.equal(dl.F_ROM_BC, host.F_ROM_BC) and dl.F_ROM_BC.shape == (N, 25)
    assert dl.solve_info["converged"]
    with pytest.raises(ValueError):
        DataLoader(dl.X.float())
    dl.lock_physics_assembly()
    with pytest.raises(RuntimeError):
        dl.assemble(ph, bce, device=dev)
    # the labels satisfy the fine system: VO residual with V = W is ~0
    from gpde_b200.VirtualObservables import VoPlan
    plan = VoPlan.cached(ph['fom'], dev, pixel_input=True)
    r = plan.residual(dl.X.reshape(N, -1).to(dev), dl.Y.to(dev), torch.tensor(bce.constrained_dofs_values('fom'), device=dev),
                      torch.tensor(ph['W'], device=dev))
    assert float(r.abs().max()) < 1e-9
