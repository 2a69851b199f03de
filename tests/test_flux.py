"""Flux-balance constraints (bottleneck/flux.py:43-158): the product's vectorised, FEniCS-free implementation
(gpde_b200/flux.py) against the oracle's loop restatement (oracle/flux_ref.py), closed-form known answers, and -- on the
GPU -- the batched device assembly and a virtual-observable ensemble built with ``flux=True``."""
import numpy as np
import pytest
import torch

from conftest import rel_err


def _physics(nx, refines, ptype="NDP"):
    import gpde_b200  # noqa: F401
    from gpde_b200.physics import setup_physics
    return setup_physics(nx, nx, refines, ptype)


def _oracle_meshes(nx, refines):
    from oracle import fem_p1
    cc, cells_c = fem_p1.unit_square_mesh(nx, nx)
    nf = nx * 2 ** refines
    cf, cells_f = fem_p1.unit_square_mesh(nf, nf)
    bc, _, free = fem_p1.dirichlet_left_right(cf, "ND")
    return cc, cells_c, cf, cells_f, bc, free


@pytest.mark.parametrize("nx,refines", [(2, 1), (2, 2), (4, 1), (3, 2)])
def test_flux_matrix_matches_oracle(nx, refines):
    from oracle import flux_ref
    from gpde_b200.flux import FluxConstraintReducedOrderModel
    from gpde_b200.physics import BoundaryConditionEnsemble
    ph = _physics(nx, refines)
    cc, cells_c, cf, cells_f, bc, free = _oracle_meshes(nx, refines)
    assert np.array_equal(ph['fom'].mesh.cells, cells_f) and np.array_equal(ph['fom'].free_dofs, free)
    rng = np.random.RandomState(nx * 10 + refines)
    x = np.exp(rng.normal(0.4, 0.8, size=cells_f.shape[0]))
    fc = FluxConstraintReducedOrderModel(ph)
    assert not fc.initialized
    fc.create_measures()
    assert fc.initialized and fc.N == cells_c.shape[0] and fc.tdim == 2
    G = fc._assemble(x)
    G0 = flux_ref.flux_gamma(cc, cells_c, cf, cells_f, x)
    assert G.shape == G0.shape == (cf.shape[0], cells_c.shape[0])
    assert rel_err(G, G0) < 1e-13
    bce = BoundaryConditionEnsemble(ph, 3, "NDP", rng=rng)
    for fix in (False, True):
        fcx = FluxConstraintReducedOrderModel(ph, fix_alpha=fix)
        fcx.create_measures()
        Gr, al = fcx.assemble_reduced(x, bce[1])
        Gr0, al0 = flux_ref.flux_reduced(G0, bc, free, bce[1].constrained_dofs_values('fom'), fix_alpha=fix)
        assert Gr.shape == (fc.N, free.size) and rel_err(Gr, Gr0) < 1e-13
        if fix:
            assert rel_err(al, al0) < 1e-13 and np.abs(al0).max() > 0
        else:
            assert np.all(al == 0) and np.all(al0 == 0)      # the reference's alpha is identically zero (flux.py:153)


def test_flux_known_answers():
    """Closed forms, on the oracle AND on the product (uniform medium alpha = 1):
    (1) u = x carries no net flux out of any coarse cell (divergence theorem; the omitted Neumann edges are parallel to the
        gradient); constants carry none either (column sums vanish);
    (2) u = y: closed coarse cells balance; a cell with an edge on y = 0 misses the flux -1/nx through it, so the rest sums to
        +1/nx (and -1/nx for cells with an edge on y = 1);
    (3) linearity in alpha."""
    from oracle import flux_ref
    from gpde_b200.flux import FluxConstraintReducedOrderModel
    nx, refines = 4, 1
    ph = _physics(nx, refines)
    cc, cells_c, cf, cells_f, bc, free = _oracle_meshes(nx, refines)
    fc = FluxConstraintReducedOrderModel(ph)
    fc.create_measures()
    ones = np.ones(cells_f.shape[0])
    for G in (fc._assemble(ones), flux_ref.flux_gamma(cc, cells_c, cf, cells_f, ones)):
        assert np.abs(G.T @ cf[:, 0]).max() < 1e-13                       # (1)
        assert np.abs(G.sum(axis=0)).max() < 1e-13
        flux_y = G.T @ cf[:, 1]                                            # (2)
        n_bottom = n_top = 0
        for n in range(cells_c.shape[0]):
            ys = cc[cells_c[n], 1]
            if (ys == 0).sum() == 2:
                assert abs(flux_y[n] - 1.0 / nx) < 1e-13
                n_bottom += 1
            elif (ys == 1).sum() == 2:
                assert abs(flux_y[n] + 1.0 / nx) < 1e-13
                n_top += 1
            else:
                assert abs(flux_y[n]) < 1e-13
        assert n_bottom == n_top == nx
    rng = np.random.RandomState(0)                                         # (3)
    x1, x2 = np.exp(rng.normal(size=cells_f.shape[0])), np.exp(rng.normal(size=cells_f.shape[0]))
    assert rel_err(fc._assemble(2.0 * x1 + x2), 2.0 * fc._assemble(x1) + fc._assemble(x2)) < 1e-14


def test_flux_single_edge_sign_and_size():
    """1 x 1 coarse mesh (two coarse cells), fine mesh 4 x 4 (h = 1/4), alpha = 1, u = 1 at the nodes on x = 1 and 0 elsewhere.
    Coarse cell 0 = lower-right triangle: its Dirichlet edge x = 1 carries grad(u).n |e| = (1/h)(1)(1) = 4 (outward normal
    (1,0)); its bottom edge is a Neumann edge (omitted); on its diagonal only the last fine edge sees a gradient:
    (1/h, 0).(-1, 1)/sqrt(2) * h sqrt(2) = -1.  Net 3.  Coarse cell 1 = upper-left triangle: left edge 0, top omitted, the
    same diagonal edge from the other side: (1/h, 0).(1, -1)/sqrt(2) * h sqrt(2) = +1."""
    from oracle import flux_ref
    from gpde_b200.flux import FluxConstraintReducedOrderModel
    ph = _physics(1, 2)
    cc, cells_c, cf, cells_f, bc, free = _oracle_meshes(1, 2)
    assert np.array_equal(cells_c[0], [0, 1, 3]) and np.array_equal(cells_c[1], [0, 2, 3])
    ones = np.ones(cells_f.shape[0])
    fc = FluxConstraintReducedOrderModel(ph)
    fc.create_measures()
    u = (cf[:, 0] == 1.0).astype(np.float64)
    for G in (fc._assemble(ones), flux_ref.flux_gamma(cc, cells_c, cf, cells_f, ones)):
        assert abs(G[:, 0] @ u - 3.0) < 1e-13
        assert abs(G[:, 1] @ u - 1.0) < 1e-13
    # the intended alpha (fix_alpha) is minus the Dirichlet part of that flux: Gamma_free y - alpha = the net flux of (y, g)
    from gpde_b200.physics import BoundaryConditionEnsemble
    bce = BoundaryConditionEnsemble(ph, 1, "ND")
    fcx = FluxConstraintReducedOrderModel(ph, fix_alpha=True)
    fcx.create_measures()
    Gr, al = fcx.assemble_reduced(ones, bce[0])
    full = np.zeros(cf.shape[0])
    full[bc] = bce[0].constrained_dofs_values('fom')
    y = np.random.RandomState(1).normal(size=free.size)
    full[free] = y
    assert rel_err(Gr @ y - al, fcx._assemble(ones).T @ full) < 1e-13


@pytest.mark.gpu
def test_flux_on_device_and_in_an_ensemble():
    """Batched device assembly == host assembly; QuerryEnsemble.FromQuerryPointEnsemble(flux=True) builds CGR + flux
    queries (VirtualObservables.py:514-527) and the ensemble update conditions on them (dense route: flux samplers have no
    weighting matrix), matching the reference's Gaussian conditioning restated in the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    from oracle import vo_ref
    from gpde_b200 import VirtualObservables as VO
    from gpde_b200.flux import FluxConstraintReducedOrderModel
    from gpde_b200.physics import BoundaryConditionEnsemble
    ph = _physics(2, 2)
    fom = ph['fom']
    rng = np.random.RandomState(3)
    N = 5
    X = rng.normal(0.4, 0.8, size=(N, fom.dim_in))
    bce = BoundaryConditionEnsemble(ph, N, "NDP", rng=rng)
    for fix in (False, True):
        fc = FluxConstraintReducedOrderModel(ph, fix_alpha=fix)
        fc.create_measures()
        a = torch.tensor(np.exp(X), device=dev)
        g = torch.tensor(bce.constrained_dofs_values('fom'), device=dev)
        Gd, ad = fc.assemble_reduced_batched(a, g, dev)
        for n in range(N):
            Gh, ah = fc.assemble_reduced(np.exp(X[n]), bce[n])
            assert rel_err(Gd[n].cpu(), Gh) < 1e-13
            assert np.abs(ad[n].cpu().numpy() - ah).max() < 1e-13
    qpe = VO.QuerryPointEnsemble.FromArrays(X, bce, fom, device=dev)
    qe = VO.QuerryEnsemble.FromQuerryPointEnsemble(qpe, ph, True, True, 0, 0, dtype=torch.float64, device=dev)
    n_c, E_c = ph['rom'].mesh.num_nodes, ph['rom'].mesh.num_cells
    assert qe[0].m == n_c + E_c
    assert np.array_equal(qe[0].precision_mask, np.concatenate([-np.ones(n_c), np.ones(E_c)]))
    ens = VO.VirtualObservablesEnsemble(qpe, qe, dtype=torch.float64, device=dev)
    G = torch.tensor(rng.normal(size=(N, fom.dim_out)), device=dev)
    P = torch.tensor(rng.uniform(0.5, 2.0, size=(N, fom.dim_out)), device=dev)
    ens.update(G, P, 0)
    noise = ens._mean_vo_variances.cpu().numpy()
    for n in range(N):
        mean0, vars0 = vo_ref.virtual_observable_update(qe[n].Gamma.cpu(), qe[n].alpha.cpu(), torch.tensor(noise), G[n].cpu(), P[n].cpu())
        assert rel_err(ens.mean[n].cpu(), mean0) < 1e-9
        assert rel_err(ens.vars[n].cpu(), vars0) < 1e-8
    ens.update(G, P, 1)           # second update: learnable precisions of the flux block move (update_vo_precision)
    assert torch.isfinite(ens.mean).all() and (ens._mean_vo_variances[n_c:] > 0).all() and (ens._mean_vo_variances[:n_c] == 0).all()
