"""Batched, matrix-free posterior update of the virtual observables (csrc/vo_posterior.cuh through the C ABI) against the
oracle's restatement of the reference's per-data-point arithmetic:
    VirtualObservable.update                      bottleneck/VirtualObservables.py:642-669  -> oracle/vo_ref.virtual_observable_update
    VirtualObservablesEnsemble.update_vo_precision                              :971-998    -> oracle/vo_ref.update_vo_precision_beta
Tolerances: 1e-9 on the posterior mean, 1e-8 on the variances (a Cholesky-inverse chain on Lambda = Gamma C Gamma^T + Sigma
whose condition number reaches 1e6 here; the reference's own golden vectors are matched at the same level)."""
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


def _oracle_system(nx, ny):
    from oracle import fem_p1, vo_ref
    c, cells = fem_p1.unit_square_mesh(nx, ny)
    bc, _, free = fem_p1.dirichlet_left_right(c, "ND")
    return vo_ref.CsrAssembler(c, cells, bc, free), fem_p1.pixel_of_cell(c, cells, nx, ny), free, bc


@pytest.mark.parametrize("nx,ny,m,N,shared", [(16, 16, 25, 5, True), (16, 16, 9, 3, False), (8, 6, 40, 4, False),
                                              (32, 32, 64, 2, True), (4, 4, 1, 7, True)])
def test_posterior_and_moments_kernels_match_oracle(nx, ny, m, N, shared, dev):
    from oracle import vo_ref
    from gpde_b200 import fem
    from gpde_b200.physics import LinearEllipticPhysics
    from gpde_b200.VirtualObservables import VoPlan
    rng = np.random.RandomState(nx * 100 + m)
    mesh = fem.P1Mesh(nx, ny)
    fom = LinearEllipticPhysics("fom", "NDP", mesh)
    plan = VoPlan(fom, dev, mesh.pixel_of_cell(), nx * ny)
    asm, pix, free, bc = _oracle_system(nx, ny)
    d = fom.dim_out
    img = rng.normal(0.3, 0.7, size=(N, nx * ny))
    gbc = rng.uniform(-0.5, 0.5, size=(N, bc.size))
    V = rng.normal(size=(d, m)) if shared else rng.normal(size=(N, d, m))
    g = rng.normal(size=(N, d))
    prec = rng.uniform(0.5, 50.0, size=(N, d))
    noise = rng.uniform(1e-4, 1e-1, size=m)
    noise[::3] = 0.0                                  # infinite-precision observables (precision_mask = -1)
    T = lambda t: torch.tensor(t, device=dev)
    a = torch.exp(T(img))
    _, rho = plan.residual(a, T(g), T(gbc), None, a_is_log=False)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    mean, vars_ = plan.posterior(a, T(V), rho, T(noise), T(g), T(prec), info=info)
    assert int(info.item()) == 0
    r_k, s2_k = plan.moments(a, T(V), rho, T(1.0 / prec))
    for n in range(N):
        K, f = asm.assemble(np.exp(img[n][pix]), gbc[n])
        Vn = V if shared else V[n]
        Gam, alp = vo_ref.construct_querry_weak_galerkin(K, f, Vn)
        mean0, vars0 = vo_ref.virtual_observable_update(Gam, alp, noise, g[n], prec[n])
        assert rel_err(mean[n].cpu(), mean0) < 1e-9, n
        assert rel_err(vars_[n].cpu(), vars0) < 1e-8, n
        assert rel_err(r_k[n].cpu(), Gam @ g[n] - alp) < 1e-10, n
        assert rel_err(s2_k[n].cpu(), (Gam ** 2) @ (1.0 / prec[n])) < 1e-10, n
    # a Lambda that is not positive definite is reported through the info word, not by a trap
    bad_noise = T(np.full(m, -1e9))
    plan.posterior(a, T(V), rho, bad_noise, T(g), T(prec), info=info)
    assert int(info.item()) == 2
    with pytest.raises(Exception):
        plan.posterior(a, T(rng.normal(size=(d, 65))), rho, T(np.ones(65)), T(g), T(prec))


def _image_ensemble(dev, N=6, seed=2, learnable=False, extra_rbf=0):
    """Reference-style construction on data that come from images (both cells of a pixel share a value)."""
    from gpde_b200 import VirtualObservables as VO
    from gpde_b200.physics import setup_physics, BoundaryConditionEnsemble
    from gpde_b200 import fem
    rng = np.random.RandomState(seed)
    ph = setup_physics(4, 4, 2, "NDP")                       # 16 x 16 fine mesh
    fom = ph['fom']
    img = fem.sample_log_field(16, 16, 0.4, 0.8, 0.15, N, rng)
    X_DG = img.reshape(N, -1)[:, fom.mesh.pixel_of_cell()]
    bce = BoundaryConditionEnsemble(ph, N, "NDP", rng=rng)
    qpe = VO.QuerryPointEnsemble.FromArrays(X_DG, bce, fom, device=dev)
    np.random.seed(seed)
    qe = VO.QuerryEnsemble.FromQuerryPointEnsemble(qpe, ph, CGR=True, flux=False, N_gaussian=0, N_rbf=extra_rbf, l_rbf=0.2,
                                                   dtype=torch.double, device=dev)
    if learnable:
        for q in qe:
            q._sampler.__class__ = type("Learnable", (q._sampler.__class__,), {"precision_mask": property(lambda s: np.ones(s.m))})
    ens = VO.VirtualObservablesEnsemble(qpe, qe, torch.double, dev)
    return ph, ens, qe, X_DG, bce, rng


@pytest.mark.parametrize("extra_rbf", [0, 4])
def test_ensemble_update_runs_matrix_free_and_matches_oracle_loop(extra_rbf, dev):
    from oracle import fem_p1, vo_ref
    ph, ens, qe, X_DG, bce, rng = _image_ensemble(dev, learnable=True, extra_rbf=extra_rbf)
    N, d, m = ens.N, ens.dim_out, ens.m
    plan, a, gbc = ens._inputs()
    assert plan.n_inputs == 256                                     # per-pixel layout detected from the DG0 fields
    V = ens._weights()
    assert V is not None and (V.dim() == 2) == (extra_rbf == 0)     # V = W shared by all data points unless RBF columns are appended
    assert all("Gamma" not in q._store for q in qe)                 # nothing dense was built
    G1, P1 = rng.normal(size=(N, d)), rng.uniform(1.0, 30.0, size=(N, d))
    G2, P2 = rng.normal(size=(N, d)), rng.uniform(1.0, 30.0, size=(N, d))
    T = lambda t: torch.tensor(t, device=dev)
    ens.update(T(G1), T(P1), 0)
    ens.check()
    m1, v1 = ens.mean.clone(), ens.vars.clone()
    ens.update(T(G2), T(P2), 1)                                     # includes the precision hyper-update on (m1, v1)
    assert all("Gamma" not in q._store for q in qe)
    # oracle: the reference's loop over data points on dense Gamma
    P = fem_p1.build_problem(4, 4, 2)
    Gams, alps = [], []
    for n in range(N):
        K, f = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], np.exp(X_DG[n]), P['bc_dofs_fom'],
                                           bce[n].constrained_dofs_values('fom'), P['free_dofs_fom'])
        Gam, alp = vo_ref.construct_querry_weak_galerkin(K, f, qe[n].V.cpu().numpy())
        Gams.append(Gam); alps.append(alp)
    prec_alpha = 0.5 * N + 1e-6
    mask = torch.zeros(m, dtype=torch.bool)
    var0 = vo_ref.mean_vo_variances(torch.ones(m, dtype=torch.double), prec_alpha, mask)
    post1 = [vo_ref.virtual_observable_update(Gams[n], alps[n], var0, G1[n], P1[n]) for n in range(N)]
    assert rel_err(m1.cpu(), torch.stack([p[0] for p in post1])) < 1e-9
    assert rel_err(v1.cpu(), torch.stack([p[1] for p in post1])) < 1e-8
    beta = vo_ref.update_vo_precision_beta(Gams, alps, [p[0] for p in post1], [p[1] for p in post1])
    assert rel_err(ens._prec_beta.cpu(), beta) < 1e-8
    var1 = vo_ref.mean_vo_variances(beta, prec_alpha, mask)
    post2 = [vo_ref.virtual_observable_update(Gams[n], alps[n], var1, G2[n], P2[n]) for n in range(N)]
    assert rel_err(ens.mean.cpu(), torch.stack([p[0] for p in post2])) < 1e-8
    assert rel_err(ens.vars.cpu(), torch.stack([p[1] for p in post2])) < 1e-7
    # members read their row of the batched result; a member updated on its own keeps its own value
    assert torch.equal(ens[2].mean, ens.mean[2]) and torch.equal(ens[2].vars, ens.vars[2])
    ens[2].update(T(G1[2]), T(P1[2]), 2, ForceUpdate=True)
    want = vo_ref.virtual_observable_update(Gams[2], alps[2], var1, G1[2], P1[2])
    assert rel_err(ens[2].mean.cpu(), want[0]) < 1e-8 and rel_err(ens.mean[2].cpu(), want[0]) < 1e-8
    assert rel_err(ens.mean[3].cpu(), post2[3][0]) < 1e-8
    # lazily materialised dense tensors still match the reference API
    assert rel_err(qe[1].Gamma.cpu(), Gams[1]) < 1e-12 and rel_err(qe[1].alpha.cpu(), alps[1]) < 1e-11


def test_ensemble_residuals_pack_cache_follows_in_place_changes(dev):
    """the packed-weights cache must engage on a pixel plan (PackedWeights, not V itself), must notice an
    in-place change of V (version counter) and must not be fooled by a different tensor at a recycled address."""
    from gpde_b200.VirtualObservables import PackedWeights
    ph, ens, qe, X_DG, bce, rng = _image_ensemble(dev, N=5)
    N, d = ens.N, ens.dim_out
    T = lambda t: torch.tensor(t, device=dev)
    Y = T(rng.normal(size=(N, d)))
    W = qe[0].V
    r = ens.residuals(Y, W)
    assert isinstance(ens._packed_weights[2], PackedWeights) and ens._packed_weights[1] is W
    assert torch.equal(ens.residuals(Y, W), r)
    W.mul_(2.0)
    assert rel_err(ens.residuals(Y, W).cpu(), 2.0 * r.cpu()) < 1e-13
    W.mul_(0.5)
    V2 = W.clone() * 3.0                                # another tensor: never served from W's packed copy
    assert rel_err(ens.residuals(Y, V2).cpu(), 3.0 * r.cpu()) < 1e-13
    assert rel_err(ens.residuals(Y.cpu(), W.cpu().numpy()).cpu(), r.cpu()) < 1e-13     # host inputs are moved, not dereferenced
    # raw C-ABI calls with a tensor on the wrong device raise cleanly instead of faulting
    plan, a, gbc = ens._inputs()
    with pytest.raises(RuntimeError):
        plan.residual(a, Y.cpu(), gbc, W, a_is_log=False)


def test_dense_route_for_many_observables(dev):
    """m > 64 per data point: the ensemble conditions densely (torch), same numbers as the kernel route on a split."""
    from gpde_b200 import VirtualObservables as VO
    ph, ens, qe, X_DG, bce, rng = _image_ensemble(dev, N=3, extra_rbf=45)     # m = 25 + 45 = 70
    assert ens.m == 70 and ens._weights() is None
    N, d = ens.N, ens.dim_out
    T = lambda t: torch.tensor(t, device=dev)
    G, P = T(rng.normal(size=(N, d))), T(rng.uniform(1.0, 30.0, size=(N, d)))
    ens.update(G, P, 0)
    for n in range(N):       # infinite precision: Gamma mean = alpha
        assert (qe[n].Gamma @ ens.mean[n] - qe[n].alpha).abs().max() < 1e-7
