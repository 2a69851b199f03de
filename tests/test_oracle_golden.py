"""The oracle restatements (oracle/rom_ref.py, oracle/vo_ref.py) against vectors produced by the
reference's own classes (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import fem_p1, rom_ref, vo_ref
from conftest import load_golden, rel_err

ROM_CASES = ["rom_4x4_ndp", "rom_8x8_nd"]
VO_CASES = ["vo_4x4_32_ndp", "vo_2x2_8_nd"]


@pytest.mark.parametrize("name", ROM_CASES)
def test_rom_restatement_matches_reference(name):
    g = load_golden(name)
    M, bc = torch.tensor(g['const_M']), torch.tensor(g['const_bc_dofs_rom'])
    logX, F, gbar = torch.tensor(g['in_logX']), torch.tensor(g['in_F']), torch.tensor(g['in_gbar_u'])
    u, gX, gF = rom_ref.rom_fwd_adjoint(M, bc, logX, F, gbar)
    assert rel_err(u, g['out_u']) < 1e-13
    assert rel_err(gX, g['out_grad_logX']) < 1e-12
    assert rel_err(gF, g['out_grad_F']) < 1e-12
    _, K = rom_ref.rom_call(M, bc, torch.exp(logX) + 1e-8, F, return_stiffness=True)
    assert rel_err(K, g['out_K']) < 1e-14
    mu = rom_ref.operator_forward_mean(M, bc, torch.tensor(g['const_W']), logX, F)
    assert rel_err(mu, g['out_mu_y']) < 1e-13


@pytest.mark.parametrize("name", ROM_CASES)
def test_rom_closed_form_matches_reference(name):
    g = load_golden(name)
    u, gX, gF = rom_ref.rom_fwd_adjoint_closed_form(g['const_M'], g['const_bc_dofs_rom'], g['const_free_dofs_rom'],
                                                    g['in_logX'], g['in_F'], g['in_gbar_u'])
    assert rel_err(u, g['out_u']) < 1e-12
    assert rel_err(gX, g['out_grad_logX']) < 1e-11
    assert rel_err(gF, g['out_grad_F']) < 1e-11
    # gradient w.r.t. conductivities = gradient w.r.t. log-conductivities / exp(logX)
    assert rel_err(gX / np.exp(g['in_logX']), g['out_grad_x']) < 1e-11


@pytest.mark.parametrize("name", ROM_CASES)
def test_golden_constants_are_the_oracle_assembler(name):
    g = load_golden(name)
    P = fem_p1.build_problem(int(g['nx']), int(g['nx']), int(g['refines']))
    assert np.array_equal(P['M'], g['const_M'])
    assert np.array_equal(P['W'], g['const_W'])
    assert np.array_equal(P['bc_dofs_rom'], g['const_bc_dofs_rom'])


def test_rom_raises_on_nonpositive_conductivity():
    g = load_golden("rom_4x4_ndp")
    X = torch.ones(2, 32, dtype=torch.double)
    X[1, 3] = 1e-13
    with pytest.raises(ValueError):
        rom_ref.rom_call(torch.tensor(g['const_M']), torch.tensor(g['const_bc_dofs_rom']), X, torch.tensor(g['in_F'][:2]))


def _vo_problem(g):
    P = fem_p1.build_problem(int(g['nx']), int(g['nx']), int(g['refines']))
    Ks, fs = [], []
    for n in range(g['in_X_DG'].shape[0]):
        K, f = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], np.exp(g['in_X_DG'][n]), P['bc_dofs_fom'],
                                           g['in_g_fom'][n], P['free_dofs_fom'])
        Ks.append(K); fs.append(f)
    return P, Ks, fs


@pytest.mark.parametrize("name", VO_CASES)
def test_vo_gamma_alpha_residual(name):
    g = load_golden(name)
    _, Ks, fs = _vo_problem(g)
    for n, (K, f) in enumerate(zip(Ks, fs)):
        Gamma, alpha = vo_ref.construct_querry_weak_galerkin(K, f, g['in_V'])
        assert rel_err(Gamma, g['out_Gamma'][n]) < 1e-14
        assert rel_err(alpha, g['out_alpha'][n]) < 1e-13
    r = vo_ref.vo_residual_batch(Ks, fs, g['in_V'], g['in_Y'])
    assert rel_err(r, g['out_residual']) < 1e-13


@pytest.mark.parametrize("name", VO_CASES)
def test_vo_posterior_update_and_precision(name):
    g = load_golden(name)
    N = g['in_Y'].shape[0]
    m = g['in_V'].shape[1]
    inf_mask = torch.tensor(g['in_mask'] < 0)
    prec_alpha = 0.5 * N + 1e-6
    v0 = vo_ref.mean_vo_variances(torch.ones(m, dtype=torch.double), prec_alpha, inf_mask)
    means, varss = [], []
    for n in range(N):
        mu, va = vo_ref.virtual_observable_update(g['out_Gamma'][n], g['out_alpha'][n], v0, g['in_Y'][n], g['in_PREC1'][n])
        means.append(mu.numpy()); varss.append(va.numpy())
    assert rel_err(np.stack(means), g['out_mean1']) < 1e-10
    assert rel_err(np.stack(varss), g['out_vars1']) < 1e-9
    beta = vo_ref.update_vo_precision_beta(g['out_Gamma'], g['out_alpha'], means, varss)
    assert rel_err(beta, g['out_prec_beta']) < 1e-9
    v1 = vo_ref.mean_vo_variances(beta, prec_alpha, inf_mask)
    assert rel_err(v1, g['out_mean_vo_variances']) < 1e-9
    for n in range(N):
        mu, va = vo_ref.virtual_observable_update(g['out_Gamma'][n], g['out_alpha'][n], v1, g['in_G2'][n], g['in_PREC2'][n])
        assert rel_err(mu, g['out_mean2'][n]) < 1e-9
        assert rel_err(va, g['out_vars2'][n]) < 1e-8
    # with infinite precision the posterior mean satisfies the constraint exactly
    k = np.nonzero(g['in_mask'] < 0)[0]
    for n in range(N):
        assert np.abs(g['out_Gamma'][n][k] @ g['out_mean1'][n] - g['out_alpha'][n][k]).max() < 1e-8


def test_energy_vo_restatement_matches_reference():
    """oracle/vo_ref.energy_vo_update against EnergyVirtualObservablesEnsemble of the reference
    (tests/golden/energy_2x2_16_ndp.npz, generated by tests/golden/make_golden.py energy)."""
    from oracle import fem_p1, vo_ref
    g = load_golden("energy_2x2_16_ndp")
    P = fem_p1.build_problem(int(g['nx']), int(g['nx']), int(g['refines']))
    N, n_it = g['in_X_DG'].shape[0], int(g['n_it'])
    means = [np.zeros(len(P['free_dofs_fom'])) for _ in range(N)]
    call = 0
    for it in range(g['in_G'].shape[0]):
        T = float(g['out_temperature'][it])
        for n in range(N):
            K, f = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], np.exp(g['in_X_DG'][n]), P['bc_dofs_fom'],
                                               g['in_g_fom'][n], P['free_dofs_fom'])
            Vs = [g['in_V_seq'][call + i] for i in range(n_it)]
            call += n_it
            means[n], vars_ = vo_ref.energy_vo_update(K, f, g['in_G'][it, n], g['in_PREC'][it, n], means[n], Vs, T)
            assert rel_err(means[n], g['out_mean'][it, n]) < 1e-12
            assert rel_err(vars_, g['out_vars'][it, n]) < 1e-13


def test_elbo_restatement_matches_reference_generative_model():
    """oracle/elbo_ref.py (the caller of the hot path, generative.py:247-287, 352-392, 456-500) with the oracle's operator
    and virtual observables reproduces the ELBO and all 15 parameter gradients of the UNMODIFIED GenerativeModel run
    (tests/golden/elbo_4x4_16_ndp.npz).  The GPU drop-in test swaps the mirrored modules into the same restatement."""
    from oracle import elbo_ref, fem_p1, vo_ref
    import elbo_fixture
    G = load_golden(elbo_fixture.NAME)
    dt, dev = torch.double, torch.device("cpu")
    P = fem_p1.build_problem(int(G['nx']), int(G['nx']), int(G['refines']))
    g = elbo_ref.OracleOperator(torch.tensor(P['M']), torch.tensor(P['bc_dofs_rom']), torch.tensor(P['W']), dt, dev)
    Gam, alp = [], []
    for n in range(G['in_vo_X_DG'].shape[0]):
        K, f = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], np.exp(G['in_vo_X_DG'][n]), P['bc_dofs_fom'],
                                           G['in_vo_g_fom'][n], P['free_dofs_fom'])
        Ga, al = vo_ref.construct_querry_weak_galerkin(K, f, P['W'])
        Gam.append(Ga); alp.append(al)
    VO = elbo_ref.OracleVO(Gam, alp, torch.zeros(P['W'].shape[1], dtype=dt), G['in_vo_G'], G['in_vo_PREC'], dt)
    assert rel_err(VO.mean, G['out_vo_mean']) < 1e-9 and rel_err(VO.logsigma, G['out_vo_logsigma']) < 1e-8
    value, grads = elbo_fixture.run_elbo(g, VO, G, dt, dev)
    elbo_fixture.compare_with_reference(value, grads, G, 1e-9)
