"""The product's vectorised setup exporter (generative-physics-informed-pde_b200/fem.py, physics.py) against
the slow, obviously-correct oracle assembler (oracle/fem_p1.py).  Host-side logic, no GPU."""
import numpy as np
import pytest

import gpde_b200  # noqa: F401
from gpde_b200 import fem, physics
from oracle import fem_p1
from conftest import rel_err


@pytest.mark.parametrize("nx,ny,diag", [(4, 4, "right"), (3, 5, "right"), (4, 4, "alternating"), (2, 3, "left")])
def test_mesh_and_element_matrices(nx, ny, diag):
    mesh = fem.P1Mesh(nx, ny, diag)
    c, cells = fem_p1.unit_square_mesh(nx, ny, diag)
    assert np.array_equal(mesh.coords, c) and np.array_equal(mesh.cells, cells)
    assert np.allclose(mesh.element_stiffness(), fem_p1.element_stiffness_all(c, cells), rtol=0, atol=1e-15)
    a = np.exp(np.random.RandomState(0).normal(size=len(cells)))
    assert abs(mesh.assemble_csr(a) - fem_p1.assemble_fom_csr(c, cells, a)).max() < 1e-13


@pytest.mark.parametrize("nx,refines", [(4, 3), (8, 1), (2, 2)])
def test_physics_bundle_matches_oracle(nx, refines):
    P = fem_p1.build_problem(nx, nx, refines)
    ph = physics.setup_physics(nx, nx, refines, "ND")
    assert np.array_equal(ph['rom'].mesh.dense_element_tensor(), P['M'])
    assert np.allclose(ph['W'], P['W'], rtol=0, atol=1e-15)
    assert np.array_equal(ph['rom'].constrained_dofs, P['bc_dofs_rom'])
    assert np.array_equal(ph['rom'].free_dofs, P['free_dofs_rom'])
    assert np.array_equal(ph['fom'].constrained_dofs, P['bc_dofs_fom'])
    assert np.array_equal(ph['fom'].free_dofs, P['free_dofs_fom'])
    assert np.array_equal(ph['fom'].mesh.pixel_of_cell(), P['pixel_of_cell_fom'])


def test_boundary_ensembles_and_F():
    ph = physics.setup_physics(4, 4, 2, "NDP")
    rng = np.random.RandomState(3)
    bce = physics.BoundaryConditionEnsemble(ph, 5, "NDP", rng=rng)
    F = bce.FULL_F_WITH_APPLIED_BC('rom')
    P = fem_p1.build_problem(4, 4, 2)
    for b in range(5):
        dofs, vals, _ = fem_p1.dirichlet_left_right(P['coords_rom'], "NDP", bce.coefficients[b])
        assert np.array_equal(dofs, bce.constrained_dofs('rom'))
        assert np.allclose(F[b, dofs], vals, atol=1e-15) and np.abs(np.delete(F[b], dofs)).max() == 0
        _, vf, _ = fem_p1.dirichlet_left_right(P['coords_fom'], "NDP", bce.coefficients[b])
        assert np.allclose(bce[b].constrained_dofs_values('fom'), vf, atol=1e-15)
    nd = physics.BoundaryConditionEnsemble(ph, 2, "ND")
    _, v, _ = fem_p1.dirichlet_left_right(P['coords_rom'], "ND")
    assert np.array_equal(nd.constrained_dofs_values('rom')[1], v)


def test_assemble_system_and_direct_solve():
    ph = physics.setup_physics(2, 2, 2, "NDP")
    rng = np.random.RandomState(1)
    bce = physics.BoundaryConditionEnsemble(ph, 1, "NDP", rng=rng)
    x = np.exp(rng.normal(size=ph['fom'].dim_in))
    K, f = ph['fom'].assemble_system(x, bce[0])
    P = fem_p1.build_problem(2, 2, 2)
    K0, f0 = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], x, P['bc_dofs_fom'],
                                         bce[0].constrained_dofs_values('fom'), P['free_dofs_fom'])
    assert abs(K - K0).max() < 1e-13 and np.abs(f - f0).max() < 1e-13
    y = ph['fom'].solve_direct(x, bce[0])
    assert np.abs(K0 @ y - f0).max() < 1e-12
    with pytest.raises(ValueError):
        ph['fom'].assemble_system(-x, bce[0])


def test_random_field_statistics_and_cell_average():
    rng = np.random.RandomState(0)
    img = fem.sample_log_field(32, 32, 0.4, 0.8, 0.15, 4096, rng)
    assert abs(img.mean() - 0.4) < 0.03 and abs(img.std() - 0.8) < 0.03
    # neighbouring-pixel correlation of the squared-exponential kernel: exp(-0.5 (h/l)^2)
    h = 1 / 32
    c = np.mean((img[:, :, 1:] - 0.4) * (img[:, :, :-1] - 0.4)) / 0.64
    assert abs(c - np.exp(-0.5 * (h / 0.15) ** 2)) < 0.02
    coarse, fine = fem.P1Mesh(4, 4), fem.P1Mesh(32, 32)
    avg = fem.coarse_cell_average(np.full((2, 32, 32), 1.5), coarse, fine)
    assert avg.shape == (2, 32) and np.allclose(avg, 1.5)
    # oracle definition of the same average (tests/golden/make_golden.py) on a random image
    X_DG = img[:3].reshape(3, -1)[:, fine.pixel_of_cell()]
    mid = fine.coords[fine.cells].mean(axis=1)
    sx, sy = np.minimum((mid[:, 0] * 4).astype(int), 3), np.minimum((mid[:, 1] * 4).astype(int), 3)
    owner = 2 * (sy * 4 + sx) + ((mid[:, 1] * 4 - sy) > (mid[:, 0] * 4 - sx))
    ref = np.stack([X_DG[:, owner == e].mean(axis=1) for e in range(32)], axis=1)
    assert rel_err(fem.coarse_cell_average(img[:3], coarse, fine), ref) < 1e-14


def test_rom_size_guard():
    with pytest.raises(Exception):
        fem.P1Mesh(13, 13).dense_element_tensor()   # 338 cells > 290 (bottleneck/ROM.py:43-44)


@pytest.mark.parametrize("nx,l", [(8, 0.1), (32, 0.25)])
def test_rbf_weighting_matches_oracle_columns(nx, l):
    """fem.rbf_weighting (vectorised) against oracle/fem_p1.rbf_columns (VirtualObservables.py:184-198,
    fawkes/Expressions.py:26-31: exp(-|x - r0|^2 / l^2) interpolated at the fine free nodes)."""
    mesh = fem.P1Mesh(nx, nx)
    ph = physics.LinearEllipticPhysics('fom', 'ND', mesh)
    centres = np.random.RandomState(4).uniform(size=(7, 2))
    c, _ = fem_p1.unit_square_mesh(nx, nx)
    _, _, free = fem_p1.dirichlet_left_right(c, "ND")
    V = fem.rbf_weighting(mesh, ph.free_dofs, centres, l)
    V0 = fem_p1.rbf_columns(c, free, centres, l)
    assert V.shape == V0.shape == (ph.dim_out, 7)
    assert np.abs(V - V0).max() < 1e-15
