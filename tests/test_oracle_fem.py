"""Known-answer tests that pin the restated P1 assembler (oracle/fem_p1.py) -- the part of the oracle
that stands in for FEniCS, for which the reference holds no vectors."""
import numpy as np
import torch

from oracle import fem_p1, rom_ref, vo_ref
from conftest import rel_err


def test_right_isosceles_element_matrix():
    # right angle at the middle vertex -> 1/2 [[1,-1,0],[-1,2,-1],[0,-1,1]], independent of h
    for h in (1.0, 0.25, 1 / 64):
        Ke = fem_p1.p1_element_stiffness(np.array([[0, 0], [h, 0], [h, h]], dtype=float))
        assert np.array_equal(Ke, 0.5 * np.array([[1, -1, 0], [-1, 2, -1], [0, -1, 1.0]]))


def test_element_tensor_properties():
    P = fem_p1.build_problem(4, 4, 1)
    M = P['M']
    assert M.shape == (25, 25, 32)
    assert np.abs(M.sum(axis=1)).max() < 1e-14          # constants are in the kernel of every K_e
    assert np.array_equal(M, M.transpose(1, 0, 2))
    L = M.sum(axis=2)                                    # 5-point Laplacian at an interior vertex
    i = 2 * 5 + 2
    assert L[i, i] == 4.0 and L[i, i - 1] == -1.0 and L[i, i + 5] == -1.0 and L[i, i + 6] == 0.0
    assert ((M != 0).sum(axis=(0, 1)) == 7).all()        # 3x3 block minus the two hypotenuse zeros


def test_dof_sets_and_prolongation():
    P = fem_p1.build_problem(4, 4, 3)
    assert len(P['bc_dofs_rom']) == 10 and len(P['free_dofs_rom']) == 15
    assert len(P['bc_dofs_fom']) == 66 and len(P['free_dofs_fom']) == 1023
    W = P['W']
    assert W.shape == (1023, 25)
    assert np.abs(W.sum(axis=1) - 1).max() < 1e-13
    assert ((W != 0).sum(axis=1) <= 3).all()
    # W reproduces linear functions exactly
    for f in (lambda p: p[:, 0], lambda p: 2 * p[:, 1] - 0.3 * p[:, 0] + 1):
        assert np.abs(W @ f(P['coords_rom']) - f(P['coords_fom'][P['free_dofs_fom']])).max() < 1e-13


def test_uniform_medium_gives_u_equal_x():
    P = fem_p1.build_problem(8, 8, 0)
    _, g, _ = fem_p1.dirichlet_left_right(P['coords_rom'], 'ND')
    F = torch.tensor(fem_p1.full_F_with_applied_bc(81, P['bc_dofs_rom'], g))
    X = torch.full((1, 128), 3.7, dtype=torch.double)
    u = rom_ref.rom_call(torch.tensor(P['M']), torch.tensor(P['bc_dofs_rom']), X, F)
    assert np.abs(u[0].numpy() - P['coords_rom'][:, 0]).max() < 1e-14


def test_layered_medium_gives_harmonic_profile():
    # conductivity constant in y, piecewise constant in x: exact P1 solution u(x_k) = R(x_k)/R(1), R = int 1/a
    nx = 8
    P = fem_p1.build_problem(nx, nx, 0)
    a_col = np.array([1.0, 5.0, 0.2, 2.0, 9.0, 0.7, 1.3, 4.0])
    mid = P['coords_rom'][P['cells_rom']].mean(axis=1)
    x = a_col[np.minimum((mid[:, 0] * nx).astype(int), nx - 1)]
    _, g, _ = fem_p1.dirichlet_left_right(P['coords_rom'], 'ND')
    F = torch.tensor(fem_p1.full_F_with_applied_bc(81, P['bc_dofs_rom'], g))
    u = rom_ref.rom_call(torch.tensor(P['M']), torch.tensor(P['bc_dofs_rom']), torch.tensor(x[None]), F)[0].numpy()
    R = np.concatenate([[0], np.cumsum(1 / a_col)]) / np.sum(1 / a_col)
    assert np.abs(u - R[np.round(P['coords_rom'][:, 0] * nx).astype(int)]).max() < 1e-13


def test_exact_fom_solution_has_zero_residual_and_galerkin_identity():
    P = fem_p1.build_problem(2, 2, 2)
    rng = np.random.RandomState(0)
    a = np.exp(rng.normal(size=len(P['cells_fom'])))
    _, g, _ = fem_p1.dirichlet_left_right(P['coords_fom'], 'NDP', rng.uniform(-.5, .5, 4))
    K, f = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], a, P['bc_dofs_fom'], g, P['free_dofs_fom'])
    y = np.linalg.solve(K.toarray(), f)
    V = np.hstack([P['W'], rng.normal(size=(K.shape[0], 3))])
    assert np.abs(vo_ref.vo_residual(K, f, V, y)).max() < 1e-12
    Gamma, _ = vo_ref.construct_querry_weak_galerkin(K, f, P['W'])
    assert rel_err(Gamma @ P['W'], P['W'].T @ K.toarray() @ P['W']) < 1e-14
    # transposed application
    s = rng.normal(size=V.shape[1])
    G2, _ = vo_ref.construct_querry_weak_galerkin(K, f, V)
    assert rel_err(vo_ref.vo_residual_transposed(K, V, s), G2.T @ s) < 1e-13


def test_pixel_map_two_cells_per_pixel_top_row_first():
    P = fem_p1.build_problem(2, 2, 1)
    pix = P['pixel_of_cell_fom']
    assert np.array_equal(np.bincount(pix), np.full(16, 2))
    top_left = [c for c in range(len(pix)) if pix[c] == 0]
    mid = P['coords_fom'][P['cells_fom'][top_left]].mean(axis=1)
    assert (mid[:, 0] < 0.25).all() and (mid[:, 1] > 0.75).all()


def test_alternating_diagonal_mesh_is_consistent():
    c, cells = fem_p1.unit_square_mesh(4, 4, 'alternating')
    K = fem_p1.assemble_fom_csr(c, cells, np.ones(len(cells))).toarray()
    assert np.abs(K.sum(axis=1)).max() < 1e-14
    area = sum(0.5 * abs(np.linalg.det(np.c_[c[cl], np.ones(3)])) for cl in cells)
    assert abs(area - 1) < 1e-14
