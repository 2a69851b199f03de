"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
virtual-observable arithmetic (bottleneck/VirtualObservables.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.

Pinned against the reference itself through tests/golden/ (make_golden.py runs the
reference's QuerryPoint / LinearQuerry / VirtualObservable / VirtualObservablesEnsemble
classes under stub ``dolfin``).  The fine system (K, f) fed to it comes from
oracle/fem_p1.py ("parity unpinned" at the FEniCS boundary).
"""
import numpy as np
import torch


def construct_querry_weak_galerkin(K, f, V):
    """Gamma = V^T K (dense m x d), alpha = V^T f   (VirtualObservables.py:61-69)."""
    assert V.shape[0] == K.shape[0]
    assert V.shape[0] == f.shape[0]
    Gamma = V.T @ K
    alpha = V.T @ f
    return np.asarray(Gamma), np.asarray(alpha)


def vo_residual(K, f, V, y):
    """r = Gamma y - alpha = V^T (K_ff y - f_eff)   (VirtualObservables.py:662, 990)."""
    Gamma, alpha = construct_querry_weak_galerkin(K, f, V)
    return Gamma @ y - alpha


def vo_residual_batch(Ks, fs, V, Y):
    """Reference route, one data point at a time (Python loop as VirtualObservables.py:895, 985)."""
    return np.stack([vo_residual(K, f, V, y) for K, f, y in zip(Ks, fs, Y)])


def vo_residual_transposed(K, V, s):
    """q = Gamma^T s = K_ff (V s)   (the transposed application at VirtualObservables.py:663)."""
    return K.T @ (V @ s)


def virtual_observable_update(Gamma, alpha, vo_variances, g, prec):
    """Gaussian conditioning of N(g, diag(1/prec)) on Gamma y = alpha (+ noise vo_variances)
    -- VirtualObservable.update, VirtualObservables.py:642-669, same torch ops.
    Gamma [m,d], alpha [m], vo_variances [m], g [d], prec [d]  ->  mean [d], vars [d]."""
    Gamma = torch.as_tensor(Gamma, dtype=torch.double)
    alpha = torch.as_tensor(alpha, dtype=torch.double)
    g = torch.as_tensor(g, dtype=torch.double)
    prec = torch.as_tensor(prec, dtype=torch.double)
    vo_variances = torch.as_tensor(vo_variances, dtype=torch.double)
    GT = Gamma  # the reference's "_GammaTransposed" is Gamma.t().t() == Gamma  (:653-654)
    cov = 1 / prec
    Lambda = torch.einsum('im, m, sm -> is', [GT, cov, GT])
    Lambda = Lambda + torch.diag(vo_variances)
    L = torch.linalg.cholesky(Lambda)
    LambdaInv = torch.cholesky_inverse(L)
    solvec = LambdaInv @ (GT @ g - alpha)
    mean = g - torch.einsum('i, mi, m -> i', [cov, GT, solvec])
    A = GT * cov
    sub = torch.einsum('si, sm, mi -> i', [A, LambdaInv, A])
    return mean, cov - sub


def mean_vo_variances(prec_beta, prec_alpha, infinite_precision_mask):
    """beta/(alpha+1), zero where the precision is infinite (VirtualObservables.py:962-966)."""
    mv = prec_beta / (prec_alpha + 1)
    mv = mv.clone()
    mv[infinite_precision_mask] = 0
    return mv


def update_vo_precision_beta(Gammas, alphas, means, varss, beta_0=1e-6):
    """prec_beta = 0.5 * sum_n [(Gamma_n mean_n - alpha_n)^2 + Gamma_n^2 vars_n] + beta_0
    (VirtualObservables.py:981-992)."""
    m = Gammas[0].shape[0]
    beta = torch.zeros(m, dtype=torch.double)
    for Gamma, alpha, mean, vars_ in zip(Gammas, alphas, means, varss):
        Gamma = torch.as_tensor(Gamma, dtype=torch.double)
        beta = beta + (Gamma @ torch.as_tensor(mean) - torch.as_tensor(alpha)) ** 2 \
            + (Gamma ** 2 @ torch.as_tensor(vars_))
    return 0.5 * beta + beta_0


class CsrAssembler(object):
    """Vectorised restatement of ``assemble_system`` for timing the CPU route fairly: the sparsity
    pattern of K_ff / K_fc is fixed by the mesh, so K.data = S @ a with a precomputed sparse S
    (what FEniCS' assemble does per data point, physics/LinearElliptic.py:144-157, without the
    form-compiler overhead).  Checked against fem_p1.assemble_system_free in the tests."""

    def __init__(self, coords, cells, bc_dofs, free_dofs, Ke=None):
        import scipy.sparse as sp
        from . import fem_p1
        N, E = len(coords), len(cells)
        if Ke is None:
            Ke = fem_p1.element_stiffness_all(coords, cells)
        rows = np.repeat(cells, 3, axis=1).ravel()
        cols = np.tile(cells, (1, 3)).ravel()
        cell = np.repeat(np.arange(E), 9)
        vals = Ke.reshape(-1)
        keep = vals != 0
        rows, cols, cell, vals = rows[keep], cols[keep], cell[keep], vals[keep]
        pattern = sp.coo_matrix((np.ones_like(vals), (rows, cols)), shape=(N, N)).tocsr()
        pattern.sum_duplicates()
        pattern.sort_indices()
        # position of every (row, col) contribution inside pattern.data
        pos = np.empty(len(rows), dtype=np.int64)
        indptr, indices = pattern.indptr, pattern.indices
        for k in range(len(rows)):
            lo, hi = indptr[rows[k]], indptr[rows[k] + 1]
            pos[k] = lo + np.searchsorted(indices[lo:hi], cols[k])
        self._S = sp.coo_matrix((vals, (pos, cell)), shape=(pattern.nnz, E)).tocsr()
        self._pattern = pattern
        self._free, self._bc = np.asarray(free_dofs), np.asarray(bc_dofs)
        self._sp = sp

    def assemble(self, a_cell, g):
        K = self._sp.csr_matrix((self._S @ a_cell, self._pattern.indices, self._pattern.indptr),
                                shape=self._pattern.shape)
        Kf = K[self._free]
        return Kf[:, self._free], -(Kf[:, self._bc] @ g)


def energy_vo_update(K, f, g, prec, mean, V_list, temperature):
    """EnergyVirtualObservable.update (VirtualObservables.py:769-788), numpy, same order of operations:
        vars = 1 / (prec + K_diag / T);  A = diag(prec) + K / T;  b = f / T + prec * g
        for V in V_list:  mean -= V (V^T A V)^-1 V^T (A mean - b)
    K: scipy sparse [d,d] (free dofs), f, g, prec, mean: [d]; V_list: one [d,m] weighting matrix per iteration.
    Returns (mean, vars)."""
    inv_temperature = 1 / temperature
    vars_ = 1 / (prec + inv_temperature * K.diagonal())
    A = np.diag(prec) + inv_temperature * K
    b = inv_temperature * f + prec * g
    mean = np.array(mean, dtype=np.float64)
    for V in V_list:
        M = np.array(V.T @ A @ V)
        mean = mean - V @ np.linalg.solve(M, V.T @ np.array(A @ mean - b).flatten())
    return mean, vars_
