"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
coarse-grained model, the same torch ops in the same order.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.

Pinned against the reference itself: tests/golden/make_golden.py imports
the reference tree's bottleneck/ROM.py + components.py under stub ``dolfin`` (oracle/ref_shim.py)
and stores its outputs; tests/test_oracle_golden.py checks this restatement against them.
The constants fed to it (M, W, dof sets) come from oracle/fem_p1.py, whose FEniCS
boundary is "parity unpinned" (see its header).
"""
import numpy as np
import torch


def get_stiffness(M, x, bc_dofs, dirichlet=True):
    """K[n,n,B] = M @ x^T, Dirichlet rows -> identity rows (bottleneck/ROM.py:91-100)."""
    K = torch.matmul(M, x.t())
    if dirichlet:
        K[bc_dofs] = 0
        K[bc_dofs, bc_dofs] = 1
    return K


def rom_call(M, bc_dofs, X, F, return_stiffness=False):
    """u[B,n] = A(X)^-1 F  (bottleneck/ROM.py:65-88; torch.solve == batched LU, :59-62)."""
    if F.dim() < 3:
        F = F.unsqueeze(2)
    trunc = 1e-12
    if (X <= trunc).any().item():  # bottleneck/ROM.py:74-76
        raise ValueError('At least one of the conductivity values supplied to the ROM was smaller than {}'.format(trunc))
    K = get_stiffness(M, X, bc_dofs, True)
    y = torch.linalg.solve(K.permute(2, 0, 1), F).squeeze(2)
    if return_stiffness:
        return y, K
    return y


def operator_forward_mean(M, bc_dofs, W, effprop, F):
    """mu_y[B,d] = W @ rom(exp(effprop)+1e-8, F)  (bottleneck/components.py:296-302)."""
    return torch.einsum('sk,nk->ns', [W, rom_call(M, bc_dofs, torch.exp(effprop) + 1e-8, F)])


def rom_fwd_adjoint(M, bc_dofs, logX, F, gbar_u):
    """One 'CGM fwd+adjoint solve' per sample through autograd, exactly what the reference
    does implicitly: returns u, dL/dlogX, dL/dF for L = <gbar_u, u>."""
    logX = logX.detach().clone().requires_grad_(True)
    F = F.detach().clone().requires_grad_(True)
    u = rom_call(M, bc_dofs, torch.exp(logX) + 1e-8, F)
    u.backward(gbar_u)
    return u.detach(), logX.grad.detach(), F.grad.detach()


def rom_fwd_adjoint_closed_form(M, bc_dofs, free_dofs, logX, F, gbar_u):
    """Same quantities from the closed-form SPD statement, numpy fp64,
    sample by sample -- an independent check of the autograd route above."""
    M = np.asarray(M); logX = np.asarray(logX); F = np.asarray(F); gbar = np.asarray(gbar_u)
    B, E = logX.shape
    n = M.shape[0]
    fr, bc = np.asarray(free_dofs), np.asarray(bc_dofs)
    u = np.zeros((B, n)); gX = np.zeros((B, E)); gF = np.zeros((B, n))
    for b in range(B):
        x = np.exp(logX[b]) + 1e-8
        K = M @ x
        Kff, Kfc = K[np.ix_(fr, fr)], K[np.ix_(fr, bc)]
        g = F[b, bc]
        u[b, bc] = g
        u[b, fr] = np.linalg.solve(Kff, F[b, fr] - Kfc @ g)
        lam = np.zeros(n)
        lam[fr] = np.linalg.solve(Kff, gbar[b, fr])
        lam[bc] = gbar[b, bc] - Kfc.T @ lam[fr]
        gF[b] = lam
        lam_f = np.zeros(n); lam_f[fr] = lam[fr]
        gx = -np.einsum('i,ije,j->e', lam_f, M, u[b])
        gX[b] = gx * np.exp(logX[b])
    return u, gX, gF
