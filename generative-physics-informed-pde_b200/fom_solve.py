"""Fine-mesh label solves on the device: K_ff(a_b) y_b = f_f - K_fc(a_b) g_b for a whole data set at once.

Replaces the per-sample FEniCS solve of ``LinearEllipticPhysics.solve`` (physics/LinearElliptic.py:85-101; ``solve_direct``
:120-133 is the same system through scipy's spsolve) that ``DataLoader.assemble`` loops over (utils/data.py:96-99) -- the
reference's data preparation is that serial loop.  Batched Jacobi-preconditioned conjugate gradients, matrix-free:

  * A p = K_ff(a) p for ALL samples = one launch of the structured-grid marching kernel (``VoPlan.residual`` with y = p,
    no Dirichlet data, load ignored: its rho output);
  * everything else of an iteration = one launch of ``cg_step_kernel`` (csrc/fom_cg.cu, one CTA per sample);
  * right-hand side f_eff = -(K_fom(a) (0, g) - f)_free = minus the rho output for y = 0.
Host code only sequences launches; convergence is checked every ``check_every`` iterations (one device->host read).
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .VirtualObservables import VoPlan

_F64 = torch.float64


def _diagonal_operator(physics, plan, device):
    """Sparse D [d, n_inputs] with diag(K_ff(a)) = D a: the l-th diagonal entry of every cell's element matrix scattered to
    its node, columns mapped through the plan's cell -> input map (per pixel or per cell)."""
    key = ("diag", str(device), plan.n_inputs)
    cache = physics.__dict__.setdefault("_fom_solve_cache", {})
    if key not in cache:
        mesh = physics.mesh
        Ke = mesh.element_stiffness()
        pos = np.full(mesh.num_nodes, -1, dtype=np.int64)
        pos[physics.free_dofs] = np.arange(physics.free_dofs.size)
        inputs = plan.cell_to_input if plan.cell_to_input is not None else np.arange(mesh.num_cells)
        rows = pos[mesh.cells].ravel()
        cols = np.repeat(inputs, 3)
        vals = np.stack([Ke[:, 0, 0], Ke[:, 1, 1], Ke[:, 2, 2]], 1).ravel()
        keep = rows >= 0
        D = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(physics.free_dofs.size, plan.n_inputs)).tocsr().tocoo()
        idx = torch.tensor(np.stack([D.row, D.col]), device=device)
        cache[key] = torch.sparse_coo_tensor(idx, torch.tensor(D.data, device=device), D.shape, check_invariants=False).coalesce()
    return cache[key]


@torch.no_grad()
def solve_batched(physics, a, g, *, plan=None, tol=1e-12, max_iter=None, check_every=25, x0=None, return_info=False):
    """y [B,d] with K_ff(a_b) y_b = f_eff,b to a relative residual ``tol`` (|r| <= tol |f_eff|, per sample).

    physics : the fine LinearEllipticPhysics;  a [B, n_inputs] float64 CUDA tensor of CONDUCTIVITIES (per pixel when the
    plan is a pixel-input plan, else per cell);  g [B, n_bc] or [n_bc] Dirichlet values;  plan : VoPlan (default: the
    cached pixel-input plan when ``a`` has one value per pixel, else the per-cell plan)."""
    device = _lib.require_cuda(a.device, "fom_solve.solve_batched")
    lib = _lib.load()
    mesh = physics.mesh
    if plan is None:
        plan = VoPlan.cached(physics, device, pixel_input=(a.shape[1] == mesh.nx * mesh.ny and a.shape[1] != mesh.num_cells))
    a = a.to(_F64).contiguous()
    g = g.to(device=device, dtype=_F64).contiguous()
    B, d = int(a.shape[0]), int(plan.d)
    if (a <= 0).any():            # LinearElliptic.py:76-79
        raise ValueError('Trying to set negative or zero material values')
    dinv = (1.0 / torch.sparse.mm(_diagonal_operator(physics, plan, device), a.t().contiguous()).t()).contiguous()
    _, rho0 = plan.residual(a, None, g, None, a_is_log=False)
    rhs = rho0.neg_()
    x = torch.zeros(B, d, dtype=_F64, device=device) if x0 is None else x0.to(_F64).clone().contiguous()
    Ax = None
    if x0 is not None:
        _, Ax = plan.residual(a, x, None, None, a_is_log=False, ignore_load=True)
    r, p = torch.empty_like(x), torch.empty_like(x)
    rz, stop2, rn2 = (torch.empty(B, dtype=_F64, device=device) for _ in range(3))
    st, dev_index = _lib.stream_of(device), device.index
    P = lambda t: _lib.ptr(t, device)
    _lib.check(lib.gpde_cg_init_f64(P(rhs), P(Ax), P(dinv), d, P(r), P(p), P(rz), P(stop2), P(rn2), float(tol), d, B, dev_index, st),
               "gpde_cg_init_f64")
    max_iter = int(max_iter if max_iter is not None else 20 * int(np.sqrt(d)) + 200)
    it, done = 0, False
    while it < max_iter and not done:
        for _ in range(min(check_every, max_iter - it)):
            _, Ap = plan.residual(a, p, None, None, a_is_log=False, ignore_load=True)
            _lib.check(lib.gpde_cg_step_f64(P(Ap), P(dinv), d, P(x), P(r), P(p), P(rz), P(stop2), P(rn2), d, B, dev_index, st),
                       "gpde_cg_step_f64")
            it += 1
        done = bool((rn2 <= stop2).all().item())
    if return_info:
        rel = torch.sqrt(rn2 / (stop2 / (tol * tol) if tol > 0 else torch.ones_like(stop2)))
        return x, dict(iterations=it, converged=done, max_rel_residual=float(rel.max().item()))
    if not done:
        raise RuntimeError("fom_solve: %d CG iterations did not reach tol = %g" % (it, tol))
    return x
