"""Setup-time exporter: everything the reference asks FEniCS/DOLFIN for, as numpy arrays.

Runs ONCE at setup on the host (north_star: "FEniCS ... runs only once at setup, to export
the element stiffness tensors, load vectors, Dirichlet maps and weighting matrices").
The run-time hot path (ROM.py / VirtualObservables.py here) only sees the exported arrays.

Reference call sites this replaces:
  mesh            factories/model.py:130-133, fawkes/utils.py:9-14
  bilinear form   physics/LinearEllipticFactories.py:151-160, 209-219  (alpha grad u . grad v dx)
  Dirichlet maps  LinearEllipticFactories.py:173-179 (ND), 239-281 (NDP); fawkes/BoundaryConditions.py:131-146
  M               bottleneck/ROM.py:46-55
  W               bottleneck/components.py:38-60, fawkes/utils.py:115-192, factories/model.py:140
  F_ROM_BC        physics/BoundaryConditions.py:132-147
  K_fom, f_eff    physics/LinearElliptic.py:137-159
  pixel <-> cell  bottleneck/utils.py:41-98, 115-132
  random field    physics/RandomField.py:61-73, 162-209; factories/data.py:88, 99

Conventions: vertex id = iy*(nx+1)+ix (x fastest, y up), P1 dof = vertex id, DG0 dof = cell id,
square s = iy*nx+ix holds cells 2s, 2s+1.  DOLFIN's own dof numbering is a permutation of
this one; every quantity on the hot path is equivariant to it.
"""
import numpy as np
import scipy.sparse as sp


class P1Mesh(object):
    """Triangulated unit square, P1 (nodes) / DG0 (cells)."""

    def __init__(self, nx, ny, diagonal="right"):
        self.nx, self.ny, self.diagonal = int(nx), int(ny), diagonal
        ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1))
        self.coords = np.stack([ix.ravel() / nx, iy.ravel() / ny], axis=1)
        sx, sy = np.meshgrid(np.arange(nx), np.arange(ny))
        v00 = (sy * (nx + 1) + sx).ravel()
        v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
        if diagonal == "right":
            right = np.ones(v00.shape, dtype=bool)
        elif diagonal == "left":
            right = np.zeros(v00.shape, dtype=bool)
        elif diagonal == "alternating":
            right = ((sx + sy) % 2 == 0).ravel()
        else:
            raise ValueError("unknown diagonal pattern %r" % (diagonal,))
        lo = np.where(right[:, None], np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
        hi = np.where(right[:, None], np.stack([v00, v01, v11], 1), np.stack([v10, v01, v11], 1))
        cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
        cells[0::2], cells[1::2] = lo, hi
        self.cells = cells

    @property
    def num_nodes(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    def refine(self, num_refines):
        """Uniformly refined mesh (each refinement halves h); fawkes/utils.py:9-14."""
        f = 2 ** int(num_refines)
        return P1Mesh(self.nx * f, self.ny * f, self.diagonal)

    # -- element matrices ------------------------------------------------------------
    def element_stiffness(self):
        """Ke[E,3,3]: unit-conductivity P1 stiffness of every cell."""
        p = self.coords[self.cells]                       # [E,3,2]
        x, y = p[:, :, 0], p[:, :, 1]
        det = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
        gx = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], 1) / det[:, None]
        gy = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], 1) / det[:, None]
        area = 0.5 * np.abs(det)
        return area[:, None, None] * (gx[:, :, None] * gx[:, None, :] + gy[:, :, None] * gy[:, None, :])

    def dense_element_tensor(self):
        """M[n,n,E] as ROM.FromPhysics builds it (bottleneck/ROM.py:46-53)."""
        if self.num_cells > 290:  # bottleneck/ROM.py:43-44
            raise Exception('ROM exceeds intended maximum size')
        n, E = self.num_nodes, self.num_cells
        M = np.zeros((n, n, E))
        Ke = self.element_stiffness()
        e = np.arange(E)
        for a in range(3):
            for b in range(3):
                np.add.at(M, (self.cells[:, a], self.cells[:, b], e), Ke[:, a, b])
        return M

    def assemble_csr(self, a_cell):
        """K(a) = sum_c a_c K_c over all nodes, scipy CSR."""
        Ke = self.element_stiffness() * np.asarray(a_cell)[:, None, None]
        rows = np.repeat(self.cells, 3, axis=1).ravel()
        cols = np.tile(self.cells, (1, 3)).ravel()
        N = self.num_nodes
        return sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(N, N)).tocsr()

    # -- Dirichlet maps ----------------------------------------------------------------
    def dirichlet_dofs(self):
        """(constrained dofs ascending, free dofs ascending): Dirichlet on x=0 and x=1."""
        ixs = np.arange(self.num_nodes) % (self.nx + 1)
        on = (ixs == 0) | (ixs == self.nx)
        return np.nonzero(on)[0].astype(np.int64), np.nonzero(~on)[0].astype(np.int64)

    def dirichlet_values(self, kind="ND", u=None):
        """Values at the constrained dofs; ``u`` = [B,4] NDP coefficients (or [4])."""
        bc, _ = self.dirichlet_dofs()
        y = self.coords[bc, 1]
        left = (self.coords[bc, 0] < 0.5)
        if kind.upper() == "ND":
            return np.where(left, 0.0, 1.0)
        if kind.upper() != "NDP":
            raise NotImplementedError(kind)
        u = np.atleast_2d(np.asarray(u, dtype=np.float64))
        vl = u[:, 0:1] * (1 - y)[None] + u[:, 1:2] * y[None]
        vr = u[:, 2:3] * (1 - y)[None] + u[:, 3:4] * y[None]
        out = np.where(left[None], vl, vr)
        return out if out.shape[0] > 1 else out[0]

    # -- images --------------------------------------------------------------------------
    def pixel_of_cell(self):
        """Pixel id (image row 0 = top) of every cell; two cells per pixel."""
        mid = self.coords[self.cells].mean(axis=1)
        cx = np.minimum((mid[:, 0] * self.nx).astype(np.int64), self.nx - 1)
        cy = (self.ny - 1) - np.minimum((mid[:, 1] * self.ny).astype(np.int64), self.ny - 1)
        pix = cy * self.nx + cx
        assert np.all(np.bincount(pix, minlength=self.nx * self.ny) == 2)
        return pix


def prolongation(coarse, fine, fine_rows):
    """W[len(fine_rows), n_coarse]: coarse P1 basis functions evaluated at fine nodes."""
    pts = fine.coords[fine_rows]
    W = np.zeros((len(fine_rows), coarse.num_nodes))
    done = np.zeros(len(fine_rows), dtype=bool)
    p = coarse.coords[coarse.cells]
    x, y = p[:, :, 0], p[:, :, 1]
    det = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    for c in range(coarse.num_cells):
        todo = np.nonzero(~done)[0]
        if todo.size == 0:
            break
        q = pts[todo]
        l1 = ((q[:, 0] - x[c, 0]) * (y[c, 2] - y[c, 0]) - (x[c, 2] - x[c, 0]) * (q[:, 1] - y[c, 0])) / det[c]
        l2 = ((x[c, 1] - x[c, 0]) * (q[:, 1] - y[c, 0]) - (q[:, 0] - x[c, 0]) * (y[c, 1] - y[c, 0])) / det[c]
        l0 = 1.0 - l1 - l2
        inside = np.minimum(np.minimum(l0, l1), l2) >= -1e-12
        rows = todo[inside]
        W[rows[:, None], coarse.cells[c][None, :]] = np.stack([l0, l1, l2], 1)[inside]
        done[rows] = True
    if not done.all():
        raise Exception('No collision with mesh for requested point')
    return W


def full_F_with_applied_bc(n, bc_dofs, bc_values, load=None):
    """F[B,n] = load (zero by default) with the Dirichlet values written at bc_dofs."""
    bc_values = np.atleast_2d(bc_values)
    F = np.zeros((bc_values.shape[0], n)) if load is None else np.tile(np.asarray(load, float), (bc_values.shape[0], 1))
    F[:, bc_dofs] = bc_values
    return F


def sample_log_field(py, px, mean, stddev, corrlength, batch, rng, dtype=np.float64):
    """Gaussian field on pixel centres, covariance stddev^2 exp(-r^2/(2 l^2)) + 1e-12 I.

    The kernel is separable, C = Cy (x) Cx, so it is sampled through the two small Cholesky
    factors instead of the reference's dense (py*px)^2 factor (capped at 8192 dofs,
    physics/RandomField.py:43-44).  Same distribution, different random stream."""
    pwx, pwy = 1.0 / px, 1.0 / py
    x = np.linspace(0.5 * pwx, 1 - 0.5 * pwx, px)
    y = np.linspace(0.5 * pwx, 1 - 0.5 * pwy, py)   # sic: the reference starts y at 0.5*pixelwidth_x
    Cx = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / corrlength ** 2)
    Cy = np.exp(-0.5 * (y[:, None] - y[None, :]) ** 2 / corrlength ** 2)

    def factor(C):
        w, Q = np.linalg.eigh(C)
        return Q * np.sqrt(np.clip(w, 0.0, None))[None, :]
    Lx, Ly = factor(Cx), factor(Cy)
    out = np.empty((batch, py, px), dtype=dtype)
    chunk = max(1, (1 << 24) // (py * px))
    for b0 in range(0, batch, chunk):
        g = rng.standard_normal((min(chunk, batch - b0), py, px))
        out[b0:b0 + g.shape[0]] = mean + stddev * np.einsum('ia,bac,jc->bij', Ly, g, Lx, optimize=True)
    return out


def coarse_cell_average(images, coarse, fine):
    """Area average of a per-pixel field over every coarse cell: a deterministic stand-in for
    the decoder output X [B,E] in the synthetic workloads."""
    B = images.shape[0]
    pix = fine.pixel_of_cell()
    mid = fine.coords[fine.cells].mean(axis=1)
    # coarse cell containing each fine cell midpoint
    sx = np.minimum((mid[:, 0] * coarse.nx).astype(np.int64), coarse.nx - 1)
    sy = np.minimum((mid[:, 1] * coarse.ny).astype(np.int64), coarse.ny - 1)
    fx, fy = mid[:, 0] * coarse.nx - sx, mid[:, 1] * coarse.ny - sy
    if coarse.diagonal != "right":
        raise NotImplementedError
    upper = fy > fx
    owner = 2 * (sy * coarse.nx + sx) + upper.astype(np.int64)
    vals = images.reshape(B, -1)[:, pix]
    out = np.zeros((B, coarse.num_cells), dtype=images.dtype)
    cnt = np.bincount(owner, minlength=coarse.num_cells)
    for e in range(coarse.num_cells):
        out[:, e] = vals[:, owner == e].mean(axis=1)
    assert cnt.min() > 0
    return out


def rbf_weighting(fine, rows, centres, l):
    """exp(-|x-r0|^2/l^2) at the fine nodes ``rows`` for every centre: V[len(rows), len(centres)]."""
    p = fine.coords[rows]
    c = np.asarray(centres, dtype=np.float64)
    d2 = ((p[:, None, :] - c[None, :, :]) ** 2).sum(-1)
    return np.exp(-d2 / l ** 2)
