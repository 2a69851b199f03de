"""ORACLE (test infrastructure, not product code): FEniCS-free restatement of the
setup-time constants the reference obtains from DOLFIN 2018.1.0.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Nothing under ``generative-physics-informed-pde_b200/`` does.

Parity status: **unpinned at the FEniCS boundary** -- DOLFIN/UFL/FFC/PETSc
(readme.md:11-12, no lock file) are not installable here and the reference holds no
golden vectors, so the arrays below are anchored on closed-form known answers
(tests/test_oracle_fem.py) and on the reference's own pure-torch/numpy arithmetic
executed on top of them (oracle/ref_shim.py -> tests/golden/).

Everything is written as plain per-element numpy loops on purpose: it is the slow,
obviously-correct statement that the vectorised product exporter is checked against.

Conventions (ours; DOLFIN's dof reordering is a permutation the maths is equivariant to):
  * mesh = UnitSquareMesh(nx, ny) with the "right" diagonal (factories/model.py:132),
    vertex id = iy*(nx+1)+ix, coordinates (ix/nx, iy/ny);
  * square s = iy*nx+ix holds cells 2s = (v00, v10, v11) and 2s+1 = (v00, v01, v11);
  * P1 dof = vertex id, DG0 dof = cell id;
  * fine mesh = the coarse mesh refined ``num_refines`` times (factories/model.py:133,
    fawkes/utils.py:9-14).  Regular refinement of a right-diagonal mesh is again a
    right-diagonal mesh; ``diagonal='alternating'`` gives the right/left pattern that
    longest-edge bisection can produce instead (hard parts).
"""
import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------- mesh
def unit_square_mesh(nx, ny, diagonal="right"):
    """Vertices [N,2] and cells [2*nx*ny,3] of df.UnitSquareMesh (factories/model.py:132)."""
    coords = np.zeros(((nx + 1) * (ny + 1), 2))
    for iy in range(ny + 1):
        for ix in range(nx + 1):
            coords[iy * (nx + 1) + ix] = (ix / nx, iy / ny)
    cells = np.zeros((2 * nx * ny, 3), dtype=np.int64)
    for iy in range(ny):
        for ix in range(nx):
            v00 = iy * (nx + 1) + ix
            v10 = v00 + 1
            v01 = v00 + (nx + 1)
            v11 = v01 + 1
            s = iy * nx + ix
            right = True
            if diagonal == "alternating":
                right = ((ix + iy) % 2 == 0)
            elif diagonal == "left":
                right = False
            elif diagonal != "right":
                raise ValueError(diagonal)
            if right:
                cells[2 * s] = (v00, v10, v11)
                cells[2 * s + 1] = (v00, v01, v11)
            else:
                cells[2 * s] = (v00, v10, v01)
                cells[2 * s + 1] = (v10, v01, v11)
    return coords, cells


# ------------------------------------------------------------------ element matrices
def p1_element_stiffness(xy):
    """Unit-conductivity P1 stiffness of one triangle: int grad(phi_i).grad(phi_j) dx.

    Bilinear form ``alpha * inner(grad(u), grad(v)) * dx`` with alpha in DG0
    (physics/LinearEllipticFactories.py:151-160, 209-219) => K_e = alpha_e * this.
    """
    (x0, y0), (x1, y1), (x2, y2) = xy
    det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
    area = 0.5 * abs(det)
    # gradients of the barycentric coordinates
    g = np.array([[y1 - y2, x2 - x1],
                  [y2 - y0, x0 - x2],
                  [y0 - y1, x1 - x0]]) / det
    return area * (g @ g.T)


def element_stiffness_all(coords, cells):
    Ke = np.zeros((len(cells), 3, 3))
    for e, c in enumerate(cells):
        Ke[e] = p1_element_stiffness(coords[c])
    return Ke


def rom_element_tensor(coords, cells):
    """M[n,n,E], M[:,:,e] = assemble(derivative(a, alpha, e_e))  (bottleneck/ROM.py:46-53)."""
    n, E = len(coords), len(cells)
    if E > 290:  # bottleneck/ROM.py:43-44
        raise Exception('ROM exceeds intended maximum size')
    M = np.zeros((n, n, E))
    Ke = element_stiffness_all(coords, cells)
    for e, c in enumerate(cells):
        for a in range(3):
            for b in range(3):
                M[c[a], c[b], e] += Ke[e, a, b]
    return M


# ------------------------------------------------------------------------ Dirichlet
def dirichlet_left_right(coords, kind="ND", u=None):
    """Constrained dofs (ascending, as np.unique in fawkes/BoundaryConditions.py:139-140),
    their values, and the free dofs.

    ND : u=0 on x=0, u=1 on x=1        (physics/LinearEllipticFactories.py:173-179)
    NDP: u0(1-y)+u1 y on x=0, u2(1-y)+u3 y on x=1, u_i ~ U[-.5,.5]   (:239-281)
    """
    tol = 1e-12
    dofs, vals = [], []
    for i, (x, y) in enumerate(coords):
        if abs(x) < tol:
            dofs.append(i)
            vals.append(0.0 if kind == "ND" else u[0] * (1 - y) + u[1] * y)
        elif abs(x - 1.0) < tol:
            dofs.append(i)
            vals.append(1.0 if kind == "ND" else u[2] * (1 - y) + u[3] * y)
    dofs = np.array(dofs, dtype=np.int64)
    vals = np.array(vals)
    free = np.array(sorted(set(range(len(coords))) - set(dofs.tolist())), dtype=np.int64)
    return dofs, vals, free


def full_F_with_applied_bc(n, bc_dofs, bc_vals_batch):
    """F[N,n]: vanilla (zero) load with Dirichlet values written at the constrained dofs
    (physics/BoundaryConditions.py:132-147; zero Neumann/source LinearEllipticFactories.py:165-171)."""
    bc_vals_batch = np.atleast_2d(bc_vals_batch)
    F = np.zeros((bc_vals_batch.shape[0], n))
    for b in range(F.shape[0]):
        F[b, bc_dofs] = bc_vals_batch[b]
    return F


# --------------------------------------------------------------------- interpolation
def prolongation_W(coords_c, cells_c, coords_f, free_f):
    """W[d,n]: W[i,k] = phi_k^rom(x_i^fom) for the fine *free* dofs
    (bottleneck/components.py:38-60, fawkes/utils.py:115-192, factories/model.py:140)."""
    W = np.zeros((len(free_f), len(coords_c)))
    for row, i in enumerate(free_f):
        p = coords_f[i]
        found = False
        for c in cells_c:
            (x0, y0), (x1, y1), (x2, y2) = coords_c[c]
            det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
            l1 = ((p[0] - x0) * (y2 - y0) - (x2 - x0) * (p[1] - y0)) / det
            l2 = ((x1 - x0) * (p[1] - y0) - (p[0] - x0) * (y1 - y0)) / det
            l0 = 1.0 - l1 - l2
            if min(l0, l1, l2) >= -1e-12:
                W[row, c] = (l0, l1, l2)
                found = True
                break
        if not found:
            raise Exception('No collision with mesh for requested point')  # fawkes/utils.py:152-153
    return W


# ------------------------------------------------------------------------ fine system
def assemble_fom_csr(coords_f, cells_f, a_cell):
    """Full fine stiffness K_fom(a) = sum_c a_c K_c as scipy CSR (physics/LinearElliptic.py:144)."""
    N = len(coords_f)
    rows, cols, vals = [], [], []
    for c, dofs in enumerate(cells_f):
        Kc = a_cell[c] * p1_element_stiffness(coords_f[dofs])
        for a in range(3):
            for b in range(3):
                rows.append(dofs[a]); cols.append(dofs[b]); vals.append(Kc[a, b])
    return sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsr()


def assemble_system_free(coords_f, cells_f, a_cell, bc_dofs, bc_vals, free, f_full=None):
    """(K_ff, f_eff) exactly as LinearEllipticPhysics.assemble_system(only_free_dofs=True)
    (physics/LinearElliptic.py:137-159): f_eff = f[free] - K[free,:][:,constrained] @ g."""
    K = assemble_fom_csr(coords_f, cells_f, a_cell)
    if f_full is None:
        f_full = np.zeros(len(coords_f))
    K_coupling = K[free, :][:, bc_dofs]
    f_eff = f_full[free] - K_coupling.dot(bc_vals)
    return K[free][:, free], f_eff


# ------------------------------------------------------------------- pixel <-> cells
def pixel_of_cell(coords, cells, px, py):
    """Pixel id (row 0 = top of the image) of each DG0 cell
    (bottleneck/utils.py:69-80: cy flipped, pixel_id = cy*(Ny-1)+cx, two cells per pixel)."""
    out = np.zeros(len(cells), dtype=np.int64)
    dx, dy = 1.0 / px, 1.0 / py
    for c, dofs in enumerate(cells):
        mx, my = coords[dofs].mean(axis=0)
        cx = int(mx // dx)
        cy = (py - 1) - int(my // dy)
        out[c] = cy * px + cx
    counts = np.bincount(out, minlength=px * py)
    assert np.all(counts == 2)  # utils/data.py:109
    return out


def image_to_function(images, pix_of_cell):
    """X_DG = image.flatten()[pixel_of_cell]  (bottleneck/utils.py:123-129)."""
    flat = images.reshape(images.shape[0], -1)
    return flat[:, pix_of_cell]


# --------------------------------------------------------------------- random fields
def sample_log_field(py, px, mean, stddev, corrlength, batch, rng):
    """Gaussian random field on pixel centres with covariance
    stddev^2 exp(-r^2 / (2 l^2)) + 1e-12 I  (physics/RandomField.py:61-73, 162-172, 205-209;
    presets factories/data.py:88,99).  Dense Cholesky like the reference (<= 8192 dofs, :43-44)."""
    if py * px > 8192:
        raise RuntimeError
    pwx, pwy = 1.0 / px, 1.0 / py
    x = np.linspace(0.5 * pwx, 1 - 0.5 * pwx, px)
    y = np.linspace(0.5 * pwx, 1 - 0.5 * pwy, py)
    Xg, Yg = np.meshgrid(x, y)
    P = np.hstack([Xg.reshape(-1, 1), Yg.reshape(-1, 1)])
    r2 = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    C = stddev ** 2 * np.exp(-0.5 * r2 / corrlength ** 2) + 1e-12 * np.eye(len(P))
    L = np.linalg.cholesky(C)
    out = np.zeros((batch, py, px))
    for b in range(batch):
        out[b] = (mean + L @ rng.normal(0, 1, len(P))).reshape(py, px)
    return out


def rbf_columns(coords_f, free_f, centres, l):
    """RBF weighting functions exp(-|x-r0|^2 / l^2) interpolated at the fine nodes
    (bottleneck/VirtualObservables.py:184-198, fawkes/Expressions.py:26-31)."""
    V = np.zeros((len(free_f), len(centres)))
    for k, r0 in enumerate(centres):
        d2 = ((coords_f[free_f] - np.asarray(r0)) ** 2).sum(1)
        V[:, k] = np.exp(-d2 / l ** 2)
    return V


# ---------------------------------------------------------------------------- bundle
def build_problem(nx_rom, ny_rom, num_refines, diagonal_fom="right"):
    """All setup-time constants for one (coarse, fine) mesh pair, as a dict of numpy arrays."""
    nx_f, ny_f = nx_rom * 2 ** num_refines, ny_rom * 2 ** num_refines
    cc, cells_c = unit_square_mesh(nx_rom, ny_rom)
    cf, cells_f = unit_square_mesh(nx_f, ny_f, diagonal_fom)
    bc_c, _, free_c = dirichlet_left_right(cc, "ND")
    bc_f, _, free_f = dirichlet_left_right(cf, "ND")
    return dict(
        nx_rom=nx_rom, ny_rom=ny_rom, nx_fom=nx_f, ny_fom=ny_f,
        coords_rom=cc, cells_rom=cells_c, coords_fom=cf, cells_fom=cells_f,
        bc_dofs_rom=bc_c, free_dofs_rom=free_c, bc_dofs_fom=bc_f, free_dofs_fom=free_f,
        M=rom_element_tensor(cc, cells_c),
        W=prolongation_W(cc, cells_c, cf, free_f),
        pixel_of_cell_fom=pixel_of_cell(cf, cells_f, nx_f, ny_f),
    )
