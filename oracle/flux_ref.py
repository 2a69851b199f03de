"""ORACLE (test infrastructure only -- see oracle/__init__ and DESIGN.md section 2): loop restatement of the reference's
flux-balance constraints, bottleneck/flux.py:8-158, for P1 / DG0 on the oracle's own meshes (oracle/fem_p1.py).

PARITY UNPINNED at the FEniCS boundary: DOLFIN/UFL 2018.1.0 are not installable here and the reference holds no golden
vectors for this path.  The restatement follows the reference line by line where it is Python (facet / cell marking,
choice of measure, reduction to the free dofs, the zero alpha) and the published semantics of the three FEniCS pieces it
calls:
  * ``dot(alpha*grad(u), n)*ds(id)``           exterior facets with that marker; the only adjacent cell, outward normal;
  * ``dot(alpha('+')*grad(u)('+'), n('+'))*dS(id)``  interior facets with that marker; '+' = DOLFIN's first cell of the
    facet, swapped when the form carries cell markers and the other cell's marker is larger (Assembler.cpp,
    assemble_interior_facets) -- the zero ``dx(subdomain_data=cellfct)`` term of flux.py:34-36 supplies those markers;
  * ``assemble(derivative(form, u))``          the coefficient vector of the linear functional.
Anchors (tests/test_flux.py): divergence theorem on closed coarse cells for linear fields, sign and size of the flux of
u = x through single edges, linearity in alpha, and agreement with the product's vectorised implementation.
"""
import numpy as np


def _facets(cells):
    """(facet -> sorted vertex pair, facet -> adjacent cells in ascending cell order, cell -> its 3 facets)."""
    index, verts, adj, of_cell = {}, [], [], []
    for c, tri in enumerate(cells):
        mine = []
        for l in range(3):
            pair = tuple(sorted((int(tri[(l + 1) % 3]), int(tri[(l + 2) % 3]))))
            if pair not in index:
                index[pair] = len(verts)
                verts.append(pair)
                adj.append([])
            adj[index[pair]].append(c)
            mine.append(index[pair])
        of_cell.append(mine)
    return verts, adj, of_cell


def _contains(tri, q):
    A = np.array([[tri[1][0] - tri[0][0], tri[2][0] - tri[0][0]], [tri[1][1] - tri[0][1], tri[2][1] - tri[0][1]]])
    l12 = np.linalg.solve(A, np.asarray(q) - tri[0])
    return min(1.0 - l12.sum(), l12[0], l12[1]) >= -1e-12


def _hat_gradients(tri):
    """Rows = grad(phi_k) of the three P1 hat functions on the triangle (phi_k(x_j) = delta_kj)."""
    A = np.hstack([np.ones((3, 1)), np.asarray(tri)])       # phi_k(x) = c0 + c1 x + c2 y
    coef = np.linalg.solve(A, np.eye(3))                      # column k = coefficients of phi_k
    return coef[1:, :].T


def flux_gamma(coords_c, cells_c, coords_f, cells_f, x):
    """Gamma[V.dim(), N]: column n = d/du of the flux form of coarse cell n (flux.py:84-139) for fine conductivities x."""
    verts_c, adj_c, of_cell_c = _facets(cells_c)
    verts_f, adj_f, _ = _facets(cells_f)
    mid_f = [0.5 * (coords_f[a] + coords_f[b]) for a, b in verts_f]
    cellmid_f = coords_f[cells_f].mean(axis=1)
    Gamma = np.zeros((coords_f.shape[0], len(cells_c)))
    for n, rom_cell in enumerate(cells_c):
        tri_c = coords_c[rom_cell]
        cellfct = [1 if _contains(tri_c, cellmid_f[c]) else 0 for c in range(len(cells_f))]     # flux.py:125-131
        for rom_facet in of_cell_c[n]:
            p0, p1 = coords_c[verts_c[rom_facet][0]], coords_c[verts_c[rom_facet][1]]
            marked = [f for f in range(len(verts_f))                                              # flux.py:106-118
                      if np.linalg.norm(mid_f[f] - p0) + np.linalg.norm(mid_f[f] - p1) - np.linalg.norm(p0 - p1) < 1e-12]
            exterior = len(adj_c[rom_facet]) == 1
            dirichlet = exterior and p0[0] == p1[0] and p0[0] in (0.0, 1.0)     # mark_facets of the left / right Dirichlet sides
            for f in marked:
                cells = adj_f[f]
                if dirichlet:                      # ds: exterior facets only
                    if len(cells) != 1:
                        continue
                    plus = cells[0]
                else:                              # dS: interior facets only
                    if len(cells) != 2:
                        continue
                    plus, minus = cells[0], cells[1]
                    if cellfct[plus] < cellfct[minus]:
                        plus, minus = minus, plus
                tri = coords_f[cells_f[plus]]
                a, b = coords_f[verts_f[f][0]], coords_f[verts_f[f][1]]
                t = b - a
                normal = np.array([t[1], -t[0]]) / np.linalg.norm(t)
                if np.dot(normal, 0.5 * (a + b) - tri.mean(axis=0)) < 0:          # outward of the '+' cell
                    normal = -normal
                grads = _hat_gradients(tri)
                for k in range(3):
                    Gamma[cells_f[plus][k], n] += np.linalg.norm(t) * x[plus] * np.dot(grads[k], normal)
    return Gamma


def flux_reduced(Gamma, constrained, free, values, fix_alpha=False):
    """flux.py:141-158: (Gamma_reduced [N,d], alpha_reduced [N]); the reference reads the Dirichlet part from its
    never-filled ``self.Gamma`` (zeros) -- ``fix_alpha`` uses the assembled matrix instead."""
    N = Gamma.shape[1]
    source = Gamma if fix_alpha else np.zeros_like(Gamma)
    Gamma_reduced = np.zeros((free.size, N))
    alpha_reduced = np.zeros(N)
    for n in range(N):
        Gamma_reduced[:, n] = Gamma[free, n]
        alpha_reduced[n] = np.dot(source[constrained, n], values)
    alpha_reduced = alpha_reduced * (-1)
    return Gamma_reduced.T, alpha_reduced
