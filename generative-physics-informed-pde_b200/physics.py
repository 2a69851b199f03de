"""FEniCS-free stand-ins for the setup-time objects the hot path is constructed from.

Mirrors (names, argument meaning) of
  physics/LinearElliptic.py:8-171          LinearEllipticPhysics
  physics/BoundaryConditions.py:8-147      BoundaryConditionEnsemble (the parts the hot path reads)
  factories/model.py:106-142               ModelFactory._setup  -> ``setup_physics``
Everything here is host/numpy and runs once; nothing here is on the hot path.
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import fem


class DirichletData(object):
    """One sample's boundary condition: what QuerryPoint / assemble_system read from ``bc``
    (physics/BoundaryConditions.py:150-260): dof sets and values per function space identifier."""

    def __init__(self, ensemble, index):
        self._ensemble = ensemble
        self._index = index

    def constrained_dofs(self, identifier):
        return self._ensemble.constrained_dofs(identifier)

    def free_dofs(self, identifier):
        return self._ensemble.free_dofs(identifier)

    def constrained_dofs_values(self, identifier):
        return self._ensemble.constrained_dofs_values(identifier)[self._index]

    def assemble_vanilla_force_vector(self, identifier):
        return self._ensemble.load(identifier)


class BoundaryConditionEnsemble(object):
    """N boundary conditions on the 'rom' and 'fom' meshes (physics/BoundaryConditions.py:8-147).

    kind 'ND' : u=0 left, u=1 right for every sample          (LinearEllipticFactories.py:173-179)
    kind 'NDP': u0(1-y)+u1 y left, u2(1-y)+u3 y right, coefficients [N,4]   (:239-281)
    """

    def __init__(self, physics, N, kind="ND", coefficients=None, rng=None):
        self._meshes = {"rom": physics["rom"].mesh, "fom": physics["fom"].mesh}
        self.kind = kind.upper()
        self._N = int(N)
        if self.kind == "NDP" and coefficients is None:
            rng = rng if rng is not None else np.random
            coefficients = rng.uniform(-0.5, 0.5, size=(N, 4))   # LinearEllipticFactories.py:241-251
        self.coefficients = None if coefficients is None else np.asarray(coefficients, dtype=np.float64)
        self._values = {}

    def __len__(self):
        return self._N

    def __getitem__(self, key):
        if isinstance(key, list):
            return [DirichletData(self, k) for k in key]
        return DirichletData(self, key)

    def constrained_dofs(self, identifier):
        return self._meshes[identifier.lower()].dirichlet_dofs()[0]

    def free_dofs(self, identifier):
        return self._meshes[identifier.lower()].dirichlet_dofs()[1]

    def load(self, identifier):
        return np.zeros(self._meshes[identifier.lower()].num_nodes)   # zero source/Neumann (:165-171)

    def constrained_dofs_values(self, identifier):
        identifier = identifier.lower()
        if identifier not in self._values:
            mesh = self._meshes[identifier]
            if self.kind == "ND":
                v = np.tile(mesh.dirichlet_values("ND")[None], (self._N, 1))
            else:
                v = np.atleast_2d(mesh.dirichlet_values("NDP", self.coefficients))
            self._values[identifier] = v
        return self._values[identifier]

    def FULL_F_WITH_APPLIED_BC(self, identifier):
        """F[N, V.dim()] (physics/BoundaryConditions.py:132-147)."""
        identifier = identifier.lower()
        mesh = self._meshes[identifier]
        return fem.full_F_with_applied_bc(mesh.num_nodes, self.constrained_dofs(identifier),
                                          self.constrained_dofs_values(identifier), self.load(identifier))


class LinearEllipticPhysics(object):
    """a(u,v) = int alpha grad(u).grad(v) dx on a P1 mesh with DG0 conductivity
    (physics/LinearElliptic.py:8-171, physics/LinearEllipticFactories.py:151-160)."""

    def __init__(self, identifier, physics_id, mesh):
        if physics_id.upper() not in ("ND", "NDP"):
            raise NotImplementedError
        self.identifier = identifier
        self.ptype = physics_id.upper()
        self.mesh = mesh
        self._constrained_dofs, self._free_dofs = mesh.dirichlet_dofs()

    @property
    def free_dofs(self):
        return self._free_dofs

    @property
    def constrained_dofs(self):
        return self._constrained_dofs

    @property
    def tdim(self):
        return 2

    @property
    def dim_in(self):
        return self.mesh.num_cells

    @property
    def dim_out(self):
        return self._free_dofs.size

    @property
    def dim_out_all(self):
        return self.mesh.num_nodes

    def assemble_system(self, x, bc, *, only_free_dofs=True):
        """Host CSR (K, f) exactly as LinearElliptic.py:137-159 -- kept for API parity and for the
        label solver; the device path never forms K."""
        x = np.asarray(x)
        if np.any(x <= 0):   # LinearElliptic.py:76-79
            raise ValueError('Trying to set negative or zero material values')
        K = self.mesh.assemble_csr(x)
        f = bc.assemble_vanilla_force_vector(self.identifier)
        if not only_free_dofs:
            return K, f
        cd, fd = bc.constrained_dofs(self.identifier), bc.free_dofs(self.identifier)
        vals = bc.constrained_dofs_values(self.identifier)
        f_eff = f[fd] - K[fd, :][:, cd].dot(vals)
        return K[fd][:, fd], f_eff

    def solve_direct(self, x, bc, only_free_dofs=True):
        """Sparse direct FOM solve (LinearElliptic.py:120-133); used to make labels y."""
        K, f = self.assemble_system(x, bc)
        y_sub = spla.spsolve(K.tocsc(), f)
        if only_free_dofs:
            return y_sub
        y = np.zeros(self.mesh.num_nodes)
        y[bc.constrained_dofs(self.identifier)] = bc.constrained_dofs_values(self.identifier)
        y[bc.free_dofs(self.identifier)] = y_sub
        return y

    solve = solve_direct

    def scatter_restricted_solution(self, y, bc):
        out = np.zeros(self.dim_out_all)
        out[bc.constrained_dofs('fom')] = bc.constrained_dofs_values('fom')
        out[bc.free_dofs('fom')] = y
        return out


def setup_physics(nx_rom, ny_rom, num_refines, ptype="ND", diagonal="right"):
    """physics dict {'rom','fom','W'} as ModelFactory._setup builds it (factories/model.py:130-140)."""
    mesh_rom = fem.P1Mesh(nx_rom, ny_rom, diagonal)
    mesh_fom = mesh_rom.refine(num_refines)
    physics = dict()
    physics['fom'] = LinearEllipticPhysics('fom', ptype, mesh_fom)
    physics['rom'] = LinearEllipticPhysics('rom', ptype, mesh_rom)
    physics['W'] = fem.prolongation(mesh_rom, mesh_fom, physics['fom'].free_dofs)   # [d, n]
    return physics
