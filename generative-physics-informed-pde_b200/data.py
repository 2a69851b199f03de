"""Data preparation without FEniCS: images -> DG0 fields -> fine-mesh labels (utils/data.py:8-119, the assembly part of the
reference's ``DataLoader``; its partition / chunk bookkeeping is caller-side control logic and stays with the caller).

    X [N,py,px]  log-conductivity images (float64, CPU, as the reference requires, utils/data.py:12-30)
    X_DG [N,E_f] the same fields as DG0 functions, two cells per pixel (bottleneck/utils.py:115-132)
    Y [N,d]      fine-mesh solutions on the free dofs for a = exp(X_DG) and the sample's Dirichlet data (utils/data.py:96-99)
    F_ROM_BC [N,n]  coarse load vector with the Dirichlet values written in (physics/BoundaryConditions.py:132-147)

The reference solves the N fine systems one after the other with FEniCS; ``assemble(..., device=cuda)`` solves all of them
at once with the batched matrix-free conjugate gradients of ``fom_solve`` (B200 kernels); ``device=None`` keeps the
reference's serial sparse direct solves on the host (``LinearEllipticPhysics.solve``), for machines without a GPU."""
import numpy as np
import torch

from .physics import BoundaryConditionEnsemble


class DataLoader(object):
    def __init__(self, X, X_DG=None, Y=None, BCE=None, F_ROM_BC=None, hash=None):
        if X.dtype != torch.double:
            raise ValueError
        if BCE is not None and len(BCE) != X.shape[0]:
            raise ValueError
        if X_DG is not None:
            assert X.shape[0] == X_DG.shape[0]
        if Y is not None:
            assert Y.shape[0] == X.shape[0]
        if F_ROM_BC is not None:
            assert F_ROM_BC.shape[0] == X.shape[0]
        if X.device != torch.device('cpu'):
            raise ValueError
        self._X, self._BCE, self._X_DG, self._Y, self._F_ROM_BC = X, BCE, X_DG, Y, F_ROM_BC
        self._hash = hash
        self._lock_physics_assembly = False
        self.solve_info = None

    def lock_physics_assembly(self):
        self._lock_physics_assembly = True

    N = property(lambda self: self._X.shape[0])

    def __len__(self):
        return self._X.shape[0]

    def assemble_BCE(self, physics):
        self._BCE = BoundaryConditionEnsemble(physics, self.N, physics['fom'].ptype)

    def assemble_DG(self, physics):
        """X_DG[n, c] = X[n] at the pixel of cell c (DiscontinuousGalerkinPixelConverter.ImageToFunctionBatchedFast)."""
        mesh = physics['fom'].mesh
        assert mesh.num_cells == 2 * int(np.prod(self._X.shape[1:]))
        pix = torch.as_tensor(mesh.pixel_of_cell())
        self._X_DG = self._X.reshape(self.N, -1)[:, pix].contiguous()

    def assemble(self, physics, BCE=None, *, device=None, tol=1e-12):
        if self._lock_physics_assembly:
            raise RuntimeError
        if self._X.dim() != 3:
            raise ValueError
        if self._BCE is None:
            if BCE is not None:
                assert isinstance(BCE, BoundaryConditionEnsemble) and len(BCE) == self.N
                self._BCE = BCE
            else:
                self.assemble_BCE(physics)
        fom = physics['fom']
        self.assemble_DG(physics)
        if device is not None and torch.device(device).type == "cuda":
            from . import fom_solve
            a = torch.exp(self._X.reshape(self.N, -1).to(device))            # per pixel: the structured-grid kernels
            g = torch.as_tensor(self._BCE.constrained_dofs_values('fom'), device=device)
            Y, self.solve_info = fom_solve.solve_batched(fom, a, g, tol=tol, return_info=True)
            if not self.solve_info["converged"]:
                raise RuntimeError("label solve did not converge: %r" % (self.solve_info,))
            self._Y = Y.cpu()
        else:
            self._Y = torch.zeros(self.N, fom.dim_out, dtype=torch.double)
            for n in range(self.N):
                matprop = np.exp(self._X_DG[n, :].numpy().flatten())
                self._Y[n, :] = torch.tensor(fom.solve(x=matprop, bc=self._BCE[n]), dtype=torch.double)
        self._F_ROM_BC = torch.tensor(self._BCE.FULL_F_WITH_APPLIED_BC('rom'), dtype=torch.double)

    def _assembled(name):
        def get(self):
            value = getattr(self, name)
            if value is None:
                raise RuntimeError('Assembly has not yet been called on this dataset')
            return value
        return property(get)

    X = property(lambda self: self._X)
    X_DG = _assembled("_X_DG")
    Y = _assembled("_Y")
    F_ROM_BC = _assembled("_F_ROM_BC")
    BCE = property(lambda self: self._BCE)
    del _assembled

    @classmethod
    def FromSampler(cls, sampler, N):
        """N draws of ``sampler.sample()`` -> images [N,py,px] (utils/data.py:313-325)."""
        first = np.asarray(sampler.sample())
        X = torch.zeros(N, first.shape[0], first.shape[1], dtype=torch.double)
        X[0, :] = torch.tensor(first, dtype=torch.double)
        for n in range(1, N):
            X[n, :] = torch.tensor(np.asarray(sampler.sample()), dtype=torch.double)
        return cls(X=X)

    def __repr__(self):
        return 'DataLoader with {} random field realizations ({},{}) [Assembled = {}]'.format(
            self._X.shape[0], self._X.shape[1], self._X.shape[2], self._X_DG is not None)
