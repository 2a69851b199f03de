"""Synthetic workloads of BASELINE.json's configs, host-side numpy.

Used by bench.py, __graft_entry__.smoke() and the GPU tests to build identical, seeded inputs:
log-normal random fields (factories/data.py:88,99 presets), their coarse cell averages as the
decoder stand-in, ND / NDP boundary data, an adjoint seed, and the weighting matrix V.
"""
import numpy as np

from . import fem
from .physics import BoundaryConditionEnsemble, setup_physics

CONFIGS = {
    # name: nx_rom, num_refines, ptype, corr. length, batch, number of RBF columns appended to V=W
    "cfg1": dict(nx=4, refines=3, ptype="NDP", ell=0.15, B=64, n_rbf=0,
                 desc="example.ipynb/highres32: 4x4 CGM, 32x32 FOM, batch 64, V=W (m=25)"),
    "cfg2": dict(nx=4, refines=4, ptype="ND", ell=0.04, B=4096, n_rbf=0,
                 desc="4x4 CGM, 64x64 FOM, batch 4096 log-normal fields, V=W (m=25)"),
    "cfg4": dict(nx=4, refines=4, ptype="ND", ell=0.04, B=131072, n_rbf=0,
                 desc="config 2's mesh, batch 131072 sample-sharded over the GPUs (strong scaling), V=W (m=25)"),
    "cfg3": dict(nx=8, refines=4, ptype="ND", ell=0.04, B=16384, n_rbf=175,
                 desc="8x8 CGM, 128x128 FOM, 256 weighting functions (81 CGR + 175 RBF), batch 16384"),
}


class Workload(object):
    """Host arrays of one workload.  Conductivity input of the VO path is the log-field image,
    flattened row-major with image row 0 = top (pixel input, P = E_f/2 values per sample)."""

    def __init__(self, name, B=None, seed=0, ptype=None):
        cfg = dict(CONFIGS[name])
        self.name, self.cfg = name, cfg
        self.B = int(B if B is not None else cfg["B"])
        self.ptype = ptype or cfg["ptype"]
        rng = np.random.RandomState(seed)
        self.physics = setup_physics(cfg["nx"], cfg["nx"], cfg["refines"], self.ptype)
        rom, fom = self.physics["rom"], self.physics["fom"]
        self.n, self.E = rom.mesh.num_nodes, rom.mesh.num_cells
        self.d, self.P = fom.dim_out, fom.mesh.nx * fom.mesh.ny
        self.n_bc_fom = fom.constrained_dofs.size
        img = fem.sample_log_field(fom.mesh.ny, fom.mesh.nx, 0.4, 0.8, cfg["ell"], self.B, rng)
        self.log_image = img.reshape(self.B, -1)                                 # [B,P]
        self.logX = fem.coarse_cell_average(img, rom.mesh, fom.mesh)             # [B,E]
        self.bce = BoundaryConditionEnsemble(self.physics, self.B, self.ptype, rng=rng)
        self.F = self.bce.FULL_F_WITH_APPLIED_BC("rom")                          # [B,n]
        self.g_fom = self.bce.constrained_dofs_values("fom")                     # [B,n_bc_f]
        self.gbar_u = rng.standard_normal((self.B, self.n))
        W = self.physics["W"]
        if cfg["n_rbf"]:
            centres = rng.uniform(size=(cfg["n_rbf"], 2))
            self.V = np.hstack([W, fem.rbf_weighting(fom.mesh, fom.free_dofs, centres, 0.1)])
        else:
            self.V = W
        self.m = self.V.shape[1]
        # a solution-like fine field: the linear-in-x profile between the Dirichlet data plus noise
        xf = fom.mesh.coords[fom.free_dofs, 0]
        left = self.g_fom[:, fom.mesh.coords[fom.constrained_dofs, 0] < 0.5].mean(axis=1)
        right = self.g_fom[:, fom.mesh.coords[fom.constrained_dofs, 0] > 0.5].mean(axis=1)
        self.y = left[:, None] * (1 - xf)[None] + right[:, None] * xf[None] + 0.01 * rng.standard_normal((self.B, self.d))

    # algorithmic bytes per unit, s = bytes per scalar
    def cgm_bytes_per_solve(self, s=8):
        return s * (2 * self.E + 3 * self.n)

    def vo_bytes_per_eval(self, s=8):
        return s * (self.P + self.d + self.n_bc_fom + self.m)

    def describe(self):
        return dict(workload=self.name, desc=self.cfg["desc"], batch=self.B, coarse="%dx%d" % ((self.cfg["nx"],) * 2),
                    fine="%dx%d" % ((self.physics["fom"].mesh.nx,) * 2), n=self.n, E=self.E, d=self.d, P=self.P,
                    m=self.m, bc=self.ptype)
