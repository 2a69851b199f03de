"""Synthetic workload of BASELINE.json config 5: one semi-supervised SVI ELBO step AROUND the physics layer.

The caller's side of the hot path -- CNN encoder / decoder, the effective-property map, the per-sample variational tables
and the ELBO (bottleneck/generative.py:247-287, 352-392, 456-500, 546-585) -- is NOT part of this package's product
(the encoder/decoder, ELBO and SVI loop stay the reference's PyTorch code).  bench.py and the tests need
a caller of the right shape to time and exercise the data-parallel step, so this module holds a small stand-in written
against the same formulas:

    supervised   Z ~ q_z, X_s ~ q_X;  logL_x(f(Z), X) + logL_X(X_s | gp(Z)) + logL_y(Y | g(X_s, F)) + H[q_X] - KL[q_z]
    VO           the same with Y replaced by a draw of the virtual-observable posterior N(VO.mean, exp(2 VO.logsigma))
    unsupervised (amortised)  z ~ encoder(X);  logL_x(f(z), X) - KL

``g`` is this package's ReducedOrderModelOperator (the sm_100a kernels): either the drop-in ``g(X_s, F)`` + the diagonal
Gaussian log-likelihood in torch, or the fused epilogue ``g.log_likelihood`` (no [B,d] intermediate).  Random-init weights,
synthetic fields; sizes follow the reference's ``highres32`` preset (N_s = 128, bs_u = 64,
N_vo <= 128 per replica; 32 x 32 images, 4 x 4 coarse mesh, latent dimension 16).  No BatchNorm in the stand-in CNNs (with
the reference's CNNs: SyncBatchNorm, or per-shard statistics)."""
import math

import numpy as np
import torch
from torch import nn

from .components import ReducedOrderModelOperator
from .workloads import Workload

LOG2PI = math.log(2.0 * math.pi)


def gaussian_loglik(target, mean, logsigma):
    """sum log N(target | mean, exp(2 logsigma)) (bottleneck/utils.py:231-241 with logvars = 2 logsigma)."""
    return torch.sum(-0.5 * (2.0 * logsigma + ((target - mean) * torch.exp(-logsigma)) ** 2 + LOG2PI))


def reparametrize(mean, logsigma):
    return mean + torch.exp(logsigma) * torch.randn_like(mean)


class Decoder(nn.Module):
    """f: z [B,L] -> (mean, logsigma) of the fine log-field image, flattened [B, py*px]."""

    def __init__(self, dim_latent, py, px):
        super().__init__()
        assert py % 4 == 0 and px % 4 == 0
        self.dim_latent, self.py, self.px = dim_latent, py, px
        self.fc = nn.Linear(dim_latent, 8 * (py // 4) * (px // 4))
        self.up1 = nn.ConvTranspose2d(8, 4, 4, stride=2, padding=1)
        self.up2 = nn.ConvTranspose2d(4, 1, 4, stride=2, padding=1)
        self.logsigma = nn.Parameter(torch.full((1,), -0.5))

    def forward(self, z):
        h = torch.relu(self.fc(z)).view(z.shape[0], 8, self.py // 4, self.px // 4)
        mean = self.up2(torch.relu(self.up1(h))).reshape(z.shape[0], -1)
        return mean, self.logsigma.expand_as(mean)


class Encoder(nn.Module):
    """x [B, py*px] -> (mean, logsigma) of q(z | x)."""

    def __init__(self, dim_latent, py, px):
        super().__init__()
        self.py, self.px = py, px
        self.c1 = nn.Conv2d(1, 4, 4, stride=2, padding=1)
        self.c2 = nn.Conv2d(4, 8, 4, stride=2, padding=1)
        self.fc = nn.Linear(8 * (py // 4) * (px // 4), 2 * dim_latent)

    def forward(self, x):
        h = torch.relu(self.c2(torch.relu(self.c1(x.view(-1, 1, self.py, self.px)))))
        mean, logsigma = self.fc(h.flatten(1)).chunk(2, dim=1)
        return mean, logsigma


class EffectivePropertyMap(nn.Module):
    """gp: z -> (mean, logsigma) of the coarse log-conductivities (bottleneck/components.py:205-232, independent X)."""

    def __init__(self, dim_latent, E):
        super().__init__()
        self.fc = nn.Linear(dim_latent, E)
        self.logsigmas_X = nn.Parameter(torch.zeros(E))

    def forward(self, z):
        return self.fc(z), self.logsigmas_X.expand(z.shape[0], -1)


class Tables(nn.Module):
    """q(v_n) = N(mean_n, exp(2 logsigma_n)), one row per OWNED data point (bottleneck/components.py:70-201)."""

    def __init__(self, N, dim, init_logsigma=-1.0):
        super().__init__()
        self.mean = nn.Parameter(torch.zeros(N, dim))
        self.logsigma = nn.Parameter(torch.full((N, dim), float(init_logsigma)))

    def sample(self):
        return reparametrize(self.mean, self.logsigma)

    def kld(self):
        return -0.5 * torch.sum(1 + 2 * self.logsigma - self.mean.pow(2) - torch.exp(2 * self.logsigma))

    def entropy(self):
        return torch.sum(self.logsigma) + self.mean.shape[0] * self.mean.shape[1] * 0.5 * (LOG2PI + 1)


class SviWorkload(object):
    """This rank's shard of the config-5 step: data, the stand-in caller modules, the mirrored physics layer."""

    def __init__(self, device, dtype=torch.float32, N_s=128, N_vo=128, bs_u=64, dim_latent=16, seed=0, fused_loglik=True):
        self.device, self.dtype, self.fused = device, dtype, bool(fused_loglik)
        self.N_s, self.N_vo, self.bs_u, self.L = int(N_s), int(N_vo), int(bs_u), int(dim_latent)
        w = Workload("cfg1", B=self.N_s + self.N_vo + self.bs_u, seed=seed)
        self.w, ph = w, w.physics
        fom = ph["fom"]
        py, px = fom.mesh.ny, fom.mesh.nx
        T = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=dtype, device=device)
        s0, s1, s2 = self.N_s, self.N_s + self.N_vo, self.N_s + self.N_vo + self.bs_u
        self.sup = dict(X=T(w.log_image[:s0]), Y=T(w.y[:s0]), F=T(w.F[:s0]))
        self.vo = dict(X=T(w.log_image[s0:s1]), F=T(w.F[s0:s1]))
        self.unsup = dict(X=T(w.log_image[s1:s2]))
        torch.manual_seed(1234)                          # same shared initial weights on every rank (broadcast anyway)
        self.g = ReducedOrderModelOperator.FromPhysics(ph, dtype=dtype, device=device)
        self.g.rom.deferred_checks = True                # no host synchronisation inside the step
        self.f = Decoder(self.L, py, px).to(device=device, dtype=dtype)
        self.encoder = Encoder(self.L, py, px).to(device=device, dtype=dtype)
        self.gp = EffectivePropertyMap(self.L, w.E).to(device=device, dtype=dtype)
        mk = lambda N, dim: Tables(N, dim).to(device=device, dtype=dtype)
        self.q_z = dict(supervised=mk(self.N_s, self.L), vo=mk(self.N_vo, self.L))
        self.q_X = dict(supervised=mk(self.N_s, w.E), vo=mk(self.N_vo, w.E))
        # virtual-observable posterior of the owned VO data points (constant during a step; generative.py:311, 356)
        self.VO = None
        self.vo_mean = T(w.y[s0:s1])
        self.vo_logsigma = torch.full_like(self.vo_mean, -3.0)

    # ---- parameters
    def shared_parameters(self):
        ps = list(self.f.parameters()) + list(self.encoder.parameters()) + list(self.gp.parameters()) + [self.g.logsigmas_y]
        return ps

    def local_parameters(self):
        return [p for t in (self.q_z, self.q_X) for k in ("supervised", "vo") for p in t[k].parameters()]

    # ---- virtual observables of the owned VO data points (setup + the periodic update, generative.py:182-222)
    def build_virtual_observables(self):
        from . import VirtualObservables as VOm
        from .physics import BoundaryConditionEnsemble
        w, ph = self.w, self.w.physics
        s0, s1 = self.N_s, self.N_s + self.N_vo
        coef = None if w.bce.coefficients is None else w.bce.coefficients[s0:s1]
        bce = BoundaryConditionEnsemble(ph, self.N_vo, w.ptype, coefficients=coef)
        pix = ph["fom"].mesh.pixel_of_cell()
        qpe = VOm.QuerryPointEnsemble.FromArrays(np.ascontiguousarray(w.log_image[s0:s1][:, pix]), bce, ph["fom"], device=self.device)
        qe = VOm.QuerryEnsemble.FromQuerryPointEnsemble(qpe, ph, True, False, 0, 0, dtype=torch.double, device=self.device)
        self.VO = VOm.VirtualObservablesEnsemble(qpe, qe, self.dtype, self.device)
        return self.VO

    @torch.no_grad()
    def update_virtual_observables(self, N_mc=64, step=0):
        """generative.py:182-222 with the loops over data points replaced by batched launches: Monte-Carlo predictive
        moments of the operator for all VO data points, then the batched posterior update."""
        X_s = self.q_X["vo"].mean.unsqueeze(1) + torch.exp(self.q_X["vo"].logsigma).unsqueeze(1) * torch.randn(
            self.N_vo, N_mc, self.w.E, dtype=self.dtype, device=self.device)
        Y_mean, Y_std = self.g.predictive_moments(X_s, self.vo["F"])
        self.VO.resample()
        self.VO.update(Y_mean, 1.0 / Y_std ** 2, step)
        self.vo_mean.copy_(self.VO.mean)
        self.vo_logsigma.copy_(self.VO.logsigma)

    # ---- the ELBO of this rank's data points
    def _physics_loglik(self, X_s, F, Y):
        if self.fused:
            return self.g.log_likelihood(X_s, F, Y)
        mu_y, ls_y = self.g(X_s, F)
        return gaussian_loglik(Y, mu_y, ls_y)

    def _data_term(self, key, X_img, F, Y):
        Z, X_s = self.q_z[key].sample(), self.q_X[key].sample()
        logL_x = gaussian_loglik(X_img, *self.f(Z))
        logL_X = gaussian_loglik(X_s, *self.gp(Z))
        logL_y = self._physics_loglik(X_s, F, Y)
        return logL_x + logL_X + logL_y + self.q_X[key].entropy() - self.q_z[key].kld()

    def elbo(self):
        total = self._data_term("supervised", self.sup["X"], self.sup["F"], self.sup["Y"])
        y_vo = reparametrize(self.vo_mean, self.vo_logsigma)
        total = total + self._data_term("vo", self.vo["X"], self.vo["F"], y_vo)
        mean, logsigma = self.encoder(self.unsup["X"])
        z = reparametrize(mean, logsigma)
        kld = -0.5 * torch.sum(1 + 2 * logsigma - mean.pow(2) - torch.exp(2 * logsigma))
        return total + gaussian_loglik(self.unsup["X"], *self.f(z)) - kld

    def samples_per_step(self):
        return self.N_s + self.N_vo + self.bs_u

    def cgm_solves_per_step(self):
        return self.N_s + self.N_vo
