"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's ELBO around the physics layer --
the CALLER of the hot path (bottleneck/generative.py), used to prove that the mirrored modules are drop-ins for it.

Only ``tests/`` may import this module (plus tests/golden/make_golden.py, which pins it against the reference itself:
the UNMODIFIED bottleneck/generative.py GenerativeModel is run under the stub-dolfin shim on the same components, inputs
and noise, and its ELBO and parameter gradients are stored in tests/golden/elbo_4x4_16_ndp.npz).

What is restated (same torch ops in the same order; ``writer`` / logging branches dropped):
    reparametrize                          bottleneck/utils.py:215-218
    DiagonalGaussianLogLikelihood          bottleneck/utils.py:231-241
    UnitGaussianKullbackLeiblerDivergence  bottleneck/utils.py:245-247
    VariationalApproximation               bottleneck/components.py:70-201   (sample / KLD / entropy)
    EffectivePropertyMap (independent_X)   bottleneck/components.py:205-232
    GenerativeModel.elbo                   bottleneck/generative.py:247-287  (supervised + vo data sets)
        ._elbo_supervised_freeX            :456-500
        ._elbo_virtual_observables_freeX   :352-392
        .random_field_likelihood           :231-244  (tuple branch, reconstruct_log_eff_property)
The decoder ``f`` of the reference is a CNN (bottleneck/Decoder.py) outside the hot path; the fixtures use TinyDecoder
below in its place on both sides (the reference's GenerativeModel takes any module with ``dim_latent`` that returns
(mean, logsigma)).

All randomness goes through torch.randn_like, as in the reference; tests replay a recorded NoiseTape through it so that
CPU and GPU runs see the same draws.
"""
import numpy as np
import torch

LOG2PI = 1.8378770664093453


def reparametrize(mean, logsigma):
    std = torch.exp(logsigma)
    return mean + std * torch.randn_like(std)


def diagonal_gaussian_log_likelihood(target, mean, logvars):
    sigma = logvars.mul(0.5).exp_()
    part2 = ((target - mean) / sigma) ** 2
    return torch.sum(-0.5 * (logvars + part2 + LOG2PI))


def unit_gaussian_kld(mean, logvars):
    return -0.5 * torch.sum(1 + logvars - mean.pow(2) - logvars.exp())


class NoiseTape(object):
    """Stand-in for torch.randn_like: records the draws of one run (mode 'record') or replays them (mode 'replay'),
    moved to the device / dtype of the tensor they are asked for."""

    def __init__(self, draws=None):
        self.draws = [] if draws is None else [torch.as_tensor(d) for d in draws]
        self.replay = draws is not None
        self.pos = 0
        self._orig = torch.randn_like

    def __call__(self, t, **kw):
        if self.replay:
            d = self.draws[self.pos]
            self.pos += 1
            assert tuple(d.shape) == tuple(t.shape), (tuple(d.shape), tuple(t.shape))
            return d.to(device=t.device, dtype=t.dtype)
        d = self._orig(t, **kw)
        self.draws.append(d.detach().cpu().double().clone())
        return d

    def __enter__(self):
        torch.randn_like = self
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


class VariationalApproximation(torch.nn.Module):
    """q(z_n) = N(mean_n, diag(exp(2 logsigma_n))) per data point (bottleneck/components.py:70-201)."""

    def __init__(self, dim, N, dtype, device):
        super().__init__()
        self._logsigma = torch.nn.Parameter(torch.zeros(N, dim, dtype=dtype, device=device))
        self._mean = torch.nn.Parameter(torch.zeros(N, dim, dtype=dtype, device=device))
        self.N, self.dim = N, dim

    def sample(self):
        eps = torch.randn_like(self._logsigma)
        return self._mean + torch.exp(self._logsigma) * eps

    def KLD(self):
        return unit_gaussian_kld(self._mean, 2 * self._logsigma)

    def entropy(self, sample):
        const = self.N * 0.5 * (np.log(2 * np.pi) + 1)
        return torch.sum(self._logsigma) + const


class EffectivePropertyMap(torch.nn.Module):
    """gp: z -> (mean, logsigma) of the coarse log-conductivities, independent_X (bottleneck/components.py:205-232)."""

    independent_X = True

    def __init__(self, latent_dim, dim_effective_property, dtype, device):
        super().__init__()
        self.fc = torch.nn.Linear(latent_dim, dim_effective_property)
        self.logsigmas_X = torch.nn.Parameter(torch.ones(dim_effective_property))
        self.to(dtype=dtype, device=device)

    def forward(self, z):
        return self.fc(z), self.logsigmas_X.expand(z.shape[0], -1)


class TinyDecoder(torch.nn.Module):
    """Stand-in for the CNN decoder f: z -> (mean, logsigma) of the fine log-field (flattened image)."""

    def __init__(self, dim_latent, n_pixels, dtype, device):
        super().__init__()
        self.dim_latent = dim_latent
        self.fc = torch.nn.Linear(dim_latent, n_pixels)
        self.logsigma = torch.nn.Parameter(torch.full((n_pixels,), -0.5))
        self.to(dtype=dtype, device=device)

    def forward(self, z):
        m = self.fc(z)
        return m, self.logsigma.expand(m.shape[0], -1)


class OracleOperator(torch.nn.Module):
    """ReducedOrderModelOperator.forward from the restated ROM ops (bottleneck/components.py:260-311)."""

    def __init__(self, M, bc_dofs, W, dtype, device):
        super().__init__()
        self.M, self.bc, self.W = M, bc_dofs, W
        self.logsigmas_y = torch.nn.Parameter(torch.ones(W.shape[0], dtype=dtype, device=device))

    dim_effective_property = property(lambda self: self.M.shape[2])
    dim_out = property(lambda self: self.W.shape[0])

    def forward(self, effprop, F):
        from . import rom_ref
        return rom_ref.operator_forward_mean(self.M, self.bc, self.W, effprop, F), self.logsigmas_y.repeat(effprop.shape[0], 1)


class OracleVO(object):
    """What the ELBO reads from a virtual-observable ensemble: .mean and .logsigma [N,d] (detached pseudo-data,
    VirtualObservables.py:852, 868), produced by the restated per-data-point update (oracle/vo_ref.py)."""

    def __init__(self, Gammas, alphas, noise_var, G, PREC, dtype):
        from . import vo_ref
        post = [vo_ref.virtual_observable_update(Ga, al, noise_var, g, p) for Ga, al, g, p in zip(Gammas, alphas, G, PREC)]
        self.mean = torch.stack([p[0] for p in post]).to(dtype)
        self.vars = torch.stack([p[1] for p in post]).to(dtype)
        self.logsigma = 0.5 * torch.log(self.vars)


def random_field_likelihood(predict, target):
    return diagonal_gaussian_log_likelihood(target, predict[0], 2 * predict[1])          # generative.py:233-236


def elbo_supervised_freeX(f, gp, g, q_z, q_X, X, Y, F_ROM_BC):
    """generative.py:456-500 (independent_X, no normalisation, no y preprocessing)."""
    Z_sample = q_z.sample()
    X_sample = q_X.sample()
    predict_x = f(Z_sample)
    logL_x = random_field_likelihood(predict_x, X.detach())
    mu_X, logsigmas_X = gp(Z_sample)
    logL_X = diagonal_gaussian_log_likelihood(X_sample, mu_X, 2 * logsigmas_X)
    mu_y, logsigmas_y = g(X_sample, F_ROM_BC)
    logL_y = diagonal_gaussian_log_likelihood(Y.detach(), mu_y, 2 * logsigmas_y)
    DKL = q_z.KLD()
    entropy = q_X.entropy(X_sample)
    return logL_x + logL_y + logL_X + entropy - DKL


def elbo_virtual_observables_freeX(f, gp, g, q_z, q_X, VO, X, F_ROM_BC):
    """generative.py:352-392 (independent_X, holdoff = False)."""
    Z_sample = q_z.sample()
    DKL = q_z.KLD()
    predict_x = f(Z_sample)
    logL_x = random_field_likelihood(predict_x, X.detach())
    X_sample = q_X.sample()
    mu_X, logsigmas_X = gp(Z_sample)
    logL_X = diagonal_gaussian_log_likelihood(X_sample, mu_X, 2 * logsigmas_X)
    mu_y, logsigmas_y = g(X_sample, F_ROM_BC)
    y_sample = reparametrize(VO.mean, VO.logsigma)
    logL_y = diagonal_gaussian_log_likelihood(y_sample, mu_y, 2 * logsigmas_y)
    entropy = q_X.entropy(X_sample)
    return logL_x + logL_y + logL_X + entropy - DKL


def elbo(f, gp, g, q_z, q_X, VO, data):
    """GenerativeModel.elbo (generative.py:247-287) for a supervised and a virtual-observable data set:
    data = {'supervised': {X, Y, F_ROM_BC}, 'vo': {X, F_ROM_BC}}."""
    total = 0
    s, v = data['supervised'], data['vo']
    total = total + elbo_supervised_freeX(f, gp, g, q_z['supervised'], q_X['supervised'], s['X'], s['Y'], s['F_ROM_BC'])
    total = total + elbo_virtual_observables_freeX(f, gp, g, q_z['vo'], q_X['vo'], VO, v['X'], v['F_ROM_BC'])
    return total


def named_parameters(f, gp, g, q_z, q_X):
    """Flat name -> parameter map in a fixed order (what the fixtures store initial values and gradients under)."""
    out = {}
    for prefix, mod in (("f", f), ("gp", gp), ("g", g)):
        for n, p in mod.named_parameters():
            out["%s.%s" % (prefix, n)] = p
    for key in ("supervised", "vo"):
        out["q_z.%s.mean" % key], out["q_z.%s.logsigma" % key] = q_z[key]._mean, q_z[key]._logsigma
        out["q_X.%s.mean" % key], out["q_X.%s.logsigma" % key] = q_X[key]._mean, q_X[key]._logsigma
    return out
