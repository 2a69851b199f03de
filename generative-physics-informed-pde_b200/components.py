"""ReducedOrderModelOperator -- drop-in for bottleneck/components.py:260-323.

forward(effprop, F) = (W @ rom(exp(effprop)+1e-8, F), logsigmas_y.repeat(B,1)), with
  * exp(.)+1e-8, assembly, solve and (in backward) the adjoint in the two ROM kernels (ROM.py here),
  * the prolongation einsum('sk,nk->ns', W, u) as a CSR gather kernel (csrc/prolong.cu): W has <= 3
    non-zeros per row, the reference multiplies by it as a dense [d,n] matrix.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ROM import ROM


class _ProlongPlan(object):

    def __init__(self, W, device):
        lib = _lib.load()
        device = _lib.require_cuda(device, "ReducedOrderModelOperator")
        W_host = np.ascontiguousarray(W.detach().to("cpu", torch.float64).numpy())
        d, n = W_host.shape
        handle = ctypes.c_void_p()
        rc = lib.gpde_prolong_plan_create(ctypes.byref(handle), d, n, W_host.ctypes.data_as(ctypes.c_void_p),
                                          device.index)
        _lib.check(rc, "gpde_prolong_plan_create")
        self.handle, self.device, self._lib, self.d, self.n = handle, device, lib, d, n

    def __del__(self):
        try:
            if self.handle:
                self._lib.gpde_prolong_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def apply(self, u, transpose=False):
        sfx = _lib.suffix(u.dtype)
        B = u.shape[0]
        out = torch.empty((B, self.n if transpose else self.d), dtype=u.dtype, device=u.device)
        fn = getattr(self._lib, "gpde_prolong_apply_%s%s" % ("T_" if transpose else "", sfx))
        _lib.check(fn(self.handle, _lib.ptr(u, self.device), _lib.ptr(out, self.device), B, _lib.stream_of(self.device)),
                   "gpde_prolong_apply")
        return out


class ProlongFn(torch.autograd.Function):
    """y[B,d] = u[B,n] W^T ; backward gbar_u = gbar_y W."""

    @staticmethod
    def forward(ctx, u, plan):
        ctx.plan = plan
        return plan.apply(u.contiguous(), transpose=False)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        return ctx.plan.apply(gy.contiguous(), transpose=True), None


class ReducedOrderModelOperator(torch.nn.Module):

    def __init__(self, rom, W, *, dtype=None, device=None):

        super(ReducedOrderModelOperator, self).__init__()

        self.W = W
        self.rom = rom

        self._dtype = dtype
        self._device = device

        self.logsigmas_y = torch.nn.Parameter(torch.ones(W.shape[0], requires_grad=True))
        self.to(dtype=dtype, device=device)
        self._prolong = None

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    @property
    def dim_effective_property(self):
        return self.rom.Vc_dim

    @property
    def dim_in(self):
        return self.dim_effective_property

    @property
    def dim_out(self):
        return self.W.shape[0]

    def _prolong_plan(self):
        if self._prolong is None:
            self._prolong = _ProlongPlan(self.W, self.W.device)
        return self._prolong

    def forward(self, effprop, F):

        return self.forward_mean(effprop, F), self.logsigmas_y.repeat(effprop.shape[0], 1)

    def forward_mean(self, effprop, F):

        u = self.rom.solve_log(effprop, F)
        return ProlongFn.apply(u, self._prolong_plan())

    def propagate_samples(self, effprops, F):

        means, logsigmas = self.forward(effprops, F)

        if means.shape != logsigmas.shape:
            raise RuntimeError('Implementation assumes that full logsigmas matrix is given; check for broadcasting')

        return means + torch.exp(logsigmas) * torch.randn_like(logsigmas)

    @classmethod
    def FromPhysics(cls, physics, *, dtype=None, device=None):

        W = torch.tensor(physics['W'].T, dtype=dtype, device=device).t()

        if W.shape[0] < W.shape[1]:
            raise ValueError

        rom = ROM.FromPhysics(physics['rom'], dtype=dtype, device=device)
        return cls(rom, W, dtype=dtype, device=device)
