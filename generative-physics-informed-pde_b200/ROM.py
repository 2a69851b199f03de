"""Coarse-grained model -- drop-in for the reference's ``bottleneck/ROM.py``.

Same class name, constructor, properties and call signature as bottleneck/ROM.py:5-104; underneath,
``ROM.__call__`` is one ``torch.autograd.Function`` (RomSolveFn) whose forward/backward are single
launches of the sm_100a kernels in csrc/rom.cu through the C ABI (include/gpde_b200.h):
fused assemble -> banded LDL^T -> solve, and the adjoint solve with the stashed factor.

The reference has no custom autograd.Function (autograd differentiates matmul/index_put/solve,
SURVEY.md 8 a6); gradients returned here are the same quantities:
    dL/dX = -(lambda_f^T K_e u)_e,   dL/dF = lambda = A^-T gbar.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class _RomPlan(object):
    """Owner of one gpde_rom_plan (device constants built from M and the Dirichlet map)."""

    def __init__(self, M, bc_dofs, device):
        lib = _lib.load()
        device = _lib.require_cuda(device, "ROM")
        M_host = np.ascontiguousarray(M.detach().to("cpu", torch.float64).numpy())
        bc = np.ascontiguousarray(np.asarray(bc_dofs, dtype=np.int64))
        n, n2, E = M_host.shape
        assert n == n2
        handle = ctypes.c_void_p()
        rc = lib.gpde_rom_plan_create(ctypes.byref(handle), n, E, M_host.ctypes.data_as(ctypes.c_void_p),
                                      bc.ctypes.data_as(ctypes.c_void_p), bc.size, device.index)
        _lib.check(rc, "gpde_rom_plan_create")
        self.handle, self.device, self._lib = handle, device, lib
        info = (ctypes.c_int64 * 8)()
        _lib.check(lib.gpde_rom_plan_info(handle, info), "gpde_rom_plan_info")
        self.n, self.E, self.n_free, self.half_bandwidth, self.factor_doubles, self.n_contrib, self.lanes = \
            [int(v) for v in info[:7]]
        # lanes == 2: windowed thread-per-sample kernels -- the forward call itself streams the factor through the stash
        self.factor_required = self.lanes == 2

    def factor_buffer(self, B, device):
        """Opaque factor stash of a batch of B samples (layout is the kernels' business), or None when the plan keeps none."""
        nbytes = int(self._lib.gpde_rom_factor_bytes(self.handle, int(B)))
        return torch.empty(nbytes // 8, dtype=torch.float64, device=device) if nbytes else None

    def __del__(self):
        try:
            if self.handle:
                self._lib.gpde_rom_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _launch_forward(plan, X, F, x_is_log, want_factor=True, info=None):
    lib, B = plan._lib, X.shape[0]
    sfx = _lib.suffix(X.dtype)
    u = torch.empty((B, plan.n), dtype=X.dtype, device=X.device)
    factor = plan.factor_buffer(B, X.device) if (want_factor or plan.factor_required) else None
    fn = getattr(lib, "gpde_rom_forward_" + sfx)
    dev = plan.device
    rc = fn(plan.handle, _lib.ptr(X, dev), int(bool(x_is_log)), _lib.ptr(F, dev), _lib.ptr(u, dev), _lib.ptr(factor, dev),
            _lib.ptr(info, dev), B, _lib.stream_of(dev))
    _lib.check(rc, "gpde_rom_forward_" + sfx)
    return u, factor


def _launch_adjoint(plan, X, u, factor, gbar, x_is_log, want_gradF=True):
    lib, B = plan._lib, X.shape[0]
    sfx = _lib.suffix(X.dtype)
    gX = torch.empty_like(X)
    gF = torch.empty((B, plan.n), dtype=X.dtype, device=X.device) if want_gradF else None
    fn = getattr(lib, "gpde_rom_adjoint_" + sfx)
    dev = plan.device
    rc = fn(plan.handle, _lib.ptr(X, dev), int(bool(x_is_log)), _lib.ptr(u, dev), _lib.ptr(factor, dev), _lib.ptr(gbar, dev),
            _lib.ptr(gX, dev), _lib.ptr(gF, dev), B, _lib.stream_of(dev))
    _lib.check(rc, "gpde_rom_adjoint_" + sfx)
    return gX, gF


class RomSolveFn(torch.autograd.Function):
    """u = A(x)^-1 F per sample; x = X (conductivities) or exp(X)+1e-8 (x_is_log).

    forward saves (X, u, banded LDL^T factor); backward is one adjoint launch reusing that factor."""

    @staticmethod
    def forward(ctx, X, F, rom, x_is_log):
        plan = rom._get_plan()
        Xc, Fc = X.contiguous(), F.contiguous()
        need_grad = X.requires_grad or F.requires_grad
        info = rom._info_word(Xc.device)
        u, factor = _launch_forward(plan, Xc, Fc, x_is_log, want_factor=need_grad, info=info)
        ctx.plan, ctx.x_is_log = plan, bool(x_is_log)
        ctx.save_for_backward(Xc, u, factor if factor is not None else torch.empty(0, device=Xc.device))
        return u

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gbar):
        X, u, factor = ctx.saved_tensors
        factor = factor if factor.numel() else None
        gX, gF = _launch_adjoint(ctx.plan, X, u, factor, gbar.contiguous(), ctx.x_is_log,
                                 want_gradF=ctx.needs_input_grad[1])
        return (gX if ctx.needs_input_grad[0] else None), gF, None, None


class ROM(object):
    """Reference surface: bottleneck/ROM.py:5-104."""

    trunc = 1e-12   # ROM.py:73

    def __init__(self, physics, M, dtype, device):

        self.M = M
        self.dtype = dtype
        self.device = device

        self._bc_dofs = torch.tensor(np.asarray(physics.constrained_dofs).copy(), dtype=torch.long, device=device)
        self._free_dofs = torch.tensor(np.asarray(physics.free_dofs).copy(), dtype=torch.long, device=device)
        self._bc_dofs_host = np.asarray(physics.constrained_dofs, dtype=np.int64).copy()

        self._plan = None
        self._info = None
        # The reference synchronises in every call (`(X <= trunc).any().item()`, ROM.py:75).  With
        # deferred_checks=True the same test runs on the device inside the solve kernel and the flag
        # is only read by check(), keeping the hot path free of host syncs.
        self.deferred_checks = False

    @property
    def physics(self):
        raise DeprecationWarning

    @property
    def V_dim(self):
        return self.M.shape[0]

    @property
    def Vc_dim(self):
        return self.M.shape[2]

    @property
    def dim_in(self):
        return self.Vc_dim

    @property
    def dim_out(self):
        return self.V_dim

    @classmethod
    def FromPhysics(cls, physics, dtype=torch.double, device=torch.device("cpu")):
        """physics: the coarse LinearEllipticPhysics (FEniCS-free exporter, physics.py).  The element
        tensor M[:,:,e] is what df.assemble(derivative(a, alpha, e_e)) gives (ROM.py:46-53)."""
        if physics.mesh.num_cells > 290:
            raise Exception('ROM exceeds intended maximum size')
        M = torch.tensor(physics.mesh.dense_element_tensor(), dtype=dtype, device=device)
        return cls(physics, M, dtype=dtype, device=device)

    # ------------------------------------------------------------------------------ internals
    def _get_plan(self):
        if self._plan is None:
            self._plan = _RomPlan(self.M, self._bc_dofs_host, self.device)
        return self._plan

    def _info_word(self, device):
        if self._info is None or self._info.device != device:
            self._info = torch.zeros(1, dtype=torch.int32, device=device)
        return self._info

    def check(self):
        """Reads and clears the device info word; raises like the reference (ROM.py:74-76)."""
        if self._info is None:
            return
        flag = int(self._info.item())
        if flag:
            self._info.zero_()
        if flag & 1:
            raise ValueError('At least one of the conductivity values supplied to the ROM was smaller than {}'.format(self.trunc))
        if flag & 2:
            raise RuntimeError('ROM stiffness matrix is not positive definite (non-positive pivot)')

    def _solve(self, X, F, x_is_log):
        if X.dim() < 2:
            X = X.unsqueeze(0)
        if F.dim() == 3:   # the reference accepts F[B,n,1] as well (ROM.py:71-72)
            F = F.squeeze(2)
        if F.dim() < 2:
            F = F.unsqueeze(0)
        if F.shape[0] != X.shape[0]:
            F = F.expand(X.shape[0], -1)
        if F.dtype != X.dtype:
            F = F.to(X.dtype)
        u = RomSolveFn.apply(X, F, self, x_is_log)
        if not self.deferred_checks:
            self.check()
        return u

    # ------------------------------------------------------------------------------ reference API
    def __call__(self, X, F=None, ReturnStiffness=False):

        # F is assumed to already have been modified correctly for the B.C.
        if F is None:
            # the reference dereferences F before its own None check (ROM.py:71 vs :80-81)
            raise AttributeError("'NoneType' object has no attribute 'dim'")

        y_rom = self._solve(X, F, x_is_log=False)

        if ReturnStiffness:
            return y_rom, self.GetStiffness(X, DirichletBC=True)
        else:
            return y_rom

    def solve_log(self, logX, F):
        """rom(exp(logX) + 1e-8, F) with the exponential fused into the kernel and the gradient
        returned w.r.t. logX (what ReducedOrderModelOperator.forward needs, components.py:298)."""
        return self._solve(logX, F, x_is_log=True)

    def GetStiffness(self, x, DirichletBC=True):
        """K_batched[n,n,B], batch last (ROM.py:91-100).  Not differentiable (diagnostic path)."""
        plan = self._get_plan()
        x64 = x.detach().to(torch.float64).contiguous()
        if x64.dim() < 2:
            x64 = x64.unsqueeze(0)
        B = x64.shape[0]
        K = torch.empty((plan.n, plan.n, B), dtype=torch.float64, device=x64.device)
        rc = plan._lib.gpde_rom_stiffness_f64(plan.handle, _lib.ptr(x64, plan.device), _lib.ptr(K, plan.device),
                                              int(bool(DirichletBC)), B, _lib.stream_of(plan.device))
        _lib.check(rc, "gpde_rom_stiffness_f64")
        return K.to(x.dtype)

    def __repr__(self):
        return "This is the ROM \n Maps: {} -> {}".format(self.Vc_dim, self.V_dim)
