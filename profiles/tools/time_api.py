"""Times the module-level calls of the mirrored API (operator forward/backward, VO ensemble update) on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpde_b200  # noqa
from gpde_b200 import VirtualObservables as VO
from gpde_b200.components import ReducedOrderModelOperator
from gpde_b200.workloads import Workload

dev = torch.device("cuda", 0)


def timeit(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


# operator forward + backward at cfg2 (B = 4096, d = 4095)
w = Workload("cfg2", seed=0)
op = ReducedOrderModelOperator.FromPhysics(w.physics, dtype=torch.float64, device=dev)
logX = torch.tensor(w.logX, device=dev, requires_grad=True)
F = torch.tensor(w.F, device=dev)
gy = torch.randn(w.B, w.d, dtype=torch.float64, device=dev)


def fwd_bwd():
    mu, ls = op.forward(logX, F)
    mu.backward(gy)
    logX.grad = None

print("operator forward (no grad)   cfg2: %.3f ms" % timeit(lambda: op.forward(logX.detach(), F)))
print("operator forward + backward  cfg2: %.3f ms" % timeit(fwd_bwd))

# VO ensemble at the notebook's scale: highres32, N = 128 data points, CGR weighting (m = 25)
w1 = Workload("cfg1", B=128, seed=0)
fom = w1.physics["fom"]
X_DG = w1.log_image[:, fom.mesh.pixel_of_cell()]
t0 = time.perf_counter()
qpe = VO.QuerryPointEnsemble.FromArrays(X_DG, w1.bce, fom, device=dev)
qe = VO.QuerryEnsemble.FromQuerryPointEnsemble(qpe, w1.physics, CGR=True, flux=False, N_gaussian=0, N_rbf=0,
                                               dtype=torch.double, device=dev)
ens = VO.VirtualObservablesEnsemble(qpe, qe, torch.double, dev)
torch.cuda.synchronize()
print("ensemble construction N=128: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
G = torch.tensor(w1.y, device=dev)
PREC = torch.full_like(G, 100.0)
ens.update(G, PREC, 0)
print("ensemble.update N=128 m=%d d=%d: %.3f ms" % (ens.m, ens.dim_out, timeit(lambda: ens.update(G, PREC, 1), n=5, warm=1)))
print("ensemble.mean/logsigma: %.3f ms" % timeit(lambda: (ens.flush_cache(), ens.mean, ens.logsigma)))

# north_star kernel (a) as the reference formulates it: dense assembly K = matmul(M, x^T) ([n,n,E] x [E,B], cuBLAS FP64)
# + Dirichlet overwrite + batched LU (torch.linalg.solve), all on the GPU, against the fused sparse-band kernels
wb = Workload("cfg2", B=32768, seed=1)
opb = ReducedOrderModelOperator.FromPhysics(wb.physics, dtype=torch.float64, device=dev)
rom = opb.rom
Xb = torch.exp(torch.tensor(wb.logX, device=dev)) + 1e-8
Fb = torch.tensor(wb.F, device=dev)
M = torch.tensor(wb.physics["rom"].mesh.dense_element_tensor(), device=dev)
bc = torch.tensor(wb.physics["rom"].constrained_dofs, device=dev)


def dense_route():
    K = torch.matmul(M, Xb.t())                       # [n,n,B]   (bottleneck/ROM.py:93)
    K[bc] = 0
    K[bc, bc] = 1                                     # (:97-98)
    return torch.linalg.solve(K.permute(2, 0, 1), Fb.unsqueeze(2)).squeeze(2)   # (:61, :83)

u_dense = dense_route()
u_fused = rom(Xb, Fb)
print("dense GEMM assembly + batched LU (torch/cuBLAS/cuSOLVER on the GPU), B=32768: %.3f ms   max|du| %.1e"
      % (timeit(dense_route, n=5), (u_dense - u_fused).abs().max().item()))
print("  of which matmul(M, x^T) alone: %.3f ms ; GetStiffness kernel (dense K out): %.3f ms"
      % (timeit(lambda: torch.matmul(M, Xb.t()), n=5), timeit(lambda: rom.GetStiffness(Xb), n=5)))
print("fused sparse-band assemble + LDL^T + solve kernel, B=32768: %.3f ms" % timeit(lambda: rom(Xb, Fb), n=10))
