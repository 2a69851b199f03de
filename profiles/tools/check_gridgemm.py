"""One-kernel route for many weighting functions (vo_gridgemm.cuh) against the two-kernel route (rho through HBM + vo_gemm)
on the same inputs: pixel meshes the lean kernels serve, ragged batches, shared / per-sample fields, log / conductivity input.
    python profiles/tools/check_gridgemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.physics import LinearEllipticPhysics
from gpde_b200 import fem

dev = torch.device("cuda", 0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_vo import _grid_case  # noqa

worst = 0.0
for nx, ny, B in [(16, 4, 5), (32, 32, 37), (64, 64, 130), (128, 16, 19), (128, 128, 64)]:
    plan, fom, a, y, g, rng = _grid_case(nx, ny, "NDP", B, nx * 1000 + ny, dev, load=False)
    old = plan.variant(GPDE_VO_GRIDGEMM="0")
    T = lambda t: torch.tensor(t, device=dev)
    for m in (33, 70, 130, 256, 300):
        V = rng.normal(size=(fom.dim_out, m))
        V[:, : m // 3] *= (rng.uniform(size=(fom.dim_out, m // 3)) < 0.02)     # sparse columns: tiles get skipped
        V = T(V)
        for kw in (dict(a=T(a), y=T(y), g=T(g)), dict(a=T(a[0]), y=T(y), g=T(g[0])),
                   dict(a=T(np.exp(a)), y=T(y), g=None, a_is_log=False)):
            r1 = plan.residual(V=V, **kw)
            r0 = old.residual(V=V, **kw)
            torch.cuda.synchronize()
            e = ((r1 - r0).abs().max() / r0.abs().max()).item()
            worst = max(worst, e)
            if not e < 1e-12:
                print("MISMATCH", nx, ny, B, m, list(kw), e)
        for dt in (torch.float32,):
            r1 = plan.residual(T(a).to(dt), T(y).to(dt), T(g).to(dt), V.to(dt))
            r0 = old.residual(T(a).to(dt), T(y).to(dt), T(g).to(dt), V.to(dt))
            e = ((r1 - r0).abs().max() / r0.abs().max()).item()
            if not e < 1e-5:
                print("MISMATCH f32", nx, ny, B, m, e)
    print("mesh", nx, ny, "B", B, "ok; worst so far %.2e" % worst)
print("worst rel diff fused vs two-kernel: %.3e" % worst)
