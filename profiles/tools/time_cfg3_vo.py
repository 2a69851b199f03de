"""Device time of the VO residual at BASELINE config 3 (128 x 128, m = 256) and of its two halves, device-generated inputs.
    python profiles/tools/time_cfg3_vo.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.workloads import Workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
w = Workload("cfg3", B=8, seed=0)
plan = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
gen = torch.Generator(device=dev).manual_seed(0)
a = torch.exp(0.4 + 0.8 * torch.randn(B, w.P, generator=gen, device=dev, dtype=torch.float64))
y = torch.randn(B, w.d, generator=gen, device=dev, dtype=torch.float64)
g = torch.tensor(w.g_fom[0], device=dev)
V = torch.tensor(w.V, device=dev)

def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("cfg3 B=%d: residual (rho + contraction) %.3f ms; rho only %.3f ms" % (
    B, timed(lambda: plan.residual(a, y, g, V, a_is_log=False)), timed(lambda: plan.residual(a, y, g, None, a_is_log=False))))
