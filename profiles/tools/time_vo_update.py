"""Where the virtual-observable update of BASELINE config 5 spends its time (generative.py:182-222 on the mirrors):
device-synchronised wall time of each stage, N_vo = 128 data points, N_mc = 64 Monte-Carlo samples, FP32 model dtype."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200.svi_workload import SviWorkload
dev = torch.device("cuda", 0)
wl = SviWorkload(dev, torch.float32, seed=0)
wl.build_virtual_observables()
for _ in range(3):
    wl.update_virtual_observables(N_mc=64, step=1)
torch.cuda.synchronize()

def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, out

N_mc = 64
ms_s, X_s = t(lambda: wl.q_X["vo"].mean.unsqueeze(1) + torch.exp(wl.q_X["vo"].logsigma).unsqueeze(1) * torch.randn(wl.N_vo, N_mc, wl.w.E, dtype=wl.dtype, device=dev))
ms_m, (Ym, Ys) = t(lambda: wl.g.predictive_moments(X_s.detach(), wl.vo["F"]))
ms_r, _ = t(lambda: wl.VO.resample())
ms_p, _ = t(lambda: wl.VO.update_vo_precision(1))
ms_u, _ = t(lambda: wl.VO.update(Ym, 1.0 / Ys ** 2, 1))
ms_c, _ = t(lambda: (wl.vo_mean.copy_(wl.VO.mean), wl.vo_logsigma.copy_(wl.VO.logsigma)))
ms_all, _ = t(lambda: wl.update_virtual_observables(N_mc=64, step=1))
print("VO update N_vo=%d N_mc=%d: sample %.3f | predictive moments %.3f | resample %.3f | precision %.3f | update (incl. precision) %.3f | "
      "copy out %.3f | whole %.3f ms" % (wl.N_vo, N_mc, ms_s, ms_m, ms_r, ms_p, ms_u, ms_c, ms_all))

if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        wl.update_virtual_observables(N_mc=64, step=1)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
