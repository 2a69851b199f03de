import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cProfile, pstats
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.workloads import Workload
w = Workload("cfg2", seed=0)
dev = torch.device("cuda", 0)
plan = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
a = torch.tensor(w.log_image, device=dev); V = torch.tensor(w.V, device=dev)
s = torch.randn(w.B, w.m, dtype=torch.float64, device=dev)
for _ in range(3): plan.residual_T(a, V, s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): q = plan.residual_T(a, V, s)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("enqueue per call %.3f ms, total per call %.3f ms" % ((t1 - t0) * 50, (t2 - t0) * 50))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): q = plan.residual_T(a, V, s)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
