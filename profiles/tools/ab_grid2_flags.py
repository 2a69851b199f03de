"""A/B of vo_grid2_kernel experiment switches on ONE box: device time of the VO residual call at cfg 2 (graph of 20 calls,
replayed 5 times) for GPDE_GRID2_FLAGS = 0, 2, ... alternating, three rounds.
    python profiles/tools/ab_grid2_flags.py [flags ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.workloads import Workload
flags = [int(x) for x in sys.argv[1:]] or [0, 2]
dev = torch.device("cuda", 0)
w = Workload("cfg2", B=4096, seed=0)
base = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
T = lambda t: torch.tensor(t, device=dev)
a_log, y, g, V = T(w.log_image), T(w.y), T(w.g_fom[0]), T(w.V)
a = torch.exp(a_log)
plans = {f: base.variant(GPDE_GRID2_FLAGS=str(f)) for f in flags}
ref = plans[flags[0]].residual(a, y, g, V, a_is_log=False)

def timed(plan, log):
    fn = lambda: plan.residual(a_log if log else a, y, g, V, a_is_log=log)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = [fn() for _ in range(20)]
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1e3

for f, p in plans.items():
    assert (p.residual(a, y, g, V, a_is_log=False) - ref).abs().max() <= 1e-9 * ref.abs().max(), f
for rnd in range(3):
    for f in flags:
        print("round %d flags %d: conductivity input %.2f us, log input %.2f us" % (rnd, f, timed(plans[f], False), timed(plans[f], True)), flush=True)
