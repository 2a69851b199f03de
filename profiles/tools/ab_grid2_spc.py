"""A/B of the samples a CTA of vo_grid2_kernel takes (GPDE_GRID2_SPC; 0 = automatic) on ONE box: device time of the VO
residual call (graph of 20 calls, replayed 5 times), three rounds, FP64 and FP32 I/O.
    python profiles/tools/ab_grid2_spc.py [B] [spc ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.workloads import Workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spcs = [int(x) for x in sys.argv[2:]] or [32, 0, 30, 28, 26, 24]
dev = torch.device("cuda", 0)
w = Workload("cfg2", B=min(B, 4096), seed=0)
base = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
T = lambda t: torch.tensor(t, device=dev)
rep = (B + w.B - 1) // w.B
a_log, y = T(w.log_image).repeat(rep, 1)[:B].contiguous(), T(w.y).repeat(rep, 1)[:B].contiguous()
g, V = T(w.g_fom[0]), T(w.V)
a = torch.exp(a_log)
plans = {c: base.variant(GPDE_GRID2_SPC=str(c)) for c in spcs}
ref = plans[spcs[0]].residual(a, y, g, V, a_is_log=False)

def timed(plan, log, f32=False):
    c = (lambda t: t.float()) if f32 else (lambda t: t)
    aa, yy, gg, VV = c(a_log if log else a), c(y), c(g), c(V)
    fn = lambda: plan.residual(aa, yy, gg, VV, a_is_log=log)
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

for c, p in plans.items():
    assert torch.equal(p.residual(a, y, g, V, a_is_log=False), ref), c
for rnd in range(3):
    for c in spcs:
        print("round %d B %d spc %2d: conductivity %.2f us, log input %.2f us, FP32 I/O %.2f us" %
              (rnd, B, c, timed(plans[c], False), timed(plans[c], True), timed(plans[c], False, True)), flush=True)
