"""Times VoPlan.residual_T (q = K_ff V s) and the autograd backward of the residual at a workload's shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200.VirtualObservables import VoPlan
from gpde_b200.workloads import Workload

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = Workload(name, B=B, seed=0)
dev = torch.device("cuda", 0)
plan = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)
a = torch.tensor(w.log_image, device=dev)
V = torch.tensor(w.V, device=dev)
y = torch.tensor(w.y, device=dev)
g = torch.tensor(w.g_fom[0] if w.ptype == "ND" else w.g_fom, device=dev)
s = torch.randn(w.B, w.m, dtype=torch.float64, device=dev)
def timed(fn, n=20):
    """device time per call from CUDA-graph replays (no host launch gaps); eager time beside it"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / n
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, eager


for fn, label in ((lambda: plan.residual_T(a, V, s), "residual_T"), (lambda: plan.residual(a, y, g, V), "residual")):
    tg, te = timed(fn)
    print("%s %s B=%d m=%d d=%d: %.3f ms (graph replay), %.3f ms (eager launches)" % (name, label, w.B, w.m, w.d, tg, te))
