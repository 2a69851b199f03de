"""Device time of the coarse-grained-model kernels alone: 20 launches captured in one CUDA graph, replayed 5 times.
    python profiles/tools/time_rom.py [B] [cfg] [f32]       (GPDE_ROM_PATH=coop for the cooperative kernels)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200 import ROM as rom_mod
from gpde_b200.components import ReducedOrderModelOperator
from gpde_b200.workloads import Workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
w = Workload(cfg, B=B, seed=0)
dev = torch.device("cuda", 0)
op = ReducedOrderModelOperator.FromPhysics(w.physics, dtype=dt, device=dev)
rom = op.rom
plan = rom._get_plan()
logX = torch.tensor(w.logX, device=dev, dtype=dt); F = torch.tensor(w.F, device=dev, dtype=dt); gb = torch.tensor(w.gbar_u, device=dev, dtype=dt)
def fwd():
    return rom_mod._launch_forward(plan, logX, F, True, want_factor=True, info=rom._info_word(dev))
u, fac = fwd()
def adj():
    return rom_mod._launch_adjoint(plan, logX, u, fac, gb, True, want_gradF=False)
def adjF():
    return rom_mod._launch_adjoint(plan, logX, u, fac, gb, True, want_gradF=True)
for fn, name in ((fwd, "forward"), (adj, "adjoint"), (adjF, "adjoint+gradF")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = [fn() for _ in range(20)]
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 100 * 1e3
    print("rom %s %s lanes=%d B=%d: %.1f us  (%.1f M/s)" % (name, cfg, plan.lanes, B, t, B / t))
