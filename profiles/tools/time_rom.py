import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpde_b200  # noqa
from gpde_b200 import ROM as rom_mod
from gpde_b200.components import ReducedOrderModelOperator
from gpde_b200.workloads import Workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
w = Workload("cfg2", B=B, seed=0)
dev = torch.device("cuda", 0)
op = ReducedOrderModelOperator.FromPhysics(w.physics, dtype=torch.float64, device=dev)
rom = op.rom
plan = rom._get_plan()
logX = torch.tensor(w.logX, device=dev); F = torch.tensor(w.F, device=dev); gb = torch.tensor(w.gbar_u, device=dev)
def fwd():
    return rom_mod._launch_forward(plan, logX, F, True, want_factor=True, info=rom._info_word(dev))
u, fac = fwd()
def adj():
    return rom_mod._launch_adjoint(plan, logX, u, fac, gb, True, want_gradF=False)
for fn, name in ((fwd, "forward"), (adj, "adjoint")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("rom %s B=%d: %.1f us" % (name, B, e0.elapsed_time(e1) / 20 * 1e3))
