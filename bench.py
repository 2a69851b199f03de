#!/usr/bin/env python
"""bench.py -- the hot path of the physics layer on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--dtype f64]
    python bench.py --impl reference ...          # the reference's CPU arithmetic (oracle port), host cores

A "step" is one pass of the physics layer over one batch of synthetic log-normal fields:
    coarse-grained model forward (logX, F -> u), its adjoint (gbar_u -> dL/dlogX),
    and the virtual-observable residual r = V^T (K_fom(a) y~ - f) for the same samples.
``value`` = samples/s with inputs resident in HBM (each sample = 1 CGM fwd+adjoint solve + 1 VO residual
evaluation); per-unit rates are reported beside it.  ``e2e`` is the same pass through the public
module API from pinned HOST buffers, copies included.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "CGM fwd+adjoint solves/s & VO residual evals/s"
UNIT = "samples/s (1 CGM fwd+adjoint solve + 1 VO residual eval per sample)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-overlap", action="store_true", help="keep the ROM and VO kernels on one stream")
    ap.add_argument("--no-split-pack", action="store_true",
                    help="let the VO residual call pack V itself (no ordering of the ROM kernels behind the packing)")
    return ap.parse_args()


# ------------------------------------------------------------------------------- CPU arm (oracle port)
class CpuReference(object):
    """The reference's own arithmetic on the host cores: bottleneck/ROM.py ops through autograd
    (oracle/rom_ref.py) and the per-data-point VO route K -> Gamma = V^T K -> Gamma y - alpha
    (oracle/vo_ref.py), all torch/MKL threads.  FEniCS assembly itself cannot be timed here (not
    installable); its place is taken by a precomputed-pattern CSR assembly."""

    def __init__(self, w, sample):
        import numpy as np
        import torch
        from oracle import rom_ref, vo_ref
        self.np, self.torch, self.rom_ref, self.vo_ref = np, torch, rom_ref, vo_ref
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.w, self.sample = w, int(min(sample, w.B))
        rom, fom = w.physics["rom"], w.physics["fom"]
        self.M = torch.tensor(rom.mesh.dense_element_tensor())
        self.bc = torch.tensor(rom.constrained_dofs)
        self.asm = vo_ref.CsrAssembler(fom.mesh.coords, fom.mesh.cells, fom.constrained_dofs, fom.free_dofs,
                                       Ke=fom.mesh.element_stiffness())
        self.pix = fom.mesh.pixel_of_cell()

    def cgm(self, n):
        t = self.torch
        w = self.w
        t0 = time.perf_counter()
        self.rom_ref.rom_fwd_adjoint(self.M, self.bc, t.tensor(w.logX[:n]), t.tensor(w.F[:n]), t.tensor(w.gbar_u[:n]))
        return time.perf_counter() - t0

    def vo(self, n):
        np_, w = self.np, self.w
        t0 = time.perf_counter()
        for b in range(n):   # Python loop over data points, as VirtualObservables.py:895 / :985
            K, f = self.asm.assemble(np_.exp(w.log_image[b][self.pix]), w.g_fom[b])
            Gamma, alpha = self.vo_ref.construct_querry_weak_galerkin(K, f, w.V)
            _ = Gamma @ w.y[b] - alpha
        return time.perf_counter() - t0

    def step(self):
        """One bounded sample: CGM fwd+adjoint on the whole batch, VO on ``sample`` data points."""
        t_cgm = self.cgm(self.w.B)
        t_vo = self.vo(self.sample)
        per_sample = t_cgm / self.w.B + t_vo / self.sample
        return per_sample, t_cgm / self.w.B, t_vo / self.sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gpde_b200  # noqa: F401
    from gpde_b200.workloads import Workload
    from gpde_b200.workloads import CONFIGS
    # bounded sample: the CPU arm never builds more than 4096 samples of the workload (cfg4 has 131072)
    w = Workload(args.workload, B=min(int(args.batch or CONFIGS[args.workload]["B"]), 4096), seed=0)
    sample = 64 if w.d <= 4095 else 8
    ref = CpuReference(w, sample)
    for _ in range(min(args.warmup, 2)):
        ref.step()
    acc, acc_c, acc_v = 0.0, 0.0, 0.0
    t0 = time.perf_counter()
    steps = 0
    for _ in range(max(1, args.steps)):
        p, c, v = ref.step()
        acc += p; acc_c += c; acc_v += v
        steps += 1
        if time.perf_counter() - t0 > 120.0:
            break
    per = acc / steps
    value = 1.0 / per
    sample_desc = "per step: CGM fwd+adjoint on %d samples (torch, %d threads) + VO route on %d data points" % (
        w.B, ref.cores, sample)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * per * w.B, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": dict(w.describe(), parallelism="host cores only"),
        "components": {"cgm_solves_per_s": steps / acc_c, "vo_evals_per_s": steps / acc_v},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                self.samples.append((time.perf_counter(), ln.strip()))
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        rows = [s for (t, s) in self.samples if t0 <= t <= t1] or [s for (_, s) in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for nm, val in zip(names, p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- B200 arm
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload, dtype):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get("%s_%s" % (workload, dtype))
    except Exception:
        return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import gpde_b200  # noqa: F401
    from gpde_b200 import ROM as rom_mod
    from gpde_b200.components import ReducedOrderModelOperator
    from gpde_b200.VirtualObservables import VoPlan
    from gpde_b200.workloads import Workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if args.dtype == "f64" else 4

    # per-GPU shard of the sample-sharded batch.  Default: weak scaling, every rank owns a full workload batch.
    # cfg4 (BASELINE config 4): ONE batch of 131072 cut into contiguous shards (gpde_b200.sharding), strong scaling.
    from gpde_b200.sharding import shard_range
    strong = args.workload == "cfg4"
    if strong:
        from gpde_b200.workloads import CONFIGS
        total = int(args.batch if args.batch is not None else CONFIGS["cfg4"]["B"])
        lo, hi = shard_range(total, rank, world)
        w = Workload(args.workload, B=hi - lo, seed=rank)
    else:
        w = Workload(args.workload, B=args.batch, seed=rank)
    B = w.B
    op = ReducedOrderModelOperator.FromPhysics(w.physics, dtype=tdt, device=dev)
    rom = op.rom
    rom.deferred_checks = True
    plan = rom._get_plan()
    vplan = VoPlan.cached(w.physics["fom"], dev, pixel_input=True)

    host = dict(logX=w.logX, F=w.F, gbar=w.gbar_u, a=w.log_image, y=w.y,
                g=(w.g_fom[0] if w.ptype == "ND" else w.g_fom), V=w.V)
    pinned = {k: torch.tensor(v, dtype=tdt).pin_memory() for k, v in host.items()}
    d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
    torch.cuda.synchronize()

    vo_stream = torch.cuda.Stream(device=dev)
    split_pack = not args.no_split_pack and vplan.kernel_path(w.m, tdt) == 2 and tdt == torch.float64
    packed, ev_packed = [None], torch.cuda.Event()

    def step_resident(overlap=False):
        """One pass of the physics layer.  The coarse-grained model (forward -> adjoint) and the VO residual are
        independent: with ``overlap`` the VO kernels go to a second stream (fork / join by events), so the small
        ROM kernels fill the SMs the VO grid leaves idle (128 CTAs on 148 SMs at this batch)."""
        cur = torch.cuda.current_stream(dev)
        if overlap:
            vo_stream.wait_stream(cur)
            with torch.cuda.stream(vo_stream):
                if split_pack:
                    # V packed by its own call; the ROM kernels are held back until the packing has run, so that
                    # the VO kernel's CTAs (one whole SM each) are placed first and the ROM CTAs take the SMs left
                    pw = vplan.pack_weights(d["V"], B, out=packed[0])
                    if packed[0] is None and hasattr(pw, "buf"):
                        packed[0] = pw
                    ev_packed.record(vo_stream)
                    r = vplan.residual(d["a"], d["y"], d["g"], pw)
                else:
                    r = vplan.residual(d["a"], d["y"], d["g"], d["V"])
            if split_pack:
                cur.wait_event(ev_packed)
        u, factor = rom_mod._launch_forward(plan, d["logX"], d["F"], True, want_factor=True, info=rom._info_word(dev))
        gX, _ = rom_mod._launch_adjoint(plan, d["logX"], u, factor, d["gbar"], True, want_gradF=False)
        if overlap:
            cur.wait_stream(vo_stream)
        else:
            r = vplan.residual(d["a"], d["y"], d["g"], d["V"])
        return u, gX, r

    launches_per_step = 1 + 1 + vplan.launches_per_residual(w.m, tdt)    # rom_forward, rom_adjoint, vo residual

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    K = args.steps

    # ---- (1) eager launches: host enqueue cost and the eager step time (events on the launching stream)
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e_start.record()
    for k in range(K):
        step_resident()
    e_end.record()
    t_host1 = time.perf_counter()
    torch.cuda.synchronize()
    rom.check()
    eager_ms = e_start.elapsed_time(e_end) / K
    host_enqueue_ms = (t_host1 - t_host0) * 1e3 / K

    # ---- (1b) per-kernel-group device times, free of host launch gaps: each group (ROM forward, ROM adjoint,
    # VO residual) is captured in its own CUDA graph and replayed K times between two events on its stream
    def timed_group(fn):
        try:
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
            run = gr.replay
        except Exception:   # noqa: BLE001
            keep, run = None, fn
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(K):
            run()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / K, keep

    u_ref, f_ref = rom_mod._launch_forward(plan, d["logX"], d["F"], True, want_factor=True, info=rom._info_word(dev))
    t_fwd, _k1 = timed_group(lambda: rom_mod._launch_forward(plan, d["logX"], d["F"], True, want_factor=True,
                                                             info=rom._info_word(dev)))
    t_adj, _k2 = timed_group(lambda: rom_mod._launch_adjoint(plan, d["logX"], u_ref, f_ref, d["gbar"], True,
                                                             want_gradF=False))
    t_vo, _k3 = timed_group(lambda: vplan.residual(d["a"], d["y"], d["g"], d["V"]))
    rom.check()

    # ---- (2) the timed region: K steps, each step = the same launches replayed from one CUDA graph (the
    # step is launch-bound from Python at this batch); falls back to eager launches if capture is refused
    graph, mode = None, "eager"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_resident(overlap=not args.no_overlap)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_out = step_resident(overlap=not args.no_overlap)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            mode = "cuda_graph"
        except Exception as exc:   # noqa: BLE001 -- report, keep measuring eagerly
            sys.stderr.write("bench: CUDA graph capture failed (%s); timing eager launches\n" % (exc,))
            graph = None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    start.record()
    for k in range(K):
        if graph is not None:
            graph.replay()
        else:
            step_resident(overlap=not args.no_overlap)
    end.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    rom.check()
    elapsed_ms = start.elapsed_time(end)
    if world > 1:
        tt = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    ms_per_step = elapsed_ms / K

    # ---- end to end through the public module API from pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        outs = dict(u=torch.empty((B, w.n), dtype=tdt).pin_memory(), gX=torch.empty((B, w.E), dtype=tdt).pin_memory(),
                    r=torch.empty((B, w.m), dtype=tdt).pin_memory())
        names_in = ("logX", "F", "gbar", "a", "y", "g")
        bytes_in = sum(pinned[k].numel() * pinned[k].element_size() for k in names_in)
        bytes_out = sum(t.numel() * t.element_size() for t in outs.values())

        def step_e2e():
            dd = {k: pinned[k].to(dev, non_blocking=True) for k in names_in}
            lX = dd["logX"].requires_grad_(True)
            uu = rom.solve_log(lX, dd["F"])            # public API: autograd.Function forward
            uu.backward(dd["gbar"])                      # ... and its adjoint
            rr = vplan.residual(dd["a"], dd["y"], dd["g"], d["V"])
            outs["u"].copy_(uu.detach(), non_blocking=True)
            outs["gX"].copy_(lX.grad, non_blocking=True)
            outs["r"].copy_(rr, non_blocking=True)

        Ke = max(3, min(K, 10))
        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Ke):
            step_e2e()
        e1.record()
        torch.cuda.synchronize()
        e_ms = e0.elapsed_time(e1) / Ke
        if world > 1:
            tt = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        e2e = {"value": world * B / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(bytes_in),
               "d2h_bytes_per_step": int(bytes_out), "ms_per_step": e_ms, "steps": Ke}
    if sampler:
        sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    vo_bytes = w.vo_bytes_per_eval(s) * B
    achieved = vo_bytes / (t_vo * 1e-3) / 1e9
    cgm_bytes = w.cgm_bytes_per_solve(s) * B
    path = vplan.kernel_path(w.m, tdt)
    # FP64 work of one VO evaluation: exp 11 x 1.25, fluxes 10, contraction m (FMA = 2 flop)
    vo_flops = 2.0 * w.d * (11 * 1.25 + 10 + w.m) * B
    fp64_peak = 37.0   # TFLOP/s, DMMA/DFMA rate measured on this pool: profiles/r1_fp64_peak.txt
    line = {
        "metric": METRIC, "value": (total if strong else world * B) / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": dict(w.describe(), per_gpu_batch=B, global_batch=(total if strong else world * B),
                       parallelism="sample-sharded x%d, no data-path collective" % world,
                       l2="inputs larger than L2: %.0f MB streamed per step per GPU" % ((vo_bytes + cgm_bytes) / 1e6)),
        "components": {
            "cgm_solves_per_s": world * B / ((t_fwd + t_adj) * 1e-3), "vo_evals_per_s": world * B / (t_vo * 1e-3),
            "ms_rom_forward": t_fwd, "ms_rom_adjoint": t_adj, "ms_vo_residual": t_vo,
            "launch_mode": mode + ("" if args.no_overlap else " + ROM/VO on two streams"), "ms_per_step_eager": eager_ms, "ms_host_enqueue_per_step": host_enqueue_ms,
            "cgm_hbm_frac": cgm_bytes / ((t_fwd + t_adj) * 1e-3) / 1e9 / peak,
        },
        "roofline": ({"kernel": "vo_grid2_kernel<rho> + vo_gemm_kernel (FP64 DMMA contraction dominates)", "bound": "tensor",
                      "achieved": vo_flops / (t_vo * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": vo_flops / (t_vo * 1e-3) / 1e12 / fp64_peak, "traffic": recorded_traffic(args.workload, args.dtype),
                      "peak_source": "measured FP64 mma.sync rate (profiles/r1_fp64_peak.txt); MEASURED_PEAKS.json has no FP64 entry",
                      "algorithmic_flops_per_launch": vo_flops, "hbm_frac": achieved / peak}
                     if path in (0, 3) and w.m > 32 else
                     {"kernel": {2: "vo_grid2_kernel (+ vo_grid2_pack_kernel)", 1: "vo_fused_kernel"}.get(path, "vo_matvec_kernel + vo_gemm_kernel"),
                      "bound": "hbm", "achieved": achieved, "peak": peak,
                      "unit": "GB/s", "frac": achieved / peak, "traffic": recorded_traffic(args.workload, args.dtype),
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": vo_bytes,
                      "fp64_pipe_frac": vo_flops / (t_vo * 1e-3) / 1e12 / fp64_peak}),
        "gpu_launches": launches_per_step * K,
        "clocks": sampler.summary(t_wall0, t_wall1) if sampler else None,
        "e2e": e2e,
    }
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(w, 32 if w.d <= 4095 else 4)
        ref.step()
        t0, n, acc = time.perf_counter(), 0, 0.0
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 20):
            acc += ref.step()[0]
            n += 1
        line["cpu_baseline"] = {
            "value": n / acc, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "sample": "%d x (CGM fwd+adjoint on %d samples via torch autograd + VO per-data-point route on %d samples)"
                      % (n, B, ref.sample)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
