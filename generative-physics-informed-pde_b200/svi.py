"""Data-parallel plumbing for the stochastic-variational-inference step that calls the physics layer
(the reference's loop: training.py:393-462 ``Trainer.run`` -- zero_grad, ``model.elbo``, ``(-elbo).backward()``,
``optimizer.step()`` -- single process, single device).  SURVEY.md section 8(e), BASELINE config 5:

  * one process per GPU; the supervised / virtual-observable data points are sharded by OWNER rank, and with them their
    per-sample variational tables q_z / q_X (bottleneck/components.py:70-201: one mean / logsigma row per data point) --
    those rows only ever see their own sample's gradient, so they need no communication;
  * the SHARED parameters (decoder, encoder, effective-property map, logsigmas_y: ~10.9 k values) receive the SUM of the
    ranks' gradients (the ELBO is a sum over data points) through ONE all-reduce of ONE flat buffer: their ``.grad``
    tensors are views into it, so there is no per-parameter bucket logic and no copy before or after the collective.
    43 kB over NVLink / NVSwitch is latency-bound: one NCCL launch is the cost;
  * the whole step (NN kernels, the physics-layer kernels, the all-reduce, the optimizer) can be captured into one CUDA graph
    (``GraphedStep``): at these sizes the eager step is bound by Python launch overhead, not by the GPU.

The ELBO itself is the caller's (``elbo_fn``): this module never looks inside it.
"""
import torch
import torch.distributed as dist

from .sharding import shard_range


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradientBucket(object):
    """The gradients of ``params`` as views of one contiguous buffer (same dtype and device for all of them)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradientBucket: no trainable parameter")
        p0 = self.params[0]
        if any(p.dtype != p0.dtype or p.device != p0.device for p in self.params):
            raise ValueError("FlatGradientBucket: parameters must share dtype and device")
        self.group = group
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=p0.dtype, device=p0.device)
        self._views = []
        offset = 0
        for p in self.params:
            view = self.flat[offset:offset + p.numel()].view_as(p)
            p.grad = view                      # autograd accumulates in place into an existing .grad
            self._views.append(view)
            offset += p.numel()

    numel = property(lambda self: self.flat.numel())
    nbytes = property(lambda self: self.flat.numel() * self.flat.element_size())

    def zero_(self):
        """Use this instead of ``optimizer.zero_grad()`` (whose set_to_none default would drop the views)."""
        self.flat.zero_()

    def intact(self):
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self._views))

    def allreduce_(self, async_op=False):
        """SUM over the ranks, in place, one collective."""
        if _world(self.group) == 1:
            return None
        if not self.intact():
            raise RuntimeError("FlatGradientBucket: a parameter's .grad no longer aliases the flat buffer "
                               "(zero_grad(set_to_none=True)?); use bucket.zero_()")
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)


def broadcast_parameters(params, src=0, group=None):
    """Same initial values of the shared parameters on every rank (one flat broadcast)."""
    params = list(params)
    if _world(group) == 1 or not params:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    offset = 0
    with torch.no_grad():
        for p in params:
            p.copy_(flat[offset:offset + p.numel()].view_as(p))
            offset += p.numel()


def owner_rows(N, rank=None, world=None, group=None):
    """[lo, hi): the data points (and rows of the per-sample variational tables) this rank owns."""
    if world is None:
        world = _world(group)
    if rank is None:
        rank = dist.get_rank(group) if world > 1 else 0
    return shard_range(N, rank, world)


class DataParallelSVI(object):
    """One SVI step = local ELBO of this rank's data points -> backward -> one flat all-reduce of the shared gradients ->
    optimizer steps (shared parameters: identical update on every rank; local tables: this rank only)."""

    def __init__(self, shared_params, local_params, elbo_fn, *, lr=1e-3, group=None, capturable=False, optimizer=None):
        self.shared, self.local = list(shared_params), list(local_params)
        self.elbo_fn, self.group = elbo_fn, group
        broadcast_parameters(self.shared, 0, group)
        self.bucket = FlatGradientBucket(self.shared, group)
        make = optimizer if optimizer is not None else (lambda ps: torch.optim.Adam(ps, lr=lr, capturable=capturable))
        self.opt_shared = make(self.shared)
        self.opt_local = make(self.local) if self.local else None
        self.last_elbo = None

    def step(self):
        self.bucket.zero_()
        for p in self.local:
            if p.grad is not None:
                p.grad.zero_()
        elbo = self.elbo_fn()
        (-elbo).backward()
        self.bucket.allreduce_()
        self.opt_shared.step()
        if self.opt_local is not None:
            self.opt_local.step()
        self.last_elbo = elbo.detach()
        return self.last_elbo

    def global_elbo(self):
        """Sum of the ranks' ELBO values of the last step (logging; one scalar all-reduce)."""
        v = self.last_elbo.clone()
        if _world(self.group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        return v


class GraphedStep(object):
    """``DataParallelSVI.step`` captured into one CUDA graph (kernels of the NN and of the physics layer, the NCCL
    all-reduce, capturable Adam).  The ELBO must be free of host synchronisation (``rom.deferred_checks = True``) and draw
    its noise with torch's CUDA generator (graph-safe).  ``replay()`` runs one step; the ELBO value of the step is in
    ``svi.last_elbo`` (a static tensor)."""

    def __init__(self, svi, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        self.svi = svi
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # allocates optimizer state, grads of the local tables, workspaces
                svi.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            svi.step()

    def replay(self):
        self.graph.replay()
        return self.svi.last_elbo
