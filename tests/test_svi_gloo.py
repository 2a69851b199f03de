"""Data-parallel SVI plumbing (gpde_b200/svi.py; the reference's single-process loop is training.py:393-462) with two gloo
ranks on the CPU: owner-sharded per-sample tables + ONE flat all-reduce (SUM) of the shared gradients reproduce the
single-process step on the union of the data; the flat bucket keeps aliasing the parameters' gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy(N, rows, seed=0):
    """Shared linear map + per-sample table rows [lo,hi); 'ELBO' = -sum_n |A z_n - y_n|^2 - 0.1 |z_n|^2 over the owned rows."""
    g = torch.Generator().manual_seed(seed)
    Y = torch.randn(N, 5, generator=g, dtype=torch.float64)
    Z0 = torch.randn(N, 3, generator=g, dtype=torch.float64)
    lo, hi = rows
    lin = torch.nn.Linear(3, 5).double()
    with torch.no_grad():
        lin.weight.copy_(torch.randn(5, 3, generator=g, dtype=torch.float64))
        lin.bias.copy_(torch.randn(5, generator=g, dtype=torch.float64))
    table = torch.nn.Parameter(Z0[lo:hi].clone())

    def elbo():
        return -((lin(table) - Y[lo:hi]) ** 2).sum() - 0.1 * (table ** 2).sum()
    return lin, table, elbo


def _worker(rank, world, port, N, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpde_b200  # noqa: F401
        from gpde_b200 import svi
        torch.set_num_threads(1)
        rows = svi.owner_rows(N)
        lin, table, elbo = _toy(N, rows)
        if rank == 1:                                       # a rank that starts from different shared values ...
            with torch.no_grad():
                lin.weight.add_(1.0)
        opt = lambda ps: torch.optim.SGD(ps, lr=0.01)
        dp = svi.DataParallelSVI(lin.parameters(), [table], elbo, optimizer=opt)     # ... is overwritten by rank 0's
        assert dp.bucket.numel == 20 and dp.bucket.intact()
        values = [float(dp.step()) for _ in range(3)]
        assert dp.bucket.intact()
        total = float(dp.global_elbo())
        lin.zero_grad(set_to_none=True)                     # dropping the views must be noticed, not silently un-reduced
        try:
            dp.bucket.allreduce_()
            noticed = False
        except RuntimeError:
            noticed = True
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), w=lin.weight.detach().numpy(), b=lin.bias.detach().numpy(),
                 table=table.detach().numpy(), rows=np.array(rows), values=np.array(values), total=total, noticed=noticed)
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_step(tmp_path):
    import gpde_b200  # noqa: F401
    from gpde_b200 import svi
    N, world = 11, 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    # single process on all rows
    lin, table, elbo = _toy(N, (0, N))
    dp = svi.DataParallelSVI(lin.parameters(), [table], elbo, optimizer=lambda ps: torch.optim.SGD(ps, lr=0.01))
    serial = [float(dp.step()) for _ in range(3)]
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    for k in range(world):
        assert np.allclose(r[k]["w"], lin.weight.detach().numpy(), rtol=1e-12, atol=1e-14)
        assert np.allclose(r[k]["b"], lin.bias.detach().numpy(), rtol=1e-12, atol=1e-14)
        lo, hi = r[k]["rows"]
        assert np.allclose(r[k]["table"], table.detach().numpy()[lo:hi], rtol=1e-12, atol=1e-14)
        assert bool(r[k]["noticed"])
    assert np.array_equal(r[0]["rows"], [0, 6]) and np.array_equal(r[1]["rows"], [6, 11])
    assert np.allclose(r[0]["values"] + r[1]["values"], serial, rtol=1e-12)
    assert abs(float(r[0]["total"]) - serial[-1]) < 1e-9 * abs(serial[-1])


def test_flat_bucket_single_process():
    import gpde_b200  # noqa: F401
    from gpde_b200 import svi
    lin = torch.nn.Linear(4, 2)
    frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
    b = svi.FlatGradientBucket(list(lin.parameters()) + [frozen])
    assert b.numel == 10 and b.nbytes == 40 and b.intact()
    lin(torch.ones(1, 4)).sum().backward()
    assert b.intact() and float(b.flat.abs().sum()) > 0            # autograd accumulated INTO the flat buffer
    assert b.allreduce_() is None                                  # no process group: nothing to do
    b.zero_()
    assert float(lin.weight.grad.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        svi.FlatGradientBucket([torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(2, dtype=torch.float64))])
    with pytest.raises(ValueError):
        svi.FlatGradientBucket([frozen])
    assert svi.owner_rows(10, 1, 3) == (4, 7)
