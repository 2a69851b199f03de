"""Multi-process (world_size 2, gloo, CPU) tests of the sample-sharding host logic:
shards partition the batch, per-shard results reassemble to the single-process result (the oracle stands in
for the kernels on the CPU), and the one real exchange of the path -- the data-point sum of the VO precision
hyper-update -- all-reduces to the serial value."""
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


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpde_b200  # noqa: F401
        from gpde_b200 import sharding
        from gpde_b200.workloads import Workload
        from oracle import fem_p1, rom_ref, vo_ref
        torch.set_num_threads(1)
        w = Workload("cfg1", B=B, seed=0)                       # every rank builds the same seeded workload ...
        full = dict(logX=torch.tensor(w.logX), F=torch.tensor(w.F), gbar=torch.tensor(w.gbar_u),
                    a=torch.tensor(w.log_image), y=torch.tensor(w.y), g=torch.tensor(w.g_fom), V=torch.tensor(w.V))
        mine = sharding.shard_batch(full, B, rank, world)         # ... and keeps its contiguous slice
        lo, hi = sharding.shard_range(B, rank, world)
        assert mine["logX"].shape[0] == hi - lo and mine["V"].shape == full["V"].shape
        P = fem_p1.build_problem(4, 4, 3)
        M, bc = torch.tensor(P["M"]), torch.tensor(P["bc_dofs_rom"])
        u, gX, _ = rom_ref.rom_fwd_adjoint(M, bc, mine["logX"], mine["F"], mine["gbar"])
        r = []
        for n in range(hi - lo):
            X_DG = fem_p1.image_to_function(mine["a"][n:n + 1].numpy().reshape(1, 32, 32), P["pixel_of_cell_fom"])[0]
            K, f = fem_p1.assemble_system_free(P["coords_fom"], P["cells_fom"], np.exp(X_DG), P["bc_dofs_fom"],
                                               mine["g"][n].numpy(), P["free_dofs_fom"])
            r.append(vo_ref.vo_residual(K, f, w.V, mine["y"][n].numpy()))
        r = torch.tensor(np.stack(r)) if r else torch.zeros((0, w.m), dtype=torch.float64)
        u_all = sharding.gather_batch(u, B)
        gX_all = sharding.gather_batch(gX, B)
        r_all = sharding.gather_batch(r, B)
        beta = sharding.vo_precision_sums(r)
        t_max = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), u=u_all.numpy(), gX=gX_all.numpy(), r=r_all.numpy(),
                     beta=beta.numpy(), t_max=t_max)
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    import gpde_b200  # noqa: F401
    from gpde_b200.sharding import shard_range
    for B in (0, 1, 7, 8, 131072, 131073):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


@pytest.mark.parametrize("B", [7, 8])
def test_two_ranks_reproduce_the_single_process_result(tmp_path, B):
    import gpde_b200  # noqa: F401
    from gpde_b200.workloads import Workload
    from oracle import fem_p1, rom_ref, vo_ref
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    w = Workload("cfg1", B=B, seed=0)
    P = fem_p1.build_problem(4, 4, 3)
    u, gX, _ = rom_ref.rom_fwd_adjoint(torch.tensor(P["M"]), torch.tensor(P["bc_dofs_rom"]), torch.tensor(w.logX),
                                       torch.tensor(w.F), torch.tensor(w.gbar_u))
    # samples are independent; the CPU oracle's batched LAPACK calls may round differently per batch size
    assert np.allclose(got["u"], u.numpy(), rtol=1e-12, atol=1e-14)
    assert np.allclose(got["gX"], gX.numpy(), rtol=1e-11, atol=1e-13)
    r = []
    for n in range(B):
        X_DG = fem_p1.image_to_function(w.log_image[n:n + 1].reshape(1, 32, 32), P["pixel_of_cell_fom"])[0]
        K, f = fem_p1.assemble_system_free(P["coords_fom"], P["cells_fom"], np.exp(X_DG), P["bc_dofs_fom"],
                                           w.g_fom[n], P["free_dofs_fom"])
        r.append(vo_ref.vo_residual(K, f, w.V, w.y[n]))
    r = np.stack(r)
    assert np.allclose(got["r"], r, rtol=1e-12, atol=1e-14)
    assert np.allclose(got["beta"], (r ** 2).sum(axis=0), rtol=1e-13, atol=0)
    assert float(got["t_max"]) == 11.0
