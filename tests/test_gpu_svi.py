"""BASELINE config 5 on one GPU: the SVI step around the mirrored physics layer (gpde_b200/svi.py + svi_workload.py).
The two-rank data-parallel logic is covered on the CPU by tests/test_svi_gloo.py; with >= 2 GPUs the NCCL path runs here too."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def test_fused_and_dropin_elbo_agree_and_the_step_learns(dev):
    from gpde_b200 import svi
    from gpde_b200.svi_workload import SviWorkload
    vals, grads = {}, {}
    for fused in (True, False):
        wl = SviWorkload(dev, torch.float64, N_s=16, N_vo=8, bs_u=8, seed=1, fused_loglik=fused)
        torch.manual_seed(7)
        e = wl.elbo()
        e.backward()
        wl.g.rom.check()
        vals[fused] = float(e)
        grads[fused] = {n: p.grad.detach().cpu().clone() for n, p in
                        zip(("ls_y", "gp_w", "qX_mean"), (wl.g.logsigmas_y, wl.gp.fc.weight, wl.q_X["supervised"].mean))}
    assert abs(vals[True] - vals[False]) < 1e-10 * abs(vals[False])
    for n in grads[True]:
        assert rel_err(grads[True][n], grads[False][n]) < 1e-9, n
    # a few steps of the (single-rank) data-parallel driver raise the ELBO
    wl = SviWorkload(dev, torch.float32, N_s=16, N_vo=8, bs_u=8, seed=1)
    dp = svi.DataParallelSVI(wl.shared_parameters(), wl.local_parameters(), wl.elbo, lr=1e-2)
    assert dp.bucket.numel == sum(p.numel() for p in wl.shared_parameters()) and dp.bucket.intact()
    first = np.mean([float(dp.step()) for _ in range(3)])
    for _ in range(40):
        dp.step()
    last = np.mean([float(dp.step()) for _ in range(3)])
    wl.g.rom.check()
    assert dp.bucket.intact() and np.isfinite(last) and last > first


def test_graphed_step_and_vo_update(dev):
    """The whole step replayed from one CUDA graph keeps training (same modules, static tensors), and the batched
    virtual-observable update (predictive moments + posterior) feeds the VO term."""
    from gpde_b200 import svi
    from gpde_b200.svi_workload import SviWorkload
    wl = SviWorkload(dev, torch.float32, N_s=16, N_vo=8, bs_u=8, seed=2)
    wl.build_virtual_observables()
    wl.update_virtual_observables(N_mc=16, step=0)
    wl.VO.check()
    assert torch.isfinite(wl.vo_mean).all() and torch.isfinite(wl.vo_logsigma).all()
    dp = svi.DataParallelSVI(wl.shared_parameters(), wl.local_parameters(), wl.elbo, lr=1e-2, capturable=True)
    gs = svi.GraphedStep(dp)
    before = wl.gp.fc.weight.detach().clone()
    vals = []
    for _ in range(30):
        vals.append(float(gs.replay()))
    wl.g.rom.check()
    assert np.all(np.isfinite(vals)) and np.mean(vals[-5:]) > np.mean(vals[:5])
    assert not torch.equal(before, wl.gp.fc.weight.detach())
    wl.update_virtual_observables(N_mc=16, step=1)            # second update: the precision hyper-update path
    wl.VO.check()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import gpde_b200  # noqa: F401
        from gpde_b200 import svi
        from gpde_b200.svi_workload import SviWorkload
        wl = SviWorkload(dev, torch.float32, N_s=16, N_vo=8, bs_u=8, seed=10 + rank)
        dp = svi.DataParallelSVI(wl.shared_parameters(), wl.local_parameters(), wl.elbo, lr=1e-2)
        for _ in range(5):
            dp.step()
        flat = torch.cat([p.detach().reshape(-1) for p in wl.shared_parameters()]).cpu().numpy()
        np.save(os.path.join(out_dir, "shared%d.npy" % rank), flat)
    finally:
        dist.destroy_process_group()


def test_two_gpu_ranks_keep_the_shared_parameters_identical(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the gloo test covers the logic on the CPU)")
    import torch.multiprocessing as mp
    mp.start_processes(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    a, b = (np.load(os.path.join(str(tmp_path), "shared%d.npy" % k)) for k in range(2))
    assert np.array_equal(a, b)
