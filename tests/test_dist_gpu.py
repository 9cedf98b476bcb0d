"""2-GPU NCCL run of the landmark-partitioned sweep against the single-GPU engine (needs >= 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
           eta_damping=0.4, prior_std_weaker_factor=50.0)


def _worker(rank, world, port, out_dir, p2p=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    from gbp_b200.dist import PartitionedBAGraph
    from gbp_b200.synthetic import make_synthetic
    prob = make_synthetic(40, 6000, 8, seed=5)
    pg = PartitionedBAGraph(prob, CFG, rank=rank, world=world, device=rank, stream=s.cuda_stream, dist=dist, torch_stream=s,
                            p2p=p2p)
    pg.generate_priors_var(50.0)
    pg.update_all_beliefs()
    trace = []
    for i in range(25):
        if i in (3, 8):
            pg.fill_iters(1)
        trace.append(pg.metrics())
        if i == 12:
            assert pg.capture(local_relin=True, robustify=True)   # from here on an iteration is one graph replay; no state change
        pg.synchronous_iteration(robustify=True, local_relin=True)
    trace.append(pg.metrics())
    status = pg.adapter.p2p_status() if p2p else (0, 0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), means=pg.get_means(), trace=np.array(trace), p2p_status=np.array(status),
             n_iterations=pg.n_iterations)
    pg.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True], ids=["nccl", "p2p"])
def test_two_gpu_partition_matches_single_gpu(tmp_path, p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from gbp_b200.dist import PartitionedBAGraph
    from gbp_b200.synthetic import make_synthetic
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), p2p), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["means"], r1["means"]) and np.array_equal(r0["trace"], r1["trace"])
    assert int(r0["n_iterations"]) == 25
    if p2p:      # one exchange per belief update (1 initial + 25 iterations), no wait timed out
        assert r0["p2p_status"][0] == 26 and r0["p2p_status"][1] == 0 and r1["p2p_status"][1] == 0
    prob = make_synthetic(40, 6000, 8, seed=5)
    pg = PartitionedBAGraph(prob, CFG)
    pg.generate_priors_var(50.0)
    pg.update_all_beliefs()
    trace = []
    for i in range(25):
        if i in (3, 8):
            pg.fill_iters(1)
        trace.append(pg.metrics())
        pg.synchronous_iteration(robustify=True, local_relin=True)
    trace.append(pg.metrics())
    trace = np.array(trace)
    assert np.array_equal(trace[:, 2], r0["trace"][:, 2])
    assert relerr(r0["trace"][:, :2], trace[:, :2]) < 1e-9
    assert relerr(r0["means"], pg.get_means()) < 1e-9
    pg.close()
