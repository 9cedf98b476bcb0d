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


def _worker(rank, world, port, out_dir):
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
    pg = PartitionedBAGraph(prob, CFG, rank=rank, world=world, device=rank, stream=s.cuda_stream, dist=dist, torch_stream=s)
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
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), means=pg.get_means(), trace=np.array(trace), n_iterations=pg.n_iterations)
    pg.close()
    from gbp_b200.dist import shutdown
    shutdown()
    dist.destroy_process_group()


def test_two_gpu_partition_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from gbp_b200.dist import PartitionedBAGraph
    from gbp_b200.synthetic import make_synthetic
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["means"], r1["means"]) and np.array_equal(r0["trace"], r1["trace"])
    assert int(r0["n_iterations"]) == 25
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
    # The same layout choices and the same landmark chunking on one GPU: the keyframe-side sums are associated chunk by chunk
    # everywhere, so the 2-GPU run and the 1-GPU run are the SAME floating-point program: bit-identical means.
    from gbp_b200.dist import global_layout
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L_
    layout, k_total = global_layout(prob, 2)
    lanes = layout.pop("belief_lanes")
    g = create_ba_graph(prob, CFG, chunks=(k_total, 0, k_total, 0, prob.n_points), **layout)
    g._eng.tune(L_.TUNE_BELIEF_LANES, lanes)
    g.generate_priors_var(50.0)
    g.update_all_beliefs()
    for i in range(25):
        if i in (3, 8):
            g.reset_iters_since_relin(1)
        g.synchronous_iteration(robustify=True, local_relin=True)
    assert np.array_equal(g.get_means(), r0["means"])
    g.close()


def test_native_exchange_world_of_one_matches_plain_engine(built_library):
    """The library's own NCCL path (gbp_ba_attach_comm) on ONE GPU: a communicator of one rank runs the complete distributed
    iteration (side stream, chunk sums, ncclAllGather, keyframe update, cross-rank metric sums) and must reproduce the plain
    engine bit for bit.  Runs on the driver's single-GPU box, where the 2-GPU test is skipped."""
    from gbp_b200.engine import BAEngine, Communicator
    from gbp_b200.synthetic import make_synthetic
    from gbp_b200 import _lib as L
    prob = make_synthetic(30, 3000, 8, seed=7)
    comm = Communicator(Communicator.unique_id(), 0, 1, device=0)
    args = (prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG)
    chunks = (4, 0, 4, 0, prob.n_points)
    out = []
    for with_comm in (False, True):
        e = BAEngine(*args, chunks=chunks)
        if with_comm:
            e.attach_comm(comm)
            assert e.comm_info() == (0, 1)
            with pytest.raises(L.GbpError):
                comm.destroy()                                 # refused while a graph is attached
        e.generate_priors(50.0)
        e.update_beliefs()
        e.iterate(3, robustify=True, local_relin=True)
        e.fill_iters(1)
        e.iterate(17, robustify=True, local_relin=True)        # REPS-of-8 graphs and single-iteration graphs
        m = e.metrics()
        out.append((e.read(L.F_CAM_MU).copy(), e.read(L.F_LMK_MU).copy(), m, e.read(L.F_CAM_BELIEF).copy()))
        if with_comm:
            n0 = e.launch_count()
            e.iterate(1, robustify=True, local_relin=True)
            assert e.launch_count() - n0 == 4                  # sweep, keyframe chunk sums, landmark beliefs, keyframe update
            e.detach_comm()
            assert e.comm_info() == (0, 1)
            e.iterate(1, robustify=True, local_relin=True)     # back to the plain two-kernel iteration
        e.close()
    comm.destroy()
    (c0, l0, m0, b0), (c1, l1, m1, b1) = out
    assert np.array_equal(c0, c1) and np.array_equal(l0, l1) and np.array_equal(b0, b1)
    assert m0 == m1


def test_communicator_argument_checks(built_library):
    from gbp_b200.engine import Communicator
    from gbp_b200._lib import GbpError
    with pytest.raises(ValueError):
        Communicator(b"short", 0, 1)
    with pytest.raises(GbpError):
        Communicator(b"\0" * 128, 2, 2)        # rank out of range: refused before NCCL is asked
