"""examples/shard_client.cpp: a C++ program that shards one bundle-adjustment problem over N GPUs (one process per GPU) through
the C ABI alone -- gbp_bal_*, gbp_ba_create with its share of the global landmark chunking, gbp_comm_create / gbp_ba_attach_comm
(NCCL inside the library), gbp_ba_iterate, gbp_ba_metrics.  Its printed numbers must equal the single-GPU Python engine's with
the same chunking (the multi-GPU iteration is the same floating-point program)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)


@pytest.fixture(scope="module")
def client(tmp_path_factory, built_library):
    exe = str(tmp_path_factory.mktemp("shard") / "shard_client")
    lib_dir = os.path.join(ROOT, "gbp_b200", "lib")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "examples", "shard_client.cpp"), "-I" + os.path.join(ROOT, "include"),
                    "-L" + lib_dir, "-lgbp_b200", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("nranks", [1, 2])
def test_cpp_client_shards_like_the_python_engine(client, tmp_path, nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    from gbp_b200 import balio
    from gbp_b200 import _lib as L
    from gbp_b200.engine import BAEngine
    G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk_vsmall.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
    bal = str(tmp_path / "problem.txt")
    balio.write_bal(bal, prob)
    iters = 20
    env = dict(os.environ)                       # the client is not a torch process: the system's NCCL or $GBP_NCCL_LIB, either will do
    procs = [subprocess.Popen([client, "--rank", str(r), "--nranks", str(nranks), "--id-file", str(tmp_path / "nccl.id"), "--bal", bal,
                               "--iters", str(iters)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for r in range(nranks)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se
    text = outs[0][0]
    m = re.search(r"after\s+%d ARE ([-\d.e+]+)\s+energy ([-\d.e+]+)\s+relinearising (\d+)" % iters, text)
    assert m, text
    are_c, energy_c, nrel_c = float(m.group(1)), float(m.group(2)), int(m.group(3))
    mu_c = np.array([float(x) for x in re.search(r"keyframe 0 mean (.*)", text).group(1).split()])

    # the same solve on ONE GPU through the Python binding, with the chunking the ranks used (nranks chunks)
    e = BAEngine(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG,
                 chunks=(nranks, 0, nranks, 0, prob.n_points))
    e.generate_priors(50.0)
    e.update_beliefs()
    e.iterate(3, robustify=True, local_relin=True)
    e.fill_iters(1)
    e.iterate(5, robustify=True, local_relin=True)
    e.fill_iters(1)
    e.iterate(iters - 8, robustify=True, local_relin=True)
    a, en, nrel = e.metrics()
    mu = e.read(L.F_CAM_MU)[0]
    e.close()
    assert nrel_c == nrel
    assert abs(are_c - a / prob.n_edges) <= 2e-12 * max(1.0, abs(are_c))          # printed with 12 decimals
    assert abs(energy_c - en) <= 2e-9 * max(1.0, abs(en))
    assert np.allclose(mu_c, mu, rtol=0, atol=2e-12)
