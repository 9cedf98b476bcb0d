#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE (joeaortiz/gbp).

This script is test infrastructure.  It is the only file in the repo that imports
``/root/reference``; it runs in the build container (the GPU box has no reference
checkout), and the ``.npz`` files it writes are committed next to it.

It drives the reference exactly as ``ba.py:51-105`` does (minus the three viewer
lines ``ba.py:79-81,103`` - trimesh/pyglet are not installed) and records

* the parsed BAL inputs (``utils/read_balfile.py:4-37``),
* linearised factors after ``create_ba_graph`` (``gbp/gbp_ba.py:97-150``),
* priors after ``generate_priors_var`` (``gbp/gbp_ba.py:20-34``),
* per-outer-iteration ARE / energy / relinearisation count (``ba.py:95-101``),
* beliefs, messages, linearisation points and per-factor control state after
  selected sweeps of ``synchronous_iteration`` (``gbp/gbp.py:86-92``).

Usage:  python tests/golden/make_golden.py [case ...]
Cases:  vsmall vsmall_huber vsmall_constant vsmall_float fr1desk posegraph synth_small fr2robot2 fr1xyz_av fr1desk_small
"""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np

REF = os.environ.get("GBP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.simplefilter("ignore")
    from gbp import gbp_ba  # noqa: the REFERENCE package, not ours
    from utils import read_balfile
    assert os.path.realpath(gbp_ba.__file__).startswith(os.path.realpath(REF))
    return gbp_ba, read_balfile


def default_configs(**over):
    # ba.py:14-44 defaults
    cfg = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6,
               min_linear_iters=8, eta_damping=0.4, prior_std_weaker_factor=50.0)
    cfg.update(over)
    return cfg


def pack_vars(nodes):
    eta = np.concatenate([n.belief.eta for n in nodes])
    lam = np.concatenate([np.asarray(n.belief.lam).ravel() for n in nodes])
    mu = np.concatenate([np.asarray(n.mu, dtype=float) for n in nodes])
    return eta, lam, mu


def snapshot(graph, fsample):
    """State after a sweep: all beliefs, sampled messages/linpoints, all control ints."""
    out = {}
    ce, cl, cm = pack_vars(graph.cam_nodes)
    le, ll, lm = pack_vars(graph.lmk_nodes)
    out.update(cam_eta=ce, cam_lam=cl, cam_mu=cm, lmk_eta=le, lmk_lam=ll, lmk_mu=lm)
    fs = [graph.factors[i] for i in fsample]
    out["msg_cam_eta"] = np.stack([f.messages[0].eta for f in fs])
    out["msg_cam_lam"] = np.stack([f.messages[0].lam for f in fs])
    out["msg_lmk_eta"] = np.stack([f.messages[1].eta for f in fs])
    out["msg_lmk_lam"] = np.stack([f.messages[1].lam for f in fs])
    out["linpoint"] = np.stack([np.asarray(f.linpoint, dtype=float) for f in fs])
    out["factor_eta"] = np.stack([f.factor.eta for f in fs])
    out["factor_lam"] = np.stack([f.factor.lam for f in fs])
    out["iters_since_relin"] = np.array([f.iters_since_relin for f in graph.factors], dtype=np.int32)
    out["eta_damping"] = np.array([f.eta_damping for f in graph.factors], dtype=np.float64)
    out["adaptive_var"] = np.array([f.adaptive_gauss_noise_var for f in graph.factors], dtype=np.float64)
    return out


def run_ba(name, bal_file, n_iters, checkpoints, float_impl=False, nsample=64, **cfg_over):
    gbp_ba, read_balfile = _import_reference()
    cfg = default_configs(**cfg_over)
    t0 = time.time()
    (n_kf, n_pts, n_edges, cam_means, lmk_means, meas, cam_ids, lmk_ids, K) = read_balfile.read_balfile(bal_file)
    G = {}
    G["in_cam_id"] = np.asarray(cam_ids, dtype=np.int32)
    G["in_lmk_id"] = np.asarray(lmk_ids, dtype=np.int32)
    G["in_z"] = np.asarray(meas, dtype=np.float64)
    G["in_cam0"] = np.asarray(cam_means, dtype=np.float64)
    G["in_lmk0"] = np.asarray(lmk_means, dtype=np.float64)
    G["in_K"] = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    G["cfg_keys"] = np.array(sorted(cfg.keys()))
    G["cfg_vals"] = np.array([str(cfg[k]) for k in sorted(cfg.keys())])
    G["float_impl"] = np.array(int(float_impl))

    graph = gbp_ba.create_ba_graph(bal_file, cfg)
    F = len(graph.factors)
    # factor order of the reference (camera-major scan, gbp_ba.py:128-130) as (cam, lmk) ids
    n_cam = len(graph.cam_nodes)
    G["factor_cam"] = np.array([f.adj_vIDs[0] for f in graph.factors], dtype=np.int32)
    G["factor_lmk"] = np.array([f.adj_vIDs[1] - n_cam for f in graph.factors], dtype=np.int32)
    fsample = np.unique(np.linspace(0, F - 1, nsample).astype(int))
    G["fsample"] = fsample
    G["init_factor_eta"] = np.stack([graph.factors[i].factor.eta for i in fsample])
    G["init_factor_lam"] = np.stack([graph.factors[i].factor.lam for i in fsample])
    G["init_linpoint"] = np.stack([np.asarray(graph.factors[i].linpoint, float) for i in fsample])
    # max entry of every factor's Lambda (what generate_priors_var consumes)
    G["init_factor_lam_max"] = np.array([np.max(f.factor.lam) for f in graph.factors])

    graph.generate_priors_var(weaker_factor=cfg["prior_std_weaker_factor"])
    G["prior_cam_lam00"] = np.array([n.prior.lam[0, 0] for n in graph.cam_nodes])
    G["prior_lmk_lam00"] = np.array([n.prior.lam[0, 0] for n in graph.lmk_nodes])
    graph.update_all_beliefs()
    for k, v in snapshot(graph, fsample).items():
        G[f"s_init_{k}"] = v

    # ba.py:62-65
    final_weaker, n_weak = 100.0, 5
    weakening_factor = np.log10(final_weaker) / n_weak

    are_t, en_t, nrel_t = [], [], []
    for i in range(n_iters):
        if float_impl and (i + 1) % 2 == 0 and (i < n_weak * 2):      # ba.py:86-88
            graph.weaken_priors(weakening_factor)
        if i == 3 or i == 8:                                           # ba.py:91-93
            for factor in graph.factors:
                factor.iters_since_relin = 1
        are_t.append(graph.are())                                      # ba.py:95-100
        en_t.append(graph.energy())
        nrel_t.append(sum(1 for f in graph.factors if f.iters_since_relin == 0))
        graph.synchronous_iteration(robustify=True, local_relin=True)  # ba.py:105
        if i in checkpoints:
            for k, v in snapshot(graph, fsample).items():
                G[f"s{i}_{k}"] = v
        if i % 10 == 0:
            print(f"[{name}] iter {i} ARE {are_t[-1]:.6f} energy {en_t[-1]:.4f} relin {nrel_t[-1]} "
                  f"({time.time() - t0:.0f}s)", flush=True)
    # metrics after the last sweep as well
    are_t.append(graph.are())
    en_t.append(graph.energy())
    nrel_t.append(sum(1 for f in graph.factors if f.iters_since_relin == 0))
    G["are"] = np.array(are_t)
    G["energy"] = np.array(en_t)
    G["n_relin"] = np.array(nrel_t, dtype=np.int64)
    G["checkpoints"] = np.array(sorted(checkpoints), dtype=np.int64)
    G["n_iters"] = np.array(n_iters)
    out = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(out, **G)
    print(f"[{name}] wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB) in {time.time() - t0:.0f}s", flush=True)


def run_posegraph(name, argv):
    """Run the unmodified ndim_posegraph.py and capture its printed trace (config 1)."""
    import subprocess
    env = dict(os.environ, PYTHONPATH=REF)
    res = subprocess.run([sys.executable, os.path.join(REF, "ndim_posegraph.py")] + argv,
                         cwd=REF, env=env, capture_output=True, text=True, check=True)
    lines = [l for l in res.stdout.splitlines() if l.startswith("Iteration")]
    energy = np.array([float(l.split("Energy")[1].split("//")[0]) for l in lines])
    dist = np.array([float(l.split("MAP")[1]) for l in lines])
    out = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(out, argv=np.array(argv), energy=energy, dist=dist, stdout=np.array(res.stdout))
    print(f"[{name}] wrote {out}: energy[0]={energy[0]} energy[-1]={energy[-1]} dist[-1]={dist[-1]}")


def main():
    cases = sys.argv[1:] or ["vsmall", "vsmall_huber", "vsmall_constant", "vsmall_float", "posegraph", "fr1desk"]
    data = os.path.join(REF, "data")
    for c in cases:
        if c == "vsmall":
            run_ba("fr1desk_vsmall", f"{data}/fr1desk_vsmall.txt", 60, {0, 1, 2, 7, 14, 15, 16, 24, 59})
        elif c == "vsmall_huber":
            run_ba("fr1desk_vsmall_huber", f"{data}/fr1desk_vsmall.txt", 40, {0, 1, 15, 16, 39}, loss="huber")
        elif c == "vsmall_constant":
            run_ba("fr1desk_vsmall_constant", f"{data}/fr1desk_vsmall.txt", 40, {0, 1, 15, 16, 39}, loss="constant")
        elif c == "vsmall_float":
            run_ba("fr1desk_vsmall_float", f"{data}/fr1desk_vsmall.txt", 30, {0, 1, 9, 16, 29}, float_impl=True)
        elif c == "fr1desk":
            run_ba("fr1desk", f"{data}/fr1desk.txt", 200, {0, 15, 16, 99, 199}, nsample=48)
        elif c == "posegraph":
            run_posegraph("posegraph_n50_d3", ["--n_varnodes", "50", "--dim", "3"])
            run_posegraph("posegraph_default", [])
        elif c == "fr2robot2":      # another sequence: a robot-mounted camera, other intrinsics and motion
            run_ba("fr2robot2", f"{data}/fr2robot2.txt", 40, {0, 1, 15, 16, 39})
        elif c == "fr1xyz_av":      # another sequence: pure translation (small rotations)
            run_ba("fr1xyz_av", f"{data}/fr1xyz_av.txt", 30, {0, 1, 15, 16, 29}, nsample=48)
        elif c == "fr1desk_small":
            run_ba("fr1desk_small", f"{data}/fr1desk_small.txt", 40, {0, 1, 15, 16, 39})
        elif c == "synth_small":
            bal = os.path.join(HERE, "synth_small.txt")
            run_ba("synth_small", bal, 30, {0, 1, 15, 16, 29})
        else:
            raise SystemExit(f"unknown case {c}")


if __name__ == "__main__":
    main()
