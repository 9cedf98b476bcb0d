#!/usr/bin/env python
"""Bundle adjustment with GBP on the B200 engine: a client with the same flags, call sequence and
printed trace as the reference's ba.py (ba.py:10-105), written against the reference-named packages
in gbp_b200/compat.  (The reference script itself also runs unmodified:
`python -m gbp_b200.run /path/to/reference/ba.py --bal_file ...`; it is not shipped here.)

    python examples/ba_client.py --bal_file problem.txt [--n_iters 200] [--loss huber] [--float_implementation]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gbp_b200", "compat"), ROOT]

from gbp import gbp_ba  # noqa: E402
import vis  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--bal_file", required=True)
    ap.add_argument("--n_iters", type=int, default=200)
    ap.add_argument("--gauss_noise_std", type=int, default=2)
    ap.add_argument("--loss", default=None)
    ap.add_argument("--Nstds", type=float, default=3.)
    ap.add_argument("--beta", type=float, default=0.01)
    ap.add_argument("--num_undamped_iters", type=int, default=6)
    ap.add_argument("--min_linear_iters", type=int, default=8)
    ap.add_argument("--eta_damping", type=float, default=0.4)
    ap.add_argument("--prior_std_weaker_factor", type=float, default=50.)
    ap.add_argument("--float_implementation", action="store_true", default=False)
    ap.add_argument("--final_prior_std_weaker_factor", type=float, default=100.)
    ap.add_argument("--num_weakening_steps", type=int, default=5)
    args = ap.parse_args(argv)
    print("Configs: \n", args)
    configs = {k: getattr(args, k) for k in ("gauss_noise_std", "loss", "Nstds", "beta", "num_undamped_iters",
                                             "min_linear_iters", "eta_damping", "prior_std_weaker_factor")}
    weakening = np.log10(args.final_prior_std_weaker_factor) / args.num_weakening_steps

    graph = gbp_ba.create_ba_graph(args.bal_file, configs)
    print(f"\nData: {args.bal_file}\n")
    print(f"Number of keyframes: {len(graph.cam_nodes)}")
    print(f"Number of landmarks: {len(graph.lmk_nodes)}")
    print(f"Number of measurement factors: {len(graph.factors)}\n")
    graph.generate_priors_var(weaker_factor=args.prior_std_weaker_factor)
    graph.update_all_beliefs()
    scene = vis.ba_vis.create_scene(graph)
    viewer = vis.ba_vis.TrimeshSceneViewer(scene=scene, resolution=scene.camera.resolution)
    viewer.show()

    trace = []
    for i in range(args.n_iters):
        if args.float_implementation and (i + 1) % 2 == 0 and i < args.num_weakening_steps * 2:
            print("Weakening priors")
            graph.weaken_priors(weakening)
        if i == 3 or i == 8:
            for factor in graph.factors:
                factor.iters_since_relin = 1
        are, energy = graph.are(), graph.energy()
        n_relins = sum(1 for factor in graph.factors if factor.iters_since_relin == 0)
        print(f"Iteration {i} // ARE {are:.4f} // Energy {energy:.4f} // Num factors relinearising {n_relins}")
        trace.append((are, energy, n_relins))
        viewer.update(graph)
        graph.synchronous_iteration(robustify=True, local_relin=True)
    return graph, trace


if __name__ == "__main__":
    main()
