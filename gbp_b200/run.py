"""Run an UNMODIFIED reference client script (ba.py, ndim_posegraph.py) against this engine:

    python -m gbp_b200.run /path/to/reference/ba.py --bal_file data/fr1desk.txt

puts gbp_b200/compat (the reference-named packages gbp / utils / vis) first on sys.path and
executes the script with runpy, so `from gbp import gbp_ba` / `import vis` resolve here."""
import os
import runpy
import sys

COMPAT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, COMPAT):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [COMPAT, root]
    for name in [m for m in sys.modules if m in ("gbp", "utils", "vis") or m.startswith(("gbp.", "utils.", "vis."))]:
        del sys.modules[name]
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
