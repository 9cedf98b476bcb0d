#!/bin/bash
# A/B of two BUILDS of the library on one box (boxes of the pool differ by up to 13 %, so only same-box numbers compare):
# alternates gbp_b200/lib/libgbp_b200_prev.so (built from the previous commit) and the in-tree build on the 10 M-factor graph.
#   REPS=2 ARGS="--synthetic --tiles 64" LIBS="gbp_b200/lib/libgbp_b200_prev.so gbp_b200/lib/libgbp_b200.so" bash scripts/ab_builds.sh
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
REPS=${REPS:-2}
ARGS=${ARGS:-"--synthetic --iters 30"}
for r in $(seq 1 $REPS); do
  for lib in ${LIBS:-gbp_b200/lib/libgbp_b200_prev.so gbp_b200/lib/libgbp_b200.so}; do
    echo "== $lib (rep $r)"
    timeout 300 python scripts/ab_variants.py --lib $lib $ARGS 2>&1 | cut -c1-600
  done
done
