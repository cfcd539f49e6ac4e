#!/bin/bash
# Kernel variants on the GPU box: parity suite + timings per variant, one gpurun call.
#   here:   python tools/build_variants.py legacy:-DRAST_TIGHT_TINY=0,-DRAST_SHADE_PREP=0 t2:-DRAST_SETUP_TRIS=2 ...
#   there:  tools/gpu_variants.sh r2v legacy t2        (results under gpurun_out/<tag>_*)
# A variant becomes the default only if its whole -m gpu suite is green (bit-exact against the oracle and the reference's
# golden hashes, full-size configs included) AND it is faster.
tag=${1:-r2v}
o=gpurun_out
mkdir -p $o
for v in "${@:2}"; do
  lib=build/variants/librast_b200_$v.so
  [ -f $lib ] || { echo "build $lib first (tools/build_variants.py)"; continue; }
  RAST_LIB=$lib RAST_FUZZ_SEEDS=${FUZZ_SEEDS:-400} python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $o/${tag}_pytest_$v.log
done
WORKLOADS="${WORKLOADS:-spin1080p suzanne640 tess4k tess4k_64lights overdraw8k}" tools/gpu_ab.sh $tag "${@:2}"
cat $o/${tag}_pytest_*.log $o/${tag}_summary.txt
