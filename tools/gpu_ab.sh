#!/bin/bash
# A/B runs of kernel variants on the GPU box: tools/gpu_ab.sh <tag>  (results under gpurun_out/<tag>_*)
tag=${1:-ab}
mkdir -p gpurun_out
run() { # name lib workload
  RAST_LIB=$2 python bench.py --workload $3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/${tag}_$1_$3.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['roofline']['pass_ms_per_step']
print('$1 $3 ms/step', round(d['ms_per_step'],4), 'vertex', round(p['vertex'],4), 'setup', round(p['setup'],4), 'raster', round(p['raster'],4), 'shade', round(p['shade'],4), 'checksum', d['checksum'])" | tee -a gpurun_out/${tag}_summary.txt
}
for w in ${WORKLOADS:-spin1080p tess4k tess4k_64lights}; do
  run default "" $w
  for v in "${@:2}"; do run $v build/variants/librast_b200_$v.so $w; done
done
