#!/bin/bash
# Short A/B session: tools/gpu_quick.sh <tag> "<workloads>" [pytest-args...] -- default build and every build/variants/*.so through tools/quick_ab.py
tag=$1; wls=$2; shift 2
o=gpurun_out; mkdir -p $o
if [ -n "$*" ]; then timeout 900 python -m pytest "$@" -x -q -p no:cacheprovider 2>&1 | tail -15 > $o/${tag}_pytest.log; fi
for w in $wls; do
  python tools/quick_ab.py $w >> $o/${tag}_ab.jsonl 2>> $o/${tag}_ab.err
  for lib in build/variants/librast_b200_*.so; do [ -f "$lib" ] && RAST_LIB=$lib python tools/quick_ab.py $w >> $o/${tag}_ab.jsonl 2>> $o/${tag}_ab.err; done
done
python - <<PY > $o/${tag}_ab.txt
import json
for l in open("$o/${tag}_ab.jsonl"):
    d = json.loads(l); print(d["workload"], d["lib"], d["ms_per_call"], d["pass_ms_per_call"], d["hash_rgb"], d["hash_depth"], d["hash_ids"])
PY
cat $o/${tag}_ab.txt; [ -f $o/${tag}_pytest.log ] && cat $o/${tag}_pytest.log
