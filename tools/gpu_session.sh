#!/bin/bash
# One measurement session on the GPU box:  tools/gpu_session.sh <tag> [steps...]   (everything lands in gpurun_out/<tag>_*)
#   steps: pytest smoke ab bench ref launches ncu   (default: all);  variants under build/variants/ are A/B-timed by `ab`
tag=${1:-s}; shift
steps=${*:-pytest smoke ab bench ref launches ncu}
o=gpurun_out
mkdir -p $o
has() { [[ " $steps " == *" $1 "* ]]; }
if has pytest; then timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > $o/${tag}_pytest.log; fi
if has smoke; then python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; fi
if has ab; then
  ab() { env RAST_LIB=${1:+build/variants/librast_b200_$1.so} python tools/quick_ab.py $2 >> $o/${tag}_ab.jsonl 2>> $o/${tag}_ab.err; }
  for w in spin1080p overdraw8k tess4k tess4k_64lights suzanne640; do ab "" $w; done
  for lib in build/variants/librast_b200_*.so; do
    [ -f "$lib" ] || continue
    v=$(basename $lib .so); v=${v#librast_b200_}
    for w in ${AB_WORKLOADS:-spin1080p overdraw8k tess4k}; do ab $v $w; done
  done
  python - <<PY > $o/${tag}_ab.txt
import json
for l in open("$o/${tag}_ab.jsonl"):
    d = json.loads(l); print(d["workload"], d["lib"], d["ms_per_call"], d["pass_ms_per_call"], d["hash_rgb"], d["hash_depth"], d["hash_ids"])
PY
fi
if has ref; then python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err; fi
if has bench; then python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; fi
if has launches; then
  # launch list of one bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $o/${tag}_launches_bench.log 2>&1
fi
if has ncu; then
  cap() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o $o/${tag}_$3 python tools/quick_ab.py $4 --calls 2 > $o/${tag}_ncu_$3.log 2>&1; }
  cap k_resolve_shade 4 shade spin1080p
  cap k_raster_chunks 4 raster spin1080p
  cap k_vertex 4 vertex spin1080p
  cap k_setup 4 setup_spin spin1080p
  cap k_prepare_tris 4 prepare spin1080p
  cap k_setup 5 setup tess4k
  cap k_setup 5 setup50m tess4k_64lights
  cap k_raster_tiles 3 raster_over overdraw8k
  cap k_resolve_shade 3 shade_over overdraw8k
  cap k_resolve_shade 5 shade_tess tess4k
fi
ls -la $o | grep $tag
