#!/bin/bash
# First GPU session of the next round (about 6-8 minutes): the whole suite and the benchmark on the shipped default, then the
# prepared kernel variants -- parity files and quick A/B with output hashes -- and fresh ncu captures.
#   here first:  python tools/build_variants.py legacy:-DRAST_TIGHT_TINY=0,-DRAST_SHADE_PREP=0 blockz:-DRAST_BLOCK_Z=1 ptrs:-DRAST_SHADE_PTRS=1 \
#                    blockz_ptrs:-DRAST_BLOCK_Z=1,-DRAST_SHADE_PTRS=1 g12:-DRAST_SHADE_GROUPS=12 g15:-DRAST_SHADE_GROUPS=15 g20:-DRAST_SHADE_GROUPS=20 \
#                    t256g8:-DRAST_SHADE_THREADS=256,-DRAST_SHADE_GROUPS=8
#   there:       tools/gpu_round2_first.sh r2a
t=${1:-r2a}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 > $o/${t}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke.log 2>&1
ab() { env RAST_LIB=${1:+build/variants/librast_b200_$1.so} $3 python tools/quick_ab.py $2 >> $o/${t}_ab.jsonl 2>> $o/${t}_ab.err; }
for v in "" legacy blockz ptrs blockz_ptrs; do
  [ -z "$v" ] || [ -f build/variants/librast_b200_$v.so ] || continue
  for w in spin1080p overdraw8k tess4k suzanne640; do ab "$v" $w; done
done
for v in g12 g15 g20 t256g8; do  # shade-pass geometry with the prepared records (the optimum may have moved)
  [ -f build/variants/librast_b200_$v.so ] && ab $v spin1080p
done
for v in blockz ptrs blockz_ptrs; do
  [ -f build/variants/librast_b200_$v.so ] || continue
  RAST_LIB=build/variants/librast_b200_$v.so timeout 120 python -m pytest tests/test_parity_gpu.py tests/test_parity_gpu_fuzz.py tests/test_parity_gpu_large.py -x -q -p no:cacheprovider 2>&1 | tail -3 > $o/${t}_pytest_$v.log
done
python bench.py --impl reference > $o/${t}_bench_reference.json 2> $o/${t}_bench_reference.err
python bench.py > $o/${t}_bench_n1.json 2> $o/${t}_bench_n1.err
for w in suzanne640 tess4k tess4k_64lights overdraw8k; do python bench.py --workload $w --steps 10 >> $o/${t}_workloads.jsonl 2>> $o/${t}_workloads.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${t}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $o/${t}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resolve_shade -s 12 -c 1 -f -o $o/${t}_shade python tools/quick_ab.py spin1080p --calls 2 > $o/${t}_ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_chunks -s 12 -c 1 -f -o $o/${t}_raster python tools/quick_ab.py spin1080p --calls 2 > $o/${t}_ncu_raster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_setup -s 5 -c 1 -f -o $o/${t}_setup python tools/quick_ab.py tess4k --calls 2 > $o/${t}_ncu_setup.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_chunks -s 5 -c 1 -f -o $o/${t}_raster_over python tools/quick_ab.py overdraw8k --calls 2 > $o/${t}_ncu_raster_over.log 2>&1
cat $o/${t}_pytest*.log; python - <<PY
import json
for l in open("$o/${t}_ab.jsonl"):
    d = json.loads(l); print(d["workload"], d["lib"] or "default", d["ms_per_call"], d["pass_ms_per_call"], d["hash_rgb"], d["hash_ids"])
PY
