#!/bin/bash
# Round measurement session on the GPU box: tools/gpu_profile.sh <tag>; everything lands in gpurun_out/<tag>_*
tag=${1:-r1f}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
: > $o/${tag}_workloads.jsonl
for w in suzanne640 tess4k tess4k_64lights overdraw8k; do
  python bench.py --workload $w --steps 10 >> $o/${tag}_workloads.jsonl 2>> $o/${tag}_workloads.err
done
# launch list of one bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_launches_bench.log 2>&1
# full-set captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:k_resolve_shade -s 40 -c 1 -f -o $o/${tag}_shade \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_chunks -s 40 -c 1 -f -o $o/${tag}_raster \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_ncu_raster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_setup -s 4 -c 1 -f -o $o/${tag}_setup50m \
    python bench.py --workload tess4k_64lights --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_ncu_setup.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_chunks -s 4 -c 1 -f -o $o/${tag}_raster_over \
    python bench.py --workload overdraw8k --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_ncu_raster_over.log 2>&1
ls -la $o | grep $tag
