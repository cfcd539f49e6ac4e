#!/bin/bash
# Host-side rows (loaders, output path) measured on the GPU box's own CPU, plus small A/B runs: tools/gpu_host_rows.sh <tag>
tag=${1:-r1h}
o=gpurun_out
mkdir -p $o
python tools/bench_loader.py --n 91 --keep /tmp/ldr --threads 1,4,0 --cli > $o/${tag}_loader_8Mtris.json 2> $o/${tag}_loader.err
python tools/bench_output.py > $o/${tag}_output_png_8k.json 2> $o/${tag}_output.err
for px in 1 4; do
  RAST_SHADE_PX=$px python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['roofline']['pass_ms_per_step']
print('RAST_SHADE_PX=$px ms/step', round(d['ms_per_step'],4), 'shade', round(p['shade'],4), 'raster', round(p['raster'],4))" >> $o/${tag}_shade_px.txt
done
cat $o/${tag}_shade_px.txt
tail -c 1500 $o/${tag}_loader_8Mtris.json
