for t in 32 64 128 256 512; do for w in tess4k spin1080p; do RAST_TINY_MAX=$t python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['roofline']['pass_ms_per_step']; print('tiny_max=$t $w', round(d['ms_per_step'],3), round(p['setup'],3), round(p['raster'],3))"; done; done
