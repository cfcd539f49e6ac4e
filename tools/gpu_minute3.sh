#!/bin/bash
# Third short session: the parity file with the most cases against the candidate default (tight + prep + 2 triangles per setup
# thread), then a few more timings.
t=${1:-r1x}
o=gpurun_out
mkdir -p $o
RAST_LIB=build/variants/librast_b200_cand.so timeout 50 python -m pytest tests/test_parity_gpu.py -x -q -p no:cacheprovider > $o/${t}_pytest_cand.log 2>&1
tail -3 $o/${t}_pytest_cand.log
ab() { env RAST_LIB=${1:+build/variants/librast_b200_$1.so} $3 python tools/quick_ab.py $2 --calls 3 >> $o/${t}_ab.jsonl 2>> $o/${t}_ab.err; }
ab cand spin1080p
ab cand tess4k
ab tight_t3 tess4k
ab tight_t4 tess4k
ab "" overdraw8k
ab cand overdraw8k
cat $o/${t}_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], d['lib'], d['ms_per_call'], d['pass_ms_per_call'], d['hash_rgb'], d['hash_ids'])"
