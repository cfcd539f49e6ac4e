#!/bin/bash
# Second short session: where does k_setup's time go on the 8 M-triangle mesh (ncu capture + timing probes), and a few variants.
t=${1:-r1y}
o=gpurun_out
mkdir -p $o
ab() { env RAST_LIB=${1:+build/variants/librast_b200_$1.so} $3 python tools/quick_ab.py $2 --calls 3 >> $o/${t}_ab.jsonl 2>> $o/${t}_ab.err; }
RAST_LIB=build/variants/librast_b200_tight.so timeout 40 ncu --set full --clock-control none --import-source on -k regex:k_setup -s 5 -c 1 -f -o $o/${t}_setup_tight \
    python tools/quick_ab.py tess4k --calls 2 > $o/${t}_ncu.log 2>&1
ab probe_noatom tess4k
ab probe_nowalk tess4k
ab probe_base_noatom tess4k
ab tight_t2 tess4k
ab tight_b6 tess4k
ab tight_b8 tess4k
ab tight tess4k RAST_TINY_MAX=16
ab tight tess4k RAST_TINY_MAX=256
cat $o/${t}_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'], d['ms_per_call'], d['pass_ms_per_call']['setup'], d['pass_ms_per_call']['shade'], d['hash_ids'])"
ls -la $o | grep $t
