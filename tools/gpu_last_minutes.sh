#!/bin/bash
# A very short GPU session (torch-free, ~1 minute): smoke() on the default library and on every prebuilt kernel variant
# (parity of the variants' device code on a real B200), then quick A/B timings with output hashes.  Each step writes its own
# file under gpurun_out/, most valuable first, so a call cut short still leaves results.
t=${1:-r1z}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $o/${t}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke_default.log 2>&1
for v in tight prep tight_prep frcp; do
  RAST_LIB=build/variants/librast_b200_$v.so python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke_$v.log 2>&1
done
ab() { RAST_LIB=${1:+build/variants/librast_b200_$1.so} python tools/quick_ab.py $2 >> $o/${t}_ab.jsonl 2>> $o/${t}_ab.err; }
ab "" spin1080p; ab prep spin1080p
ab "" tess4k; ab tight tess4k
ab tight_prep spin1080p; ab frcp spin1080p
ab "" suzanne640; ab tight_prep suzanne640
ab "" tess4k_64lights; ab tight tess4k_64lights
cat $o/${t}_smoke_*.log | tail -8; cat $o/${t}_ab.jsonl
