#!/bin/bash
# compute-sanitizer over the GPU suite:  tools/gpu_sanitizer.sh <tag>   -> gpurun_out/<tag>_sanitizer.txt
tag=${1:-san}; o=gpurun_out/${tag}_sanitizer.txt; mkdir -p gpurun_out
cs=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
echo "$($cs --version | head -2 | tr '\n' ' ') on one B200, python -m pytest under the tool" > $o
run() { echo "--- $1" >> $o; shift; timeout 1500 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|^=========.*(error|Error)" | tail -12 >> $o; }
run "memcheck, -m gpu suite without the full-size frames" $cs --tool memcheck python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_parity_gpu_large.py -k "not fuzz"
run "memcheck, the 8K overdraw frame through the tile schedule and a 4K tessellated frame" $cs --tool memcheck python tools/quick_ab.py overdraw8k --calls 1
run "initcheck, tests/test_parity_gpu.py tests/test_api_gpu.py" $cs --tool initcheck python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -q -p no:cacheprovider
run "racecheck, tile schedule (shared-memory keys, bins, votes) and chunk schedule soups" $cs --tool racecheck python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "tile or soup or tie"
cat $o
