#!/bin/bash
# Final measurement session of a round on one GPU: everything the bench line quotes comes from here.  tools/gpu_final.sh <tag>
tag=${1:-r2}
tools/gpu_session.sh $tag pytest smoke ref bench launches ncu
cs=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
$cs --tool memcheck python -m pytest tests/test_api_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|ERROR SUMMARY" > gpurun_out/${tag}_memcheck_api.txt
tools/micro/d2h_2d_bench > gpurun_out/${tag}_micro_d2h_2d.txt 2>&1
cat gpurun_out/${tag}_memcheck_api.txt gpurun_out/${tag}_pytest.log
