#!/bin/bash
# N-GPU bench line exactly as the driver launches it, plus the host-ingest probe:  tools/gpu_session_n8.sh <tag> <N>
tag=${1:-n8}; n=${2:-8}
o=gpurun_out; mkdir -p $o
nvidia-smi topo -m > $o/${tag}_topo.txt 2>&1; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA|Hypervisor" >> $o/${tag}_topo.txt
for g in $(seq 0 $((n-1))); do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $g $b numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)" >> $o/${tag}_topo.txt; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 tools/d2h_probe.py 2>/dev/null | grep n_gpus > $o/${tag}_d2h_probe.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err
cat $o/${tag}_d2h_probe.txt; tail -c 300 $o/${tag}_bench_n$n.err
