#!/bin/bash
# Two-GPU session: the GPU suite (incl. the peer-memory band test that needs two devices), the N=1 and N=2 bench lines.
tag=${1:-n2}
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $o/${tag}_bench_n2.json 2> $o/${tag}_bench_n2.err
ls -la $o | grep ${tag}_
