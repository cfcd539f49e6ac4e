"""The N > 1 host logic (frame partitioning, band stitching, gather to rank 0) on CPU with the gloo backend,
world sizes 2 and 3.  The rendering itself is replaced by arrays that encode (frame index) / (row index), so
the test checks exactly the plumbing bench.py uses on GPUs with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rasteriser_b200 import multi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, height, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # frames: frame k is an image filled with k
        mine = multi.frames_of_rank(n_frames, rank, world)
        local = torch.stack([torch.full((3, 4, 5), k, dtype=torch.uint8) for k in mine]) if mine else torch.zeros((0, 3, 4, 5), dtype=torch.uint8)
        out = multi.gather_frames(local, n_frames)
        ok = True
        if rank == 0:
            ok &= out.shape == (n_frames, 3, 4, 5) and all(int(out[k].min()) == k == int(out[k].max()) for k in range(n_frames))
        else:
            ok &= out is None
        # bands: row y of the image holds y (+100 per channel)
        y0, y1 = multi.band_of_rank(height, rank, world)
        band = torch.stack([torch.arange(y0, y1, dtype=torch.float32)[:, None].expand(y1 - y0, 7) + 100 * c for c in range(3)])
        img = multi.gather_bands(band.contiguous(), height)
        depth = multi.gather_bands(band[0].contiguous(), height)
        if rank == 0:
            want = torch.arange(height, dtype=torch.float32)[:, None].expand(height, 7)
            ok &= img.shape == (3, height, 7) and all(torch.equal(img[c], want + 100 * c) for c in range(3)) and torch.equal(depth, want)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= float(t) == float(world)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,height", [(2, 7, 9), (3, 10, 8), (2, 1, 3)])
def test_gather_frames_and_bands(world, n_frames, height):
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_frames, height, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_partition_properties():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 720):
            parts = [multi.frames_of_rank(n, r, world) for r in range(world)]
            assert sorted(k for p in parts for k in p) == list(range(n))
        for h in (1, 5, 1080, 4320):
            bands = [multi.band_of_rank(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h and all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))


def test_bind_host_to_gpu_splits_the_allowed_cpus():
    """multi.bind_host_to_gpu without a GPU (no NVML here): every local rank gets its own contiguous share of the CPUs this
    process may use; the shares do not overlap and the binding is really applied (child process: the test keeps its own mask)."""
    import os
    import subprocess
    import sys
    assert multi._cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    allowed = sorted(os.sched_getaffinity(0))
    world = min(4, len(allowed))
    code = ("import os, sys, json; sys.path.insert(0, %r); from rasteriser_b200 import multi; "
            "info = multi.bind_host_to_gpu(int(sys.argv[1]), int(sys.argv[2])); print(json.dumps([sorted(os.sched_getaffinity(0)), info]))" % ROOT)
    import json
    seen = []
    for rank in range(world):
        out = subprocess.run([sys.executable, "-c", code, str(rank), str(world)], capture_output=True, text=True, check=True).stdout
        cpus, info = json.loads(out.strip().splitlines()[-1])
        assert cpus and set(cpus) <= set(allowed) and len(cpus) == max(1, len(allowed) // world)
        assert not (set(cpus) & set(seen))
        assert info["how"] != "unchanged"
        seen += cpus
