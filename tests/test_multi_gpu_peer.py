"""Sort-first bands written straight into rank 0's image over peer memory (multi.PeerImage, CUDA IPC): needs two
GPUs in one box, so it is skipped on single-GPU runs; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu_peer.py`."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from rasteriser_b200 import api, multi
import scenes as S
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = S.scene("suzanne")
r = api.Renderer(local)
r.upload_mesh(sc.positions, sc.tris, sc.normals, sc.uvs)
r.upload_materials(sc.materials)
r.set_lights(S.lights("threepoint"))
W, H = 641, 483  # odd sizes: bands of unequal height, rows that are not multiples of anything
a = api.Args(W, H, tait_bryan_angles=(0.2, 0.7, -0.1), scale=1.2)
whole_f, whole_d = r.draw_frame(a)             # every rank renders the whole frame for reference
y0, y1 = multi.band_of_rank(H, rank, world)
r.set_band(y0, y1)
img = multi.PeerImage(r, W, H, frames=2)
img.draw_band(a, frame=0)
b = api.Args(W, H, tait_bryan_angles=(0.0, 2.0, 0.0))
img.draw_band(b, frame=1)
img.barrier()
got = img.read()
r.set_band(0, 0)
whole_f2, whole_d2 = r.draw_frame(b)
if rank == 0:
    rgb, depth = got
    assert np.array_equal(rgb[0], whole_f) and np.array_equal(depth[0].view(np.uint32), whole_d.view(np.uint32)), "frame 0 differs"
    assert np.array_equal(rgb[1], whole_f2) and np.array_equal(depth[1].view(np.uint32), whole_d2.view(np.uint32)), "frame 1 differs"
    print("PEER_OK", int(rgb.sum()))
else:
    assert got is None
img.close()
# a frame sequence partitioned by frame (frame k on rank k mod N), every rank's shade pass storing its frames into their slots
# of rank 0's sequence buffer (rast_set_output_frame_stride), colour only
N = 7
seq_poses = [api.Args(320, 240, tait_bryan_angles=(0.0, api.spin_angle(0.1, k, N), 0.0)) for k in range(N)]
want_seq, _ = r.draw_frames(seq_poses)
seq = multi.PeerImage(r, 320, 240, frames=N, with_depth=False)
seq.draw_sequence(seq_poses[rank::world], rank, world)
seq.barrier()
got_seq = seq.read()
if rank == 0:
    assert got_seq[1] is None and np.array_equal(got_seq[0], want_seq), "sequence differs"
    print("SEQ_OK")
seq.close()
dist.destroy_process_group()
r.close()
'''


@pytest.mark.gpu
def test_bands_written_into_rank0_image_over_peer_memory(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PEER_OK" in out.stdout and "SEQ_OK" in out.stdout
