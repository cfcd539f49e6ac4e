"""Multi-GPU partitioning of the frame path (one process per GPU, torch.distributed).

The path shards two ways, both without any data-path collective:
  * frames of the spin sequence are independent: frame k of the sequence is rendered by rank k mod N;
  * a huge single frame is split sort-first into row bands: rank g renders rows [g*H/N, (g+1)*H/N) of the
    whole scene (rast_set_band); pixels are independent, so the stitched bands equal the whole frame.
The only communication is the final gather of finished frames / band slabs to rank 0 (NCCL over NVLink on
GPUs; the same code runs over gloo on CPU tensors for the tests)."""
import torch
import torch.distributed as dist


def frames_of_rank(n_frames_total, rank, world):
    """Indices of the global sequence this rank renders (round-robin: k mod world == rank)."""
    return list(range(rank, n_frames_total, world))


def band_of_rank(height, rank, world):
    """Row band [y0, y1) of rank `rank` (bands differ by at most one row; empty when world > height)."""
    return height * rank // world, height * (rank + 1) // world


def gather_frames(local, n_frames_total, dst=0, group=None):
    """local: [n_local, ...] frames of this rank in the order of frames_of_rank().  Returns on `dst` a tensor
    [n_frames_total, ...] in sequence order, None elsewhere.  Ranks may hold different counts."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local
    counts = [len(frames_of_rank(n_frames_total, r, world)) for r in range(world)]
    n_max = max(counts)
    if local.shape[0] < n_max:  # pad so every rank contributes the same shape
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    local = local.contiguous()
    if rank == dst:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, parts, dst=dst, group=group)
        out = torch.empty((n_frames_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r in range(world):
            out[r::world] = parts[r][:counts[r]]
        return out
    dist.gather(local, None, dst=dst, group=group)
    return None


def gather_bands(local_band, height, dst=0, group=None):
    """local_band: [C, rows, W] (planar, CImg layout) or [rows, W] slab of this rank's band.  Returns on `dst`
    the stitched [C, height, W] / [height, W] image, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local_band
    planar = local_band.dim() == 3
    x = local_band if planar else local_band.unsqueeze(0)
    rows = [band_of_rank(height, r, world) for r in range(world)]
    r_max = max(y1 - y0 for y0, y1 in rows)
    if x.shape[1] < r_max:
        pad = torch.zeros((x.shape[0], r_max - x.shape[1], x.shape[2]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], 1)
    x = x.contiguous()
    if rank == dst:
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.gather(x, parts, dst=dst, group=group)
        out = torch.empty((x.shape[0], height, x.shape[2]), dtype=x.dtype, device=x.device)
        for r, (y0, y1) in enumerate(rows):
            out[:, y0:y1] = parts[r][:, :y1 - y0]
        return out if planar else out[0]
    dist.gather(x, None, dst=dst, group=group)
    return None
