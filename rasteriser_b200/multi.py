"""Multi-GPU partitioning of the frame path (one process per GPU, torch.distributed).

The path shards two ways, both without any data-path collective:
  * frames of the spin sequence are independent: frame k of the sequence is rendered by rank k mod N;
  * a huge single frame is split sort-first into row bands: rank g renders rows [g*H/N, (g+1)*H/N) of the
    whole scene (rast_set_band); pixels are independent, so the stitched bands equal the whole frame.
The only communication is the final gather of finished frames / band slabs to rank 0 (NCCL over NVLink on
GPUs; the same code runs over gloo on CPU tensors for the tests)."""
import torch
import torch.distributed as dist


def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def bind_host_to_gpu(local_rank, local_world):
    """Pin this process (and every thread it starts later: the library's background-fill pool, the pinned allocations' first
    touch) to CPUs next to GPU `local_rank`: the cores of the GPU's NUMA node (sysfs numa_node of its PCI function) -- split
    among the ranks that share the node -- or, where the platform reports no NUMA node for the device (VMs), an even share of the
    process's CPUs.  Host-buffer draws are bound by host-memory ingest at N > 1 (DMA writes + background fill); without this
    every rank's helpers roam all sockets and the ranks whose GPU sits on the far socket see a fraction of the copy rate.
    Returns a dict describing what was done (bench.py prints it)."""
    import os
    info = {"numa_node": None, "cpus": None, "how": "unchanged"}
    if not hasattr(os, "sched_setaffinity"):
        return info
    allowed = sorted(os.sched_getaffinity(0))
    node, node_cpus, sharers, my_slot = -1, None, local_world, local_rank
    try:
        import pynvml
        pynvml.nvmlInit()
        nodes = []
        for g in range(local_world):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(g)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = "/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()
            nodes.append(int(open(path).read()) if os.path.exists(path) else -1)
        node = nodes[local_rank]
        if node >= 0:
            node_cpus = [c for c in _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()) if c in allowed]
            same = [g for g in range(local_world) if nodes[g] == node]
            sharers, my_slot = len(same), same.index(local_rank)
    except Exception:
        node = -1
    pool = node_cpus if node_cpus else allowed
    share = max(1, len(pool) // max(1, sharers))
    cpus = pool[my_slot * share:(my_slot + 1) * share] or pool
    try:
        os.sched_setaffinity(0, cpus)
        info.update(numa_node=node if node >= 0 else None, cpus="%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus)), how="numa node of the GPU" if node_cpus else "even share of the allowed CPUs")
    except OSError:
        pass
    return info


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


class PeerImage:
    """A full-size image (RGB8 planes + f32 depth, `frames` of them) that lives in rank `dst`'s device memory and is
    mapped into every other rank of the node (CUDA IPC over NVLink): each rank's shade pass stores its band -- or its
    frames of a sequence -- straight into it, so the gather is the kernel's own store and needs no staging buffer, no
    stitching copy and no collective; one barrier at the end tells rank `dst` that the image is complete.

        img = PeerImage(renderer, W, H)                  # collective: every rank calls it
        img.draw_band(args)                              # this rank's rows of one frame, into rank dst's image
        img.draw_sequence(args_list, first_index, stride)  # frames first_index, +stride, ... of a sequence, batched
        img.barrier(); rgb, depth = img.read()           # rank dst: numpy copies (other ranks: None)
    """

    def __init__(self, renderer, width, height, frames=1, dst=0, group=None, with_depth=True):
        self.r, self.W, self.H, self.frames, self.dst, self.group = renderer, int(width), int(height), int(frames), dst, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        P = self.W * self.H
        self.rgb_bytes, self.depth_bytes = self.frames * 3 * P, (self.frames * 4 * P if with_depth else 0)
        self._owned = self.rank == dst
        self._flag = None
        self.depth = 0  # with_depth=False: colour only (the spin sequence gathers RGB frames)
        handles = [None, None]
        if self._owned:
            self.rgb = renderer.device_alloc(self.rgb_bytes)
            if with_depth:
                self.depth = renderer.device_alloc(self.depth_bytes)
            handles = [renderer.ipc_export(self.rgb), renderer.ipc_export(self.depth) if with_depth else None]
        if self.world > 1:
            dist.broadcast_object_list(handles, src=dst, group=group)
            if not self._owned:
                self.rgb = renderer.ipc_open(handles[0])
                if with_depth:
                    self.depth = renderer.ipc_open(handles[1])

    def draw_band(self, args, frame=0):
        """Render this rank's row band of one frame (rast_set_band must hold band_of_rank) into the shared image.
        args: one Args / RastArgs, or a one-element list / ctypes array of them."""
        import ctypes
        y0, _ = band_of_rank(self.H, self.rank, self.world)
        P = self.W * self.H
        one = args if isinstance(args, (list, tuple, ctypes.Array)) else [args]
        self.r.set_output_plane_stride(P)
        try:
            self.r.draw_frames_device(one, self.rgb + frame * 3 * P + y0 * self.W, (self.depth + (frame * P + y0 * self.W) * 4) if self.depth else None)
        finally:
            self.r.set_output_plane_stride(0)

    def draw_sequence(self, args_list, first_index, stride):
        """Render frames first_index, first_index + stride, ... of the sequence into their slots of the shared image, batched
        through the kernels like any multi-frame call: rast_set_output_frame_stride puts the i-th frame of the call into slot
        first_index + i * stride, so N ranks (first_index = rank, stride = N) fill one sequence buffer in order."""
        P = self.W * self.H
        self.r.set_output_frame_stride(stride)
        try:
            self.r.draw_frames_device(args_list, self.rgb + first_index * 3 * P, (self.depth + first_index * P * 4) if self.depth else None)
        finally:
            self.r.set_output_frame_stride(1)

    def barrier(self, sync=True):
        """After it returns on rank dst (and the current stream has passed it), every rank's band is in the image.
        sync=True first waits on the host for this rank's own kernels (the renderer may launch on a stream of its own);
        sync=False is for callers whose renderer shares torch's current stream (Renderer.set_stream): the completion is
        then ordered purely on the device by a one-element all-reduce -- NCCL runs it after each rank's preceding
        kernels -- and the host is never blocked."""
        if sync:
            self.r.sync()
        if self.world > 1:
            if self._flag is None:
                self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(self._flag, group=self.group)
            if sync:
                torch.cuda.current_stream().synchronize()

    def read(self):
        """Rank dst: (rgb [frames,3,H,W] uint8, depth [frames,H,W] float32) as numpy arrays; call after barrier()."""
        if not self._owned:
            return None
        import numpy as np
        rgb = np.empty((self.frames, 3, self.H, self.W), np.uint8)
        depth = np.empty((self.frames, self.H, self.W), np.float32) if self.depth else None
        self.r.sync()
        self.r.device_read(self.rgb, rgb)
        if self.depth:
            self.r.device_read(self.depth, depth)
        return rgb, depth

    def close(self):
        """Collective.  Peers unmap first, then the owner frees (exported memory must outlive every mapping)."""
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still write
        if not self._owned:
            self.r.ipc_close(self.rgb)
            if self.depth:
                self.r.ipc_close(self.depth)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self._owned:
            self.r.device_free(self.rgb)
            if self.depth:
                self.r.device_free(self.depth)
