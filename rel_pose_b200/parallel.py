"""Sharded and streamed inference over independent image pairs.

Every pair is independent in the forward pass (self-attention is per image, the Essential Matrix Module is
per pair, BatchNorm uses running statistics), so N GPUs run N shards with NO data-path collective; the two
images of a pair always stay on the same rank (CrossBlock relies on their adjacency, vision_transformer.py:287).
This module is host logic only: shard arithmetic, an optional gather of the [B,2,7] results, and a
double-buffered host->device pipeline that overlaps the copy of micro-batch k+1 with the kernels of
micro-batch k (the reference's callers do `images.cuda()` synchronously, demo.py:76 / train.py:143).
"""
import torch
import torch.distributed as dist


def shard_range(n_pairs, world_size, rank):
    """Contiguous [lo, hi) of pair indices owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(int(n_pairs), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def micro_batches(lo, hi, size):
    """[(a, b)] covering [lo, hi) in chunks of at most `size` pairs."""
    if size <= 0:
        raise ValueError("micro-batch size must be positive")
    return [(a, min(a + size, hi)) for a in range(lo, hi, size)]


def gather_poses(local_poses, n_pairs, group=None):
    """All ranks receive the full [n_pairs,2,7] tensor (shards in rank order).  Works with gloo (CPU
    tensors) and nccl (CUDA tensors); shards may be ragged."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_poses
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(n_pairs, world, rank)
    if local_poses.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local_poses.shape[0]} pairs, expected {hi - lo}")
    cap = -(-n_pairs // world)
    pad = torch.zeros((cap,) + tuple(local_poses.shape[1:]), dtype=local_poses.dtype, device=local_poses.device)
    pad[: hi - lo] = local_poses
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        a, b = shard_range(n_pairs, world, r)
        out.append(p[: b - a])
    return torch.cat(out, 0)


class StreamedInference:
    """Double-buffered inference from PINNED host batches.

        runner = StreamedInference(model)
        for poses in runner.run(batches):      # batches: iterable of (images, Gs_data, intrinsics) host tensors
            ...                                # poses: host tensor [b,2,7] of the corresponding batch

    The H2D copy of batch k+1 is issued on a side stream while batch k computes; results come back through
    pinned staging buffers.  Order is preserved; nothing is dropped."""

    def __init__(self, model, device=None):
        self.model = model
        self.device = device if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StreamedInference needs the model on a CUDA device (no CPU path)")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs = [None, None]
        self._result_ring = {}
        self._dev_ring = {}                  # (shape, dtype) -> [[device buffer, event of its last consumer] x 3, next index]
        self.row_selective = True            # copy only the rows the nearest resize reads
        self.last_image_h2d_bytes = 0

    def _device_buffer(self, shape, dtype):
        """Device staging buffer for the next batch, from a ring of three per (shape, dtype): one being consumed by the
        kernels, one being filled by the copy engine, one spare.  The buffers are allocated once -- a per-step
        torch.empty on the copy stream makes the caching allocator wait for / grow around blocks the compute stream still
        holds (measured: the first end-to-end leg of a run 30 % slower than the second) -- and the copy stream waits for
        the kernels that last read a buffer before overwriting it."""
        key = (tuple(shape), dtype)
        ring = self._dev_ring.get(key)
        if ring is None:
            ring = self._dev_ring[key] = [[[torch.empty(shape, dtype=dtype, device=self.device), None] for _ in range(3)], 0]
        slot = ring[0][ring[1] % 3]
        ring[1] += 1
        if slot[1] is not None:
            self.copy_stream.wait_event(slot[1])
        return slot

    def _copy_images(self, images):
        """H2D of the image batch on the copy stream.  Pinned float32 / uint8 images taller than 224 rows are
        copied row-selectively (rp_copy_rows_h2d): only the rows the nearest resize reads cross PCIe."""
        H, W = int(images.shape[-2]), int(images.shape[-1])
        if (self.row_selective and images.is_pinned() and images.is_contiguous() and H > 224 and images.dim() == 5
                and images.dtype in (torch.float32, torch.uint8)):
            import ctypes
            from math import gcd
            from . import _lib
            if 224 // gcd(H, 224) <= 64:
                slot = self._device_buffer(tuple(images.shape[:-2]) + (224, W), images.dtype)
                d = slot[0]
                planes = images.numel() // (H * W)
                dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
                rc = _lib.lib().rp_copy_rows_h2d(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(images.data_ptr()), planes, H,
                                                 W * images.element_size(), 224, dev,
                                                 ctypes.c_void_p(self.copy_stream.cuda_stream))
                if rc == 0:
                    return slot, (H, W), d.numel() * d.element_size()
                # RP_EINVAL: the float32 row map of this height is not periodic -> plain copy of the whole tensor
        slot = self._device_buffer(tuple(images.shape), images.dtype)
        slot[0].copy_(images, non_blocking=True)
        return slot, None, images.numel() * images.element_size()

    def _pinned_result(self, out):
        """Pinned host buffer for a [b,2,7] result, from a ring of three per shape (one being filled, one pending,
        one being cloned for the caller)."""
        key = (tuple(out.shape), out.dtype)
        ring = self._result_ring.setdefault(key, [[torch.empty(out.shape, dtype=out.dtype, pin_memory=True) for _ in range(3)], 0])
        buf = ring[0][ring[1] % 3]
        ring[1] += 1
        return buf

    def _stage(self, slot, batch):
        images, gs, intr = batch
        with torch.cuda.stream(self.copy_stream):
            img_slot, orig_hw, nbytes = self._copy_images(images)
            self.last_image_h2d_bytes = nbytes
            d_gs = gs.to(self.device, non_blocking=True)
            d_k = intr.to(self.device, non_blocking=True) if intr is not None else None
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._bufs[slot] = (img_slot, d_gs, d_k, ev, orig_hw)

    def run(self, batches):
        from .lietorch import SE3
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._stage(0, nxt)
        slot = 0
        pending = None                       # (host result, event) of the previous batch
        compute = torch.cuda.current_stream(self.device)
        while True:
            img_slot, d_gs, d_k, ev, orig_hw = self._bufs[slot]
            d_img = img_slot[0]
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:
                self._stage(slot ^ 1, nxt)   # overlaps with the kernels launched below
            compute.wait_event(ev)
            with torch.no_grad():
                out = self.model(d_img, SE3(d_gs), intrinsics=d_k, _orig_hw=orig_hw)[0].data
            for t in (d_gs, d_k):            # the copy stream allocated them, the compute stream used them
                if t is not None:
                    t.record_stream(compute)
            used = torch.cuda.Event()        # the image buffer may be refilled once these kernels are done
            used.record(compute)
            img_slot[1] = used
            host = self._pinned_result(out)      # pinned staging ring: cudaHostAlloc per step costs ~0.1 ms
            host.copy_(out, non_blocking=True)
            done = torch.cuda.Event()
            done.record(compute)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].clone()         # the staging buffer is reused two batches later
            pending = (host, done)
            if nxt is None:
                break
            slot ^= 1
        pending[1].synchronize()
        yield pending[0].clone()
