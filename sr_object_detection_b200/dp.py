"""Data-parallel sharding of a detection batch over the GPUs of one box.

Images are independent and the weights read-only, so the hot path shards with NO collective on the
compute path: rank r owns a contiguous slice of the global batch, runs its own `network` replica
(own CUDA graph, stream and pinned staging) and only the per-image detection lists - a few hundred
bytes per image - are gathered on the host, in image order.  The reference has no inference
data-parallelism at all (its only multi-GPU code is the training weight averaging of
network_kernels.cu:279-376, out of scope).

Everything here is host-side plumbing over `torch.distributed` (NCCL on the GPU box, gloo in the CPU
tests); the per-rank work is a callable so the same code drives the CUDA library and the tests.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_range(rank: int, world: int, n_images: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank`: the first n % world ranks get one
    image more, so every image is owned exactly once and the order is preserved by concatenation."""
    if world <= 0 or not 0 <= rank < world or n_images < 0:
        raise ValueError(f"bad shard request rank={rank} world={world} n_images={n_images}")
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local: Sequence, group=None, dst: int = 0) -> List | None:
    """Gather the per-image detection lists of every rank on `dst`, in global image order.
    `local` is this rank's list (one entry per owned image, any picklable object).  Returns the
    concatenated list on `dst`, None elsewhere.  Single-process (no process group) returns `local`."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out: List = []
    for part in bucket:
        out.extend(part)
    return out


def detect_sharded(images: np.ndarray, detect: Callable[[np.ndarray], Sequence], group=None,
                   dst: int = 0) -> List | None:
    """Run `detect` (e.g. `lambda x: darknet.network_detect_batch(net, x, thresh, nms)[0]`) on this
    rank's slice of the global `images` batch and gather the lists on `dst`.  Every rank passes the same
    global batch (or at least its own slice at the right offsets); nothing but the detection lists
    crosses ranks."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(rank, world, len(images))
    local = list(detect(images[lo:hi])) if hi > lo else []
    if len(local) != hi - lo:
        raise RuntimeError(f"rank {rank}: detector returned {len(local)} lists for {hi - lo} images")
    return gather_detections(local, group=group, dst=dst)


def gather_detection_arrays(dets: np.ndarray, counts: np.ndarray, max_det: int, cap: int, group=None,
                            dst: int = 0):
    """Fixed-size gather of one batch's detections for the steady-state loop (no pickling, one collective):
    `dets` is the flat record array network_detect_wait filled ([B * max_det] records of the 28-byte
    `y2_detection`), `counts` the per-image totals.  Every rank sends `counts` and the first `cap` records of
    each of its images; `dst` receives them in rank order = global image order.

    Returns on `dst` (dets_all [world * B][cap] records, counts_all [world * B], n_truncated) where n_truncated is
    the number of images that had more than `cap` detections (their lists are cut at `cap`, exactly as `max_det`
    cuts them in the C API); None on the other ranks.  Without a process group it returns the local arrays."""
    import torch
    import torch.distributed as dist

    counts = np.ascontiguousarray(counts, dtype=np.int32)
    B = counts.shape[0]
    rec = dets.dtype.itemsize
    cap = min(cap, max_det)
    rows = dets.reshape(B, max_det)[:, :cap]
    if not (dist.is_available() and dist.is_initialized()):
        return np.ascontiguousarray(rows), counts.copy(), int((counts > cap).sum())
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    msg = torch.empty(B * 4 + B * cap * rec, dtype=torch.uint8)
    m = msg.numpy()
    m[:B * 4] = counts.view(np.uint8)
    m[B * 4:] = np.ascontiguousarray(rows).view(np.uint8).reshape(-1)
    bucket = [torch.empty_like(msg) for _ in range(world)] if rank == dst else None
    dist.gather(msg, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    all_counts = np.concatenate([b.numpy()[:B * 4].view(np.int32) for b in bucket])
    all_dets = np.concatenate([b.numpy()[B * 4:].view(dets.dtype).reshape(B, cap) for b in bucket])
    return all_dets, all_counts, int((all_counts > cap).sum())
