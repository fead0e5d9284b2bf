"""shard.py — frame sharding across the GPUs of one box and the host-side result gather.

The reference has no multi-GPU path at all (one blocking ``perform`` per request, server/server.py:156-163,
232).  Frames are independent from preprocess through Soft-NMS (nothing in server/detector.py carries state
across frames), so a batch shards by frame with no data-path collective: rank r takes the contiguous slice
``shard_range(n, r, world)``, runs the whole pipeline on its own GPU, and only the small per-frame result
lists travel — through ``torch.distributed.gather_object`` on whatever backend the job was launched with
(NCCL on the GPU box, gloo in the CPU tests).  No NCCL data collective is used on purpose.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the frames rank `rank` owns; sizes differ by at most one frame."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError(f"bad shard request: n={n_frames} rank={rank} world={world}")
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_results(local: Sequence[list], n_frames: int, dst: int = 0, group=None) -> Optional[List[list]]:
    """Collects each rank's per-frame result lists on rank `dst` in global frame order.

    `local` holds this rank's results for its `shard_range` slice.  Returns the full list on `dst`, None elsewhere.
    Works without torch.distributed initialised (single process)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if len(local) != n_frames:
            raise ValueError("single-process gather needs results for every frame")
        return list(local)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, end = shard_range(n_frames, rank, world)
    if len(local) != end - begin:
        raise ValueError(f"rank {rank} owns {end - begin} frames but returned {len(local)} result lists")
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[list] = []
    for part in bucket:
        out.extend(part)
    assert len(out) == n_frames
    return out


def detect_sharded(detector, frames, threshold: float = 0.1, dst: int = 0, group=None):
    """Runs the detector on this rank's slice of `frames` — an [n,h,w,3] u8 array (`perform_frames`) or a list of encoded
    payloads (`perform_jpegs`), identical on every rank — and gathers the per-frame results on `dst`."""
    import torch.distributed as dist

    n = len(frames)
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    begin, end = shard_range(n, rank, world)
    encoded = n > 0 and isinstance(frames[0], (bytes, bytearray, memoryview))
    run = detector.perform_jpegs if encoded else detector.perform_frames
    local = run(frames[begin:end], threshold=threshold) if end > begin else []
    return gather_results(local, n, dst=dst, group=group)
