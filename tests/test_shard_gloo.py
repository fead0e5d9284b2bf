"""Frame sharding + host-side gather over torch.distributed (gloo, world_size 2, CPU).  The detector is a
stand-in that tags each frame, so only the N>1 host logic is exercised here; the GPU tests cover the rest."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastdet_b200 import shard


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


class TagDetector:
    """perform_frames returns one record per frame derived from its pixels (so order errors are visible)."""

    def perform_frames(self, frames, threshold=0.1):
        return [[(int(f[0, 0, 0]) + 1, 1.0, float(f[0, 0, 1]), 0.0, 1.0, 1.0)] * (int(f[0, 0, 2]) % 3) for f in frames]


    def perform_jpegs(self, datas, threshold=0.1):
        return [[(len(d), 1.0, float(d[0]), 0.0, 1.0, 1.0)] for d in datas]


def _worker(rank, world, port, n_frames, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 200, size=(n_frames, 4, 4, 3), dtype=np.uint8)
    res = shard.detect_sharded(TagDetector(), frames, 0.1)
    payloads = [bytes([i % 251]) * (i + 1) for i in range(n_frames)]  # encoded payloads shard the same way
    res2 = shard.detect_sharded(TagDetector(), payloads, 0.1)
    if rank == 0:
        want = TagDetector().perform_frames(frames)
        torch.save({"ok": res == want and res2 == TagDetector().perform_jpegs(payloads), "n": len(res)}, out_path)
    else:
        assert res is None and res2 is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 64])
def test_gather_world2(tmp_path, n_frames):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, n_frames, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["n"] == n_frames


def test_gather_single_process():
    frames = np.zeros((3, 4, 4, 3), np.uint8)
    assert shard.detect_sharded(TagDetector(), frames) == TagDetector().perform_frames(frames)
