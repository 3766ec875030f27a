"""Multi-process host logic of the data-parallel path on CPU: world_size 2
over gloo.  The per-rank detector is a deterministic stand-in (the CUDA library needs a GPU); what is
under test is the sharding, the absence of any data-path collective and the ordered gather."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from sr_object_detection_b200 import dp  # noqa: E402


def test_shard_range_partitions_every_batch():
    for n in (0, 1, 2, 7, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = dp.shard_range(r, world, n)
                assert 0 <= lo <= hi <= n
                covered.extend(range(lo, hi))
            assert covered == list(range(n)), (n, world)
            sizes = [dp.shard_range(r, world, n)[1] - dp.shard_range(r, world, n)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(2, 2, 4)


def _fake_detect(x: np.ndarray):
    """one 'detection list' per image, a pure function of the pixels"""
    return [[(int(img.sum()) % 97, float(img.flat[0]))] * (int(img.flat[1] * 3) % 3) for img in x]


def _worker(rank: int, world: int, port: int, n_images: int, out_path: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        images = rng.random((n_images, 3, 8, 8), dtype=np.float32)  # same global batch on every rank
        seen = []

        def detect(x):
            seen.append(len(x))
            return _fake_detect(x)

        got = dp.detect_sharded(images, detect)
        lo, hi = dp.shard_range(rank, world, n_images)
        assert seen == ([hi - lo] if hi > lo else []), "a rank must only touch its own slice"
        if rank == 0:
            assert got == _fake_detect(images), "gathered lists must equal the single-process result, in order"
            with open(out_path, "w") as f:
                f.write("ok")
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 5, 1])
def test_detect_sharded_world2_gloo(tmp_path, n_images):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "rank0.txt"
    mp.spawn(_worker, args=(2, port, n_images, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


_REC = np.dtype([("x", "f4"), ("y", "f4"), ("w", "f4"), ("h", "f4"), ("prob", "f4"), ("obj_id", "i4"),
                 ("box_index", "i4")])  # y2_detection


def _fake_batch(rank: int, B: int, max_det: int):
    rng = np.random.default_rng(100 + rank)
    counts = rng.integers(0, 7, B).astype(np.int32)
    counts[0] = 6  # one image above the transfer cap of the test (4)
    dets = np.zeros(B * max_det, _REC)
    for b in range(B):
        for k in range(counts[b]):
            dets[b * max_det + k] = (rng.random(), rng.random(), rng.random(), rng.random(), rng.random(),
                                     int(rng.integers(0, 20)), rank * 1000 + b * 10 + k)
    return dets, counts


def _array_worker(rank: int, world: int, port: int, out_path: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, max_det, cap = 3, 8, 4
        dets, counts = _fake_batch(rank, B, max_det)
        res = dp.gather_detection_arrays(dets, counts, max_det, cap)
        if rank != 0:
            assert res is None
            return
        all_dets, all_counts, truncated = res
        assert all_dets.shape == (world * B, cap) and all_counts.shape == (world * B,)
        for r in range(world):
            d, c = _fake_batch(r, B, max_det)
            assert np.array_equal(all_counts[r * B:(r + 1) * B], c)
            for b in range(B):
                n = min(c[b], cap)
                assert np.array_equal(all_dets[r * B + b, :n], d.reshape(B, max_det)[b, :n]), "image order / content"
        want_truncated = sum(int((_fake_batch(r, B, max_det)[1] > cap).sum()) for r in range(world))
        assert truncated == want_truncated >= world  # image 0 of every rank carries 6 > cap detections
        with open(out_path, "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_detection_arrays_world2_gloo(tmp_path):
    """the fixed-size per-step gather bench.py times inside e2e at N > 1"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "rank0.txt"
    mp.spawn(_array_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
    # single process: the local arrays come back
    dets, counts = _fake_batch(0, 3, 8)
    d, c, t = dp.gather_detection_arrays(dets, counts, 8, 4)
    assert d.shape == (3, 4) and np.array_equal(c, counts) and t == int((counts > 4).sum())
