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
