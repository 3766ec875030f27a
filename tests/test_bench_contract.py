"""bench.py contract on a box without a GPU: the reference arm (the reference's own CPU path, oracle/_ref or the
oracle port) runs anywhere and must print exactly one JSON line with the keys the driver reads; the GPU arm must
refuse to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    from sr_object_detection_b200 import build
    build.build()
    if not ((ROOT / "oracle" / "_ref" / "darknet_ref").exists() or (ROOT / "oracle" / "_build" / "y2_oracle").exists()):
        pytest.skip("oracle binaries not built")
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must set the thread count itself
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("yolo-voc.cfg 416x416 batch 64")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["cores"] == d["config"]["omp_threads"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_gpu_arm_fails_loudly_without_a_device():
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert r.stdout.strip() == "", "no bench line may be printed without a device"
