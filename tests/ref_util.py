"""Run the oracle executables (TEST INFRASTRUCTURE under oracle/) and load their dumps."""
import json
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF_BIN = ROOT / "oracle" / "_ref" / "darknet_ref"
ORACLE_BIN = ROOT / "oracle" / "_build" / "y2_oracle"


def have_ref() -> bool:
    return REF_BIN.exists()


def have_oracle() -> bool:
    return ORACLE_BIN.exists()


def build_oracle() -> None:
    """Compile oracle/y2_oracle.c (test infrastructure) if the binary is missing or stale."""
    src = ROOT / "oracle" / "y2_oracle.c"
    if ORACLE_BIN.exists() and ORACLE_BIN.stat().st_mtime >= src.stat().st_mtime:
        return
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "oracle"], check=True, capture_output=True)


def checker_bin():
    """The CPU checker for the GPU parity tests: the compiled reference when it is present,
    else the oracle port (pinned to the reference by tests/test_oracle_golden.py)."""
    if REF_BIN.exists():
        return REF_BIN
    build_oracle()
    return ORACLE_BIN


def run_raw(cmd, cwd=None, env=None):
    return _run(cmd, cwd=cwd, extra_env=env)


def _run(cmd, cwd=None, threads=None, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, cwd=cwd, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def forward(binary, cfg, weights, input_path, outdir, thresh=0.24, nms=0.4, dump_layers=True, cwd=None):
    Path(outdir).mkdir(parents=True, exist_ok=True)
    return _run([binary, "forward", cfg, weights, input_path, outdir, thresh, nms, int(dump_layers)], cwd=cwd)


def region(binary, cfg, region_in, outdir, thresh=0.24, nms=0.4, cwd=None):
    Path(outdir).mkdir(parents=True, exist_ok=True)
    return _run([binary, "region", cfg, region_in, outdir, thresh, nms], cwd=cwd)


def timeit(binary, cfg, weights, input_path, thresh, nms, warmup, iters, threads=None):
    return _run([binary, "time", cfg, weights, input_path, thresh, nms, warmup, iters], threads=threads)


def load(outdir, name, shape=None):
    a = np.fromfile(Path(outdir) / name, np.float32)
    return a.reshape(shape) if shape is not None else a
