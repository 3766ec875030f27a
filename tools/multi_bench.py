#!/usr/bin/env python
"""End-to-end throughput of the C multi-GPU entry (network_detect_submit[_u8]_multi / network_detect_wait_multi):
ONE process, one replica and one host thread per GPU, pinned staging filled in place, two global batches in flight.
Timed on the devices (CUDA events on every replica's stream, max over replicas) and by wall clock.

    python tools/multi_bench.py [--gpus N] [--steps 200] [--batch 64] [--cfg yolo-voc] [--side 416]
"""
import argparse
import ctypes as C
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sr_object_detection_b200 import _lib, synth  # noqa: E402
from sr_object_detection_b200 import darknet as dn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cfg", default="yolo-voc")
    ap.add_argument("--side", type=int, default=416)
    ap.add_argument("--head-gain", type=float, default=13.0)
    args = ap.parse_args()
    lib = dn.lib()
    ndev = C.c_int(0)
    _lib.check(lib.y2_device_count(C.byref(ndev)))
    n = args.gpus or ndev.value
    B, side, max_det = args.batch, args.side, 256
    thresh, nms = 0.24, 0.4
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        cfg_text = synth.CFGS[args.cfg](batch=B, w=side, h=side)
        (t / "n.cfg").write_text(cfg_text)
        synth.write_weights(t / "n.weights", cfg_text, seed=1234, head_gain=args.head_gain)
        gpus = (C.c_int * n)(*[i % ndev.value for i in range(n)])
        with dn._quiet_stderr():
            nets = lib.parse_network_cfg_multi(str(t / "n.cfg").encode(), str(t / "n.weights").encode(), gpus, n, B)
    total = lib.network_multi_batch(nets, n)
    dets = (dn.Detection * (total * max_det))()
    counts = (C.c_int * total)()
    images = synth.images(B, 3, side, side, seed=42)
    u8 = np.random.default_rng(42).integers(0, 256, size=(B, side, side, 3), dtype=np.uint8)
    for i in range(n):  # fill every replica's pinned staging once; submit with NULL then uploads straight from it
        for s in (0, 1):
            C.memmove(lib.network_pipeline_staging(nets[i], s), images.ctypes.data, images.nbytes)
            C.memmove(lib.network_pipeline_staging_u8(nets[i], s), u8.ctypes.data, u8.nbytes)
    ev = [[C.c_void_p(), C.c_void_p()] for _ in range(n)]
    for i in range(n):
        lib.cuda_set_device(gpus[i])
        for e in ev[i]:
            _lib.check(lib.y2_event_create(C.byref(e)))
    streams = [C.c_void_p(lib.network_stream(nets[i])) for i in range(n)]

    def run(submit, steps):
        submit()
        for _ in range(1, steps):
            submit()
            lib.network_detect_wait_multi(nets, n, dets, counts, max_det)
        lib.network_detect_wait_multi(nets, n, dets, counts, max_det)

    def timed(submit, steps):
        for i in range(n):
            lib.cuda_set_device(gpus[i])
            lib.network_sync(nets[i])
            _lib.check(lib.y2_event_record(ev[i][0], streams[i]))
        t0 = time.perf_counter()
        run(submit, steps)
        worst = 0.0
        for i in range(n):
            lib.cuda_set_device(gpus[i])
            _lib.check(lib.y2_event_record(ev[i][1], streams[i]))
            ms = C.c_float()
            _lib.check(lib.y2_event_elapsed_ms(ev[i][0], ev[i][1], C.byref(ms)))
            worst = max(worst, ms.value)
        wall = (time.perf_counter() - t0) * 1e3
        return worst, wall

    out = {"gpus": n, "devices": ndev.value, "batch_per_gpu": B, "cfg": args.cfg, "side": side, "steps": args.steps}
    for tag, submit in (("fp32", lambda: lib.network_detect_submit_multi(nets, n, None, thresh, nms, max_det)),
                        ("u8", lambda: lib.network_detect_submit_u8_multi(nets, n, None, thresh, nms, max_det))):
        run(submit, args.warmup)
        ms_dev, ms_wall = timed(submit, args.steps)
        out[f"e2e_{tag}_images_per_s"] = round(total * args.steps / (ms_dev / 1e3), 1)
        out[f"e2e_{tag}_images_per_s_wall"] = round(total * args.steps / (ms_wall / 1e3), 1)
    out["detections_per_image"] = round(float(np.mean(list(counts))), 2)
    print(json.dumps(out))
    lib.free_network_multi(nets, n)


if __name__ == "__main__":
    main()
