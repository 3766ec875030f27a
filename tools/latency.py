#!/usr/bin/env python
"""Developer tool (GPU box): latency / throughput of the drop-in calls at small batches.

    python tools/latency.py [cfg-name] [side]

For batch 1, 2, 4, 8, 16, 64: the synchronous network_detect_batch call (upload + forward + decode + NMS +
pick + download, one host synchronisation) from pinned host fp32 images, and the same from uint8 frames."""
import ctypes as C
import json
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sr_object_detection_b200 import darknet as dn, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolo-voc"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 416
tmp = Path(tempfile.mkdtemp())
dn.set_gpu_index(0)
lib = dn.lib()
for batch in (1, 2, 4, 8, 16, 64):
    cfg_text = synth.CFGS[name](batch=batch, w=side, h=side)
    (tmp / "n.cfg").write_text(cfg_text)
    if not (tmp / "n.weights").exists():
        synth.write_weights(tmp / "n.weights", cfg_text, seed=1234)
    net = dn.parse_network_cfg(tmp / "n.cfg")
    dn.load_weights(net, tmp / "n.weights")
    x = synth.images(batch, 3, side, side, seed=42)
    u8 = np.random.default_rng(1).integers(0, 256, size=(batch, side, side, 3), dtype=np.uint8)
    max_det = 256
    dets = (dn.Detection * (batch * max_det))()
    counts = (C.c_int * batch)()
    stage = lib.network_input_staging(net)
    C.memmove(stage, x.ctypes.data, x.nbytes)

    def timed(fn, reps):
        for _ in range(5):
            fn()
        lib.network_sync(net)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        lib.network_sync(net)
        return (time.perf_counter() - t0) / reps * 1e3

    reps = 200 if batch <= 8 else 50
    f32_ms = timed(lambda: lib.network_detect_batch(net, stage, 0.24, 0.4, dets, counts, max_det), reps)
    st8 = lib.network_pipeline_staging_u8(net, 0)
    C.memmove(st8, u8.ctypes.data, u8.nbytes)
    u8_ms = timed(lambda: lib.network_detect_batch_u8(net, st8, 0.24, 0.4, dets, counts, max_det), reps)
    print(json.dumps({"cfg": name, "side": side, "batch": batch,
                      "sync_f32_host_ms": round(f32_ms, 4), "sync_f32_img_s": round(batch / f32_ms * 1e3, 1),
                      "sync_u8_host_ms": round(u8_ms, 4), "sync_u8_img_s": round(batch / u8_ms * 1e3, 1)}), flush=True)
    dn.free_network(net)
