#!/usr/bin/env python
"""Developer tool (GPU box): shape / batch sweep of whole networks through the drop-in API.

For every (cfg, side, batch): parse, load seeded weights, network_predict on seeded images, then check that
the output is finite and that image 0's row equals (to 5e-3 of the row maximum) the row a batch-1 network
produces for the same image - kernels pick different tilings, schedules and tile orders per batch size; the
result per image may only move by bf16 roundings (another kernel variant may sum the same products in another order).  Also runs the batched detect call.  Prints one line per case."""
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from sr_object_detection_b200 import darknet as dn, synth  # noqa: E402

CASES = [
    ("tiny-yolo-voc", s, b) for s in (320, 416, 608) for b in (1, 3, 7, 16, 33, 64, 128)
] + [
    ("yolo-voc", s, b) for s in (320, 416, 544) for b in (1, 5, 16, 48, 96)
] + [("yolo", 608, b) for b in (4, 16, 40)] + [("darknet19_448", 448, b) for b in (8, 32)] + \
    [("resnet50", 224, b) for b in (4, 32)] + [("mini-dense", 64, b) for b in (1, 16, 256)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] == sys.argv[1]]

tmp = Path(tempfile.mkdtemp())
dn.set_gpu_index(0)
ref_rows = {}
bad = 0
for name, side, batch in CASES:
    t0 = time.time()
    wpath = tmp / f"{name}.weights"
    key = (name, side)

    def build(b):
        text = synth.CFGS[name](batch=b, w=side, h=side)
        (tmp / "n.cfg").write_text(text)
        if not wpath.exists():
            synth.write_weights(wpath, text, seed=1234)
        net = dn.parse_network_cfg(tmp / "n.cfg")
        dn.load_weights(net, wpath)
        return net

    x = synth.images(batch, 3, side, side, seed=42)
    if key not in ref_rows:
        n1 = build(1)
        ref_rows[key] = dn.network_predict(n1, np.ascontiguousarray(x[:1]))[0]
        dn.free_network(n1)
    net = build(batch)
    out = dn.network_predict(net, x)
    ok = bool(np.isfinite(out).all())
    err = float(np.abs(out[0] - ref_rows[key]).max() / max(np.abs(ref_rows[key]).max(), 1e-30))
    ndet = -1
    if net.layers[net.n - 1].type == dn.REGION:
        dets, _ = dn.network_detect_batch(net, x, 0.02, 0.4, 256)
        ndet = sum(len(d) for d in dets)
    dn.free_network(net)
    status = "ok" if ok and err <= 5e-3 else "FAIL"
    bad += status != "ok"
    print(f"{status:4s} {name:14s} {side:4d} b{batch:<4d} row0 vs batch-1: {err:.2e}  detections {ndet:6d}  {time.time() - t0:5.1f}s",
          flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
