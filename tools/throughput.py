#!/usr/bin/env python
"""Developer tool (GPU box): device-resident throughput of any BASELINE config (the default bench.py only
measures the headline one).

    python tools/throughput.py <cfg-name> <side> <batch> [steps]

A step = forward pass (CUDA graph replay) + region decode + NMS + pick for detectors, forward pass only for
classifiers; inputs resident in HBM; timed with CUDA events on the network's stream."""
import ctypes as C
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from sr_object_detection_b200 import _lib, darknet as dn, synth  # noqa: E402

name, side, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 20
tmp = Path(tempfile.mkdtemp())
kw = {}
if name == "yolo9000":
    synth.write_tree(tmp / "9k.tree")
    kw["tree"] = str(tmp / "9k.tree")
cfg_text = synth.CFGS[name](batch=batch, w=side, h=side, **kw)
(tmp / "n.cfg").write_text(cfg_text)
import os  # noqa: E402
synth.write_weights(tmp / "n.weights", cfg_text, seed=1234, head_gain=float(os.environ.get("Y2_HEAD_GAIN", "1")))
dn.set_gpu_index(0)
lib = dn.lib()
net = dn.parse_network_cfg(tmp / "n.cfg")
dn.load_weights(net, tmp / "n.weights")
x = synth.images(batch, 3, side, side, seed=42)
staging = lib.network_input_staging(net)
C.memmove(staging, x.ctypes.data, x.nbytes)
lib.network_upload_input(net, staging)
detector = net.layers[net.n - 1].type == dn.REGION
max_det = 256
dets = (dn.Detection * (batch * max_det))()
counts = (C.c_int * batch)()


def step():
    lib.network_forward_device(net)
    if detector:
        lib.network_detect_device(net, 0.24, 0.4, dets, counts, max_det)


stream = C.c_void_p(lib.network_stream(net))
ev = [C.c_void_p(), C.c_void_p()]
for e in ev:
    _lib.check(lib.y2_event_create(C.byref(e)))
for _ in range(5):
    step()
torch.cuda.synchronize()
_lib.check(lib.y2_event_record(ev[0], stream))
for _ in range(steps):
    step()
_lib.check(lib.y2_event_record(ev[1], stream))
torch.cuda.synchronize()
ms = C.c_float()
_lib.check(lib.y2_event_elapsed_ms(ev[0], ev[1], C.byref(ms)))
ms_step = ms.value / steps
gflop = lib.network_conv_flops(net) / 1e9
kernels = {}
for i in range(net.n):
    if net.layers[i].type == dn.CONVOLUTIONAL:
        k = lib.network_conv_kernel(net, i)
        kernels[k] = kernels.get(k, 0) + 1
print(json.dumps({"cfg": name, "side": side, "batch": batch, "ms_per_step": round(ms_step, 4),
                  "images_per_s": round(batch / ms_step * 1e3, 1), "gflop_per_image": round(gflop, 3),
                  "model_tflops": round(gflop * batch / ms_step, 1),
                  "conv_kernels": {"per-tap": kernels.get(0, 0), "slab": kernels.get(1, 0), "pair": kernels.get(2, 0),
                                   "conv+pool": kernels.get(3, 0), "first-layer": kernels.get(4, 0)},
                  "step": "forward + decode + NMS + pick" if detector else "forward"}), flush=True)
if "--layers" in sys.argv:
    import numpy as np
    buf = (C.c_float * net.n)()
    acc = np.zeros(net.n)
    for _ in range(3):
        lib.network_profile_layers(net, buf, net.n)
        acc += np.ctypeslib.as_array(buf)
    acc /= 3
    names = {0: "per-tap", 1: "slab", 2: "pair", 3: "conv+pool", 4: "first-layer"}
    for i in range(net.n):
        l = net.layers[i]
        k = names.get(lib.network_conv_kernel(net, i), "") if l.type == dn.CONVOLUTIONAL else ""
        print(f"  layer {i:3d} type {l.type:2d} {l.out_w:4d}x{l.out_h:<4d}x{l.out_c:<6d} {acc[i] * 1e3:9.1f} us  {k}")
dn.free_network(net)
