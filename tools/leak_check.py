#!/usr/bin/env python
"""Developer tool (GPU box): device-memory leak check - parse / predict / detect / resize / free in a loop."""
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sr_object_detection_b200 import darknet as dn, synth  # noqa: E402

tmp = Path(tempfile.mkdtemp())
dn.set_gpu_index(0)
torch.cuda.init()


def cycle(name, side, batch):
    text = synth.CFGS[name](batch=batch, w=side, h=side)
    (tmp / "n.cfg").write_text(text)
    w = tmp / f"{name}.weights"
    if not w.exists():
        synth.write_weights(w, text, seed=1234)
    net = dn.parse_network_cfg(tmp / "n.cfg")
    dn.load_weights(net, w)
    x = synth.images(batch, 3, side, side, seed=1)
    dn.network_predict(net, x)
    dn.network_detect_batch(net, x, 0.02, 0.4, 64)
    lib = dn.lib()
    import ctypes as C
    dets = (dn.Detection * (batch * 64))()
    counts = (C.c_int * batch)()
    for _ in range(2):  # two batches in flight, then drain
        slot = lib.network_pipeline_next_slot(net)
        lib.network_detect_submit(net, lib.network_pipeline_staging(net, slot), 0.02, 0.4, 64)
    lib.network_detect_wait(net, dets, counts, 64)
    lib.network_detect_wait(net, dets, counts, 64)
    dn.set_batch_network(net, max(1, batch // 2))
    dn.network_predict(net, np.ascontiguousarray(x[:max(1, batch // 2)]))
    dn.resize_network(net, side - 32, side - 32)
    dn.free_network(net)


for name, side, batch in (("tiny-yolo-voc", 416, 8), ("yolo-voc", 416, 64)):
    cycle(name, side, batch)  # warm: allocator pools, per-device scratch
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(15):
        cycle(name, side, batch)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    print(f"{name} b{batch}: free before {free0 >> 20} MiB, after 15 cycles {free1 >> 20} MiB, delta {(free0 - free1) >> 20} MiB")
