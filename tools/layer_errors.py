#!/usr/bin/env python
"""Developer tool (GPU box): per-layer max|got-ref|/max|ref| of a cfg against oracle/_ref/darknet_ref."""
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from sr_object_detection_b200 import darknet as dn, synth  # noqa: E402
from tests import ref_util as R  # noqa: E402

name, batch, side = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tmp = Path(tempfile.mkdtemp())
cfg_text = synth.CFGS[name](batch=batch, w=side, h=side)
(tmp / "n.cfg").write_text(cfg_text)
synth.write_weights(tmp / "n.weights", cfg_text, seed=1234)
x = synth.images(batch, 3, side, side, seed=42)
x.tofile(tmp / "in.f32")
R.forward(R.REF_BIN, tmp / "n.cfg", tmp / "n.weights", tmp / "in.f32", tmp / "ref", thresh=0.24, nms=0.4)
dn.set_gpu_index(0)
net = dn.parse_network_cfg(tmp / "n.cfg")
dn.load_weights(net, tmp / "n.weights")
dn.network_predict(net, x)
for i in range(net.n):
    l = net.layers[i]
    if l.type == dn.COST:
        continue
    ref = R.load(tmp / "ref", "layer_%03d.f32" % i, (batch, l.outputs))
    got = dn.get_network_output_layer(net, i)
    d = np.abs(got - ref)
    print(i, l.type, "max|ref|=%.3e" % np.abs(ref).max(), "rms|ref|=%.3e" % np.sqrt((ref ** 2).mean()),
          "maxerr/max=%.3e" % (d.max() / max(np.abs(ref).max(), 1e-30)),
          "rmserr/rms=%.3e" % (np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30)))
