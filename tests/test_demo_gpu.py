"""demo_frames: the fetch / detect video pipeline of demo.c (:57-230) on the GPU forward pass, fed by a frame callback.
Per frame the sink must receive exactly what the reference's chain produces - network_predict, the mean of the last
three outputs, get_region_boxes, do_nms(.4) (demo.c:71-107) - as dumped by oracle/_ref/ref_demo, which runs those
reference functions on the same frames (tests/golden/demo_ref.npz).  Exactly representable detector, frames of 0 / 255
bytes: bit for bit."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from sr_object_detection_b200 import darknet as dn
from sr_object_detection_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "demo_ref.npz"


def test_demo_pipeline_frames_equal_the_reference_chain(tmp_path):
    g = np.load(GOLD)
    w, h, n, thresh = int(g["w"]), int(g["h"]), int(g["n"]), float(g["thresh"])
    cfg_text = synth.exact_detector_cfg(batch=1, w=w, h=h)
    (tmp_path / "net.cfg").write_text(cfg_text)
    synth.write_exact_weights(tmp_path / "net.weights", cfg_text)
    frames = synth.binary_frames(n, h, w, seed=int(g["seed"]))
    lib = dn.lib()
    dn.set_gpu_index(0)
    lib.cuda_set_device(0)
    fp = C.POINTER(C.c_float)
    SRC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int)
    SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(dn.Box), C.POINTER(fp), C.c_int, C.c_int)
    fetched = []
    got = {}

    def source(_ctx, rgb, fw, fh):
        i = len(fetched)
        if i >= n:
            return 0
        assert (fw, fh) == (w, h)
        C.memmove(rgb, frames[i].ctypes.data, frames[i].nbytes)
        fetched.append(i)
        return 1

    def sink(_ctx, frame, boxes, probs, total, classes):
        b = np.ctypeslib.as_array(C.cast(boxes, fp), (total * 4,)).copy()
        p = np.concatenate([np.ctypeslib.as_array(probs[j], (classes,)) for j in range(total)])
        got[frame] = np.concatenate([b, p])

    src_cb, sink_cb = SRC(source), SINK(sink)
    lib.demo_frames.restype = C.c_int
    lib.demo_frames.argtypes = [C.c_char_p, C.c_char_p, C.c_float, SRC, C.c_void_p, SINK, C.c_void_p]
    with dn._quiet_stderr():
        done = lib.demo_frames(str(tmp_path / "net.cfg").encode(), str(tmp_path / "net.weights").encode(), thresh, src_cb,
                               None, sink_cb, None)
    assert done == n and sorted(got) == list(range(n))
    nonzero = 0
    for f in range(n):
        want = g[f"frame_{f}"]
        assert got[f].shape == want.shape
        assert np.array_equal(got[f].view(np.uint32), want.view(np.uint32)), \
            f"frame {f}: {(got[f] != want).sum()} of {want.size} values differ"
        nonzero += int((want[w * h * 3 * 4:] != 0).sum())
    assert nonzero > 100
