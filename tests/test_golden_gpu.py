"""GPU parity against the committed golden fixtures (outputs of the UNMODIFIED reference, see
tests/golden/make_golden.py) — needs nothing but the repo on the GPU box.

  * decode + NMS kernels on identical region inputs: bit-exact (flat softmax, softmax tree,
    tree + coco9k-style map);
  * whole toy networks through parse_network_cfg / load_weights / network_predict: every layer's
    activations within 1e-2 of the layer's max magnitude (bf16 operands, fp32 accumulate).
"""
from pathlib import Path

import numpy as np
import pytest

from sr_object_detection_b200 import darknet as dn

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
ACT_TOL = 1e-2


def _materialise(d, tmp):
    (tmp / "net.cfg").write_text(str(d["cfg"]))
    for k in d.files:
        if k.startswith("aux_"):
            (tmp / ".".join(k[4:].rsplit("_", 1))).write_text(str(d[k]))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).ravel().view(np.uint32)


@pytest.mark.parametrize("name", ["region_voc_13", "region_voc_7_lowthresh", "region_coco_9", "region_tree_220",
                                  "region_tree_220_map", "region_tree_wide"])
def test_decode_nms_bit_exact_vs_reference_golden(tmp_path, monkeypatch, name):
    d = np.load(GOLDEN / f"{name}.npz")
    _materialise(d, tmp_path)
    monkeypatch.chdir(tmp_path)  # tree= / map= paths are relative, as in the reference
    dn.set_gpu_index(0)
    # a region layer alone is not a plannable network: parse host-only, kernels run on the device
    dn.set_gpu_index(-1)
    net = dn.parse_network_cfg("net.cfg")
    dn.set_gpu_index(0)
    dn.lib().cuda_set_device(0)
    got = dn.decode_region_input(net, d["region_in"], float(d["thresh"]), float(d["nms"]), use_map=bool(d["use_map"]))
    for k in ("region_out", "boxes", "probs_pre", "probs_post", "region_after_boxes"):
        same = np.array_equal(_bits(got[k]), _bits(d[k]))
        assert same, f"{name}/{k}: {(got[k].ravel() != d[k]).sum()} of {d[k].size} values differ"
    assert ((got["probs_post"] != 0) == (d["probs_post"].reshape(got["probs_post"].shape) != 0)).all()
    dn.free_network(net)


@pytest.mark.parametrize("name", ["mini_yolo", "mini_yolo_tree", "mini_resnet"])
def test_toy_network_layers_match_reference_golden(tmp_path, monkeypatch, name):
    d = np.load(GOLDEN / f"{name}.npz")
    _materialise(d, tmp_path)
    (tmp_path / "net.weights").write_bytes(d["weights"].tobytes())
    monkeypatch.chdir(tmp_path)
    dn.set_gpu_index(0)
    net = dn.parse_network_cfg("net.cfg")
    dn.load_weights(net, "net.weights")
    x = d["input"]
    out = dn.network_predict(net, x)
    checked = 0
    for i in range(net.n):
        key = "layer_%03d" % i
        if key not in d.files:
            continue
        ref = d[key].reshape(net.batch, -1)
        got = dn.get_network_output_layer(net, i)
        err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
        assert err <= ACT_TOL, f"{name} layer {i}: max err / max|ref| = {err:.3e}"
        checked += 1
    assert checked >= 10
    ref_out = d["output"].reshape(out.shape)
    assert float(np.abs(out - ref_out).max() / np.abs(ref_out).max()) <= ACT_TOL
    dn.free_network(net)


def test_do_nms_unsorted_bit_exact_vs_reference_golden():
    """do_nms (box.c:279-297) through the drop-in entry: caller-owned boxes[total], probs[total][classes]"""
    import ctypes as C
    d = np.load(GOLDEN / "do_nms.npz")
    lib = dn.lib()
    dn.set_gpu_index(0)
    lib.cuda_set_device(0)
    fp = C.POINTER(C.c_float)
    lib.do_nms.restype = None
    lib.do_nms.argtypes = [C.POINTER(dn.Box), C.POINTER(fp), C.c_int, C.c_int, C.c_float]
    total, classes = d["probs"].shape
    for tag in ("a", "b"):
        boxes = np.ascontiguousarray(d["boxes"], np.float32).copy()
        probs = np.ascontiguousarray(d["probs"], np.float32).copy()
        rows = (fp * total)(*[C.cast(probs.ctypes.data + j * classes * 4, fp) for j in range(total)])
        lib.do_nms(boxes.ctypes.data_as(C.POINTER(dn.Box)), rows, total, classes, float(d[f"thresh_{tag}"]))
        assert np.array_equal(_bits(probs), _bits(d[f"out_{tag}"])), \
            f"do_nms {tag}: {(probs != d[f'out_{tag}']).sum()} values differ"


@pytest.mark.parametrize("name", ["region_tree_220", "region_tree_wide"])
def test_sparse_tree_detection_equals_reference_golden(tmp_path, monkeypatch, name):
    """The detection entries of a softmax-tree network do not run the dense 9418-way pass: y2_region_tree_detect walks
    only the groups on the path of classes above .5 and y2_tree_nms_collect works on one (class, value) record per
    box.  Must give exactly the reference's final detections (softmax_tree + hierarchy_predictions +
    get_region_boxes + do_nms_sort + the max_index pick, golden `dets`), incl. groups wider than a warp."""
    d = np.load(GOLDEN / f"{name}.npz")
    _materialise(d, tmp_path)
    monkeypatch.chdir(tmp_path)
    dn.set_gpu_index(-1)
    net = dn.parse_network_cfg("net.cfg")  # host description only (a region layer alone is not a plannable network)
    dn.set_gpu_index(0)
    dn.lib().cuda_set_device(0)
    B = d["region_in"].shape[0]
    dets, counts = dn.tree_detect_region_input(net, d["region_in"], float(d["thresh"]), float(d["nms"]))
    want = d["dets"].reshape(-1, 8)
    assert len(want) > 0
    got = np.array([(b, r["box_index"], r["obj_id"], r["prob"], r["x"], r["y"], r["w"], r["h"])
                    for b in range(B) for r in dets[b]], np.float32).reshape(-1, 8)
    assert got.shape == want.shape, f"{len(got)} detections, the reference has {len(want)}"
    assert np.array_equal(_bits(got), _bits(want))
    # and without NMS: every box whose hierarchy walk ends above .5 with objectness above thresh
    dets0, _ = dn.tree_detect_region_input(net, d["region_in"], float(d["thresh"]), 0.0)
    pre = d["probs_pre"].reshape(B, -1, net.layers[net.n - 1].classes)
    for b in range(B):
        obj = pre[b].argmax(1)
        p = pre[b][np.arange(len(obj)), obj]
        keep = np.nonzero(p > float(d["thresh"]))[0]
        assert [int(r["box_index"]) for r in dets0[b]] == keep.tolist()
        assert [int(r["obj_id"]) for r in dets0[b]] == obj[keep].tolist()
        assert np.array_equal(_bits(np.array([r["prob"] for r in dets0[b]], np.float32)), _bits(p[keep]))
    dn.free_network(net)
