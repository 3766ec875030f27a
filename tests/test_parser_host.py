"""CPU tests of the C host runtime (no GPU, gpu_index = -1: parse only): cfg parsing,
.weights loading and the shape bookkeeping must agree with the reference's parser — checked
against tests/golden/parser_tables.json, the tables printed by the reference's own parser
(on its own cfg files too, see tests/golden/make_golden.py)."""
import ctypes as C
import json
import struct
from pathlib import Path

import numpy as np
import pytest

from sr_object_detection_b200 import darknet as dn
from sr_object_detection_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"
FIELDS = ("type", "w", "h", "c", "out_w", "out_h", "out_c", "outputs", "n", "size", "stride", "pad")


@pytest.fixture(autouse=True)
def _host_only():
    prev = dn.gpu_index()
    dn.set_gpu_index(-1)
    yield
    dn.set_gpu_index(prev)


def _table(net):
    return [{f: int(getattr(net.layers[i], f)) for f in FIELDS} for i in range(net.n)]


CASES = [("tiny-yolo-voc", {}, "tiny-yolo-voc"), ("yolo-voc", {}, "yolo-voc"), ("yolo", {"w": 416, "h": 416}, "yolo"),
         ("yolo", {"w": 608, "h": 608}, "yolo-608"), ("darknet19_448", {}, "darknet19_448"), ("resnet50", {}, "resnet50")]


@pytest.mark.parametrize("name,kw,key", CASES, ids=[c[2] for c in CASES])
def test_layer_tables_match_reference_parser(tmp_path, name, kw, key):
    tables = json.loads((GOLDEN / "parser_tables.json").read_text())
    cfg = tmp_path / "net.cfg"
    cfg.write_text(synth.CFGS[name](batch=1, **kw))
    net = dn.parse_network_cfg(cfg)
    assert net.n == len(tables[key]["layers"])
    assert (net.w, net.h, net.c) == (tables[key]["w"], tables[key]["h"], tables[key]["c"])
    got = _table(net)
    for i, (g, w) in enumerate(zip(got, tables[key]["layers"])):
        assert g == w, f"{key} layer {i}: {g} != {w}"
    assert net.b200 is None  # no device plan without a GPU
    dn.free_network(net)


def test_yolo9000_table_with_synthetic_tree(tmp_path, monkeypatch):
    tables = json.loads((GOLDEN / "parser_tables.json").read_text())
    synth.write_tree(tmp_path / "9k.tree")
    (tmp_path / "net.cfg").write_text(synth.yolo9000_cfg(batch=1, tree="9k.tree"))
    monkeypatch.chdir(tmp_path)  # tree= is resolved against the CWD, as in the reference
    net = dn.parse_network_cfg("net.cfg")
    assert _table(net) == tables["yolo9000"]["layers"]
    l = net.layers[net.n - 1]
    t = l.softmax_tree.contents
    assert t.n == 9418 and l.classes == 9418
    # read_tree group bookkeeping (tree.c:53-103): groups are runs of equal parents
    sizes = [t.group_size[i] for i in range(t.groups)]
    assert sum(sizes) == t.n and sizes[0] == 4 and set(sizes[1:-1]) == {5}
    dn.free_network(net)


def test_cfg_reader_semantics(tmp_path):
    """batch /= subdivisions; blanks stripped everywhere; '#' and ';' comments; pad=1 means
    size/2; maxpool defaults (size=stride, padding=(size-1)/2); conv default activation is
    logistic; anchors parsed with atof/strchr (parser.c:139-171, 236-284, 359-374, 504-577)."""
    text = """[net]
 batch = 64
subdivisions=8
height=32 ; trailing text is part of the value in the reference too
width = 32
channels=3
# comment
; comment

[convolutional]
filters = 10
size=3
stride=1
pad=1

[maxpool]
stride=2

[convolutional]
filters=14
size=1
pad=1
activation=linear

[region]
anchors = 1.5,2.5,  3.25,4.75
classes=2
num=2
"""
    cfg = tmp_path / "r.cfg"
    cfg.write_text(text.replace("height=32 ; trailing text is part of the value in the reference too", "height=32"))
    net = dn.parse_network_cfg(cfg)
    assert net.batch == 8 and net.subdivisions == 8
    c0, mp, c1, rg = (net.layers[i] for i in range(4))
    assert (c0.pad, c0.out_w, c0.out_h, c0.activation) == (1, 32, 32, 0)  # LOGISTIC = 0
    assert (mp.size, mp.stride, mp.pad, mp.out_w) == (2, 2, 0, 16)
    assert (c1.pad, c1.out_w, c1.activation) == (0, 16, 3)  # 1x1 with pad=1 -> padding 0; LINEAR = 3
    assert rg.type == dn.REGION and rg.outputs == 16 * 16 * 14
    assert [rg.biases[i] for i in range(4)] == [1.5, 2.5, 3.25, 4.75]
    dn.free_network(net)


@pytest.mark.parametrize("major,minor", [(0, 1), (0, 2)])
def test_weights_loader_layout(tmp_path, major, minor):
    """parser.c:1009-1082: header int32 x3, `seen` int32 or (major*10+minor >= 2) 64-bit; per conv
    biases, [scales, mean, variance], weights — BN arrays absent for non-BN layers."""
    cfg_text = synth.mini_yolo_cfg(batch=1)
    cfg = tmp_path / "net.cfg"
    cfg.write_text(cfg_text)
    w = tmp_path / "net.weights"
    synth.write_weights(w, cfg_text, seed=77)
    raw = w.read_bytes()
    body = raw[16:]
    header = struct.pack("<iii", major, minor, 0) + (struct.pack("<q", 5) if major * 10 + minor >= 2 else struct.pack("<i", 5))
    w.write_bytes(header + body)
    net = dn.parse_network_cfg(cfg)
    dn.load_weights(net, w)
    vals = np.frombuffer(body, np.float32)
    off = 0
    for i in range(net.n):
        l = net.layers[i]
        if l.type != dn.CONVOLUTIONAL:
            continue
        n, num = l.n, l.n * l.c * l.size * l.size
        assert np.array_equal(np.ctypeslib.as_array(l.biases, (n,)), vals[off:off + n]); off += n
        if l.batch_normalize:
            for arr in (l.scales, l.rolling_mean, l.rolling_variance):
                assert np.array_equal(np.ctypeslib.as_array(arr, (n,)), vals[off:off + n]); off += n
        assert np.array_equal(np.ctypeslib.as_array(l.weights, (num,)), vals[off:off + num]); off += num
    assert off == vals.size
    dn.free_network(net)


def test_resize_network_recomputes_shapes(tmp_path):
    """resize_network (network.c:322-388): yolo.cfg says 416, the north star runs it at 608."""
    tables = json.loads((GOLDEN / "parser_tables.json").read_text())
    cfg = tmp_path / "net.cfg"
    cfg.write_text(synth.yolo_coco_cfg(batch=1, w=416, h=416))
    net = dn.parse_network_cfg(cfg)
    assert dn.resize_network(net, 608, 608) == 0
    assert _table(net) == tables["yolo-608"]["layers"]
    assert (net.w, net.h, net.inputs) == (608, 608, 3 * 608 * 608)
    assert dn.lib().get_network_output_size(net) == 19 * 19 * 425
    dn.free_network(net)


def test_predict_without_gpu_fails_loudly(tmp_path):
    """No CPU fallback: network_predict on a host-only network aborts through error()
    (utils.c:195-200 convention), it never computes on the CPU."""
    import subprocess
    import sys
    cfg = tmp_path / "net.cfg"
    cfg.write_text(synth.mini_yolo_cfg(batch=1))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from sr_object_detection_b200 import darknet as dn\n"
        "dn.set_gpu_index(-1)\n"
        "net = dn.parse_network_cfg(%r)\n"
        "dn.network_predict(net, np.zeros((1, 3, 32, 32), np.float32))\n"
        "print('COMPUTED')\n" % (str(Path(__file__).resolve().parents[1]), str(cfg)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "COMPUTED" not in r.stdout
    assert "no CPU execution path" in r.stderr


def test_host_helpers_match_reference_semantics():
    lib = dn.lib()
    a = np.array([0.1, 0.7, 0.7, -1.0], np.float32)
    assert lib.max_index(a.ctypes.data_as(C.POINTER(C.c_float)), 4) == 1  # first max wins (utils.c:533-545)
    b1, b2 = dn.Box(0.5, 0.5, 0.2, 0.2), dn.Box(0.55, 0.5, 0.2, 0.2)
    inter = np.float32(np.float32(0.15) * np.float32(0.2))
    # box.c:67-97 in float arithmetic
    l1 = np.float32(0.5) - np.float32(0.2) / np.float32(2)
    l2 = np.float32(0.55) - np.float32(0.2) / np.float32(2)
    r1 = np.float32(0.5) + np.float32(0.2) / np.float32(2)
    r2 = np.float32(0.55) + np.float32(0.2) / np.float32(2)
    w = np.float32(min(r1, r2) - max(l1, l2))
    hgt = np.float32(np.float32(0.6) - np.float32(0.4))
    inter = np.float32(w * hgt)
    union = np.float32(np.float32(np.float32(0.2) * np.float32(0.2) + np.float32(0.2) * np.float32(0.2)) - inter)
    assert lib.box_iou(b1, b2) == pytest.approx(float(inter / union), rel=1e-6)


# ---- vector / activation helpers of the reference's C surface (blas.h:16-19, activations.h:16, cuda.h:32) ----
def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_host_vector_helpers_follow_blas_semantics():
    lib = dn.lib()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(40).astype(np.float32)
    y = rng.standard_normal(40).astype(np.float32)
    y0 = y.copy()
    lib.axpy_cpu(13, 0.5, _fp(x), 3, _fp(y), 2)          # Y[i*2] += .5 * X[i*3]
    want = y0.copy()
    want[0:26:2] += np.float32(0.5) * x[0:39:3]
    assert np.array_equal(y, want)
    lib.scal_cpu(10, 3.0, _fp(y), 4)
    want[0:40:4] *= np.float32(3.0)
    assert np.array_equal(y, want)
    lib.copy_cpu(8, _fp(x), 5, _fp(y), 1)
    want[:8] = x[0:40:5]
    assert np.array_equal(y, want)
    lib.fill_cpu(7, -2.0, _fp(y), 3)
    want[0:21:3] = -2.0
    assert np.array_equal(y, want)


def _ref_activate(x, a):
    """activations.h:22-60 restated with numpy: float in, double math, one rounding to float."""
    x = x.astype(np.float32)
    xd = x.astype(np.float64)
    f32 = lambda v: np.asarray(v, dtype=np.float64).astype(np.float32)  # noqa: E731
    if a == dn.LINEAR:
        return x
    if a == dn.LOGISTIC:
        return f32(1. / (1. + np.exp(-xd)))
    if a == dn.LOGGY:
        return f32(2. / (1. + np.exp(-xd)) - 1)
    if a == dn.RELU:
        return x * (x > 0)
    if a == dn.ELU:
        return f32((xd >= 0) * xd + (xd < 0) * (np.exp(xd) - 1))
    if a == dn.RELIE:
        return np.where(x > 0, x, f32(.01 * xd))
    if a == dn.RAMP:
        return f32(xd * (xd > 0) + .1 * xd)
    if a == dn.LEAKY:
        return np.where(x > 0, x, f32(.1 * xd))
    if a == dn.TANH:
        e = np.exp((np.float32(2) * x).astype(np.float64))
        return f32((e - 1) / (e + 1))
    if a == dn.PLSE:
        return f32(np.where(xd < -4, .01 * (xd + 4), np.where(xd > 4, .01 * (xd - 4) + 1, .125 * xd + .5)))
    if a == dn.STAIR:
        n = np.floor(xd)
        frac = (x - n.astype(np.float32)).astype(np.float64)  # float - int is a float subtraction in C
        return f32(np.where(n % 2 == 0, np.floor(xd / 2.), frac + np.floor(xd / 2.)))
    if a == dn.HARDTAN:
        return np.clip(x, -1, 1)
    if a == dn.LHTAN:
        return f32(np.where(xd < 0, .001 * xd, np.where(xd > 1, .001 * (xd - 1) + 1, xd)))
    raise AssertionError(a)


@pytest.mark.parametrize("a", range(13))
def test_activate_array_matches_reference_definitions(a):
    lib = dn.lib()
    x = np.concatenate([np.linspace(-6, 6, 241), [0.0, -0.0, 1.0, -1.0, 4.0, -4.0]]).astype(np.float32)
    got = x.copy()
    lib.activate_array(_fp(got), got.size, a)
    want = _ref_activate(x, a)
    assert np.allclose(got, want, rtol=2e-7, atol=1e-9), np.abs(got - want).max()
    if a in (dn.LINEAR, dn.RELU, dn.LEAKY, dn.RELIE, dn.HARDTAN, dn.LOGISTIC):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,want", [(1, (1, 1, 1)), (512, (1, 1, 1)), (513, (2, 1, 1)), (512 * 65535, (65535, 1, 1)),
                                    (512 * 65535 + 1, (256, 256, 1)), (10 ** 9, (1398, 1398, 1))])
def test_cuda_gridsize(n, want):
    """cuda.c:51-62: blocks of BLOCK=512 threads, folded into (x, y) above 65535 blocks."""
    d = dn.lib().cuda_gridsize(n)
    assert (d.x, d.y, d.z) == want
    assert d.x * d.y * 512 >= n


def test_classifier_front_end_matches_reference_golden():
    """letterbox_image (image.c:1624-1644) and top_k (utils.c:179-193) against outputs of the compiled
    reference (tests/golden/classifier_front.npz, written by tests/golden/make_golden.py).  Bit-exact."""
    lib = dn.lib()
    d = np.load(GOLDEN / "classifier_front.npz")
    for tag in ("wide", "tall", "same"):
        im = np.ascontiguousarray(d[f"{tag}_image"])
        want = d[f"{tag}_boxed"]
        c, h, w = im.shape
        _, oh, ow = want.shape
        out = lib.letterbox_image(dn.Image(h, w, c, _fp(im)), ow, oh)
        assert (out.h, out.w, out.c) == (oh, ow, c)
        got = np.ctypeslib.as_array(out.data, shape=(c, oh, ow)).copy()
        lib.free_image(out)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), tag
    a = np.ascontiguousarray(d["topk_input"])
    for k in (1, 5, 12):
        idx = (C.c_int * k)()
        lib.top_k(_fp(a), a.size, k, idx)
        assert list(idx) == d[f"topk_{k}"].tolist()


def test_find_replace_and_read_boxes_follow_the_reference(tmp_path):
    """utils.c:158-172 (first occurrence only, output may alias the input) and data.c:135-159 ("id x y w h" records,
    corners derived in float, reading stops at the first record that does not parse)."""
    import ctypes as C
    lib = dn.lib()
    lib.find_replace.restype = None
    lib.find_replace.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    out = C.create_string_buffer(4096)
    lib.find_replace(b"data/images/images_01.png", b"images", b"labels", out)
    assert out.value == b"data/labels/images_01.png"
    lib.find_replace(out, b".png", b".txt", out)          # in place, like detector.c:413-418
    assert out.value == b"data/labels/images_01.txt"
    lib.find_replace(out, b".jpg", b".txt", out)          # no occurrence: unchanged
    assert out.value == b"data/labels/images_01.txt"

    class BoxLabel(C.Structure):
        _fields_ = [("id", C.c_int)] + [(k, C.c_float) for k in ("x", "y", "w", "h", "left", "right", "top", "bottom")]

    lib.read_boxes.restype = C.POINTER(BoxLabel)
    lib.read_boxes.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    p = tmp_path / "l.txt"
    p.write_text("3 0.5 0.25 0.2 0.1\n7 0.125 0.75 0.5 0.5\nnot a record\n1 0.1 0.1 0.1 0.1\n")
    n = C.c_int()
    b = lib.read_boxes(str(p).encode(), C.byref(n))
    assert n.value == 2
    assert (b[0].id, b[0].x, b[0].y, b[0].w, b[0].h) == (3, 0.5, 0.25, np.float32(0.2), np.float32(0.1))
    assert b[0].left == np.float32(0.5) - np.float32(0.2) / 2 and b[0].bottom == np.float32(0.25) + np.float32(0.1) / 2
    assert (b[1].id, b[1].left, b[1].right, b[1].top, b[1].bottom) == (7, -0.125, 0.375, 0.5, 1.0)
