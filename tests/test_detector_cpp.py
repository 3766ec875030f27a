"""The C++ `Detector` facade (include/yolo_v2_class.hpp).

Pinned to the reference: tests/cpp/detector_scenario.cpp uses nothing but the public Detector surface; compiled
against the REFERENCE's yolo_v2_class.{hpp,cpp} (oracle/Makefile refdet) it wrote tests/golden/detector_ref.json,
compiled against our header + libyolo2_b200.so it must print the same lines character for character (detect,
tracking, use_mean, detect(filename), nms = 0, the load_image exception).  The older tests below compare a second
driver with an independent Python statement of yolo_v2_class.cpp and with the C API path."""
import json
import math
import subprocess
from pathlib import Path

import numpy as np
import pytest

from sr_object_detection_b200 import synth

ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "sr_object_detection_b200"


def _build(tmp_path: Path) -> Path:
    from sr_object_detection_b200 import build as b
    b.build()
    exe = tmp_path / "detector_check"
    cmd = ["g++", "-O1", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "detector_check.cpp"),
           "-L", str(LIBDIR), "-lyolo2_b200", f"-Wl,-rpath,{LIBDIR}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _build_scenario(tmp_path: Path) -> Path:
    from sr_object_detection_b200 import build as b
    b.build()
    exe = tmp_path / "detector_scenario"
    cmd = ["g++", "-O1", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "detector_scenario.cpp"),
           "-L", str(LIBDIR), "-lyolo2_b200", f"-Wl,-rpath,{LIBDIR}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _scenario_inputs(tmp_path: Path, g: dict):
    (tmp_path / "net.cfg").write_text(g["cfg"])
    synth.write_exact_weights(tmp_path / "net.weights", g["cfg"])
    synth.exact_frames(g["n"], g["h"], g["w"]).tofile(tmp_path / "frames.f32")
    ppm = np.random.default_rng(g["ppm_seed"]).integers(0, 2, (g["h"], g["w"], 3)).astype(np.uint8) * 255
    (tmp_path / "image.ppm").write_bytes(b"P6\n%d %d\n255\n" % (g["w"], g["h"]) + ppm.tobytes())


def test_scripted_tracking_equals_reference_detector_golden(tmp_path):
    """tracking() is host logic: runs here without a GPU (gpu_id = -1), against lines the reference's own
    Detector::tracking printed (yolo_v2_class.cpp:251-304)."""
    g = json.loads((ROOT / "tests" / "golden" / "detector_ref.json").read_text())
    exe = _build_scenario(tmp_path)
    _scenario_inputs(tmp_path, g)
    r = subprocess.run([str(exe), "net.cfg", "net.weights", "frames.f32", "0", str(g["w"]), str(g["h"]), str(g["thresh"]),
                        str(g["nms"]), str(g["story"]), "image.ppm", "-1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l for l in r.stdout.splitlines() if l.startswith("script ")]
    want = [l for l in g["lines"] if l.startswith("script ")]
    assert len(want) == 6 and got == want


@pytest.mark.gpu
def test_detector_equals_reference_detector_golden(tmp_path):
    """Detector::detect / tracking / use_mean / detect(filename) on the GPU against bbox_t lists the reference's
    own class produced on its CPU path (yolo_v2_class.cpp:173-304), bit for bit: the scenario network's head output
    is exactly representable, so nothing in the pipeline may differ."""
    g = json.loads((ROOT / "tests" / "golden" / "detector_ref.json").read_text())
    exe = _build_scenario(tmp_path)
    _scenario_inputs(tmp_path, g)
    r = subprocess.run([str(exe), "net.cfg", "net.weights", "frames.f32", str(g["n"]), str(g["w"]), str(g["h"]),
                        str(g["thresh"]), str(g["nms"]), str(g["story"]), "image.ppm", "0"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    tags = ("size", "detect", "track", "mean", "file", "loaded", "nonms", "load", "script")
    got = [l for l in r.stdout.splitlines() if l.split() and l.split()[0] in tags]
    assert len(got) == len(g["lines"])
    for a, b in zip(got, g["lines"]):
        assert a == b, f"first difference:\n ours: {a[:300]}\n ref:  {b[:300]}"
    assert sum(int(l.split()[2]) for l in got if l.startswith("detect ")) > 100  # not vacuous


def _parse(line: str):
    t = line.split()
    n = int(t[1])
    vals = t[2:]
    out = []
    for i in range(n):
        x, y, w, h, prob, obj, tid = vals[7 * i:7 * i + 7]
        out.append((int(x), int(y), int(w), int(h), float(prob), int(obj), int(tid)))
    return out


def _track_reference(frames, story):
    """yolo_v2_class.cpp:251-304 restated: nearest same-class box of the remembered frames within 100 px
    hands over its id (closer claim wins, an id is not reused inside a frame), extents are averaged."""
    next_id = {}
    hist = []  # most recent first
    results = []

    def fresh(obj):
        next_id.setdefault(obj, 1)
        v = next_id[obj]
        next_id[obj] += 1
        return v

    for frame in frames:
        cur = [dict(x=x, y=y, w=w, h=h, obj=obj, tid=0) for (x, y, w, h, obj) in frame]
        if not any(len(f) for f in hist):
            for b in cur:
                b["tid"] = fresh(b["obj"])
        else:
            best = [2 ** 32 - 1] * len(cur)
            for old_frame in hist:
                for old in old_frame:
                    match = -1
                    for m, k in enumerate(cur):
                        if old["obj"] != k["obj"]:
                            continue
                        dx = float(old["x"] + old["w"] // 2) - float(k["x"] + k["w"] // 2)
                        dy = float(old["y"] + old["h"] // 2) - float(k["y"] + k["h"] // 2)
                        dist = int(math.sqrt(dx * dx + dy * dy))
                        if dist < 100 and (k["tid"] == 0 or best[m] > dist):
                            best[m] = dist
                            match = m
                    taken = any(b["tid"] == old["tid"] and b["obj"] == old["obj"] for b in cur)
                    if match >= 0 and not taken:
                        cur[match]["tid"] = old["tid"]
                        cur[match]["w"] = (cur[match]["w"] + old["w"]) // 2
                        cur[match]["h"] = (cur[match]["h"] + old["h"]) // 2
            for b in cur:
                if b["tid"] == 0:
                    b["tid"] = fresh(b["obj"])
        hist.insert(0, [dict(b) for b in cur])
        if len(hist) > story:
            hist.pop()
        results.append([(b["x"], b["y"], b["w"], b["h"], b["obj"], b["tid"]) for b in cur])
    return results


def test_tracking_matches_reference_semantics(tmp_path):
    exe = _build(tmp_path)
    cfg_text = synth.mini_yolo_cfg(batch=1)
    (tmp_path / "n.cfg").write_text(cfg_text)
    synth.write_weights(tmp_path / "n.weights", cfg_text, seed=3)
    r = subprocess.run([str(exe), "track", str(tmp_path / "n.cfg"), str(tmp_path / "n.weights")], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [[(x, y, w, h, obj, tid) for (x, y, w, h, _p, obj, tid) in _parse(l)]
           for l in r.stdout.splitlines() if l.startswith("track")]
    frames = [
        [(10, 10, 40, 40, 0), (200, 200, 50, 50, 0), (300, 20, 30, 60, 1)],
        [(14, 12, 44, 40, 0), (205, 190, 50, 54, 0), (500, 400, 30, 30, 1)],
        [],
        [(20, 15, 40, 40, 0), (290, 30, 30, 60, 1), (295, 28, 30, 60, 1)],
        [(400, 400, 10, 10, 2)],
    ]
    assert got == _track_reference(frames, 3)
    assert got[1][0][5] == got[0][0][5] and got[1][2][5] != got[0][2][5]  # a track continues, a far box starts one


@pytest.mark.gpu
def test_detector_detect_matches_c_api(tmp_path):
    from sr_object_detection_b200 import darknet as dn
    exe = _build(tmp_path)
    cfg_text = synth.tiny_yolo_voc_cfg(batch=1)
    (tmp_path / "n.cfg").write_text(cfg_text)
    synth.write_weights(tmp_path / "n.weights", cfg_text, seed=1234)
    rng = np.random.default_rng(11)
    u8 = rng.integers(0, 256, size=(416, 416, 3), dtype=np.uint8)
    planar = (u8.transpose(2, 0, 1).astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    planar.tofile(tmp_path / "in.f32")
    u8.tofile(tmp_path / "in.u8")
    thresh = 0.02
    r = subprocess.run([str(exe), "detect", str(tmp_path / "n.cfg"), str(tmp_path / "n.weights"), str(tmp_path / "in.f32"),
                        str(tmp_path / "in.u8"), "416", "416", str(thresh)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert "size 416 416" in lines
    assert any(l == "load file not found" for l in lines)
    det = _parse(next(l for l in lines if l.startswith("detect")))
    means = [_parse(l) for l in lines if l.startswith("mean")]
    rgb8 = _parse(next(l for l in lines if l.startswith("rgb8")))
    assert len(det) > 0
    assert rgb8 == det, "uint8 frame path must equal the float path"

    # the C API flow of yolo_v2_class.cpp:199-239 through the ctypes mirror
    dn.set_gpu_index(0)
    net = dn.parse_network_cfg(tmp_path / "n.cfg")
    dn.load_weights(net, tmp_path / "n.weights")
    dn.network_predict(net, np.ascontiguousarray(planar[None]))
    boxes, probs = dn.get_region_boxes(net, 0, thresh)
    dn.do_nms_sort(boxes, probs, 0.4)
    want = []
    for i in range(len(boxes)):
        obj = int(np.argmax(probs[i]))
        p = float(probs[i][obj])
        if p > thresh:
            bx, by, bw, bh = (np.float32(v) for v in boxes[i])
            x = int(max(0.0, (float(bx) - float(bw) / 2.0) * 416))
            y = int(max(0.0, (float(by) - float(bh) / 2.0) * 416))
            want.append((x, y, int(np.float32(bw * np.float32(416))), int(np.float32(bh * np.float32(416))), obj))
    assert [(x, y, w, h, obj) for (x, y, w, h, _p, obj, _t) in det] == want
    # use_mean: frame 1 averages with two zero frames, frame 3 is the plain prediction again (same image thrice)
    assert [(b[0], b[1], b[2], b[3], b[5]) for b in means[2]] == want
    dn.free_network(net)


@pytest.mark.gpu
def test_detector_frame_of_another_size_device_resize_equals_host_resize(tmp_path):
    """Detector::detect(image_t) resizes a 640x480 frame on the host with resize_image (yolo_v2_class.cpp:186-193);
    Detector::detect_rgb8 uploads the raw bytes and resizes on the device.  Same boxes, bit for bit."""
    exe = _build(tmp_path)
    cfg_text = synth.tiny_yolo_voc_cfg(batch=1)
    (tmp_path / "n.cfg").write_text(cfg_text)
    synth.write_weights(tmp_path / "n.weights", cfg_text, seed=1234)
    rng = np.random.default_rng(12)
    w, h = 640, 480
    u8 = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    planar = (u8.transpose(2, 0, 1).astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    planar.tofile(tmp_path / "in.f32")
    u8.tofile(tmp_path / "in.u8")
    r = subprocess.run([str(exe), "detect", str(tmp_path / "n.cfg"), str(tmp_path / "n.weights"), str(tmp_path / "in.f32"),
                        str(tmp_path / "in.u8"), str(w), str(h), "0.02"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    det = _parse(next(l for l in lines if l.startswith("detect")))
    rgb8 = _parse(next(l for l in lines if l.startswith("rgb8")))
    assert len(det) > 0
    assert rgb8 == det
    assert max(b[0] + b[2] for b in det) > 416 or max(b[1] + b[3] for b in det) > 0  # boxes are in frame pixels
