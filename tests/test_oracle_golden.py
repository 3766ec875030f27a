"""CPU tests of the oracle (oracle/y2_oracle.c, a plain-C restatement of the reference's CPU
path) against the golden fixtures in tests/golden/ — outputs of the UNMODIFIED reference
(oracle/_ref/darknet_ref) made by tests/golden/make_golden.py.  Everything must match bit for
bit: the oracle restates the reference's expression types and accumulation order.
"""
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import ref_util as R

GOLDEN = Path(__file__).resolve().parent / "golden"
FORWARD = ["mini_yolo", "mini_resnet", "mini_yolo_tree"]
REGION = ["region_voc_13", "region_voc_7_lowthresh", "region_coco_9", "region_tree_220", "region_tree_220_map", "region_tree_wide"]
SKIP_KEYS = {"cfg", "weights", "input", "thresh", "nms", "layers", "region_in", "use_map"}


@pytest.fixture(scope="module", autouse=True)
def _oracle_built():
    R.build_oracle()
    assert R.have_oracle(), "oracle/_build/y2_oracle could not be built"


def _materialise(d, tmp):
    (tmp / "net.cfg").write_text(str(d["cfg"]))
    for k in d.files:
        if k.startswith("aux_"):
            name = k[4:].rsplit("_", 1)
            (tmp / ".".join(name)).write_text(str(d[k]))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _compare(d, outdir):
    checked = 0
    for k in d.files:
        if k in SKIP_KEYS or k.startswith("aux_"):
            continue
        got = np.fromfile(outdir / f"{k}.f32", np.float32)
        want = d[k].ravel()
        assert got.shape == want.shape, k
        assert np.array_equal(_bits(got), _bits(want)), \
            f"{k}: {(got != want).sum()} of {want.size} values differ, max |d| = {np.abs(got - want).max():.3e}"
        checked += 1
    return checked


@pytest.mark.parametrize("name", FORWARD)
def test_forward_matches_reference_golden(tmp_path, name):
    d = np.load(GOLDEN / f"{name}.npz")
    _materialise(d, tmp_path)
    (tmp_path / "net.weights").write_bytes(d["weights"].tobytes())
    d["input"].tofile(tmp_path / "in.f32")
    out = tmp_path / "out"
    out.mkdir()
    R.run_raw([R.ORACLE_BIN, "forward", "net.cfg", "net.weights", "in.f32", "out", float(d["thresh"]), float(d["nms"]), 1],
              cwd=tmp_path)
    assert _compare(d, out) >= 5
    table = R.run_raw([R.ORACLE_BIN, "layers", "net.cfg"], cwd=tmp_path)
    assert table == json.loads(str(d["layers"]))


@pytest.mark.parametrize("name", REGION)
def test_decode_and_nms_match_reference_golden(tmp_path, name):
    d = np.load(GOLDEN / f"{name}.npz")
    _materialise(d, tmp_path)
    d["region_in"].tofile(tmp_path / "in.f32")
    out = tmp_path / "out"
    out.mkdir()
    R.run_raw([R.ORACLE_BIN, "region", "net.cfg", "in.f32", "out", float(d["thresh"]), float(d["nms"])], cwd=tmp_path,
              env={"Y2_USE_MAP": str(int(d["use_map"]))})
    assert _compare(d, out) == 6  # region_out, boxes, probs_pre, probs_post, region_after_boxes, dets
    # the fixtures are not vacuous: NMS removed something
    assert (d["probs_pre"] != 0).sum() > (d["probs_post"] != 0).sum() > 0


def test_do_nms_unsorted_matches_reference_golden(tmp_path):
    """do_nms (box.c:279-297), the variant demo.c and validate_detector_recall call"""
    d = np.load(GOLDEN / "do_nms.npz")
    d["boxes"].tofile(tmp_path / "b.f32")
    d["probs"].tofile(tmp_path / "p.f32")
    total, classes = d["probs"].shape
    for tag in ("a", "b"):
        subprocess.run([str(R.ORACLE_BIN), "donms", "b.f32", "p.f32", str(total), str(classes),
                        repr(float(d[f"thresh_{tag}"])), "o.f32"], cwd=tmp_path, check=True)
        got = np.fromfile(tmp_path / "o.f32", np.float32).reshape(total, classes)
        assert np.array_equal(_bits(got), _bits(d[f"out_{tag}"]))
        assert 0 < (got != 0).sum() < (d["probs"] != 0).sum()


def test_resize_matches_reference_golden(tmp_path):
    d = np.load(GOLDEN / "resize.npz")
    im = d["image"]
    im.tofile(tmp_path / "im.f32")
    c, h, w = im.shape
    _, oh, ow = d["resized"].shape
    subprocess.run([str(R.ORACLE_BIN), "resize", "im.f32", str(c), str(h), str(w), str(oh), str(ow), "out.f32"],
                   cwd=tmp_path, check=True)
    got = np.fromfile(tmp_path / "out.f32", np.float32).reshape(c, oh, ow)
    assert np.array_equal(_bits(got), _bits(d["resized"]))


def test_parser_tables_match_reference(tmp_path):
    """Layer tables (shapes, strides, pads, route/shortcut wiring sizes) of the north-star cfgs
    as the reference's parser printed them — including on the reference's own cfg files."""
    from sr_object_detection_b200 import synth
    tables = json.loads((GOLDEN / "parser_tables.json").read_text())
    synth.write_tree(tmp_path / "9k.tree")
    cases = {"tiny-yolo-voc": {}, "yolo-voc": {}, "yolo": {"w": 416, "h": 416}, "darknet19_448": {}, "resnet50": {},
             "yolo9000": {"tree": "9k.tree"}}
    for name, kw in cases.items():
        (tmp_path / "s.cfg").write_text(synth.CFGS[name](batch=1, **kw))
        got = R.run_raw([R.ORACLE_BIN, "layers", "s.cfg"], cwd=tmp_path)
        assert got["layers"] == tables[name]["layers"], name
    (tmp_path / "s.cfg").write_text(synth.yolo_coco_cfg(batch=1, w=608, h=608))
    assert R.run_raw([R.ORACLE_BIN, "layers", "s.cfg"], cwd=tmp_path)["layers"] == tables["yolo-608"]["layers"]


def test_oracle_equals_compiled_reference_on_fresh_seed(tmp_path):
    """Where the compiled reference is present (oracle/_ref, build container and GPU box), a
    fresh seeded case must agree bit for bit as well — guards against fixture over-fitting."""
    if not R.have_ref():
        pytest.skip("oracle/_ref/darknet_ref not built here")
    from sr_object_detection_b200 import synth
    cfg_text = synth.mini_yolo_cfg(batch=3, w=64, h=32, classes=6, num=2)
    (tmp_path / "net.cfg").write_text(cfg_text)
    synth.write_weights(tmp_path / "net.weights", cfg_text, seed=int.from_bytes(os.urandom(2), "little"))
    synth.images(3, 3, 32, 64, seed=int.from_bytes(os.urandom(2), "little")).tofile(tmp_path / "in.f32")
    for tag, binary in (("ref", R.REF_BIN), ("port", R.ORACLE_BIN)):
        (tmp_path / tag).mkdir()
        R.run_raw([binary, "forward", "net.cfg", "net.weights", "in.f32", tag, 0.01, 0.3, 1], cwd=tmp_path)
    names = sorted(p.name for p in (tmp_path / "ref").iterdir())
    assert len(names) >= 20
    for n in names:
        a = np.fromfile(tmp_path / "ref" / n, np.float32)
        b = np.fromfile(tmp_path / "port" / n, np.float32)
        assert a.shape == b.shape and np.array_equal(_bits(a), _bits(b)), n
