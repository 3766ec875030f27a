"""tools/voc_eval.py against the reference's own scripts/voc_eval.py: tests/golden/voc_eval_ref.json holds a synthetic
VOC-style data set (annotations with difficult objects, detections with duplicates, misses, confidence ties) and
the (recall, precision, AP) the reference's script returned for it (tests/golden/make_golden.py voc_eval).  Same
numbers, every point of the curves."""
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
spec = importlib.util.spec_from_file_location("voc_eval", ROOT / "tools" / "voc_eval.py")
voc_eval = importlib.util.module_from_spec(spec)
spec.loader.exec_module(voc_eval)


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = json.loads((ROOT / "tests" / "golden" / "voc_eval_ref.json").read_text())
    t = tmp_path_factory.mktemp("voc")
    (t / "ann").mkdir()
    for im in d["images"]:
        xml = "<annotation>" + "".join(
            "<object><name>%s</name><difficult>%d</difficult><bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax>"
            "<ymax>%d</ymax></bndbox></object>" % (o["name"], o["difficult"], *o["bbox"]) for o in d["truth"][im]) + "</annotation>"
        (t / "ann" / f"{im}.xml").write_text(xml)
    (t / "set.txt").write_text("\n".join(d["images"]) + "\n")
    for c in d["classes"]:
        (t / f"det_{c}.txt").write_text("\n".join(d["detections"][c]) + "\n")
    return d, t


@pytest.mark.parametrize("voc07", [False, True])
@pytest.mark.parametrize("ov", [0.5, 0.7])
def test_curves_and_ap_equal_the_reference_script(data, voc07, ov):
    d, t = data
    res = voc_eval.evaluate(str(t / "det_{}.txt"), str(t / "ann" / "{}.xml"), str(t / "set.txt"), d["classes"], ov, voc07)
    for c in d["classes"]:
        want = d["results"][f"{c}|{int(voc07)}|{ov}"]
        rec, prec, ap = res[c]
        assert len(rec) == len(want["rec"]) > 10
        assert np.array_equal(rec, np.array(want["rec"])) and np.array_equal(prec, np.array(want["prec"]))
        assert ap == want["ap"]


def test_annotation_parser_reads_what_the_fixture_wrote(data):
    d, t = data
    for im in d["images"][:10]:
        assert voc_eval.parse_annotation(t / "ann" / f"{im}.xml") == d["truth"][im]


def test_average_precision_on_hand_made_curves():
    # one detection, correct: recall 1, precision 1 -> AP 1 under both definitions
    assert voc_eval.average_precision(np.array([1.0]), np.array([1.0])) == 1.0
    assert abs(voc_eval.average_precision(np.array([1.0]), np.array([1.0]), voc07=True) - 1.0) < 1e-12
    # two ground truths, detections: hit, miss, hit -> recall .5 .5 1, precision 1 .5 2/3
    rec, prec = np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2.0 / 3.0])
    assert abs(voc_eval.average_precision(rec, prec) - (0.5 * 1.0 + 0.5 * 2.0 / 3.0)) < 1e-12
    want07 = (6 * 1.0 + 5 * 2.0 / 3.0) / 11.0   # thresholds 0 ... 0.5 see precision 1, 0.6 ... 1.0 see 2/3
    assert abs(voc_eval.average_precision(rec, prec, voc07=True) - want07) < 1e-12
