"""validate_detector (detector.c:244-369): the VOC / COCO / ImageNet-detection result files written through the GPU
path against the files the REFERENCE's own validate_detector wrote on its CPU path (oracle/_ref/ref_validate,
stored in tests/golden/validate_ref.npz by tests/golden/make_golden.py) - byte for byte.  The validation network is
exactly representable (synth.exact_detector_cfg + images whose bytes are 0 / 255), so its head output is bit-identical
on both sides and every later stage - region forward, get_region_boxes with the image's pixel size and the 200-class
`map`, do_nms_sort at .45, the corner clipping and the %f formatting - must agree to the last character."""
import ctypes as C
import os
from pathlib import Path

import numpy as np
import pytest

from sr_object_detection_b200 import darknet as dn
from sr_object_detection_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden" / "validate_ref.npz"


@pytest.mark.parametrize("kind,chunk", [("voc", None), ("voc", "2"), ("coco", None), ("imagenet", "3")])
def test_validation_files_equal_the_reference_files(tmp_path, monkeypatch, kind, chunk):
    gold = np.load(GOLDEN)
    synth.write_validation_set(tmp_path, kind)
    monkeypatch.chdir(tmp_path)  # the data cfg uses relative paths, like the reference's cfg/*.data files
    if chunk:
        monkeypatch.setenv("Y2_VALID_BATCH", chunk)  # several forward passes and a partial last chunk
    else:
        monkeypatch.delenv("Y2_VALID_BATCH", raising=False)
    lib = dn.lib()
    dn.set_gpu_index(0)
    lib.cuda_set_device(0)
    lib.validate_detector.restype = None
    lib.validate_detector.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    with dn._quiet_stderr():
        lib.validate_detector(b"data.cfg", b"net.cfg", b"net.weights")
    want = {k.split("/", 1)[1]: gold[k].tobytes() for k in gold.files if k.startswith(kind + "/")}
    got = {p.name: p.read_bytes() for p in (tmp_path / "results").iterdir()}
    assert sorted(got) == sorted(want)
    for name in want:
        if got[name] != want[name]:
            a, b = got[name].splitlines(), want[name].splitlines()
            first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            raise AssertionError(f"{kind}/{name}: {len(a)} lines vs {len(b)}; first difference at line {first}: "
                                 f"{a[first:first + 1]} vs {b[first:first + 1]}")
    assert sum(len(v) for v in want.values()) > 10000  # not vacuous


class _CaptureStderr:
    """file descriptor 2 of this process into a file (the library prints from C)"""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        import sys
        sys.stderr.flush()
        self.saved = os.dup(2)
        self.fd = os.open(self.path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        os.dup2(self.fd, 2)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 2)
        os.close(self.saved)
        os.close(self.fd)


@pytest.mark.parametrize("chunk", [None, "3"])
def test_recall_lines_equal_the_reference_lines(tmp_path, monkeypatch, chunk):
    """validate_detector_recall (detector.c:371-450): objectness proposals (get_region_boxes with only_objectness,
    do_nms at .4) against label files; the per-image lines - running proposals per image, mean best IoU, recall - must
    be the ones the REFERENCE's own function printed on its CPU path (tests/golden/recall_ref.json)."""
    import json
    want = json.loads((GOLDEN.parent / "recall_ref.json").read_text())["lines"]
    synth.write_recall_set(tmp_path)
    monkeypatch.chdir(tmp_path)
    if chunk:
        monkeypatch.setenv("Y2_VALID_BATCH", chunk)
    else:
        monkeypatch.delenv("Y2_VALID_BATCH", raising=False)
    lib = dn.lib()
    dn.set_gpu_index(0)
    lib.cuda_set_device(0)
    lib.validate_detector_recall.restype = None
    lib.validate_detector_recall.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    with _CaptureStderr(str(tmp_path / "stderr.txt")):
        lib.validate_detector_recall(b"data.cfg", b"net.cfg", b"net.weights")
    got = [l for l in (tmp_path / "stderr.txt").read_text().splitlines() if "RPs/Img" in l]
    assert got == want
    assert len(want) == 7 and any("Recall:0.00%" not in l for l in want)  # not vacuous
