"""predict_classifier (classifier.c:676-730) through the GPU forward pass against the lines the REFERENCE's own
predict_classifier printed on its CPU path (oracle/_ref/ref_classify -> tests/golden/classifier_ref.json).  The
classifier is exactly representable (synth.exact_classifier_cfg; a network-sized image of 0 / 255 bytes), so avgpool,
softmax or the WordTree softmax layer, hierarchy_predictions, top_k and the %f lines must agree character for character;
an image of another size goes through letterbox_image's resize (values no longer exact in bf16) and is compared by
ranking and value."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from sr_object_detection_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
GOLD = json.loads((ROOT / "tests" / "golden" / "classifier_ref.json").read_text())

RUN = """
import sys
sys.path.insert(0, {root!r})
import ctypes as C
from sr_object_detection_b200 import darknet as dn
lib = dn.lib()
dn.set_gpu_index(0)
lib.cuda_set_device(0)
lib.predict_classifier.restype = None
lib.predict_classifier.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
lib.predict_classifier(b"data.cfg", b"net.cfg", b"net.weights", {image!r}.encode(), {top})
"""


def _run(tmp_path, image, top):
    r = subprocess.run([sys.executable, "-c", RUN.format(root=str(ROOT), image=image, top=top)], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if "Predicted in" not in l]


@pytest.mark.parametrize("kind", ["flat", "tree"])
def test_predict_classifier_lines_equal_the_reference(tmp_path, kind):
    synth.write_classifier_set(tmp_path, kind)
    for top in (0, 5):
        assert _run(tmp_path, "image.ppm", top) == GOLD[f"{kind}/image.ppm/{top}"]
    # another image size: letterboxed (resize + grey bars); same ranking, probabilities within 2 % (bf16 input rounding)
    got, want = _run(tmp_path, "other.ppm", 4), GOLD[f"{kind}/other.ppm/4"]
    assert got[0] == want[0] and len(got) == len(want) == 5

    def parse(line):
        name = line.split(":")[0]
        return name, float(line.split(":")[1].split(",")[0])

    for a, b in zip(got[1:], want[1:]):
        (na, pa), (nb, pb) = parse(a), parse(b)
        assert na == nb and abs(pa - pb) <= 0.02 * max(pb, 1e-3) + 1e-6, (a, b)
