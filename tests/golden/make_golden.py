#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/darknet_ref = /root/reference/src_yolo2 compiled by oracle/Makefile) on seeded
synthetic inputs.  The reference ships no golden vectors, weights or tests of its own
(SURVEY.md section 4), so these reference outputs are what pins the oracle (and, through it,
the CUDA path).  Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each .npz is self-contained: cfg text, weights bytes, inputs and every reference output.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from sr_object_detection_b200 import synth  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "darknet_ref"
OUT = Path(__file__).resolve().parent
REF_CFG = Path("/root/reference/cfg")


def run(args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([str(REF), *map(str, args)], cwd=cwd, env=e, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"darknet_ref {args[0]} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def f32(path):
    return np.fromfile(path, np.float32)


def forward_case(name, cfg_text, batch, side, thresh, nms, aux_files=None):
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        (t / "net.cfg").write_text(cfg_text)
        for fn, text in (aux_files or {}).items():
            (t / fn).write_text(text)
        synth.write_weights(t / "net.weights", cfg_text, seed=1234)
        x = synth.images(batch, 3, side, side, seed=42)
        x.tofile(t / "in.f32")
        (t / "out").mkdir()
        run(["forward", "net.cfg", "net.weights", "in.f32", "out", thresh, nms, 1], cwd=t)
        table = run(["layers", "net.cfg"], cwd=t)
        d = {"cfg": np.array(cfg_text), "weights": np.frombuffer((t / "net.weights").read_bytes(), np.uint8),
             "input": x, "thresh": np.float32(thresh), "nms": np.float32(nms), "layers": np.array(json.dumps(table))}
        for fn, text in (aux_files or {}).items():
            d["aux_" + fn.replace(".", "_")] = np.array(text)
        for p in sorted((t / "out").iterdir()):
            d[p.stem] = f32(p)
        np.savez_compressed(OUT / f"{name}.npz", **d)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k.startswith(("output", "probs"))})


def region_case(name, cfg_text, x, thresh, nms, aux_files=None, use_map=False):
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        (t / "net.cfg").write_text(cfg_text)
        for fn, text in (aux_files or {}).items():
            (t / fn).write_text(text)
        x.tofile(t / "in.f32")
        (t / "out").mkdir()
        run(["region", "net.cfg", "in.f32", "out", thresh, nms], cwd=t, env={"Y2_USE_MAP": "1" if use_map else "0"})
        d = {"cfg": np.array(cfg_text), "region_in": x, "thresh": np.float32(thresh), "nms": np.float32(nms),
             "use_map": np.int32(use_map)}
        for fn, text in (aux_files or {}).items():
            d["aux_" + fn.replace(".", "_")] = np.array(text)
        for p in sorted((t / "out").iterdir()):
            d[p.stem] = f32(p)
        np.savez_compressed(OUT / f"{name}.npz", **d)
        pre, post = d["probs_pre"], d["probs_post"]
        print(name, "nonzero probs pre/post NMS:", int((pre != 0).sum()), int((post != 0).sum()))


def region_only_cfg(batch, side, n, classes, anchors, extra=""):
    """[net] whose input already is the region layer's input (a region layer alone)."""
    return (f"[net]\nbatch={batch}\nsubdivisions=1\nheight={side}\nwidth={side}\nchannels={n * (classes + 5)}\n\n"
            + synth._region(anchors, classes, n, extra))


def tree_text(n, fanout, roots):
    with tempfile.NamedTemporaryFile("r", suffix=".tree") as f:
        synth.write_tree(f.name, n=n, fanout=fanout, roots=roots)
        return Path(f.name).read_text()


def tree_region_inputs(batch, n, classes, side, fanout, roots, seed, shared_path=False):
    """Region inputs for a softmax tree: on hot cells boost one root-to-leaf path so that the
    hierarchical product exceeds .5 somewhere below the root."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((batch, n, 5 + classes, side, side))).astype(np.float32)
    x[:, :, 4] -= 2.0
    x[:, :, 2:4] *= 0.25
    hot = rng.random((batch, side, side)) < 0.2
    bi, hi, wi = np.nonzero(hot)
    # shared_path: hot cells pick among four leaves only, so overlapping boxes of neighbouring cells meet in the NMS
    cell_node = rng.choice(rng.integers(roots + fanout * roots, classes, 4), len(bi)) if shared_path else None
    for a in range(n):
        x[bi, a, 4, hi, wi] += 6.0
        node = cell_node.copy() if shared_path else rng.integers(roots + fanout * roots, classes, len(bi))
        for _ in range(8):
            x[bi, a, 5 + node, hi, wi] += 9.0
            parent = np.where(node < roots, node, (node - roots) // fanout)
            node = parent
    return x.reshape(batch, n * (5 + classes), side, side)


def parser_tables():
    """Layer tables of the reference parser on its OWN cfg files, next to the tables of the
    synthetic cfg text used everywhere else: they must agree (checked here, at generation
    time, because /root/reference does not exist where the tests run)."""
    out = {}
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        for name, real, kw in (("tiny-yolo-voc", "tiny-yolo-voc.cfg", {}), ("yolo-voc", "yolo-voc.cfg", {}),
                               ("yolo", "yolo.cfg", {"w": 416, "h": 416}), ("darknet19_448", "darknet19_448.cfg", {}),
                               ("resnet50", "resnet50.cfg", {})):
            (t / "s.cfg").write_text(synth.CFGS[name](batch=1, **kw))
            mine = run(["layers", "s.cfg"], cwd=t)
            theirs = run(["layers", REF_CFG / real], cwd=t)
            assert mine["layers"] == theirs["layers"], f"synthetic {name} cfg differs from the reference's {real}"
            assert (mine["w"], mine["h"], mine["c"]) == (theirs["w"], theirs["h"], theirs["c"])
            out[name] = mine
        synth.write_tree(t / "9k.tree")
        (t / "y9k.cfg").write_text(synth.yolo9000_cfg(batch=1, tree="9k.tree"))
        out["yolo9000"] = run(["layers", "y9k.cfg"], cwd=t)
        (t / "y608.cfg").write_text(synth.yolo_coco_cfg(batch=1, w=608, h=608))
        out["yolo-608"] = run(["layers", "y608.cfg"], cwd=t)
    (OUT / "parser_tables.json").write_text(json.dumps(out, indent=0))
    print("parser tables:", {k: len(v["layers"]) for k, v in out.items()})


def main():
    if not REF.exists():
        raise SystemExit("oracle/_ref/darknet_ref missing: run `make -C oracle ref` first")
    # 1. whole networks, every layer's activations + decode + NMS
    forward_case("mini_yolo", synth.mini_yolo_cfg(batch=2), 2, 32, 0.05, 0.4)
    forward_case("mini_resnet", synth.mini_resnet_cfg(batch=2), 2, 32, 0.0, 0.0)
    tt = tree_text(30, 3, 3)
    forward_case("mini_yolo_tree", synth.mini_yolo_cfg(batch=1, classes=30, num=2, extra="tree=t.tree\n"), 1, 32, 0.05,
                 0.4, aux_files={"t.tree": tt})
    # 2. decode + NMS on crafted region inputs (non-trivial keep sets)
    region_case("region_voc_13", region_only_cfg(1, 13, 5, 20, synth.VOC_ANCHORS),
                synth.region_inputs(1, 5, 20, 13, 13, seed=11, hot_fraction=0.08), 0.24, 0.4)
    region_case("region_voc_7_lowthresh", region_only_cfg(2, 7, 5, 20, synth.VOC_ANCHORS),
                synth.region_inputs(2, 5, 20, 7, 7, seed=12, hot_fraction=0.1), 0.005, 0.45)
    region_case("region_coco_9", region_only_cfg(1, 9, 5, 80, synth.COCO_ANCHORS),
                synth.region_inputs(1, 5, 80, 9, 9, seed=13, hot_fraction=0.1), 0.24, 0.4)
    tt = tree_text(220, 5, 4)
    xt = tree_region_inputs(2, 3, 220, 5, 5, 4, seed=14)
    region_case("region_tree_220", region_only_cfg(2, 5, 3, 220, synth.Y9K_ANCHORS, "tree=t.tree\n"), xt, 0.24, 0.4,
                aux_files={"t.tree": tt})
    # groups wider than a warp (fanout 40; the root group has 3 nodes, every other group 40 or the tail)
    tw = tree_text(330, 40, 3)
    xw = tree_region_inputs(2, 3, 330, 4, 40, 3, seed=15, shared_path=True)
    # nms .05: the concentric boxes of a cell's three anchors (IoU .07 ... .3) suppress one another
    region_case("region_tree_wide", region_only_cfg(2, 4, 3, 330, synth.Y9K_ANCHORS, "tree=t.tree\n"), xw, 0.24, 0.05,
                aux_files={"t.tree": tw})
    with tempfile.NamedTemporaryFile("r", suffix=".map") as f:
        synth.write_map(f.name, 220)
        mt = Path(f.name).read_text()
    region_case("region_tree_220_map", region_only_cfg(2, 5, 3, 220, synth.Y9K_ANCHORS, "tree=t.tree\nmap=t.map\n"), xt,
                0.05, 0.4, aux_files={"t.tree": tt, "t.map": mt}, use_map=True)
    # 3. resize_image
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        im = np.random.default_rng(3).random((3, 20, 30), dtype=np.float32)
        im.tofile(t / "im.f32")
        subprocess.run([str(REF), "resize", "im.f32", "3", "20", "30", "13", "17", "out.f32"], cwd=t, check=True)
        np.savez_compressed(OUT / "resize.npz", image=im, resized=f32(t / "out.f32").reshape(3, 13, 17))
    # 3b. do_nms (the unsorted variant, box.c:279-297) on the decoded boxes of a region case plus exact ties
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        d = np.load(OUT / "region_voc_7_lowthresh.npz")
        boxes = d["boxes"].reshape(2, -1, 4)[0].copy()
        probs = d["probs_pre"].reshape(2, boxes.shape[0], -1)[0].copy()
        probs[5] = probs[6]            # equal rows: the `<` of box.c:290 decides
        boxes[6] = boxes[5]
        probs[40:60, 3] = 0.5          # ties inside a class
        boxes.tofile(t / "b.f32")
        probs.tofile(t / "p.f32")
        for tag, thr in (("a", 0.4), ("b", 0.1)):
            subprocess.run([str(REF), "donms", "b.f32", "p.f32", str(boxes.shape[0]), str(probs.shape[1]), str(thr),
                            f"o{tag}.f32"], cwd=t, check=True)
        np.savez_compressed(OUT / "do_nms.npz", boxes=boxes, probs=probs, thresh_a=np.float32(0.4),
                            out_a=f32(t / "oa.f32").reshape(probs.shape), thresh_b=np.float32(0.1),
                            out_b=f32(t / "ob.f32").reshape(probs.shape))
        print("do_nms: nonzero before/after", int((probs != 0).sum()), int((f32(t / "oa.f32") != 0).sum()),
              int((f32(t / "ob.f32") != 0).sum()))
    # 3c. the reference's C++ Detector (yolo_v2_class.cpp compiled without GPU / OPENCV over the same CPU objects)
    detector_golden()
    # 3d. the reference's validate_detector result files (VOC per-class files, COCO json, ImageNet-detection)
    validation_golden()
    # 3e. the reference's predict_classifier lines (flat softmax and WordTree softmax + hierarchy_predictions)
    classifier_golden()
    # 3f. the per-frame chain of demo.c's detect_in_thread over the reference's functions
    demo_golden()
    # 4. classifier front end: letterbox_image and top_k (classifier.c:676-730)
    classifier_front()
    # 5. parser tables, incl. the reference's own cfg files
    parser_tables()


def detector_golden():
    """tests/cpp/detector_scenario.cpp compiled against the REFERENCE's yolo_v2_class.{hpp,cpp} (oracle/Makefile
    refdet) on the exactly representable detector of synth.exact_detector_cfg: detect, tracking, use_mean,
    detect(filename), nms = 0, the load_image exception.  The GPU test compiles the same caller against our
    header and library and must print the same lines."""
    det = ROOT / "oracle" / "_ref" / "detector_ref"
    if not det.exists():
        raise SystemExit("oracle/_ref/detector_ref missing: run `make -C oracle refdet` first")
    w, h, n = 32, 24, 6
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        cfg_text = synth.exact_detector_cfg(batch=1, w=w, h=h)
        (t / "net.cfg").write_text(cfg_text)
        synth.write_exact_weights(t / "net.weights", cfg_text)
        frames = synth.exact_frames(n, h, w)
        frames.tofile(t / "frames.f32")
        ppm = np.random.default_rng(21).integers(0, 2, (h, w, 3)).astype(np.uint8) * 255  # byte/255. in {0, 1}: exact
        (t / "image.ppm").write_bytes(b"P6\n%d %d\n255\n" % (w, h) + ppm.tobytes())
        params = {"thresh": 0.5, "nms": 0.4, "story": 3}
        r = subprocess.run([str(det), "net.cfg", "net.weights", "frames.f32", str(n), str(w), str(h), str(params["thresh"]),
                            str(params["nms"]), str(params["story"]), "image.ppm", "0"], cwd=t, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("detector_ref failed:\n" + r.stderr[-2000:])
        lines = [l for l in r.stdout.splitlines() if l.split()[0] in
                 ("size", "detect", "track", "mean", "file", "loaded", "nonms", "load", "script")]
        out = {"cfg": cfg_text, "w": w, "h": h, "n": n, **params, "ppm_seed": 21, "lines": lines}
        (OUT / "detector_ref.json").write_text(json.dumps(out, indent=0))
        counts = {k: [int(l.split()[2]) for l in lines if l.startswith(k + " ")] for k in ("detect", "mean", "file", "nonms")}
        print("detector_ref:", counts)


def validation_golden():
    """oracle/_ref/ref_validate = the reference's validate_detector (detector.c:244-369) on its CPU path, run on
    synth.write_validation_set for the three eval types; every file it writes goes into validate_ref.npz."""
    exe = ROOT / "oracle" / "_ref" / "ref_validate"
    if not exe.exists():
        raise SystemExit("oracle/_ref/ref_validate missing: run `make -C oracle refval` first")
    out = {}
    for kind in synth.VALIDATION_CASES:
        with tempfile.TemporaryDirectory() as t:
            t = Path(t)
            synth.write_validation_set(t, kind)
            r = subprocess.run([str(exe), "data.cfg", "net.cfg", "net.weights"], cwd=t, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(f"ref_validate {kind} failed:\n" + r.stderr[-2000:])
            files = sorted((t / "results").iterdir())
            for f in files:
                out[f"{kind}/{f.name}"] = np.frombuffer(f.read_bytes(), np.uint8)
            print("validate", kind, {f.name: len(f.read_bytes().splitlines()) for f in files})
    np.savez_compressed(OUT / "validate_ref.npz", **out)


def recall_golden():
    """oracle/_ref/ref_validate recall = the reference's validate_detector_recall (detector.c:371-450) on its CPU path
    over synth.write_recall_set; its per-image stderr lines (running proposals per image, mean best IoU, recall) are
    the golden tests/golden/recall_ref.json."""
    exe = ROOT / "oracle" / "_ref" / "ref_validate"
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        synth.write_recall_set(t)
        r = subprocess.run([str(exe), "recall", "data.cfg", "net.cfg", "net.weights"], cwd=t, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("ref_validate recall failed:\n" + r.stderr[-2000:])
        lines = [l for l in r.stderr.splitlines() if "RPs/Img" in l]
    (OUT / "recall_ref.json").write_text(json.dumps({"lines": lines}))
    print("recall", lines)


def demo_golden():
    """oracle/_ref/ref_demo: network_predict -> 3-frame mean -> get_region_boxes -> do_nms(.4) per frame (demo.c:71-107)
    on the exactly representable detector and frames of 0 / 255 bytes."""
    exe = ROOT / "oracle" / "_ref" / "ref_demo"
    if not exe.exists():
        raise SystemExit("oracle/_ref/ref_demo missing: run `make -C oracle refdemo` first")
    w, h, n, thresh = 32, 24, 5, 0.3
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        cfg_text = synth.exact_detector_cfg(batch=1, w=w, h=h)
        (t / "net.cfg").write_text(cfg_text)
        synth.write_exact_weights(t / "net.weights", cfg_text)
        frames = synth.binary_frames(n, h, w, seed=31)
        frames.tofile(t / "frames.u8")
        (t / "out").mkdir()
        r = subprocess.run([str(exe), "net.cfg", "net.weights", "frames.u8", str(n), str(thresh), "out"], cwd=t,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("ref_demo failed:\n" + r.stderr[-2000:])
        out = {"w": np.int32(w), "h": np.int32(h), "n": np.int32(n), "thresh": np.float32(thresh), "seed": np.int32(31)}
        for f in range(n):
            out[f"frame_{f}"] = f32(t / "out" / f"frame_{f:03d}.f32")
        np.savez_compressed(OUT / "demo_ref.npz", **out)
        total = w * h * 3
        print("demo_ref: nonzero probs per frame", [int((out[f"frame_{f}"][total * 4:] != 0).sum()) for f in range(n)])


def classifier_golden():
    """oracle/_ref/ref_classify = the reference's predict_classifier (classifier.c:676-730) on its CPU path, on
    synth.write_classifier_set; the lines it prints (without the timing line) go into classifier_ref.json."""
    exe = ROOT / "oracle" / "_ref" / "ref_classify"
    if not exe.exists():
        raise SystemExit("oracle/_ref/ref_classify missing: run `make -C oracle refcls` first")
    out = {}
    for kind in ("flat", "tree"):
        with tempfile.TemporaryDirectory() as t:
            t = Path(t)
            synth.write_classifier_set(t, kind)
            for image, top in (("image.ppm", 0), ("image.ppm", 5), ("other.ppm", 4)):
                r = subprocess.run([str(exe), "data.cfg", "net.cfg", "net.weights", image, str(top)], cwd=t,
                                   capture_output=True, text=True)
                if r.returncode != 0:
                    raise SystemExit(f"ref_classify {kind} failed:\n" + r.stderr[-2000:])
                out[f"{kind}/{image}/{top}"] = [l for l in r.stdout.splitlines() if "Predicted in" not in l]
    (OUT / "classifier_ref.json").write_text(json.dumps(out, indent=0))
    print("classifier_ref:", {k: len(v) for k, v in out.items()}, out["tree/image.ppm/5"][:3])


def classifier_front():
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        rng = np.random.default_rng(8)
        out = {}
        for tag, (h, w, oh, ow) in {"wide": (20, 30, 32, 32), "tall": (31, 17, 24, 40), "same": (16, 16, 16, 16)}.items():
            im = rng.random((3, h, w), dtype=np.float32)
            im.tofile(t / "im.f32")
            subprocess.run([str(REF), "letterbox"] + [str(v) for v in ("im.f32", 3, h, w, oh, ow, "out.f32")],
                           cwd=t, check=True)
            out[f"{tag}_image"] = im
            out[f"{tag}_boxed"] = f32(t / "out.f32").reshape(3, oh, ow)
        # top_k incl. exact ties and k > number of distinct values
        a = rng.random(50, dtype=np.float32)
        a[[3, 17, 29]] = 0.75
        a[[5, 6]] = a.max() + 1
        a.tofile(t / "a.f32")
        for k in (1, 5, 12):
            r = subprocess.run([str(REF), "topk", "a.f32", "50", str(k)], cwd=t, check=True, capture_output=True, text=True)
            out[f"topk_{k}"] = np.array(json.loads(r.stdout.strip().splitlines()[-1]), dtype=np.int32)
        out["topk_input"] = a
        np.savez_compressed(OUT / "classifier_front.npz", **out)


def voc_eval_golden():
    """scripts/voc_eval.py of the reference (a Python 2 file) run on a synthetic VOC-style data set: annotations with
    difficult objects, detections with duplicates, misses, ties and images without objects.  The script's text is
    read where it lies, made importable under Python 3 / numpy 2 in memory (print statements, cPickle, np.bool, pickle
    file modes - nothing else) and executed; the fixture holds the inputs and the (recall, precision, AP) it
    returned, for tools/voc_eval.py to reproduce (tests/test_voc_eval.py)."""
    import re
    src = Path("/root/reference/scripts/voc_eval.py").read_text()
    src = src.replace("import cPickle", "import pickle as cPickle")
    src = re.sub(r"print '([^']*)'\.format\(\s*([^)]*)\)", r"print('\1'.format(\2))", src, flags=re.S)
    src = src.replace("np.bool", "bool").replace("open(cachefile, 'w')", "open(cachefile, 'wb')")
    src = src.replace("open(cachefile, 'r')", "open(cachefile, 'rb')")
    ns = {}
    exec(compile(src, "reference scripts/voc_eval.py", "exec"), ns)
    rng = np.random.default_rng(2024)
    classes = ["cat", "dog", "car"]
    images = ["%06d" % i for i in range(40)]
    truth = {}
    for im in images:
        objs = []
        for _ in range(int(rng.integers(0, 5))):
            x0, y0 = int(rng.integers(1, 300)), int(rng.integers(1, 200))
            w, h = int(rng.integers(20, 180)), int(rng.integers(20, 150))
            objs.append({"name": classes[int(rng.integers(0, 3))], "difficult": int(rng.random() < 0.2),
                         "bbox": [x0, y0, x0 + w, y0 + h]})
        truth[im] = objs
    dets = {c: [] for c in classes}
    for im in images:
        for o in truth[im]:
            if rng.random() < 0.8:   # a hit, sometimes twice (duplicate), sometimes badly localised
                for _ in range(1 + int(rng.random() < 0.25)):
                    j = rng.normal(0, 12 if rng.random() < 0.7 else 60, 4)
                    b = [o["bbox"][k] + j[k] for k in range(4)]
                    dets[o["name"]].append((im, round(float(rng.random()), 2 if rng.random() < 0.3 else 6), *b))
        for _ in range(int(rng.integers(0, 3))):  # false alarms
            x0, y0 = rng.uniform(1, 300), rng.uniform(1, 200)
            dets[classes[int(rng.integers(0, 3))]].append((im, float(rng.random()) * 0.6, x0, y0, x0 + rng.uniform(10, 150),
                                                           y0 + rng.uniform(10, 150)))
    out = {"classes": classes, "images": images, "truth": truth, "detections": {}, "results": {}}
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        (t / "ann").mkdir()
        for im in images:
            xml = "<annotation>" + "".join(
                "<object><name>%s</name><difficult>%d</difficult><bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax>"
                "<ymax>%d</ymax></bndbox></object>" % (o["name"], o["difficult"], *o["bbox"]) for o in truth[im]) + "</annotation>"
            (t / "ann" / f"{im}.xml").write_text(xml)
        (t / "set.txt").write_text("\n".join(images) + "\n")
        for c in classes:
            lines = ["%s %f %f %f %f %f" % d for d in dets[c]]   # the format print_detector_detections writes
            (t / f"det_{c}.txt").write_text("\n".join(lines) + "\n")
            out["detections"][c] = lines
            for voc07 in (False, True):
                for ov in (0.5, 0.7):
                    cache = t / f"cache_{c}_{voc07}_{ov}"
                    rec, prec, ap = ns["voc_eval"](str(t / "det_{}.txt"), str(t / "ann" / "{}.xml"), str(t / "set.txt"), c,
                                                   str(cache), ov, voc07)
                    out["results"][f"{c}|{int(voc07)}|{ov}"] = {"rec": [float(v) for v in rec],
                                                               "prec": [float(v) for v in prec], "ap": float(ap)}
    (OUT / "voc_eval_ref.json").write_text(json.dumps(out))
    print("voc_eval_ref.json:", {k: round(v["ap"], 4) for k, v in out["results"].items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "voc_eval":
        voc_eval_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "recall":
        recall_golden()
    else:
        main()
