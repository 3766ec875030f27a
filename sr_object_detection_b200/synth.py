"""Synthetic inputs for the benchmark and the parity tests: Darknet `.cfg` text for the
north-star networks, seeded `.weights` files, seeded images and a WordTree.

The network definitions are generated from compact specs (they describe the public
YOLOv2 / Darknet-19 / ResNet-50 architectures; layer graphs are cross-checked against the
reference's parser output in tests/test_oracle.py).  Nothing here touches a GPU.

Weight statistics follow SURVEY.md section 8d: weights ~ U(-s, s), s = sqrt(2/(k*k*c))
(the reference's own init scale, convolutional_layer.c:207-208), biases/means ~ U(-.1,.1),
scales/variances ~ U(.5, 1.5) (the reference leaves rolling_variance = 0, which blows up
without a weights file).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

VOC_ANCHORS = "1.3221,1.73145, 3.19275,4.00944, 5.05587,8.09892, 9.47112,4.84053, 11.2364,10.0071"
COCO_ANCHORS = "0.57273,0.677385, 1.87446,2.06253, 3.33843,5.47434, 7.88282,3.52778, 9.77052,9.16828"
TINY_VOC_ANCHORS = "1.08,1.19, 3.42,4.41, 6.63,11.38, 9.42,5.11, 16.62,10.52"
Y9K_ANCHORS = "0.77871,1.14074, 3.00525,4.31277, 9.22725,9.61974"


def _net(batch, w, h, c=3):
    return f"[net]\nbatch={batch}\nsubdivisions=1\nheight={h}\nwidth={w}\nchannels={c}\n\n"


def _conv(filters, size, stride=1, bn=1, act="leaky", pad=1):
    s = "[convolutional]\n"
    if bn:
        s += "batch_normalize=1\n"
    return s + f"filters={filters}\nsize={size}\nstride={stride}\npad={pad}\nactivation={act}\n\n"


def _maxpool(size=2, stride=2):
    return f"[maxpool]\nsize={size}\nstride={stride}\n\n"


def _region(anchors, classes, num, extra=""):
    return (f"[region]\nanchors = {anchors}\nbias_match=1\nclasses={classes}\ncoords=4\nnum={num}\n"
            f"softmax=1\njitter=.2\nrescore=1\nobject_scale=5\nnoobject_scale=1\nclass_scale=1\n"
            f"coord_scale=1\nabsolute=1\nthresh = .6\nrandom=0\n{extra}\n")


def _darknet19_trunk():
    """Layers 0-22 of yolo-voc.cfg / yolo9000.cfg / darknet19_448.cfg."""
    s = _conv(32, 3) + _maxpool() + _conv(64, 3) + _maxpool()
    s += _conv(128, 3) + _conv(64, 1) + _conv(128, 3) + _maxpool()
    s += _conv(256, 3) + _conv(128, 1) + _conv(256, 3) + _maxpool()
    s += _conv(512, 3) + _conv(256, 1) + _conv(512, 3) + _conv(256, 1) + _conv(512, 3) + _maxpool()
    s += _conv(1024, 3) + _conv(512, 1) + _conv(1024, 3) + _conv(512, 1) + _conv(1024, 3)
    return s


def tiny_yolo_voc_cfg(batch=1, w=416, h=416):
    s = _net(batch, w, h)
    for f in (16, 32, 64, 128, 256):
        s += _conv(f, 3) + _maxpool()
    s += _conv(512, 3) + _maxpool(2, 1)
    s += _conv(1024, 3) + _conv(1024, 3) + _conv(125, 1, bn=0, act="linear")
    return s + _region(TINY_VOC_ANCHORS, 20, 5)


def _yolo2(batch, w, h, head, anchors, classes):
    s = _net(batch, w, h) + _darknet19_trunk()
    s += _conv(1024, 3) + _conv(1024, 3)
    s += "[route]\nlayers=-9\n\n" + _conv(64, 1) + "[reorg]\nstride=2\n\n[route]\nlayers=-1,-4\n\n"
    s += _conv(1024, 3) + _conv(head, 1, bn=0, act="linear")
    return s + _region(anchors, classes, 5)


def yolo_voc_cfg(batch=1, w=416, h=416):
    return _yolo2(batch, w, h, 125, VOC_ANCHORS, 20)


def yolo_coco_cfg(batch=1, w=608, h=608):
    return _yolo2(batch, w, h, 425, COCO_ANCHORS, 80)


def yolo9000_cfg(batch=1, w=544, h=544, tree="9k.tree", map_file=None, classes=9418):
    s = _net(batch, w, h) + _darknet19_trunk()
    s += _conv(3 * (classes + 5), 1, bn=0, act="linear")
    extra = f"tree={tree}\n" + (f"map = {map_file}\n" if map_file else "")
    return s + _region(Y9K_ANCHORS, classes, 3, extra)


def darknet19_448_cfg(batch=1, w=448, h=448):
    s = _net(batch, w, h) + _darknet19_trunk()
    s += _conv(1000, 1, bn=0, act="linear")
    return s + "[avgpool]\n\n[softmax]\ngroups=1\n\n[cost]\ntype=sse\n\n"


def resnet50_cfg(batch=1, w=256, h=256):
    s = _net(batch, w, h) + _conv(64, 7, stride=2) + _maxpool(2, 2)
    for a, stride, blocks in ((64, 1, 3), (128, 2, 4), (256, 2, 6), (512, 2, 3)):
        for b in range(blocks):
            st = stride if b == 0 else 1
            s += _conv(a, 1) + _conv(a, 3, stride=st) + _conv(4 * a, 1, act="linear")
            s += "[shortcut]\nfrom=-4\nactivation=leaky\n\n"
    s += _conv(1000, 1, bn=0, act="linear")
    return s + "[avgpool]\n\n[softmax]\ngroups=1\n\n[cost]\ntype=sse\n\n"


def mini_yolo_cfg(batch=2, w=32, h=32, classes=4, num=3, extra=""):
    """A 17-layer detector with every layer kind and quirk of yolo-voc.cfg / tiny-yolo-voc.cfg at
    toy size (golden-fixture network): 3x3 and 1x1 convs, 2/2 pools, the 2/1 pool with pad 0,
    a single-input route, reorg, a two-input route, a linear head and a region layer."""
    s = _net(batch, w, h)
    s += _conv(8, 3) + _maxpool() + _conv(16, 3) + _maxpool() + _conv(16, 1) + _conv(32, 3)   # 0-5
    s += _maxpool() + _conv(32, 3) + _maxpool(2, 1) + _conv(64, 3)                            # 6-9
    s += "[route]\nlayers=-5\n\n" + _conv(8, 1) + "[reorg]\nstride=2\n\n[route]\nlayers=-1,-4\n\n"  # 10-13
    s += _conv(32, 3) + _conv(num * (classes + 5), 1, bn=0, act="linear")                     # 14-15
    anchors = ",".join(f"{0.6 + 0.7 * i:.2f},{0.8 + 0.5 * i:.2f}" for i in range(num))
    return s + _region(anchors, classes, num, extra)


def mini_resnet_cfg(batch=2, w=32, h=32):
    """Toy version of resnet50.cfg (golden-fixture network): 7x7/2 stem, pool, two bottlenecks
    whose shortcuts exercise the channel-mismatch and the stride-2 sampling branches of
    shortcut_cpu (blas.c:57-81), then the classifier tail conv -> avgpool -> softmax -> cost."""
    s = _net(batch, w, h) + _conv(16, 7, stride=2) + _maxpool(2, 2)
    for a, st in ((8, 1), (16, 2)):
        s += _conv(a, 1) + _conv(a, 3, stride=st) + _conv(4 * a, 1, act="linear")
        s += "[shortcut]\nfrom=-4\nactivation=leaky\n\n"
    s += _conv(10, 1, bn=0, act="linear")
    return s + "[avgpool]\n\n[softmax]\ngroups=1\n\n[cost]\ntype=sse\n\n"


def mini_dense_cfg(batch=2, w=64, h=64, classes=4, num=3):
    """Toy DenseNet-style detector: every block's 3x3 output (growth 32) is concatenated with the running
    feature map by a two-input route whose second input is itself a route (densenet201.cfg's pattern), so
    the in-place concat has to fall back to a copy for the nested input."""
    s = _net(batch, w, h) + _conv(32, 3) + _maxpool()                                   # 0-1
    s += _conv(64, 1) + _conv(32, 3) + "[route]\nlayers=-1,-3\n\n"                      # 2-4: 32 + 32
    s += _conv(64, 1) + _conv(32, 3) + "[route]\nlayers=-1,-3\n\n"                      # 5-7: 32 + 64
    s += _conv(64, 1) + _conv(32, 3) + "[route]\nlayers=-1,-3\n\n"                      # 8-10: 32 + 96
    s += _maxpool() + _conv(64, 3) + _conv(num * (classes + 5), 1, bn=0, act="linear")  # 11-13
    anchors = ",".join(f"{0.6 + 0.7 * i:.2f},{0.8 + 0.5 * i:.2f}" for i in range(num))
    return s + _region(anchors, classes, num)


def mini_dense_odd_cfg(batch=2, w=64, h=64, classes=4, num=3):
    """The same pattern with channel counts that are not whole storage groups (48, 80, 24 filters: stored as 64,
    128, 32 channels): the concat has no holes in the reference (route_layer.c:73-86), so the real channels have
    to be packed next to each other by copies; one concat is itself an input of the next."""
    s = _net(batch, w, h) + _conv(48, 3) + _maxpool()                                   # 0-1: 48
    s += _conv(80, 1) + _conv(24, 3) + "[route]\nlayers=-1,-3\n\n"                      # 2-4: 24 + 48 = 72
    s += _conv(64, 1) + _conv(40, 3) + "[route]\nlayers=-1,-3\n\n"                      # 5-7: 40 + 72 = 112
    s += _maxpool() + _conv(64, 3) + _conv(num * (classes + 5), 1, bn=0, act="linear")  # 8-10
    anchors = ",".join(f"{0.6 + 0.7 * i:.2f},{0.8 + 0.5 * i:.2f}" for i in range(num))
    return s + _region(anchors, classes, num)


EXACT_ANCHORS = "1.0,1.5, 2.5,2.0, 4.0,5.0"


def exact_detector_cfg(batch=1, w=32, h=24, classes=4, num=3, extra=""):
    """A detector whose head output is EXACTLY representable: one 3x3 linear convolution (no batchnorm) straight
    into the region layer.  With write_exact_weights and exact_frames every product k/64 * m/32 and every partial
    sum is a dyadic rational of at most 17 bits, so fp32 accumulation on the CPU (any order) and bf16 x bf16 ->
    fp32 on the tensor cores give the same bits: everything behind the head can be compared with the
    reference bit for bit through whole-pipeline entry points (Detector::detect)."""
    return (_net(batch, w, h) + _conv(num * (classes + 5), 3, bn=0, act="linear")
            + _region(EXACT_ANCHORS, classes, num, extra))


def binary_ppm(path, w, h, seed):
    """A P6 file whose bytes are 0 or 255: byte / 255. is exactly 0 or 1, so the loaders' float image is exact in
    bf16 as well (used with exact_detector_cfg for byte-identical validation files)."""
    px = np.random.default_rng(seed).integers(0, 2, (h, w, 3)).astype(np.uint8) * 255
    Path(path).write_bytes(b"P6\n%d %d\n255\n" % (w, h) + px.tobytes())
    return px


def binary_frames(n, h, w, seed):
    """n uint8 RGB frames [n][h][w][3] of 0 / 255 bytes; consecutive frames share most pixels (a block moves)"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 2, (h, w, 3)).astype(np.uint8) * 255
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        fr = base.copy()
        x0 = (5 * i) % max(w - 8, 1)
        fr[4:12, x0:x0 + 8] = rng.integers(0, 2, (8, 8, 3)).astype(np.uint8) * 255
        out[i] = fr
    return out


VALIDATION_CASES = {
    # eval type -> (classes, region extras, image file names)
    "voc": (4, "", ["2008_000001", "2008_000002", "img_c", "img_d", "img_e"]),
    # (the reference starts four loader threads on paths[0..3] unconditionally: a list needs >= 4 images)
    "coco": (80, "", ["COCO_val2014_000000000042", "COCO_val2014_000000000073", "COCO_val2014_000000000139",
                      "COCO_val2014_000000000785", "COCO_val2014_000000000872"]),
    "imagenet": (220, "tree=t.tree\nmap=t.map\n", [f"ILSVRC2013_val_0000000{i}" for i in range(1, 6)]),
}


def write_validation_set(root, kind, w=16, h=12):
    """Everything validate_detector reads for one eval type, under `root` (paths inside the files are relative
    to `root`: run with cwd = root): net.cfg, net.weights, images/*.ppm, valid.list, names.list, data.cfg,
    results/ (+ t.tree / t.map for the ImageNet-detection case)."""
    root = Path(root)
    classes, extra, stems = VALIDATION_CASES[kind]
    (root / "images").mkdir(parents=True, exist_ok=True)
    (root / "results").mkdir(exist_ok=True)
    cfg_text = exact_detector_cfg(batch=1, w=w, h=h, classes=classes, num=3, extra=extra)
    (root / "net.cfg").write_text(cfg_text)
    write_exact_weights(root / "net.weights", cfg_text, seed=17, classes=classes, num=3)
    for i, stem in enumerate(stems):
        binary_ppm(root / "images" / f"{stem}.ppm", w, h, seed=300 + i)
    (root / "valid.list").write_text("".join(f"images/{s}.ppm\n" for s in stems))
    (root / "names.list").write_text("".join(f"class{j}\n" for j in range(classes)))
    data = f"classes={classes}\nvalid=valid.list\nnames=names.list\nresults=results\neval={kind}\n"
    if kind == "imagenet":
        write_tree(root / "t.tree", n=classes, fanout=5, roots=4)
        write_map(root / "t.map", classes)
        data += "map=t.map\n"
    (root / "data.cfg").write_text(data)
    return stems


def write_recall_set(root, n_images=7, w=16, h=12, classes=4):
    """Everything validate_detector_recall (detector.c:371-450) reads, under `root` (run with cwd = root): the exactly
    representable detector, images/*.png (binary PPM bytes under a .png name: the label path is derived by replacing
    "images" -> "labels" and ".png" -> ".txt"; both loaders decode by content), labels/*.txt with 1-3 truth boxes per
    image ("id x y w h", relative units), valid.list and data.cfg."""
    root = Path(root)
    (root / "images").mkdir(parents=True, exist_ok=True)
    (root / "labels").mkdir(exist_ok=True)
    cfg_text = exact_detector_cfg(batch=1, w=w, h=h, classes=classes, num=3)
    (root / "net.cfg").write_text(cfg_text)
    write_exact_weights(root / "net.weights", cfg_text, seed=23, classes=classes, num=3)
    rng = np.random.default_rng(77)
    stems = [f"frame_{i:03d}" for i in range(n_images)]
    for i, stem in enumerate(stems):
        binary_ppm(root / "images" / f"{stem}.png", w, h, seed=500 + i)
        lines = []
        for _ in range(int(rng.integers(1, 4))):
            bw, bh = rng.uniform(0.2, 0.9), rng.uniform(0.2, 0.9)
            lines.append("%d %.6f %.6f %.6f %.6f" % (int(rng.integers(0, classes)), rng.uniform(0.2, 0.8),
                                                      rng.uniform(0.2, 0.8), bw, bh))
        (root / "labels" / f"{stem}.txt").write_text("\n".join(lines) + "\n")
    (root / "valid.list").write_text("".join(f"images/{s}.png\n" for s in stems))
    (root / "data.cfg").write_text(f"classes={classes}\nvalid=valid.list\n")
    return stems


def write_exact_weights(path, cfg_text, seed=5, classes=4, num=3):
    """weights m/32 (m/256 for the tw/th filters, so that boxes stay a few cells wide), biases m/64, |m| <= 32"""
    (sp,) = conv_specs_from_cfg(cfg_text)
    rng = np.random.default_rng(seed)
    n, k, c = sp.filters, sp.size, sp.channels
    biases = (rng.integers(-32, 33, n) / 64.0).astype(np.float32)
    w = (rng.integers(-32, 33, (n, c * k * k)) / 32.0).astype(np.float32)
    for a in range(num):
        for j in (2, 3):
            w[a * (classes + 5) + j] /= 8.0
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 0, 1, 0, 0))
        f.write(biases.tobytes())
        f.write(w.tobytes())


def exact_frames(n, h, w, seed=9):
    """n frames [3][h][w] with values k/64 (exact in bf16); consecutive frames differ in a moving block only,
    so detections persist from frame to frame (tracking has something to follow)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 65, (3, h, w))
    out = np.empty((n, 3, h, w), np.float32)
    for i in range(n):
        fr = base.copy()
        x0 = (3 * i) % max(w - 8, 1)
        fr[:, 4:12, x0:x0 + 8] = rng.integers(0, 65, (3, 8, 8))
        out[i] = fr / 64.0
    return out


def exact_classifier_cfg(batch=1, w=16, h=12, classes=10, extra=""):
    """Classifier counterpart of exact_detector_cfg: one 3x3 linear convolution -> avgpool -> softmax (-> cost), head
    output exactly representable, so the whole predict_classifier flow can be compared with the reference bit for bit.
    extra: e.g. "tree=t.tree\n" for a WordTree softmax (net.hierarchy)."""
    return (_net(batch, w, h) + _conv(classes, 3, bn=0, act="linear") + "[avgpool]\n\n[softmax]\ngroups=1\n" + extra
            + "\n[cost]\ntype=sse\n\n")


def write_classifier_set(root, kind, w=16, h=12):
    """Everything predict_classifier reads, under `root`: net.cfg, net.weights, image.ppm (network-sized, bytes 0/255),
    other.ppm (another size: goes through the letterbox resize), names.list, data.cfg (+ t.tree for kind "tree")."""
    root = Path(root)
    classes = 30 if kind == "tree" else 10
    extra = ""
    if kind == "tree":
        write_tree(root / "t.tree", n=classes, fanout=3, roots=3)
        extra = "tree=t.tree\n"
    cfg_text = exact_classifier_cfg(batch=1, w=w, h=h, classes=classes, extra=extra)
    (root / "net.cfg").write_text(cfg_text)
    (sp,) = conv_specs_from_cfg(cfg_text)
    rng = np.random.default_rng(23)
    biases = (rng.integers(-32, 33, sp.filters) / 64.0).astype(np.float32)
    wts = (rng.integers(-32, 33, (sp.filters, sp.channels * 9)) / 256.0).astype(np.float32)
    with open(root / "net.weights", "wb") as f:
        f.write(struct.pack("<iiii", 0, 1, 0, 0))
        f.write(biases.tobytes())
        f.write(wts.tobytes())
    binary_ppm(root / "image.ppm", w, h, seed=400)
    binary_ppm(root / "other.ppm", w + 9, h + 4, seed=401)
    (root / "names.list").write_text("".join(f"label{j}\n" for j in range(classes)))
    (root / "data.cfg").write_text(f"classes={classes}\nnames=names.list\ntop=3\n")


def region_only_cfg(batch, cells_w, cells_h, anchors, classes, num, extra=""):
    """A network that is ONLY the region layer over a [num*(5+classes)][cells_h][cells_w] input: lets the CPU
    checkers decode a given head output without allocating the whole detector (parse_region asserts
    l.outputs == inputs, parser.c:243, which holds for exactly this input shape)."""
    return _net(batch, cells_w, cells_h, c=num * (classes + 5)) + _region(anchors, classes, num, extra)


REGION_PARAMS = {  # anchors, classes, num of the detector cfgs above
    "tiny-yolo-voc": (TINY_VOC_ANCHORS, 20, 5),
    "yolo-voc": (VOC_ANCHORS, 20, 5),
    "yolo": (COCO_ANCHORS, 80, 5),
    "yolo9000": (Y9K_ANCHORS, 9418, 3),
}


def mini_reorg_reverse_cfg(batch=2, w=32, h=32, classes=4, num=3):
    """Toy detector with a reverse reorg layer (depth-to-space, reorg_layer.c:80-81): 8x8x64 -> 16x16x16."""
    s = _net(batch, w, h) + _conv(16, 3) + _maxpool() + _conv(32, 3) + _maxpool() + _conv(64, 1)
    s += "[reorg]\nstride=2\nreverse=1\n\n" + _conv(32, 3) + _conv(num * (classes + 5), 1, bn=0, act="linear")
    anchors = ",".join(f"{0.6 + 0.7 * i:.2f},{0.8 + 0.5 * i:.2f}" for i in range(num))
    return s + _region(anchors, classes, num)


def mini_alexnet_cfg(batch=2, w=32, h=32, classes=10):
    """Toy classifier in the style of cfg/alexnet.cfg: relu convolutions (an activation the tensor-core epilogue does
    not implement), maxpools, connected layers (one with batchnorm and a non-epilogue activation), dropout, softmax."""
    s = _net(batch, w, h) + _conv(16, 3, act="relu") + _maxpool() + _conv(32, 3, bn=0, act="relu") + _maxpool()
    s += _conv(24, 3, act="elu")
    s += "[connected]\noutput=96\nactivation=relu\n\n[dropout]\nprobability=.5\n\n"
    s += "[connected]\nbatch_normalize=1\noutput=64\nactivation=tanh\n\n[dropout]\nprobability=.5\n\n"
    s += f"[connected]\noutput={classes}\nactivation=linear\n\n"
    return s + "[softmax]\ngroups=1\n\n[cost]\ntype=sse\n\n"


CFGS = {
    "mini-alexnet": mini_alexnet_cfg,
    "mini-reorg-reverse": mini_reorg_reverse_cfg,
    "mini-dense": mini_dense_cfg, "mini-dense-odd": mini_dense_odd_cfg,
    "mini-yolo": mini_yolo_cfg,
    "mini-resnet": mini_resnet_cfg,
    "tiny-yolo-voc": tiny_yolo_voc_cfg,
    "yolo-voc": yolo_voc_cfg,
    "yolo": yolo_coco_cfg,
    "yolo9000": yolo9000_cfg,
    "darknet19_448": darknet19_448_cfg,
    "resnet50": resnet50_cfg,
}


@dataclass
class ConvSpec:
    filters: int
    size: int
    channels: int
    batch_normalize: bool
    kind: str = "conv"   # "conv": weights[n][c][k][k]; "fc": a connected layer, weights[filters][channels]


def conv_specs_from_cfg(cfg_text: str) -> list[ConvSpec]:
    """Walk a cfg the way parse_network_cfg does (parser.c:585-700), tracking the extent, to know what a
    .weights file for it must contain: one entry per convolutional / connected layer, in order."""
    sections: list[tuple[str, dict]] = []
    for raw in cfg_text.splitlines():
        line = "".join(raw.split())
        if not line or line[0] in "#;":
            continue
        if line.startswith("["):
            sections.append((line, {}))
        else:
            k, _, v = line.partition("=")
            sections[-1][1][k] = v
    assert sections[0][0] == "[net]"
    net = sections[0][1]
    c, h, w = int(net.get("channels", 3)), int(net.get("height", 0)), int(net.get("width", 0))
    out_c: list[int] = []
    specs: list[ConvSpec] = []
    for idx, (name, opt) in enumerate(sections[1:]):
        if name == "[convolutional]":
            n, k, st = int(opt.get("filters", 1)), int(opt.get("size", 1)), int(opt.get("stride", 1))
            pad = k // 2 if int(opt.get("pad", 0)) else int(opt.get("padding", 0))
            specs.append(ConvSpec(n, k, c, bool(int(opt.get("batch_normalize", 0)))))
            c, h, w = n, (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
        elif name == "[connected]":
            n = int(opt.get("output", 1))
            specs.append(ConvSpec(n, 1, c * max(h, 1) * max(w, 1), bool(int(opt.get("batch_normalize", 0))), "fc"))
            c, h, w = n, 1, 1
        elif name == "[maxpool]":
            st = int(opt.get("stride", 1))
            size = int(opt.get("size", st))
            pad = int(opt.get("padding", (size - 1) // 2))
            h, w = (h + 2 * pad) // st, (w + 2 * pad) // st
        elif name == "[avgpool]":
            h, w = 1, 1
        elif name == "[route]":
            layers = [int(v) for v in opt["layers"].split(",")]
            c = sum(out_c[(idx + l) if l < 0 else l] for l in layers)
        elif name == "[reorg]":
            st = int(opt.get("stride", 1))
            if int(opt.get("reverse", 0)):
                c, h, w = c // (st * st), h * st, w * st
            else:
                c, h, w = c * st * st, h // st, w // st
        # shortcut / dropout / softmax / cost / region keep the extent (spatial sizes behind a route are not tracked:
        # no cfg here puts a connected layer there)
        out_c.append(c)
    return specs


def write_weights(path: str | Path, cfg_text: str, seed: int = 1234, head_gain: float = 1.0) -> int:
    """Seeded synthetic `.weights` in the v0.1 layout read by load_weights_upto
    (parser.c:1009-1082): int32 major, minor, revision, seen; per conv: biases, [scales,
    rolling_mean, rolling_variance], weights[n][c][k][k].

    head_gain multiplies the weights of the LAST convolution (the detection head).  With the plain
    init every box of a random network scores objectness ~0.5 x class ~1/classes, far below the
    detector's 0.24 threshold, so decode and NMS would run on an empty candidate set; a gain of a few
    units spreads the head's logits so that a few per cent of the boxes clear the threshold and NMS
    has overlapping candidates to work on."""
    specs = conv_specs_from_cfg(cfg_text)
    total = 0
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 0, 1, 0, 0))
        for li, sp in enumerate(specs):
            rng = np.random.default_rng(seed + li)
            n, k, c = sp.filters, sp.size, sp.channels
            biases = rng.uniform(-0.1, 0.1, n).astype(np.float32)
            f.write(biases.tobytes())
            if sp.batch_normalize:
                scales = rng.uniform(0.5, 1.5, n).astype(np.float32)
                mean = rng.uniform(-0.1, 0.1, n).astype(np.float32)
                var = rng.uniform(0.5, 1.5, n).astype(np.float32)
                f.write(scales.tobytes() + mean.tobytes() + var.tobytes())
                total += 3 * n
            s = np.sqrt(2.0 / (k * k * c))
            w = rng.uniform(-s, s, n * c * k * k).astype(np.float32)
            if head_gain != 1.0 and li == len(specs) - 1:
                w = (w * np.float32(head_gain)).astype(np.float32)
            if sp.kind == "fc":
                # load_connected_weights (parser.c:897-913): biases, weights, THEN the batchnorm vectors
                f.seek(-(3 * n * 4 if sp.batch_normalize else 0), 1)
                f.write(w.tobytes())
                if sp.batch_normalize:
                    f.write(scales.tobytes() + mean.tobytes() + var.tobytes())
            else:
                f.write(w.tobytes())
            total += n + w.size
    return total


def images(batch: int, c: int, h: int, w: int, seed: int = 42) -> np.ndarray:
    """U(0,1) fp32 planar CHW images, one generator per image (seed + index)."""
    out = np.empty((batch, c, h, w), np.float32)
    for b in range(batch):
        out[b] = np.random.default_rng(seed + b).random((c, h, w), dtype=np.float32)
    return out


def write_tree(path: str | Path, n: int = 9418, fanout: int = 5, roots: int = 4) -> None:
    """A well-formed WordTree in the `name parent` format of read_tree (tree.c:53-103):
    `roots` top-level nodes, then parent = (i - roots) // fanout, so the children of a node
    are contiguous and every parent precedes its children.  (cfg/9k.tree in the reference is
    corrupt - SURVEY.md section 8c.)"""
    with open(path, "w") as f:
        for i in range(n):
            parent = -1 if i < roots else (i - roots) // fanout
            f.write(f"n{i:08d} {parent}\n")


def write_map(path: str | Path, classes: int, n: int = 200, seed: int = 5) -> None:
    """A `map=` file as read by read_map (utils.c:17-33): n class indices, one per line."""
    rng = np.random.default_rng(seed)
    idx = rng.permutation(classes)[:n] if classes >= n else rng.integers(0, classes, n)
    Path(path).write_text("".join(f"{int(i)}\n" for i in idx))


def region_inputs(batch: int, n: int, classes: int, h: int, w: int, seed: int = 7,
                  hot_fraction: float = 0.05) -> np.ndarray:
    """Seeded region-layer inputs [B][n*(5+classes)][h][w] (the conv output layout).  About
    `hot_fraction` of the cells are "objects": all anchors of such a cell (and of its right
    neighbour) get a boosted objectness and the same boosted class, so that several strongly
    overlapping boxes compete in NMS and the keep set is non-trivial."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((batch, n, 5 + classes, h, w)) * 2).astype(np.float32)
    x[:, :, 4] -= 3.0
    x[:, :, 2:4] *= 0.25  # keep exp(tw), exp(th) moderate so neighbouring boxes overlap
    hot = rng.random((batch, h, w)) < hot_fraction
    cls = rng.integers(0, classes, (batch, h, w))
    bi, hi, wi = np.nonzero(hot)
    for dx in (0, 1):
        wj = np.minimum(wi + dx, w - 1)
        for a in range(n):
            x[bi, a, 4, hi, wj] += 6.0 + rng.random(len(bi)).astype(np.float32)
            x[bi, a, 5 + cls[bi, hi, wi], hi, wj] += 8.0
    return x.reshape(batch, n * (5 + classes), h, w)
