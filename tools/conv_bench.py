#!/usr/bin/env python
"""Per-layer micro-benchmark of the convolution kernels (developer tool, GPU box only).

Times y2_conv_plan_launch for the convolution shapes of a cfg (default: yolo-voc 416, batch 64)
with CUDA events on the launching stream and prints TFLOP/s per shape.  Used to compare kernel
variants (Y2_CONV_VARIANT=pertap) and to pick one launch for `ncu --set full`.

    python tools/conv_bench.py                 # all yolo-voc shapes
    python tools/conv_bench.py --only 4 --reps 3   # a single shape, few launches (for ncu)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from sr_object_detection_b200 import _lib  # noqa: E402

# (name, cin, hw, cout, ksize) of yolo-voc.cfg at 416x416
YOLO_VOC = [
    ("L2", 32, 208, 64, 3), ("L4", 64, 104, 128, 3), ("L5", 128, 104, 64, 1), ("L8", 128, 52, 256, 3),
    ("L9", 256, 52, 128, 1), ("L12", 256, 26, 512, 3), ("L13", 512, 26, 256, 1), ("L18", 512, 13, 1024, 3),
    ("L19", 1024, 13, 512, 1), ("L23", 1024, 13, 1024, 3), ("L26", 512, 26, 64, 1), ("L29", 1280, 13, 1024, 3),
    ("L30", 1024, 13, 125, 1),
]


# wide-image shapes: yolo.cfg at 608 (152-position rows) and yolo9000 at 544 (136), batch 32
WIDE = [("W4_608", 64, 152, 128, 3), ("W4_544", 64, 136, 128, 3), ("W8_608", 128, 76, 256, 3), ("W2_608", 32, 304, 64, 3)]


def storage_channels(c):
    return (c + 31) // 32 * 32 if c < 64 else (c + 63) // 64 * 64


def pick_block_n(cout):
    return 32 if cout <= 32 else 64 if cout <= 64 else 128 if cout <= 128 else 256


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default=None, help="comma separated shape names, e.g. L4,L18")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    only = set(args.only.split(",")) if args.only else None
    rows = []
    for name, cin, hw, cout, k in YOLO_VOC + WIDE:
        if (only and name not in only) or (not only and name.startswith("W")):
            continue
        B = args.batch
        cin_pad = storage_channels(cin)
        f32 = name == "L30"
        cpad = cout if f32 else storage_channels(cout)
        bn = int(os.environ.get('Y2_BENCH_BN', 0)) or pick_block_n(cout)
        npad = (max(cout, cpad) + bn - 1) // bn * bn
        bk = 64 if cin_pad % 64 == 0 else 32
        x = (torch.rand(B, hw + 1, hw + 1, cin_pad, device=dev) - 0.5).to(torch.bfloat16)
        x[:, hw, :, :] = 0
        x[:, :, hw, :] = 0
        wt = ((torch.rand(npad, k * k * cin_pad, device=dev) - 0.5) * (2.0 / (k * k * cin)) ** 0.5).to(torch.bfloat16)
        alpha = torch.ones(npad, device=dev)
        beta = torch.zeros(npad, device=dev)
        if f32:
            out = torch.zeros(B, hw * hw, cpad, dtype=torch.float32, device=dev)
        else:
            out = torch.zeros(B, hw + 1, hw + 1, cpad, dtype=torch.bfloat16, device=dev)
        d = _lib.ConvDesc()
        d.in_ = x.data_ptr(); d.in_cs = cin_pad; d.cin = cin_pad
        d.batch = B; d.h = hw; d.w = hw; d.ksize = k
        d.wt = wt.data_ptr(); d.cout = cout if f32 else cpad; d.npad = npad; d.block_n = bn; d.block_k = bk
        d.alpha = alpha.data_ptr(); d.beta = beta.data_ptr(); d.act = 1
        d.out = out.data_ptr(); d.out_cs = cpad; d.out_mode = 1 if f32 else 0
        plan = C.c_void_p()
        _lib.check(lib.y2_conv_plan_create(C.byref(d), C.byref(plan)), "plan")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(args.warmup):
            _lib.check(lib.y2_conv_plan_launch(plan, st), "launch")
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            _lib.check(lib.y2_conv_plan_launch(plan, st), "launch")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        flops = 2.0 * cout * k * k * cin * hw * hw * B
        row = {"layer": name, "cin": cin, "hw": hw, "cout": cout, "k": k, "ms": round(ms, 4),
               "tflops": round(flops / ms / 1e9, 1), "variant": lib.y2_conv_plan_variant(plan)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        lib.y2_conv_plan_destroy(plan)
        del x, wt, out
    if args.json:
        Path(args.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
