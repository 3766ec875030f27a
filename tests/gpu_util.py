"""Helpers for the GPU parity tests: torch owns device memory, the kernels are called
through the C-ABI (ctypes) exactly as the C host runtime calls them."""
import ctypes as C

import torch

from sr_object_detection_b200 import _lib


def ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def to_padded_nhwc(x: torch.Tensor, cs: int | None = None) -> torch.Tensor:
    """fp32 NCHW -> bf16 [B][H+1][W+1][cs] with zero pad row/col (torch-side reference packing)."""
    b, c, h, w = x.shape
    cs = cs or c
    out = torch.zeros(b, h + 1, w + 1, cs, dtype=torch.bfloat16, device=x.device)
    out[:, :h, :w, :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def from_padded_nhwc(t: torch.Tensor, c: int, h: int, w: int) -> torch.Tensor:
    return t[:, :h, :w, :c].permute(0, 3, 1, 2).float().contiguous()


def pack_weights(wt: torch.Tensor, cin_pad: int, npad: int) -> torch.Tensor:
    """[n][c][k][k] fp32 -> bf16 [npad][k*k*cin_pad], K index = (r*k+s)*cin_pad + c."""
    n, c, k, _ = wt.shape
    out = torch.zeros(npad, k * k, cin_pad, dtype=torch.bfloat16, device=wt.device)
    out[:n, :, :c] = wt.permute(0, 2, 3, 1).reshape(n, k * k, c).to(torch.bfloat16)
    return out.reshape(npad, k * k * cin_pad).contiguous()


def run_conv(x_p, in_cs, cin, batch, h, w, ksize, wt_p, cout, npad, block_n, block_k,
             alpha, beta, act, out, out_cs, out_mode, repeat=1):
    lib = _lib.load()
    d = _lib.ConvDesc()
    d.in_ = ptr(x_p); d.in_cs = in_cs; d.cin = cin
    d.batch = batch; d.h = h; d.w = w; d.ksize = ksize
    d.wt = ptr(wt_p); d.cout = cout; d.npad = npad
    d.block_n = block_n; d.block_k = block_k
    d.alpha = ptr(alpha); d.beta = ptr(beta); d.act = act
    d.out = ptr(out); d.out_cs = out_cs; d.out_mode = out_mode
    plan = C.c_void_p()
    _lib.check(lib.y2_conv_plan_create(C.byref(d), C.byref(plan)), "conv_plan_create")
    try:
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(repeat):
            _lib.check(lib.y2_conv_plan_launch(plan, C.c_void_p(stream)), "conv_plan_launch")
        torch.cuda.synchronize()
    finally:
        lib.y2_conv_plan_destroy(plan)
