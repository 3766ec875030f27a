"""GPU parity tests of the individual sm_100a kernels, called through the C-ABI.

The convolution is a floating-point kernel, so it is compared with a plain PyTorch fp32
convolution of the same bf16-rounded operands (tolerances stated per test); the byte/index
kernels (maxpool, reorg, pack/unpack, route copy) must match bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from sr_object_detection_b200 import _lib  # noqa: E402
from tests import gpu_util as G  # noqa: E402

pytestmark = pytest.mark.gpu

ACT_LINEAR, ACT_LEAKY = 0, 1
OUT_BF16, OUT_F32 = 0, 1


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ref_conv(x, wt, alpha, beta, act, ksize):
    xb = x.to(torch.bfloat16).float()
    wb = wt.to(torch.bfloat16).float()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = torch.nn.functional.conv2d(xb, wb, padding=ksize // 2)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    y = y * alpha.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    if act == ACT_LEAKY:
        y = torch.where(y > 0, y, 0.1 * y)
    return y


CONV_CASES = [
    # (batch, cin, h, w, cout, ksize, block_n, block_k, act)
    (2, 64, 13, 13, 64, 1, 64, 64, ACT_LEAKY),
    (2, 64, 13, 13, 128, 3, 128, 64, ACT_LEAKY),
    (4, 128, 13, 13, 256, 3, 256, 64, ACT_LEAKY),
    (3, 32, 20, 28, 64, 3, 64, 32, ACT_LEAKY),
    (2, 256, 26, 26, 512, 3, 256, 64, ACT_LEAKY),
    (2, 512, 13, 13, 1024, 3, 128, 64, ACT_LEAKY),
    (1, 1024, 13, 13, 512, 1, 256, 64, ACT_LINEAR),
    (2, 32, 52, 52, 32, 1, 32, 32, ACT_LEAKY),
    (64, 64, 13, 13, 64, 3, 64, 64, ACT_LEAKY),
    # wide rows: the halo slab spans several TMA loads (104 -> 2 loads, 208 -> 3 loads of 64-byte rows)
    (2, 64, 104, 104, 128, 3, 128, 64, ACT_LEAKY),
    (1, 32, 208, 208, 64, 3, 64, 32, ACT_LEAKY),
    (3, 128, 52, 52, 256, 3, 256, 64, ACT_LEAKY),
    (5, 128, 17, 23, 64, 1, 64, 64, ACT_LINEAR),
    # more tiles than CTA pairs with a badly filled last wave: the pair kernel cuts the work into pieces of 64..256
    # filters and balances them over the pairs; the "roundrobin" variant runs whole 256-filter tiles instead
    (64, 256, 13, 13, 512, 3, 256, 64, ACT_LEAKY),
    (64, 128, 13, 13, 1024, 3, 256, 64, ACT_LINEAR),
    (48, 192, 13, 13, 768, 3, 256, 64, ACT_LEAKY),
    # balanced pieces with a filter count that is not a multiple of 64: direct (non-TMA) stores, clipped last piece
    (64, 128, 13, 13, 456, 3, 256, 64, ACT_LEAKY),
    # wide 1x1 layers (yolo-voc L13, L19): slab kernel by default, one-tap pair kernel under the "pair" variant
    (64, 512, 26, 26, 256, 1, 256, 64, ACT_LEAKY),
    (64, 1024, 13, 13, 512, 1, 256, 64, ACT_LEAKY),
]


@pytest.fixture(params=["auto", "slab", "pertap", "pair", "roundrobin"])
def conv_variant(request, monkeypatch):
    """All three implicit-GEMM kernels stay under test: the default choice (CTA-pair kernel for wide
    layers), the halo-slab kernel alone, the per-tap fallback, and the pair kernel wherever its shape
    constraints allow it, however small the layer (Y2_CONV_VARIANT is read at plan creation)."""
    monkeypatch.delenv("Y2_PAIR_BALANCE", raising=False)
    if request.param == "auto":  # CTA-pair kernel where it fits, then slab, then per-tap
        monkeypatch.delenv("Y2_CONV_VARIANT", raising=False)
    elif request.param == "roundrobin":  # default choice, pair kernel on whole tiles round-robin
        monkeypatch.delenv("Y2_CONV_VARIANT", raising=False)
        monkeypatch.setenv("Y2_PAIR_BALANCE", "0")
    else:
        monkeypatch.setenv("Y2_CONV_VARIANT", request.param)
    return request.param


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "b%d_c%d_%dx%d_n%d_k%d_bn%d_bk%d_a%d" % c)
def test_conv_bf16_matches_fp32_reference(case, conv_variant):
    batch, cin, h, w, cout, ksize, block_n, block_k, act = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234 + cin + cout)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev) * 2 - 1
    s = (2.0 / (ksize * ksize * cin)) ** 0.5
    wt = (torch.rand(cout, cin, ksize, ksize, generator=g).to(dev) * 2 - 1) * s
    npad = ((cout + block_n - 1) // block_n) * block_n
    alpha = torch.ones(npad, device=dev)
    beta = torch.zeros(npad, device=dev)
    alpha[:cout] = torch.rand(cout, generator=g).to(dev) + 0.5
    beta[:cout] = torch.rand(cout, generator=g).to(dev) * 0.2 - 0.1

    x_p = G.to_padded_nhwc(x)
    wt_p = G.pack_weights(wt, cin, npad)
    out = torch.full((batch, h + 1, w + 1, cout), 7.0, dtype=torch.bfloat16, device=dev)
    # three launches of the same plan: nothing a launch leaves behind (work lists, barriers) may change the next one
    G.run_conv(x_p, cin, cin, batch, h, w, ksize, wt_p, cout, npad, block_n, block_k, alpha, beta, act,
               out, cout, OUT_BF16, repeat=3)
    got = G.from_padded_nhwc(out, cout, h, w)
    ref = _ref_conv(x, wt, alpha[:cout], beta[:cout], act, ksize)
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    # bf16 output rounding is 2^-9 relative; fp32 accumulate. Tolerance: 1e-2 of the layer max.
    assert err <= 1e-2, f"max err / max|ref| = {err:.3e}"
    # pad row / column must be written as zeros
    assert out[:, h, :, :].abs().max().item() == 0
    assert out[:, :, w, :].abs().max().item() == 0


def test_conv_f32_flat_head_125(conv_variant):
    """1x1 linear head with 125 filters written as fp32 [B][H*W][125] (region-layer input)."""
    dev = torch.device("cuda:0")
    batch, cin, h, w, cout = 3, 1024, 13, 13, 125
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev) * 2 - 1
    wt = (torch.rand(cout, cin, 1, 1, generator=g).to(dev) * 2 - 1) * (2.0 / cin) ** 0.5
    npad = 128
    alpha = torch.ones(npad, device=dev)
    beta = torch.zeros(npad, device=dev)
    beta[:cout] = torch.rand(cout, generator=g).to(dev) - 0.5
    x_p = G.to_padded_nhwc(x)
    wt_p = G.pack_weights(wt, cin, npad)
    out = torch.zeros(batch, h * w, cout, dtype=torch.float32, device=dev)
    G.run_conv(x_p, cin, cin, batch, h, w, 1, wt_p, cout, npad, 128, 64, alpha, beta, ACT_LINEAR, out, cout,
               OUT_F32)
    ref = _ref_conv(x, wt, alpha[:cout], beta[:cout], ACT_LINEAR, 1)
    got = out.view(batch, h, w, cout).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    # operands are identical bf16 values on both sides; only fp32 summation order differs
    assert err <= 1e-4, f"max err / max|ref| = {err:.3e}"


@pytest.mark.parametrize("cout,batch,hw", [(1101, 3, 17), (2050, 9, 13), (1101, 64, 13)])
def test_conv_f32_flat_wide_head(conv_variant, cout, batch, hw):
    """Wide 1x1 linear head written as fp32 [B][H*W][cs] with staged row-contiguous stores (yolo9000's 28 269-filter
    head in small): filter counts that are not multiples of 4 / 64 / 256, row stride padded to 4 floats, the
    one-tap CTA-pair kernel by default and the slab / per-tap kernels under their variants."""
    dev = torch.device("cuda:0")
    cin, h, w = 128, hw, hw
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev) * 2 - 1
    wt = (torch.rand(cout, cin, 1, 1, generator=g).to(dev) * 2 - 1) * (2.0 / cin) ** 0.5
    npad = (cout + 255) // 256 * 256
    cs = (cout + 3) // 4 * 4
    alpha = torch.ones(npad, device=dev)
    beta = torch.zeros(npad, device=dev)
    alpha[:cout] = torch.rand(cout, generator=g).to(dev) + 0.5
    beta[:cout] = torch.rand(cout, generator=g).to(dev) - 0.5
    x_p = G.to_padded_nhwc(x)
    wt_p = G.pack_weights(wt, cin, npad)
    out = torch.full((batch, h * w, cs), 7.0, dtype=torch.float32, device=dev)
    G.run_conv(x_p, cin, cin, batch, h, w, 1, wt_p, cout, npad, 256, 64, alpha, beta, ACT_LINEAR, out, cs, OUT_F32)
    ref = _ref_conv(x, wt, alpha[:cout], beta[:cout], ACT_LINEAR, 1)
    got = out[:, :, :cout].view(batch, h, w, cout).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-4, f"max err / max|ref| = {err:.3e}"
    if cs > cout:  # the row padding is nobody's to write
        assert (out[:, :, cout:] == 7.0).all()


def test_conv_channel_slice_in_concat_buffer(conv_variant):
    """Output written at a channel offset of a wider buffer (in-place route) and input read
    from a channel slice; neighbouring channels must be untouched."""
    dev = torch.device("cuda:0")
    batch, cin, h, w, cout = 2, 64, 13, 13, 128
    g = torch.Generator(device="cpu").manual_seed(99)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev) * 2 - 1
    wt = (torch.rand(cout, cin, 3, 3, generator=g).to(dev) * 2 - 1) * (2.0 / (9 * cin)) ** 0.5
    alpha = torch.ones(cout, device=dev)
    beta = torch.zeros(cout, device=dev)
    in_cs, in_off = 192, 64
    buf_in = torch.full((batch, h + 1, w + 1, in_cs), 3.0, dtype=torch.bfloat16, device=dev)
    buf_in[:, :, :, in_off:in_off + cin] = G.to_padded_nhwc(x)
    out_cs, out_off = 320, 64
    buf_out = torch.full((batch, h + 1, w + 1, out_cs), 5.0, dtype=torch.bfloat16, device=dev)
    wt_p = G.pack_weights(wt, cin, cout)
    lib = _lib.load()
    d = _lib.ConvDesc()
    d.in_ = buf_in.data_ptr() + in_off * 2; d.in_cs = in_cs; d.cin = cin
    d.batch = batch; d.h = h; d.w = w; d.ksize = 3
    d.wt = wt_p.data_ptr(); d.cout = cout; d.npad = cout; d.block_n = 128; d.block_k = 64
    d.alpha = alpha.data_ptr(); d.beta = beta.data_ptr(); d.act = ACT_LEAKY
    d.out = buf_out.data_ptr() + out_off * 2; d.out_cs = out_cs; d.out_mode = OUT_BF16
    plan = C.c_void_p()
    _lib.check(lib.y2_conv_plan_create(C.byref(d), C.byref(plan)))
    _lib.check(lib.y2_conv_plan_launch(plan, _stream()))
    torch.cuda.synchronize()
    lib.y2_conv_plan_destroy(plan)
    ref = _ref_conv(x, wt, alpha, beta, ACT_LEAKY, 3)
    got = buf_out[:, :h, :w, out_off:out_off + cout].permute(0, 3, 1, 2).float()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-2
    assert (buf_out[..., :out_off] == 5.0).all() and (buf_out[..., out_off + cout:] == 5.0).all()


def test_first_layer_patches_path():
    """3-channel stem: patch gather (K=27 -> 32) + 1x1 GEMM == 3x3 convolution."""
    dev = torch.device("cuda:0")
    batch, cin, h, w, cout = 2, 3, 32, 48, 32
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev)
    wt = (torch.rand(cout, cin, 3, 3, generator=g).to(dev) * 2 - 1) * (2.0 / 27) ** 0.5
    alpha = torch.rand(cout, generator=g).to(dev) + 0.5
    beta = torch.rand(cout, generator=g).to(dev) * 0.2
    lib = _lib.load()
    patches = torch.empty(batch, h + 1, w + 1, 32, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_pack_patches_f32(x.data_ptr(), patches.data_ptr(), batch, cin, h, w, 3, 32, _stream()))
    # reference patches from torch unfold: K index = c*9 + r*3 + s
    unf = torch.nn.functional.unfold(x, 3, padding=1).view(batch, 27, h, w)
    assert torch.equal(patches[:, :h, :w, :27].permute(0, 3, 1, 2), unf.to(torch.bfloat16))
    assert patches[:, :h, :w, 27:].abs().max().item() == 0
    wt_p = torch.zeros(cout, 32, dtype=torch.bfloat16, device=dev)
    wt_p[:, :27] = wt.reshape(cout, 27).to(torch.bfloat16)
    out = torch.zeros(batch, h + 1, w + 1, cout, dtype=torch.bfloat16, device=dev)
    G.run_conv(patches, 32, 32, batch, h, w, 1, wt_p, cout, cout, 32, 32, alpha, beta, ACT_LEAKY, out, cout,
               OUT_BF16)
    ref = _ref_conv(x, wt, alpha, beta, ACT_LEAKY, 3)
    got = G.from_padded_nhwc(out, cout, h, w)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-2


@pytest.mark.parametrize("c,h,w,k,stride,pad", [(3, 32, 48, 7, 2, 3), (3, 31, 29, 11, 4, 2), (1, 20, 20, 5, 1, 2),
                                                (3, 16, 16, 3, 2, 1)])
def test_gather_patches_f32_matches_im2col(c, h, w, k, stride, pad):
    """General first-layer gather (any size / stride / padding) against torch's unfold, which walks the
    same [c][kh][kw] K order as the reference's im2col_cpu (im2col.c:16-39).  Bit-exact."""
    dev = torch.device("cuda:0")
    batch = 2
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    kreal = c * k * k
    kpad = 32 if kreal <= 32 else (kreal + 63) // 64 * 64
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.rand(batch, c, h, w, generator=g).to(dev)
    lib = _lib.load()
    patches = torch.full((batch, oh + 1, ow + 1, kpad), 7.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_gather_patches_f32(x.data_ptr(), patches.data_ptr(), batch, c, h, w, k, stride, pad, oh, ow,
                                         kpad, _stream()))
    unf = torch.nn.functional.unfold(x, k, padding=pad, stride=stride).view(batch, kreal, oh, ow)
    assert torch.equal(patches[:, :oh, :ow, :kreal].permute(0, 3, 1, 2), unf.to(torch.bfloat16))
    assert patches[:, :oh, :ow, kreal:].abs().max().item() == 0
    assert patches[:, oh].abs().max().item() == 0 and patches[:, :, ow].abs().max().item() == 0


@pytest.mark.parametrize("c,cs,h,w,k,stride,pad", [(64, 64, 16, 16, 3, 2, 1), (32, 96, 13, 17, 3, 2, 1),
                                                   (64, 64, 9, 9, 1, 2, 0), (32, 32, 12, 12, 5, 1, 2)])
def test_gather_patches_bf16_strided_conv(c, cs, h, w, k, stride, pad):
    """Mid-network gather (bf16 padded NHWC in, K = tap*cin_pad + c): rows bit-exact against unfold, and
    gather + 1x1 GEMM == the strided convolution within the bf16 tolerance."""
    dev = torch.device("cuda:0")
    batch, cout = 2, 64
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (torch.rand(batch, c, h, w, generator=g).to(dev) - 0.5)
    xin = torch.zeros(batch, h + 1, w + 1, cs, dtype=torch.bfloat16, device=dev)
    xin[:, :h, :w, :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    xin[..., c:] = 3.0  # a neighbour's channels in a concat buffer must not leak in
    xin[:, h, :, c:] = 0
    lib = _lib.load()
    kk = k * k
    patches = torch.full((batch, oh + 1, ow + 1, kk * c), 7.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_gather_patches_bf16(xin.data_ptr(), cs, c, h, w, patches.data_ptr(), batch, k, stride, pad, oh,
                                          ow, _stream()))
    unf = torch.nn.functional.unfold(x.to(torch.bfloat16).float(), k, padding=pad, stride=stride)
    unf = unf.view(batch, c, kk, oh, ow).permute(0, 3, 4, 2, 1).reshape(batch, oh, ow, kk * c)
    assert torch.equal(patches[:, :oh, :ow], unf.to(torch.bfloat16))
    assert patches[:, oh].abs().max().item() == 0 and patches[:, :, ow].abs().max().item() == 0
    wt = (torch.rand(cout, c, k, k, generator=g).to(dev) * 2 - 1) * (2.0 / (kk * c)) ** 0.5
    alpha = torch.rand(cout, generator=g).to(dev) + 0.5
    beta = torch.rand(cout, generator=g).to(dev) * 0.2
    wt_p = wt.permute(0, 2, 3, 1).reshape(cout, kk * c).to(torch.bfloat16).contiguous()
    out = torch.zeros(batch, oh + 1, ow + 1, cout, dtype=torch.bfloat16, device=dev)
    bk = 64 if (kk * c) % 64 == 0 else 32
    G.run_conv(patches, kk * c, kk * c, batch, oh, ow, 1, wt_p, cout, cout, 64, bk, alpha, beta, ACT_LEAKY, out, cout,
               OUT_BF16)
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    ref = torch.nn.functional.conv2d(xb, wb, padding=pad, stride=stride) * alpha.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    ref = torch.where(ref > 0, ref, 0.1 * ref)
    got = G.from_padded_nhwc(out, cout, oh, ow)
    assert (got - ref).abs().max().item() / ref.abs().max().item() <= 1e-2


def test_pack_unpack_roundtrip_bit_exact():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, c, h, w, cpad, cs = 3, 20, 17, 23, 32, 64
    x = torch.randn(batch, c, h, w, device=dev)
    packed = torch.full((batch, h + 1, w + 1, cs), 9.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_pack_nchw_f32(x.data_ptr(), packed.data_ptr(), batch, c, h, w, cpad, cs, _stream()))
    ref = G.to_padded_nhwc(x, cpad)
    assert torch.equal(packed[..., :cpad], ref)
    assert (packed[..., cpad:] == 9.0).all()
    back = torch.empty(batch, c, h, w, device=dev)
    _lib.check(lib.y2_unpack_to_nchw_f32(packed.data_ptr(), back.data_ptr(), batch, c, h, w, cs, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(back, x.to(torch.bfloat16).float())


@pytest.mark.parametrize("size,stride,h,w", [(2, 2, 26, 26), (2, 1, 13, 13), (2, 2, 416, 416), (3, 2, 31, 30)])
def test_maxpool_bit_exact(size, stride, h, w):
    """maxpool_layer.c:79-114 semantics: padding=(size-1)/2, out=(w+2p)/stride, window starts
    at -padding, cells outside the image are skipped."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, c = 2, 32
    pad = (size - 1) // 2
    oh, ow = (h + 2 * pad) // stride, (w + 2 * pad) // stride
    x = torch.randn(batch, c, h, w, device=dev).to(torch.bfloat16).float()
    x_p = G.to_padded_nhwc(x)
    out = torch.full((batch, oh + 1, ow + 1, c), 3.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_maxpool(x_p.data_ptr(), c, out.data_ptr(), c, batch, c, h, w, oh, ow, size, stride, pad,
                              _stream()))
    torch.cuda.synchronize()
    neg = torch.finfo(torch.float32).min
    # window covers rows [i*stride - pad, i*stride - pad + size): pad left by `pad`, right generously
    xp = torch.nn.functional.pad(x, (pad, size, pad, size), value=neg)
    ref = torch.nn.functional.max_pool2d(xp, size, stride)[:, :, :oh, :ow]
    got = G.from_padded_nhwc(out, c, oh, ow)
    assert torch.equal(got, ref)
    assert out[:, oh].abs().max().item() == 0 and out[:, :, ow].abs().max().item() == 0


def _reorg_reference(x: np.ndarray, stride: int) -> np.ndarray:
    """reorg_cpu(x, w, h, c, batch, stride, forward=0, out) of blas.c:8-29, vectorised:
    out.flat[in_index] = x.flat[out_index] per image."""
    b, c, h, w = x.shape
    out_c = c // (stride * stride)
    k, j, i = np.meshgrid(np.arange(c), np.arange(h), np.arange(w), indexing="ij")
    in_index = i + w * (j + h * k)
    c2 = k % out_c
    offset = k // out_c
    w2 = i * stride + offset % stride
    h2 = j * stride + offset // stride
    out_index = w2 + w * stride * (h2 + h * stride * c2)
    out = np.empty_like(x).reshape(b, -1)
    xf = x.reshape(b, -1)
    out[:, in_index.ravel()] = xf[:, out_index.ravel()]
    return out.reshape(b, c * stride * stride, h // stride, w // stride)


def test_reorg_matches_reference_index_map():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, c, h, w, stride = 3, 64, 26, 26, 2
    x = torch.randn(batch, c, h, w, device=dev).to(torch.bfloat16).float()
    x_p = G.to_padded_nhwc(x)
    oc, oh, ow = c * stride * stride, h // stride, w // stride
    out_cs, off = 1280, 0
    out = torch.full((batch, oh + 1, ow + 1, out_cs), 2.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_reorg(x_p.data_ptr(), c, out.data_ptr() + off * 2, out_cs, batch, c, h, w, stride, _stream()))
    torch.cuda.synchronize()
    ref = torch.from_numpy(_reorg_reference(x.cpu().numpy(), stride)).to(dev)
    got = out[:, :oh, :ow, off:off + oc].permute(0, 3, 1, 2).float()
    assert torch.equal(got, ref)
    assert (out[..., oc:] == 2.0).all()
    # the table-driven gather the network schedule uses must produce the same bytes
    table = torch.zeros(oh * ow * oc, dtype=torch.int32, device=dev)
    out2 = torch.full((batch, oh + 1, ow + 1, out_cs), 2.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_reorg_table(table.data_ptr(), c, c, h, w, stride, _stream()))
    _lib.check(lib.y2_reorg_gather(x_p.data_ptr(), c, out2.data_ptr() + off * 2, out_cs, table.data_ptr(), batch, c, h,
                                   w, stride, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out2, out)


def test_copy_channels_route_fallback():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, c, h, w = 2, 64, 13, 13
    src = torch.randn(batch, h + 1, w + 1, 96, device=dev).to(torch.bfloat16)
    dst = torch.zeros(batch, h + 1, w + 1, 160, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.y2_copy_channels(src.data_ptr() + 16 * 2, 96, dst.data_ptr() + 32 * 2, 160, batch, c, h, w,
                                    _stream()))
    torch.cuda.synchronize()
    assert torch.equal(dst[..., 32:96], src[..., 16:80])
    assert dst[..., :32].abs().max().item() == 0 and dst[..., 96:].abs().max().item() == 0


@pytest.mark.parametrize("batch,c,h,w,n", [(2, 3, 50, 38, 32), (3, 3, 64, 96, 16), (1, 1, 33, 35, 8), (4, 3, 416, 416, 32)])
def test_stem_conv_pool_matches_fp32_reference(batch, c, h, w, n):
    """Fused first layer: conv3x3 + affine + leaky + 2x2/2 maxpool from fp32 NCHW, vs PyTorch fp32 on
    the same bf16-rounded operands.  Tolerance 1e-2 of the tensor max (bf16 output rounding)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(99 + h + n)
    x = torch.rand(batch, c, h, w, generator=g).to(dev)
    s = (2.0 / (9 * c)) ** 0.5
    wt = ((torch.rand(n, c, 3, 3, generator=g) * 2 - 1) * s).to(dev)
    alpha = torch.zeros(32, device=dev)
    beta = torch.zeros(32, device=dev)
    alpha[:n] = torch.rand(n, generator=g).to(dev) + 0.5
    beta[:n] = torch.rand(n, generator=g).to(dev) * 0.4 - 0.2
    wt_p = torch.zeros(32, 32, dtype=torch.bfloat16, device=dev)
    wt_p[:n, :c * 9] = wt.reshape(n, c * 9).to(torch.bfloat16)
    oh, ow = h // 2, w // 2
    out = torch.zeros(batch, oh + 1, ow + 1, 32, dtype=torch.bfloat16, device=dev)
    lib = _lib.load()
    _lib.check(lib.y2_stem_conv_pool(x.data_ptr(), batch, c, h, w, wt_p.data_ptr(), 32, alpha.data_ptr(),
                                     beta.data_ptr(), ACT_LEAKY, out.data_ptr(), 32, _stream()), "stem")
    torch.cuda.synchronize()
    ref = _ref_conv(x, wt, alpha[:n], beta[:n], ACT_LEAKY, 3)
    ref = torch.nn.functional.max_pool2d(ref, 2, 2)
    got = G.from_padded_nhwc(out, n, oh, ow)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-2, f"max err / max|ref| = {err:.3e}"
    # pads and unused channels stay zero
    assert out[:, oh, :, :].abs().max().item() == 0 and out[:, :, ow, :].abs().max().item() == 0
    if n < 32:
        assert out[..., n:].abs().max().item() == 0


@pytest.mark.parametrize("batch,cin,h,w,cout", [(2, 32, 208, 208, 64), (2, 64, 104, 104, 128), (1, 32, 50, 38, 64),
                                                (3, 64, 26, 30, 64), (2, 32, 13, 13, 128),
                                                # many patches per CTA with an ODD number of row pairs per patch
                                                # (tiny-yolo-voc / yolo.cfg 608 shapes at serving batch sizes): the
                                                # TMEM slots must stay with the same epilogue group across patches
                                                (64, 32, 104, 104, 64), (64, 64, 52, 52, 128), (16, 32, 304, 304, 64),
                                                (24, 64, 152, 152, 128)])
def test_conv_pool_fused_matches_fp32_reference(batch, cin, h, w, cout):
    """3x3 conv + affine + leaky + 2x2/2 maxpool in one launch (Y2_OUT_BF16_POOLED) against PyTorch
    fp32 on the same bf16-rounded operands.  Tolerance 1e-2 of the tensor max (bf16 output rounding)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5 + cin + cout + h)
    x = torch.rand(batch, cin, h, w, generator=g).to(dev) * 2 - 1
    wt = (torch.rand(cout, cin, 3, 3, generator=g).to(dev) * 2 - 1) * (2.0 / (9 * cin)) ** 0.5
    alpha = (torch.rand(cout, generator=g) + 0.5).to(dev)
    beta = (torch.rand(cout, generator=g) * 0.2 - 0.1).to(dev)
    x_p = G.to_padded_nhwc(x)
    wt_p = G.pack_weights(wt, cin, cout)
    oh, ow = h // 2, w // 2
    out_cs = cout + 64
    out = torch.zeros(batch, oh + 1, ow + 1, out_cs, dtype=torch.bfloat16, device=dev)
    G.run_conv(x_p, cin, cin, batch, h, w, 3, wt_p, cout, cout, cout, cin, alpha, beta, ACT_LEAKY, out, out_cs, 2)
    ref = torch.nn.functional.max_pool2d(_ref_conv(x, wt, alpha, beta, ACT_LEAKY, 3), 2, 2)
    got = G.from_padded_nhwc(out, cout, oh, ow)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-2, f"max err / max|ref| = {err:.3e}"
    # pads and the neighbouring channels of the wider buffer stay untouched
    assert out[:, oh].abs().max().item() == 0 and out[:, :, ow].abs().max().item() == 0
    assert out[..., cout:].abs().max().item() == 0


@pytest.mark.parametrize("batch,h,w", [(2, 64, 128), (3, 416, 416), (1, 50, 160)])
def test_stem_u8_input_is_bit_identical_to_the_float_path(batch, h, w):
    """uint8 interleaved RGB input (the decoded image) against the fp32 planar input the reference's
    loaders would have produced from it ((float)byte / 255., yolo_v2_class.cpp:129-149): identical bytes out."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1000 + h)
    u8 = torch.randint(0, 256, (batch, h, w, 3), generator=g, dtype=torch.uint8)
    planar = (u8.permute(0, 3, 1, 2).to(torch.float32).to(torch.float64) / 255.0).to(torch.float32).contiguous()
    wt = ((torch.rand(32, 3, 3, 3, generator=g) * 2 - 1) * (2.0 / 27) ** 0.5)
    wt_p = torch.zeros(32, 32, dtype=torch.bfloat16)
    wt_p[:, :27] = wt.reshape(32, 27).to(torch.bfloat16)
    alpha = torch.rand(32, generator=g) + 0.5
    beta = torch.rand(32, generator=g) * 0.4 - 0.2
    u8, planar, wt_p, alpha, beta = (t.to(dev) for t in (u8.contiguous(), planar, wt_p, alpha, beta))
    oh, ow = h // 2, w // 2
    out_f = torch.zeros(batch, oh + 1, ow + 1, 32, dtype=torch.bfloat16, device=dev)
    out_u = torch.zeros_like(out_f)
    lib = _lib.load()
    _lib.check(lib.y2_stem_conv_pool(planar.data_ptr(), batch, 3, h, w, wt_p.data_ptr(), 32, alpha.data_ptr(),
                                     beta.data_ptr(), ACT_LEAKY, out_f.data_ptr(), 32, _stream()), "stem f32")
    _lib.check(lib.y2_stem_conv_pool_u8(u8.data_ptr(), batch, h, w, wt_p.data_ptr(), 32, alpha.data_ptr(),
                                        beta.data_ptr(), ACT_LEAKY, out_u.data_ptr(), 32, _stream()), "stem u8")
    torch.cuda.synchronize()
    assert out_f.abs().max().item() > 0
    assert torch.equal(out_u.view(torch.int16), out_f.view(torch.int16))


def test_device_vector_helpers_and_activate_array_ongpu():
    """fill/copy/axpy/scal_ongpu and activate_array_ongpu on cuda_make_array buffers (blas.h:43-53,
    activations.h:18): strided and unit-stride (float4) paths; axpy/scal without fma -> bit-exact."""
    from sr_object_detection_b200 import darknet as dn
    lib = dn.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    for n, incx, incy in [(1000, 1, 1), (1003, 1, 1), (333, 3, 2), (5, 1, 1)]:
        x = torch.randn(n * incx + 3, generator=g).to(dev)
        y = torch.randn(n * incy + 3, generator=g).to(dev)
        want = y.clone()
        want[0:n * incy:incy] = want[0:n * incy:incy] + (0.37 * x[0:n * incx:incx])
        lib.axpy_ongpu(n, 0.37, x.data_ptr(), incx, y.data_ptr(), incy)
        torch.cuda.synchronize()
        assert torch.equal(y, want)
        lib.scal_ongpu(n, 1.7, y.data_ptr(), incy)
        want[0:n * incy:incy] *= 1.7
        torch.cuda.synchronize()
        assert torch.equal(y, want)
        lib.copy_ongpu(n, x.data_ptr(), incx, y.data_ptr(), incy)
        want[0:n * incy:incy] = x[0:n * incx:incx]
        torch.cuda.synchronize()
        assert torch.equal(y, want)
        lib.fill_ongpu(n, -3.5, y.data_ptr(), incy)
        want[0:n * incy:incy] = -3.5
        torch.cuda.synchronize()
        assert torch.equal(y, want)
    # unaligned base pointer: falls back to the scalar kernel
    z = torch.zeros(64, device=dev)
    lib.fill_ongpu(40, 2.0, z.data_ptr() + 4, 1)
    torch.cuda.synchronize()
    assert z[0].item() == 0 and (z[1:41] == 2).all() and (z[41:] == 0).all()
    # activations: host activate_array (double math) is the checker for the device kernel (float math)
    x = torch.linspace(-6, 6, 4097)
    for a in range(13):
        host = x.clone().numpy()
        lib.activate_array(host.ctypes.data_as(C.POINTER(C.c_float)), host.size, a)
        d = x.to(dev)
        lib.activate_array_ongpu(d.data_ptr(), d.numel(), a)
        torch.cuda.synchronize()
        assert np.allclose(d.cpu().numpy(), host, rtol=2e-6, atol=2e-7), a
    # cuda_make_array / push / pull round trip
    h = np.arange(100, dtype=np.float32)
    fp = h.ctypes.data_as(C.POINTER(C.c_float))
    dptr = lib.cuda_make_array(fp, h.size)
    lib.scal_ongpu(h.size, 2.0, dptr, 1)
    back = np.zeros_like(h)
    lib.cuda_pull_array(dptr, back.ctypes.data_as(C.POINTER(C.c_float)), h.size)
    assert np.array_equal(back, h * 2)
    lib.cuda_free(dptr)


@pytest.mark.parametrize("sw,sh,w,h", [(640, 480, 416, 416), (300, 200, 416, 416), (416, 416, 416, 416),
                                       (1920, 1080, 608, 608), (37, 53, 64, 32), (1, 9, 32, 32)])
def test_device_resize_is_bit_identical_to_load_image_plus_resize_image(tmp_path, sw, sh, w, h):
    """y2_resize_u8_to_f32 (uint8 HWC frame -> fp32 planar network input) against the CPU chain it
    replaces: byte/255. (yolo_v2_class.cpp:129-149) then resize_image (image.c:1950-1993), computed by the
    oracle (pinned to the reference by tests/golden/resize.npz) and by the library's own host resize_image."""
    import subprocess
    from sr_object_detection_b200 import darknet as dn
    from tests import ref_util as R
    R.build_oracle()
    dev = torch.device("cuda:0")
    batch = 2
    rng = np.random.default_rng(sw * 7 + sh)
    u8 = rng.integers(0, 256, size=(batch, sh, sw, 3), dtype=np.uint8)
    planar = (u8.transpose(0, 3, 1, 2).astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    lib = _lib.load()
    src = torch.from_numpy(u8).to(dev)
    dst = torch.full((batch, 3, h, w), -1.0, device=dev)
    _lib.check(lib.y2_resize_u8_to_f32(src.data_ptr(), dst.data_ptr(), batch, sw, sh, w, h, _stream()))
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    dl = dn.lib()
    for b in range(batch):
        np.ascontiguousarray(planar[b]).tofile(tmp_path / "im.f32")
        subprocess.run([str(R.ORACLE_BIN), "resize", "im.f32", "3", str(sh), str(sw), str(h), str(w), "out.f32"],
                       check=True, cwd=tmp_path, capture_output=True)
        want = np.fromfile(tmp_path / "out.f32", dtype=np.float32).reshape(3, h, w)
        assert np.array_equal(got[b].view(np.uint32), want.view(np.uint32)), \
            f"frame {b}: {(got[b] != want).sum()} of {want.size} values differ from the oracle"
        src_img = np.ascontiguousarray(planar[b])
        im = dn.Image(sh, sw, 3, src_img.ctypes.data_as(C.POINTER(C.c_float)))
        out = dl.resize_image(im, w, h)
        host = np.ctypeslib.as_array(out.data, shape=(3, h, w)).copy()
        dl.free_image(out)
        assert np.array_equal(host.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("classes", [20, 80, 300, 1001])
def test_collect_final_pick_matches_max_index(classes):
    """y2_collect = the final loop of Detector::detect (yolo_v2_class.cpp:221-239): per box max_index
    (utils.c:533-545: the FIRST maximum), keep if prob > thresh, emit in box order.  Rows of >= 256 classes
    take the warp-per-box scan (box_argmax_kernel); ties and all-zero rows are planted on purpose."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, total, max_det, thresh = 3, 150, 64, 0.3
    rng = np.random.default_rng(classes)
    probs = np.zeros((batch, total, classes), dtype=np.float32)
    hot = rng.random((batch, total)) < 0.3
    for b, i in zip(*np.nonzero(hot)):
        k = rng.integers(1, 4)
        idx = rng.choice(classes, k, replace=False)
        probs[b, i, idx] = rng.random(k).astype(np.float32)
        if rng.random() < 0.3:  # an exact tie for the maximum: the lower index must win
            j = rng.choice(classes, 2, replace=False)
            probs[b, i, j] = probs[b, i].max() + np.float32(0.1)
    boxes = rng.random((batch, total, 4)).astype(np.float32)
    d_boxes, d_probs = torch.from_numpy(boxes).to(dev), torch.from_numpy(probs).to(dev)
    det = torch.zeros(batch * max_det * 7, dtype=torch.int32, device=dev)
    cnt = torch.zeros(batch, dtype=torch.int32, device=dev)
    _lib.check(lib.y2_collect(d_boxes.data_ptr(), d_probs.data_ptr(), batch, total, classes, thresh, det.data_ptr(),
                              cnt.data_ptr(), max_det, _stream()))
    torch.cuda.synchronize()
    raw = det.cpu().numpy().reshape(batch, max_det, 7)
    counts = cnt.cpu().numpy()
    for b in range(batch):
        arg = probs[b].argmax(axis=1)  # numpy argmax returns the first maximum, like max_index
        best = probs[b, np.arange(total), arg]
        keep = np.nonzero(best > np.float32(thresh))[0]
        assert counts[b] == len(keep)
        n = min(len(keep), max_det)
        got = raw[b, :n]
        assert np.array_equal(got[:, 6], keep[:n])                      # box_index, in box order
        assert np.array_equal(got[:, 5], arg[keep[:n]])                 # obj_id
        assert np.array_equal(got[:, 4].view(np.float32), best[keep[:n]])
        assert np.array_equal(got[:, :4].view(np.float32), boxes[b, keep[:n]])


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_ongpu_helper_matches_fp32_reference(ta, tb):
    """gemm_ongpu of the helper surface (gemm.c:173-183): C = ALPHA op(A) op(B) + BETA C, row-major, fp32."""
    import ctypes as C
    from sr_object_detection_b200 import darknet as dn
    lib = dn.lib()
    M, N, K = 70, 45, 133
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    Cm = torch.randn(M, N, generator=g)
    want = 0.5 * ((A.t() if ta else A).double() @ (B.t() if tb else B).double()) + 2.0 * Cm.double()
    dA, dB, dC = A.cuda().contiguous(), B.cuda().contiguous(), Cm.cuda().contiguous()
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.gemm_ongpu.restype = None
    lib.gemm_ongpu.argtypes = [i, i, i, i, i, f, vp, i, vp, i, f, vp, i]
    torch.cuda.synchronize()
    lib.gemm_ongpu(ta, tb, M, N, K, 0.5, dA.data_ptr(), A.shape[1], dB.data_ptr(), B.shape[1], 2.0, dC.data_ptr(), N)
    torch.cuda.synchronize()
    assert torch.allclose(dC.cpu().double(), want, rtol=1e-5, atol=1e-4)


def test_im2col_ongpu_helper_matches_unfold():
    """im2col_ongpu (im2col_kernels.cu:48-61): rows ordered c*k*k + i*k + j, zero outside the image."""
    import ctypes as C
    from sr_object_detection_b200 import darknet as dn
    lib = dn.lib()
    c, h, w, k, stride, pad = 5, 13, 17, 3, 2, 1
    x = torch.randn(c, h, w)
    want = torch.nn.functional.unfold(x[None], k, padding=pad, stride=stride)[0]
    dx = x.cuda().contiguous()
    col = torch.empty_like(want, device="cuda")
    vp, i = C.c_void_p, C.c_int
    lib.im2col_ongpu.restype = None
    lib.im2col_ongpu.argtypes = [vp, i, i, i, i, i, i, vp]
    torch.cuda.synchronize()
    lib.im2col_ongpu(dx.data_ptr(), c, h, w, k, stride, pad, col.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(col.cpu(), want)


def _shortcut_reference(x, add, act):
    """shortcut_layer.c:39-59 -> blas.c:57-81 (shortcut_cpu) + activate_array, on NCHW fp32 tensors."""
    b, c2, h2, w2 = x.shape
    _, c1, h1, w1 = add.shape
    stride, sample = max(w1 // w2, 1), max(w2 // w1, 1)
    minw, minh, minc = min(w1, w2), min(h1, h2), min(c1, c2)
    out = x.clone()
    out[:, :minc, 0:minh * sample:sample, 0:minw * sample:sample] += add[:, :minc, 0:minh * stride:stride, 0:minw * stride:stride]
    if act == ACT_LEAKY:
        out = torch.where(out > 0, out, 0.1 * out)
    return out


@pytest.mark.parametrize("case", [
    # (out_c, out_h, out_w, add_c, add_h, add_w, act, fp32 stream in, fp32 stream out)
    (256, 16, 16, 256, 16, 16, ACT_LEAKY, True, True),     # resnet50's common block: the division-free kernel
    (256, 16, 16, 256, 16, 16, ACT_LEAKY, False, True),
    (192, 9, 13, 192, 9, 13, ACT_LINEAR, True, False),     # 24 channel groups: not a power of two
    (64, 8, 8, 128, 8, 8, ACT_LEAKY, False, False),        # `from` has more channels than the running tensor
    (256, 16, 16, 64, 16, 16, ACT_LEAKY, False, True),     # first block of a stage: 64 channels into 256 (general)
    (128, 8, 8, 64, 16, 16, ACT_LEAKY, True, True),        # downsampling block: stride 2 (general)
], ids=lambda c: "c%d_%dx%d_from_c%d_%dx%d_a%d_%d%d" % c)
def test_shortcut_matches_reference(case, monkeypatch):
    out_c, oh, ow, add_c, ah, aw, act, in32, out32 = case
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch = 3
    g = torch.Generator(device="cpu").manual_seed(7)
    x = (torch.randn(batch, out_c, oh, ow, generator=g)).to(dev).to(torch.bfloat16).float()
    add_f = torch.randn(batch, add_c, ah, aw, generator=g).to(dev)
    add_b = add_f.to(torch.bfloat16)
    x_p, add_p = G.to_padded_nhwc(x), G.to_padded_nhwc(add_b.float())
    add32_p = torch.zeros(batch, ah + 1, aw + 1, add_c, device=dev)
    add32_p[:, :ah, :aw, :] = add_f.permute(0, 2, 3, 1)
    for general in (False, True):
        if general:
            monkeypatch.setenv("Y2_SHORTCUT_GENERAL", "1")
        out = torch.full((batch, oh + 1, ow + 1, out_c), 3.0, dtype=torch.bfloat16, device=dev)
        o32 = torch.zeros(batch, oh + 1, ow + 1, out_c, device=dev) if out32 else None
        _lib.check(lib.y2_shortcut(x_p.data_ptr(), out_c, add_p.data_ptr(), add_c, add_c, ah, aw, out.data_ptr(), out_c,
                                   out_c, out_c, oh, ow, batch, act, add32_p.data_ptr() if in32 else None, add_c,
                                   o32.data_ptr() if out32 else None, _stream()))
        torch.cuda.synchronize()
        ref = _shortcut_reference(x, add_f if in32 else add_b.float(), act)
        got = G.from_padded_nhwc(out, out_c, oh, ow)
        assert torch.equal(got, ref.to(torch.bfloat16).float()), f"general={general}"
        assert out[:, oh].abs().max().item() == 0 and out[:, :, ow].abs().max().item() == 0   # pads stay zero
        if out32:
            assert torch.equal(o32[:, :oh, :ow, :].permute(0, 3, 1, 2), ref)
            assert o32[:, oh].abs().max().item() == 0 and o32[:, :, ow].abs().max().item() == 0


@pytest.mark.parametrize("rows,n", [(64, 1000), (3, 257), (5, 4096), (2, 100)])
def test_softmax_rows_bit_exact_in_both_forms(rows, n, monkeypatch):
    """softmax_layer.c:49-61 -> blas.c:205-221: exp in double rounded to float, float sum in index order, divide.
    The block-per-row kernel (wide rows) and the warp-per-row kernel must both reproduce that sequence bit for bit."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(rows, n, generator=g) * 3).contiguous()
    xn = x.numpy()
    largest = xn.max(axis=1, keepdims=True)
    e = np.exp((xn / np.float32(1.0) - largest / np.float32(1.0)).astype(np.float64)).astype(np.float32)
    s = np.cumsum(e, axis=1, dtype=np.float32)[:, -1:]   # sequential float accumulation
    ref = e / s
    xd = x.to(dev)
    for force_warp in (False, True):
        if force_warp:
            monkeypatch.setenv("Y2_SOFTMAX_WARP_ROWS", "1")
        out = torch.zeros(rows, n, device=dev)
        _lib.check(lib.y2_softmax_rows(xd.data_ptr(), out.data_ptr(), rows, n, 1.0, _stream()))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32)), f"warp form: {force_warp}"


def test_avgpool_flat_sequential_sum():
    """avgpool_layer.c:40-55: per (image, channel) a sequential float sum over h*w, then / (h*w)."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    batch, hw, c, cs = 3, 196, 1000, 1000
    g = torch.Generator(device="cpu").manual_seed(6)
    x = torch.randn(batch, hw, cs, generator=g).contiguous()
    ref = (np.cumsum(x.numpy()[:, :, :c], axis=1, dtype=np.float32)[:, -1, :] / np.float32(hw)).astype(np.float32)
    out = torch.zeros(batch, c, device=dev)
    _lib.check(lib.y2_avgpool_flat(x.to(dev).data_ptr(), out.data_ptr(), batch, hw, c, cs, _stream()))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))
