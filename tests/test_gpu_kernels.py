"""GPU parity tests (-m gpu) of the individual kernels, all called through the C ABI
(edvr_b200/_lib.py -> libedvr_b200.so).  Checkers: the C oracle (oracle/dcn_oracle.c), the golden
vectors of the unmodified reference CUDA extension, and plain fp32 PyTorch ops for the dense stages.

Tolerances (stated per test): tensor-core stages use fp16 operands (10-bit mantissa, the same operand
precision as the TF32 path of the reference's cuDNN convs) with fp32 accumulation and fp16 storage, so
the bound is 1e-3 on max|a-b|/max|b| and on the relative L2 error (north_star: 1e-3 rel).  Index maps
(pixel shuffle, layout converters) are bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from edvr_b200 import ops as o
    return o


def test_library_loaded_and_selftest(ops):
    """tcgen05 descriptor convention: D = A.B^T through the same smem layout the kernels use."""
    from edvr_b200 import _lib as L
    assert os.path.exists(L.LIB_PATH)
    for N, K in [(128, 64), (64, 32), (256, 16), (96, 128)]:
        A = torch.randn(128, K, device="cuda").half()
        B = torch.randn(N, K, device="cuda").half()
        D = torch.empty(128, N, device="cuda")
        L.check(L.lib().eb_selftest_umma(L.ptr(A), L.ptr(B), L.ptr(D), N, K, 0, L.stream_ptr()))
        torch.cuda.synchronize()
        assert rel_err(D.cpu(), (A.float() @ B.float().t()).cpu())[0] < 1e-5


def test_layout_roundtrip_bit_exact(ops):
    x = torch.randn(2, 24, 7, 13, device="cuda").half().float()
    v = ops.nchw_to_nhwc(x)
    assert torch.equal(v.t.permute(0, 3, 1, 2).float(), x)
    assert torch.equal(ops.nhwc_to_nchw(v), x)


CONV_CASES = [
    # name, N, H, W, cins, cout, k, act, res, out_mode, maps
    ("3x3_tile_exact", 1, 16, 16, [64], 64, 3, "none", False, "same", None),
    ("3x3_ragged", 2, 21, 37, [64], 64, 3, "none", False, "same", None),
    ("3x3_c128_lrelu", 2, 45, 80, [128], 128, 3, "lrelu", False, "same", None),
    ("1x1_relu", 1, 33, 50, [128], 128, 1, "relu", False, "same", None),
    ("3x3_two_sources_residual", 3, 24, 40, [128, 128], 128, 3, "none", True, "same", None),
    ("3x3_broadcast_source", 4, 20, 24, [64, 64], 64, 3, "none", False, "same", [None, (2, 2, 0, 1, 4)]),
    ("3x3_pixel_shuffle", 1, 18, 20, [128], 512, 3, "lrelu", False, "pixshuf", None),
    ("3x3_stride2_even", 2, 22, 30, [64], 64, 3, "lrelu", False, "stride2", None),
    ("3x3_stride2_odd", 1, 45, 27, [64], 64, 3, "none", False, "stride2", None),
    ("1x1_c896_o256", 1, 20, 32, [896], 256, 1, "lrelu", False, "same", None),
    ("3x3_cout96", 1, 20, 20, [128], 96, 3, "none", False, "same", None),
    ("3x3_single_pixel_rows", 1, 1, 40, [64], 64, 3, "none", False, "same", None),
    ("3x3_c128_tall_ragged", 1, 70, 19, [128], 128, 3, "relu", True, "same", None),
    ("3x3_c128_o256_stride2", 1, 38, 26, [128], 256, 3, "lrelu", False, "stride2", None),
    ("1x1_c128_o128_pixshuf", 2, 9, 11, [128], 128, 1, "none", False, "pixshuf", None),
]


@pytest.fixture(params=["v2_transposed", "v1_pixel_major", "cta_pair"])
def conv_variant(request, monkeypatch):
    """Three kernels serve the dense convolutions and must pass the same cases: conv_igemm2 (channel-major accumulator,
    Cout tiles of 128; forced with EDVR_B200_CONV_V2=1), the pixel-major kernel (EDVR_B200_CONV_V1=1, every tile width)
    and the CTA-pair kernel with resident weights (conv_pair.cuh; default whenever the weights fit in shared memory,
    switched off with EDVR_B200_CONV_PAIR=0 - shapes it does not cover fall through to the automatic choice)."""
    monkeypatch.setenv("EDVR_B200_CONV_PAIR", "1" if request.param == "cta_pair" else "0")
    monkeypatch.setenv("EDVR_B200_CONV_V1", "1" if request.param == "v1_pixel_major" else "0")
    monkeypatch.setenv("EDVR_B200_CONV_V2", "1" if request.param == "v2_transposed" else "0")
    return request.param


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_vs_torch_fp32(ops, case, conv_variant):
    _, N, H, W, cins, cout, k, act, res, out_mode, maps = case
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(N if (maps is None or maps[i] is None) else maps[i][4], c, H, W, device="cuda", generator=g)
          for i, c in enumerate(cins)]
    w = torch.randn(cout, sum(cins), k, k, device="cuda", generator=g) / (sum(cins) * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    pc = ops.pack_conv(w, b)
    views = [ops.nchw_to_nhwc(x) for x in xs]
    rt = torch.randn(N, cout, H, W, device="cuda", generator=g) if res else None
    r16 = ops.nchw_to_nhwc(rt) if res else None
    torch.backends.cudnn.allow_tf32 = False

    def reference(rnd):
        """fp64 convolution of rnd(operands): rnd = fp16 rounding checks the kernel's arithmetic alone, rnd = identity
        is the distance to the reference's fp32 convolution (operand rounding included)."""
        xr = []
        for i, x in enumerate(xs):
            xh = rnd(x)
            if maps is not None and maps[i] is not None:
                div, mul, keep, add, _ = maps[i]
                xh = xh[torch.tensor([(n // div) * mul + (n % div) * keep + add for n in range(N)], device="cuda")]
            xr.append(xh)
        y = F.conv2d(torch.cat(xr, 1).double(), rnd(w).double(), b.double(), 1, k // 2).float()
        y = {"none": y, "relu": F.relu(y), "lrelu": F.leaky_relu(y, 0.1)}[act]
        if res:
            y = y + rnd(rt)
        if out_mode == "pixshuf":
            y = F.pixel_shuffle(y, 2)
        elif out_mode == "stride2":
            y = y[:, :, ::2, ::2]
        return y

    y = reference(lambda t: t.half().float())
    y_fp32 = reference(lambda t: t)
    mode = {"same": ops.OUT_SAME, "pixshuf": ops.OUT_PIXSHUF2, "stride2": ops.OUT_STRIDE2}[out_mode]
    if out_mode == "pixshuf":
        out = ops.new_act(N, 2 * H, 2 * W, cout // 4)
    elif out_mode == "stride2":
        out = ops.new_act(N, (H + 1) // 2, (W + 1) // 2, cout)
    else:
        out = ops.new_act(N, H, W, cout)
    out.t.fill_(float("nan"))
    sm = None if maps is None else [None if m is None else m[:4] for m in maps]
    a = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU}[act]
    ops.conv2d(pc, views, out16=out, act=a, res16=r16, out_mode=mode, src_maps=sm, N=N)
    got = ops.nhwc_to_nchw(out)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any()
    e = rel_err(got.cpu(), y.cpu())
    assert e[0] < TOL and e[1] < TOL, e
    e32 = rel_err(got.cpu(), y_fp32.cpu())            # vs the un-rounded fp32 convolution: the north_star bound itself
    assert e32[0] < TOL and e32[1] < TOL, ("vs fp32 operands", e32)


def test_conv_fp32_residual_stream_and_dual_output(ops, conv_variant):
    """Trunk block epilogue: out32 = res32 + conv, out16 = half(out32), in place on the fp32 stream."""
    g = torch.Generator(device="cuda").manual_seed(3)
    N, C, H, W = 2, 128, 37, 21
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 34
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    stream = torch.randn(N, H, W, C, device="cuda", generator=g)
    want = stream.permute(0, 3, 1, 2) + F.conv2d(x.half().double(), w.half().double(), b.double(), 1, 1).float()
    out16 = ops.new_act(N, H, W, C)
    ops.conv2d(ops.pack_conv(w, b), [ops.nchw_to_nhwc(x)], out16=out16, out32=stream, res32=stream)
    assert rel_err(stream.permute(0, 3, 1, 2).cpu(), want.cpu())[0] < 1e-5
    assert rel_err(ops.nhwc_to_nchw(out16).cpu(), want.cpu())[0] < TOL


def test_blocked_fp32_stream_roundtrip(ops):
    """The trunk's private tile-blocked fp32 layout: written by tsa_modulate, updated in place by a conv epilogue,
    un-blocked here with plain index math (ragged H, W exercise partial tiles)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    N, C, H, W = 2, 128, 37, 21
    feat, attn, add = (torch.randn(N, C, H, W, device="cuda", generator=g).half().float() for _ in range(3))
    stream = ops.Blocked32(N, H, W, C)
    o16 = ops.new_act(N, H, W, C)
    ops.tsa_modulate(ops.nchw_to_nhwc(feat), ops.nchw_to_nhwc(attn), ops.nchw_to_nhwc(add), out16=o16, out32=stream)
    want0 = feat * torch.sigmoid(attn) * 2 + add
    assert rel_err(stream.to_nhwc().permute(0, 3, 1, 2).cpu(), want0.cpu())[0] < 1e-5
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 34
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    out16 = ops.new_act(N, H, W, C)
    ops.conv2d(ops.pack_conv(w, b), [ops.nchw_to_nhwc(x)], out16=out16, out32=stream, res32=stream)
    want = stream_ref = want0 + F.conv2d(x.half().double(), w.half().double(), b.double(), 1, 1).float()
    assert rel_err(stream.to_nhwc().permute(0, 3, 1, 2).cpu(), want.cpu())[0] < 1e-5
    assert rel_err(ops.nhwc_to_nchw(out16).cpu(), want.cpu())[0] < TOL


def test_conv_pair_resident_and_streamed_weights_agree(ops, monkeypatch):
    """The CTA-pair kernel keeps the weights resident in shared memory when they fit and streams them with the activation
    stages otherwise; EDVR_B200_DBG=128 forces the streamed mode on a layer that fits, so both modes run the SAME layer:
    identical MMA order => bit-identical outputs; both within 1e-3 of an fp64 convolution on the fp16-rounded operands.
    Also covers a tile count that is not a multiple of the cluster count and an odd number of 16-pixel tile columns."""
    from edvr_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    N, C, H, W = 3, 128, 50, 41
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 34
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    pc = ops.pack_conv(w, b)
    assert L.lib().eb_conv2d_pair_supported(C, 3, pc.BN, pc.n_tiles) == 1 and pc.wpair is not None
    assert L.lib().eb_conv2d_pair_supported(256, 3, 128, 1) == 2 and L.lib().eb_conv2d_pair_supported(896, 1, 128, 2) == 2
    assert L.lib().eb_conv2d_pair_supported(96, 3, 128, 1) == 0            # channels in multiples of 64 only
    want = F.relu(F.conv2d(x.half().double(), w.half().double(), b.double(), 1, 1)).float()
    outs = []
    for dbg in ("0", "128"):
        monkeypatch.setenv("EDVR_B200_DBG", dbg)
        out = ops.new_act(N, H, W, C)
        out.t.fill_(float("nan"))
        ops.conv2d(pc, [ops.nchw_to_nhwc(x)], out16=out, act=ops.ACT_RELU)
        torch.cuda.synchronize()
        outs.append(out.t.clone())
        e = rel_err(ops.nhwc_to_nchw(out).cpu(), want.cpu())
        assert e[0] < TOL and e[1] < TOL, (dbg, e)
    assert torch.equal(outs[0], outs[1])


def test_add_base_matches_torch_bilinear(ops):
    """eb_add_base = the `out + F.interpolate(x_center, scale_factor=4, mode='bilinear', align_corners=False)` of
    edvr_arch.py:414-419 (and `+ x_center` for hr_in), applied after the tensor-core conv_last."""
    base = torch.rand(2, 5, 3, 9, 13, device="cuda")[:, 2]            # strided centre-frame view, like the engine passes it
    out = torch.randn(2, 3, 36, 52, device="cuda")
    want = out + F.interpolate(base.contiguous(), scale_factor=4, mode="bilinear", align_corners=False)
    ops.add_base(base, 5 * 3 * 9 * 13, 4, out)
    assert rel_err(out.cpu(), want.cpu())[0] < 1e-6
    out1 = torch.randn(2, 3, 9, 13, device="cuda")
    want1 = out1 + base
    ops.add_base(base, 5 * 3 * 9 * 13, 1, out1)
    assert torch.equal(out1, want1)


def test_mma_rate_probe_reaches_the_tensor_pipe_floor(ops):
    """Hardware probe used for the design decisions in DESIGN.md: a 128x128x16 MMA (and the 256x128x16 CTA-pair MMA) issue
    back to back at 64 cycles each in the kernels' no-swizzle operand layout."""
    import ctypes
    from edvr_b200 import _lib as L
    for cg, M in ((1, 128), (2, 256)):
        cyc = torch.zeros(160, dtype=torch.int64, device="cuda")
        n = ctypes.c_int(0)
        L.check(L.lib().eb_selftest_mma_rate(cg, M, 128, 0, 2048, 128, 2048 // cg, 128, 4096, 2048, L.ptr(cyc), ctypes.byref(n),
                                             L.stream_ptr()))
        torch.cuda.synchronize()
        c = cyc[:n.value].double()
        per_mma = float(c[c > 0].mean()) / 2048
        assert 60.0 < per_mma < 80.0, (cg, per_mma)


def test_stride2_conv_matches_torch_stride2(ops):
    """OUT_STRIDE2 must equal a real stride-2/pad-1 conv (edvr_arch.py:329,331)."""
    x = torch.randn(2, 64, 26, 34, device="cuda")
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    b = torch.zeros(64, device="cuda")
    out = ops.new_act(2, 13, 17, 64)
    ops.conv2d(ops.pack_conv(w, b), [ops.nchw_to_nhwc(x)], out16=out, out_mode=ops.OUT_STRIDE2)
    y = F.conv2d(x.half().double(), w.half().double(), None, 2, 1).float()
    e = rel_err(ops.nhwc_to_nchw(out).cpu(), y.cpu())
    assert e[0] < TOL, e


def test_pixel_shuffle_index_map_bit_exact(ops):
    """Identity 1x1 conv + fused PixelShuffle(2) store == nn.PixelShuffle(2) exactly (north_star)."""
    C = 128
    x = torch.randn(1, C, 12, 20, device="cuda").half().float()
    w = torch.eye(C, device="cuda").reshape(C, C, 1, 1)
    out = ops.new_act(1, 24, 40, C // 4)
    ops.conv2d(ops.pack_conv(w, None), [ops.nchw_to_nhwc(x)], out16=out, out_mode=ops.OUT_PIXSHUF2)
    assert torch.equal(ops.nhwc_to_nchw(out), F.pixel_shuffle(x, 2))


def _dcn_case(N, C, H, W, Cout, dg, seed=0, off_scale=2.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g)
    off = torch.randn(N, dg * 18, H, W, generator=g) * off_scale
    mask = torch.sigmoid(torch.randn(N, dg * 9, H, W, generator=g))
    w = (torch.rand(Cout, C, 3, 3, generator=g) * 2 - 1) / (C * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    go = torch.randn(N, Cout, H, W, generator=g)
    return x, off, mask, w, b, go


DCN_SHAPES = {"cfg1_single_layer": (1, 64, 64, 64, 64, 8), "c128_ragged": (2, 128, 19, 27, 128, 8),
              "c64_dg4_cout32": (1, 64, 9, 11, 32, 4), "tiny_1x1_image": (1, 64, 1, 1, 64, 8),
              "large_offsets": (1, 64, 20, 20, 64, 8)}


@pytest.mark.parametrize("name", list(DCN_SHAPES))
def test_mdcn_forward_vs_oracle(ops, name):
    from oracle import dcn_oracle
    shape = DCN_SHAPES[name]
    x, off, mask, w, b, _ = _dcn_case(*shape, off_scale=30.0 if name == "large_offsets" else 2.0)
    ref = dcn_oracle.forward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), b.numpy(), 1, 1, 1, 1, shape[5])
    got = ops.mdcn_forward(x.cuda(), off.cuda(), mask.cuda(), w.cuda(), b.cuda(), 1, 1, 1, 1, shape[5])
    e = rel_err(got.cpu(), ref)
    assert e[0] < TOL and e[1] < TOL, e


def test_mdcn_forward_stride_dilation_nobias(ops):
    from oracle import dcn_oracle
    N, C, H, W, Cout, dg, stride, pad, dil = 1, 64, 17, 21, 64, 8, 2, 2, 2
    Ho, Wo = dcn_oracle.out_hw(H, W, 3, 3, stride, pad, dil)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, C, H, W, generator=g)
    off = torch.randn(N, dg * 18, Ho, Wo, generator=g) * 2
    mask = torch.rand(N, dg * 9, Ho, Wo, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / 24
    ref = dcn_oracle.forward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), None, stride, pad, dil, 1, dg)
    got = ops.mdcn_forward(x.cuda(), off.cuda(), mask.cuda(), w.cuda(), None, stride, pad, dil, 1, dg)
    e = rel_err(got.cpu(), ref)
    assert e[0] < TOL and e[1] < TOL, e


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dcn_ref_cuda_*.npz"))))
def test_mdcn_forward_backward_vs_reference_cuda_golden(ops, path):
    """Our kernels vs outputs recorded from the unmodified reference CUDA extension (fwd + all 5 grads)."""
    from edvr_b200.dcn import mdcn_backward
    z = np.load(path)
    N, C, H, W, Cout, dg, stride, pad, dil, groups = (int(v) for v in z["meta"])
    t = {k: torch.from_numpy(z[k]).cuda() for k in ("x", "offset", "mask", "weight", "bias", "grad_out")}
    y = ops.mdcn_forward(t["x"], t["offset"], t["mask"], t["weight"], t["bias"], stride, pad, dil, groups, dg)
    e = rel_err(y.cpu(), z["out"])
    assert e[0] < TOL and e[1] < TOL, ("out", e)
    grads = mdcn_backward(t["x"], t["offset"], t["mask"], t["weight"], t["grad_out"], True, stride, pad, dil, groups, dg)
    for name, got in zip(("grad_x", "grad_offset", "grad_mask", "grad_weight", "grad_bias"), grads):
        e = rel_err(got.cpu(), z[name])
        assert e[0] < TOL and e[1] < TOL, (name, e)


def test_mdcn_backward_vs_oracle_and_accumulates(ops):
    """Backward vs the C oracle on a ragged shape; grad_weight/grad_bias accumulate like the reference
    (deform_conv_cuda.cpp:659-671)."""
    from oracle import dcn_oracle
    from edvr_b200.dcn import mdcn_backward
    shape = (2, 64, 13, 18, 64, 8)
    x, off, mask, w, b, go = _dcn_case(*shape, seed=3)
    ref = dcn_oracle.backward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), go.numpy(), True, 1, 1, 1, 1, 8)
    got = mdcn_backward(x.cuda(), off.cuda(), mask.cuda(), w.cuda(), go.cuda(), True, 1, 1, 1, 1, 8)
    for name, g, r in zip(("gx", "goff", "gmask", "gw", "gb"), got, ref):
        e = rel_err(g.cpu(), r)
        assert e[0] < TOL and e[1] < TOL, (name, e)


@pytest.mark.parametrize("scale", [1e-6, 1e3])
def test_mdcn_backward_tiny_and_huge_grad_out(ops, scale):
    """grad_out far outside the fp16 range of the backward kernels' operands (mean-reduced losses give ~1e-6 per pixel):
    the power-of-two pre-scaling keeps all five gradients within 1e-3 of the oracle."""
    from oracle import dcn_oracle
    from edvr_b200.dcn import mdcn_backward
    shape = (1, 64, 11, 13, 64, 8)
    x, off, mask, w, b, go = _dcn_case(*shape, seed=5)
    go = go * scale
    ref = dcn_oracle.backward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), go.numpy(), True, 1, 1, 1, 1, 8)
    got = mdcn_backward(x.cuda(), off.cuda(), mask.cuda(), w.cuda(), go.cuda(), True, 1, 1, 1, 1, 8)
    for name, g, r in zip(("gx", "goff", "gmask", "gw", "gb"), got, ref):
        e = rel_err(g.cpu(), r)
        assert e[0] < TOL and e[1] < TOL, (name, scale, e)


TRAIN_CONV_CASES = [   # name, N, H, W, Cin, Cout, k, stride, act
    ("3x3_c128", 2, 20, 28, 128, 128, 3, 1, "lrelu"),
    ("3x3_c256_o128_ragged", 1, 19, 37, 256, 128, 3, 1, "none"),
    ("3x3_stride2", 2, 24, 32, 64, 64, 3, 2, "lrelu"),
    ("3x3_first_cin3", 2, 16, 24, 3, 64, 3, 1, "lrelu"),
    ("3x3_last_cout3", 1, 32, 40, 64, 3, 3, 1, "none"),
    ("3x3_conv_offset_cout216", 1, 18, 22, 128, 216, 3, 1, "none"),
    ("1x1_c640_o128", 1, 16, 20, 640, 128, 1, 1, "relu"),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("case", TRAIN_CONV_CASES, ids=[c[0] for c in TRAIN_CONV_CASES])
def test_training_conv_function_vs_torch_autograd(ops, case, dtype):
    """The training step's convolution Function (edvr_b200/train.py: tcgen05 forward, dgrad = forward kernel on flipped
    weights, split-K tcgen05 wgrad) against fp32 autograd of F.conv2d.  Bars: fp16 operands 2e-3, bf16 (8-bit mantissa,
    the dtype BASELINE cfg 5 names) 1e-2 on max-rel and rel-L2 - operand rounding alone is 2^-9 per bf16 value."""
    from edvr_b200 import train as T
    _, N, H, W, cin, cout, k, stride, act = case
    tol = 2e-3 if dtype == torch.float16 else 1e-2
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g)
    m = torch.nn.Conv2d(cin, cout, k, stride, k // 2).cuda()
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, device="cuda", generator=g) / (cin * k * k) ** 0.5)
        m.bias.copy_(torch.randn(cout, device="cuda", generator=g) * 0.1)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    go = torch.randn(N, cout, Ho, Wo, device="cuda", generator=g)
    xq = x.to(dtype).float()                                   # the activations the kernel sees
    a = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU}[act]
    f = {"none": lambda t: t, "relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.1)}[act]
    # ours
    xo = xq.permute(0, 2, 3, 1).to(dtype).contiguous().requires_grad_(True)
    yo = T.conv(xo, m, a)
    assert yo.shape == (N, Ho, Wo, cout) and yo.dtype == dtype
    yo.backward(go.permute(0, 2, 3, 1).to(dtype).contiguous())
    gw_ours, gb_ours = m.weight.grad.clone(), m.bias.grad.clone()
    m.zero_grad()
    # reference: fp32 autograd of the convolution; the activation's derivative is taken at OUR forward output - (Leaky)ReLU'
    # is discontinuous at 0, so the ~4e-5 of the pre-activations whose sign flips under 16-bit rounding would otherwise
    # contribute O(1) differences that say nothing about the dgrad / wgrad kernels (measured: 0.9 % rel-L2 from 4e-5 flips)
    xr = xq.clone().requires_grad_(True)
    pre = F.conv2d(xr, m.weight, m.bias, stride, k // 2)
    yr = f(pre)
    y_ours = yo.detach().permute(0, 3, 1, 2).float()
    slope = {"none": None, "relu": 0.0, "lrelu": 0.1}[act]
    gpre = go if slope is None else go * torch.where(y_ours > 0, 1.0, slope)
    assert float(((y_ours > 0) != (pre.detach() > 0)).float().mean()) < 5e-3        # the flipped fraction stays tiny
    pre.backward(gpre)
    gw_ref, gb_ref, gx_ref = m.weight.grad.clone(), m.bias.grad.clone(), xr.grad.clone()
    for name, got, want in (("y", yo.detach().permute(0, 3, 1, 2).float(), yr.detach()),
                            ("grad_x", xo.grad.permute(0, 3, 1, 2).float(), gx_ref),
                            ("grad_weight", gw_ours, gw_ref), ("grad_bias", gb_ours, gb_ref)):
        e = rel_err(got.cpu(), want.cpu())
        assert e[0] < tol * 2 and e[1] < tol, (name, e)        # max-rel on a heavy-tailed gradient: 2 x the L2 bar


def test_autograd_function_matches_oracle(ops):
    from oracle import dcn_oracle
    from edvr_b200.dcn import modulated_deform_conv
    shape = (1, 64, 10, 12, 64, 8)
    x, off, mask, w, b, go = _dcn_case(*shape, seed=8)
    ts = [t.cuda().requires_grad_(True) for t in (x, off, mask, w, b)]
    y = modulated_deform_conv(*ts, 1, 1, 1, 1, 8)
    y.backward(go.cuda())
    ref = dcn_oracle.backward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), go.numpy(), True, 1, 1, 1, 1, 8)
    for t, r in zip(ts, ref):
        assert rel_err(t.grad.cpu(), r)[0] < TOL
    with pytest.raises(NotImplementedError):     # reference behaviour on CPU tensors (deform_conv.py:133-134)
        modulated_deform_conv(x, off, mask, w, b, 1, 1, 1, 1, 8)


@pytest.mark.parametrize("groups,dg", [(2, 2), (2, 4), (2, 1)])
def test_weight_groups_match_oracle(ops, groups, dg):
    """Weight groups > 1 (deform_conv_cuda.cpp:536-568: one GEMM per weight group over the rows of the shared im2col matrix):
    DCNv2 forward + all five gradients through the autograd Function, and DCNv1 forward + gradients, against the C oracle
    (itself checked against torchvision's grouped deform_conv2d in tests/test_oracle.py).  dg a multiple of groups, and
    several weight groups sharing one deformable group (their offset / mask gradients add)."""
    from oracle import dcn_oracle
    from edvr_b200.dcn import deform_conv, modulated_deform_conv
    N, C, H, W, Cout = 2, 128, 11, 13, 64
    g = torch.Generator().manual_seed(21)
    x = torch.randn(N, C, H, W, generator=g)
    off = torch.randn(N, dg * 18, H, W, generator=g) * 2
    mask = torch.sigmoid(torch.randn(N, dg * 9, H, W, generator=g))
    w = (torch.rand(Cout, C // groups, 3, 3, generator=g) * 2 - 1) / (9 * C // groups) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    go = torch.randn(N, Cout, H, W, generator=g)
    ts = [t.cuda().requires_grad_(True) for t in (x, off, mask, w, b)]
    y = modulated_deform_conv(*ts, 1, 1, 1, groups, dg)
    ref = dcn_oracle.forward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), b.numpy(), 1, 1, 1, groups, dg)
    e = rel_err(y.detach().cpu(), ref)
    assert e[0] < TOL and e[1] < TOL, e
    y.backward(go.cuda())
    grads = dcn_oracle.backward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), go.numpy(), True, 1, 1, 1, groups, dg)
    for nm, t, r in zip(("gx", "goff", "gmask", "gw", "gb"), ts, grads):
        e = rel_err(t.grad.cpu(), r)
        assert e[0] < TOL and e[1] < TOL, (nm, e)
    t1 = [t.cuda().requires_grad_(True) for t in (x, off, w)]
    y1 = deform_conv(*t1, 1, 1, 1, groups, dg)
    e = rel_err(y1.detach().cpu(), dcn_oracle.forward_v1(x.numpy(), off.numpy(), w.numpy(), (1, 1), (1, 1), (1, 1), groups, dg))
    assert e[0] < TOL and e[1] < TOL, e
    y1.backward(go.cuda())
    grads1 = dcn_oracle.backward_v1(x.numpy(), off.numpy(), w.numpy(), go.numpy(), (1, 1), (1, 1), (1, 1), groups, dg)
    for nm, t, r in zip(("gx", "goff", "gw"), t1, grads1):
        e = rel_err(t.grad.cpu(), r)
        assert e[0] < TOL and e[1] < TOL, ("v1 " + nm, e)


@pytest.mark.parametrize("shape,geo", [((2, 64, 12, 14, 64, 8), (1, 1, 1)), ((1, 128, 15, 11, 96, 4), (2, 1, 1)),
                                       ((1, 64, 13, 13, 32, 1), (1, 2, 2))])
def test_half_precision_operator_entry(ops, shape, geo):
    """at::Half tensors at the B1 boundary (the reference instantiates its kernels for them: deform_conv_cuda_kernel.cu:781)
    go through eb_mdcn_forward_f16 without casts.  Checked against the C oracle fed the same fp16-rounded tensors: the
    kernel's 1e-3 plus one rounding of the result to fp16 (2^-11 relative) - bar 1.5e-3 of max |out|; fp64 tensors take the
    conversion path of the shim."""
    from oracle import dcn_oracle
    from edvr_b200 import deform_conv_ext as ext
    N, C, H, W, Cout, dg = shape
    stride, pad, dil = geo
    Ho, Wo = dcn_oracle.out_hw(H, W, 3, 3, stride, pad, dil)
    g = torch.Generator().manual_seed(17)
    x = torch.randn(N, C, H, W, generator=g).half()
    off = (torch.randn(N, dg * 18, Ho, Wo, generator=g) * 2).half()
    mask = torch.sigmoid(torch.randn(N, dg * 9, Ho, Wo, generator=g)).half()
    w = ((torch.rand(Cout, C, 3, 3, generator=g) * 2 - 1) / (9 * C) ** 0.5).half()
    b = (torch.randn(Cout, generator=g) * 0.1).half()
    ref = dcn_oracle.forward(x.float().numpy(), off.float().numpy(), mask.float().numpy(), w.float().numpy(),
                             b.float().numpy(), stride, pad, dil, 1, dg)
    out = torch.full((N, Cout, Ho, Wo), float("nan"), dtype=torch.float16, device="cuda")
    launches = ops.LAUNCHES[0]
    ext.modulated_deform_conv_forward(x.cuda(), w.cuda(), b.cuda(), None, off.cuda(), mask.cuda(), out, None, 3, 3, stride,
                                      stride, pad, pad, dil, dil, 1, dg, True)
    assert out.dtype == torch.float16 and not torch.isnan(out).any()
    e = rel_err(out.float().cpu(), ref)
    assert e[0] < 1.5e-3 and e[1] < 1.5e-3, e
    out64 = torch.empty(N, Cout, Ho, Wo, dtype=torch.float64, device="cuda")
    ext.modulated_deform_conv_forward(x.double().cuda(), w.double().cuda(), b.double().cuda(), None, off.double().cuda(),
                                      mask.double().cuda(), out64, None, 3, 3, stride, stride, pad, pad, dil, dil, 1, dg, True)
    assert rel_err(out64.float().cpu(), ref)[0] < TOL


# ---- DCNv1: deform_conv_forward / backward_input / backward_parameters -------------------------
V1_GPU_CASES = {  # N, C, H, W, Cout, dg, (kh, kw), stride, padding, dilation, offset scale
    "iso_dg8": (2, 64, 12, 14, 64, 8, (3, 3), (1, 1), (1, 1), (1, 1), 2.0),
    "aniso_dg1": (2, 64, 15, 13, 32, 1, (3, 3), (2, 1), (1, 2), (1, 2), 3.0),
    "k1x3_c128": (1, 128, 9, 11, 128, 4, (1, 3), (1, 1), (0, 1), (1, 1), 1.5),
    "ragged_big_offsets": (3, 64, 21, 37, 64, 8, (3, 3), (1, 1), (1, 1), (1, 1), 25.0),
}


def _v1_case(N, C, H, W, Cout, dg, k, s, p, d, osc, seed=11):
    g = torch.Generator().manual_seed(seed)
    Ho = (H + 2 * p[0] - (d[0] * (k[0] - 1) + 1)) // s[0] + 1
    Wo = (W + 2 * p[1] - (d[1] * (k[1] - 1) + 1)) // s[1] + 1
    x = torch.randn(N, C, H, W, generator=g)
    off = torch.randn(N, dg * 2 * k[0] * k[1], Ho, Wo, generator=g) * osc
    w = (torch.rand(Cout, C, *k, generator=g) * 2 - 1) / (C * k[0] * k[1]) ** 0.5
    go = torch.randn(N, Cout, Ho, Wo, generator=g)
    return x, off, w, go


@pytest.mark.parametrize("name", list(V1_GPU_CASES))
def test_dcn1_autograd_vs_oracle(ops, name):
    """deform_conv (autograd Function over the three eb_dcn1_* entry points) vs the C oracle; tolerance 1e-3."""
    from oracle import dcn_oracle
    from edvr_b200.dcn import deform_conv
    N, C, H, W, Cout, dg, k, s, p, d, osc = V1_GPU_CASES[name]
    x, off, w, go = _v1_case(*V1_GPU_CASES[name])
    ts = [t.cuda().requires_grad_(True) for t in (x, off, w)]
    y = deform_conv(*ts, s, p, d, 1, dg)
    ref = dcn_oracle.forward_v1(x.numpy(), off.numpy(), w.numpy(), s, p, d, 1, dg)
    e = rel_err(y.detach().cpu(), ref)
    assert y.shape == ref.shape and e[0] < TOL and e[1] < TOL, e
    y.backward(go.cuda())
    grads = dcn_oracle.backward_v1(x.numpy(), off.numpy(), w.numpy(), go.numpy(), s, p, d, 1, dg)
    for nm, t, r in zip(("gx", "goff", "gw"), ts, grads):
        e = rel_err(t.grad.cpu(), r)
        assert e[0] < TOL and e[1] < TOL, (nm, e)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dcn1_ref_cuda_*.npz"))))
def test_dcn1_ext_vs_reference_cuda_golden(ops, path):
    """The B1 shim's three v1 functions, called with the reference's argument order (kW before kH ...), against
    outputs recorded from the unmodified reference CUDA extension; grad_weight accumulates scale * dW."""
    from edvr_b200 import deform_conv_ext as ext
    z = np.load(path)
    N, C, H, W, Cout, dg, kh, kw, sh, sw, ph, pw, dh, dw = (int(v) for v in z["meta"])
    scale = float(z["scale"])
    x, off, w, go = (torch.from_numpy(z[k]).cuda() for k in ("x", "offset", "weight", "grad_out"))
    geom = (kw, kh, sw, sh, pw, ph, dw, dh, 1, dg)
    out = x.new_empty(0)
    assert ext.deform_conv_forward(x, w, off, out, x.new_empty(0), x.new_empty(0), *geom, N) == 1
    e = rel_err(out.cpu(), z["out"])
    assert out.shape == z["out"].shape and e[0] < TOL and e[1] < TOL, ("out", e)
    gx, goff = torch.full_like(x, 7.0), torch.full_like(off, 7.0)        # overwritten, not accumulated
    ext.deform_conv_backward_input(x, off, go, gx, goff, w, x.new_empty(0), *geom, N)
    gw = torch.ones_like(w)                                              # accumulated into
    ext.deform_conv_backward_parameters(x, off, go, gw, x.new_empty(0), x.new_empty(0), *geom, scale, N)
    for nm, got, ref in (("grad_x", gx, z["grad_x"]), ("grad_offset", goff, z["grad_offset"]),
                         ("grad_weight", gw - 1.0, z["grad_weight"])):
        e = rel_err(got.cpu(), ref)
        assert e[0] < TOL and e[1] < TOL, (nm, e)


def test_dcn1_modules_and_error_behaviour(ops):
    """DeformConvPack: zero-initialised conv_offset => equals a plain convolution; DeformConv pads inputs smaller than
    the kernel (deform_conv.py:232-247); reference error behaviour for bad arguments."""
    from edvr_b200 import deform_conv_ext as ext
    from edvr_b200.dcn import DeformConv, DeformConvPack, deform_conv
    torch.manual_seed(0)
    m = DeformConvPack(64, 64, 3, stride=1, padding=1, deformable_groups=8).cuda()
    assert sorted(k for k, _ in m.named_parameters()) == ["conv_offset.bias", "conv_offset.weight", "weight"]
    x = torch.randn(2, 64, 10, 13, device="cuda")
    e = rel_err(m(x).detach().cpu(), F.conv2d(x, m.weight, None, padding=1).detach().cpu())
    assert e[0] < TOL and e[1] < TOL, e
    m(x).sum().backward()
    assert m.weight.grad is not None and m.conv_offset.weight.grad is not None
    small = DeformConv(64, 64, 3, padding=1, deformable_groups=1).cuda()
    y = small(torch.randn(1, 64, 2, 2, device="cuda"), torch.zeros(1, 18, 2, 2, device="cuda"))
    assert y.shape == (1, 64, 2, 2)
    with pytest.raises(NotImplementedError):                             # CPU tensors (deform_conv.py:43-44)
        deform_conv(torch.zeros(1, 64, 4, 4), torch.zeros(1, 18, 4, 4), torch.zeros(64, 64, 3, 3), 1, 1, 1, 1, 1)
    with pytest.raises(ValueError, match="Expected 4D tensor"):
        deform_conv(torch.zeros(64, 4, 4, device="cuda"), torch.zeros(1, 18, 4, 4, device="cuda"),
                    torch.zeros(64, 64, 3, 3, device="cuda"))
    with pytest.raises(AssertionError, match="im2col step must divide batchsize"):
        deform_conv(torch.zeros(3, 64, 4, 4, device="cuda"), torch.zeros(3, 18, 4, 4, device="cuda"),
                    torch.zeros(64, 64, 3, 3, device="cuda"), 1, 1, 1, 1, 1, 2)
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError, match="invalid number of channels of offset"):
        ext.deform_conv_forward(z(1, 64, 4, 4), z(64, 64, 3, 3), z(1, 20, 4, 4), z(0), z(0), z(0), 3, 3, 1, 1, 1, 1, 1, 1,
                                1, 1, 1)
    with pytest.raises(RuntimeError, match="invalid spatial size of offset"):
        ext.deform_conv_forward(z(1, 64, 4, 4), z(64, 64, 3, 3), z(1, 18, 5, 4), z(0), z(0), z(0), 3, 3, 1, 1, 1, 1, 1, 1,
                                1, 1, 1)
    with pytest.raises(RuntimeError, match="not implemented on CPU"):
        ext.deform_conv_forward(torch.zeros(1, 64, 4, 4), torch.zeros(64, 64, 3, 3), torch.zeros(1, 18, 4, 4),
                                torch.zeros(0), torch.zeros(0), torch.zeros(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1)
    # unbatched (3-D) call, deform_conv_cuda.cpp:176-183,230-234
    x3, w3, o3 = torch.randn(64, 6, 7, device="cuda"), torch.randn(64, 64, 3, 3, device="cuda") / 24, z(18, 6, 7)
    out3 = z(0)
    ext.deform_conv_forward(x3, w3, o3, out3, z(0), z(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1)
    e = rel_err(out3.cpu(), F.conv2d(x3[None], w3, None, padding=1)[0].cpu())
    assert out3.shape == (64, 6, 7) and e[0] < TOL, e


def test_error_codes_not_printf(ops):
    from edvr_b200 import _lib as L
    x = torch.zeros(1, 60, 8, 8, device="cuda")          # C % 64 != 0 -> unsupported, reported loudly
    with pytest.raises(RuntimeError, match="status -2"):
        ops.mdcn_forward(x, torch.zeros(1, 72, 8, 8, device="cuda"), torch.zeros(1, 36, 8, 8, device="cuda"),
                         torch.zeros(64, 60, 3, 3, device="cuda"), None, 1, 1, 1, 1, 4)
    rc = L.lib().eb_mdcn_forward(None, None, None, None, None, None, 1, 64, 8, 8, 64, 3, 3, 1, 1, 1, 1, 8, None, 0, None)
    assert rc == -5 and b"null" in L.lib().eb_last_error()


def test_empty_batch_is_noop(ops):
    y = ops.mdcn_forward(torch.zeros(0, 64, 8, 8, device="cuda"), torch.zeros(0, 144, 8, 8, device="cuda"),
                         torch.zeros(0, 72, 8, 8, device="cuda"), torch.zeros(64, 64, 3, 3, device="cuda"), None,
                         1, 1, 1, 1, 8)
    assert y.shape == (0, 64, 8, 8)


@pytest.mark.parametrize("sigma", [0.02, 3.0, 10.0])
@pytest.mark.parametrize("mode,shape", [("pair", (2, 128, 18, 23, 8)),      # EDVR-L geometry, two channel chunks per offset half
                                        ("pair", (1, 64, 16, 24, 8)),       # EDVR-M: 8 channels per group, odd tile count (dead tile)
                                        ("pair", (3, 64, 21, 9, 4)),        # 4 groups, ragged tiles
                                        ("fused", (2, 128, 18, 23, 8))])    # single-CTA form (odd dg / several output tiles)
def test_dcn_site_fused_path_matches_oracle(ops, sigma, mode, shape):
    """Production DCN site (ops.DcnSite: conv_offset -> offsets + sigmoid(mask) -> deformable conv) at sampling offsets of
    ~N(0, sigma^2) pixels - near zero like a fresh model, and multi-pixel like a trained one (the reference warns at a mean
    of 50, arch_util.py:249-253).  The oracle gets the offsets the kernel computes (fp16-rounded conv_offset operands, exact
    accumulation), so the bound covers everything from the offset record on: 1e-3 (north_star)."""
    from oracle import dcn_oracle
    N, C, H, W, dg = shape
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, C, H, W, generator=g)
    feat = torch.randn(N, C, H, W, generator=g)
    wo = torch.randn(dg * 27, C, 3, 3, generator=g) / (C * 9) ** 0.5
    wo[:dg * 18] *= sigma
    bo = torch.randn(dg * 27, generator=g) * 0.5
    w = (torch.rand(C, C, 3, 3, generator=g) * 2 - 1) / (C * 9) ** 0.5
    b = torch.randn(C, generator=g) * 0.1
    raw = F.conv2d(feat.half().double(), wo.half().double(), bo.double(), padding=1).float()
    off, mask = raw[:, :dg * 18].contiguous(), torch.sigmoid(raw[:, dg * 18:]).contiguous()
    ref = dcn_oracle.forward(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    site = ops.DcnSite(wo.cuda(), bo.cuda(), w.cuda(), b.cuda(), dg, mode=mode)
    assert site.mode == mode
    acc = torch.zeros(1, device="cuda")
    out = ops.new_act(N, H, W, C)
    site(ops.nchw_to_nhwc(x.cuda()), ops.nchw_to_nhwc(feat.cuda()), out, absmean=acc)
    e = rel_err(ops.nhwc_to_nchw(out).cpu(), ref)
    assert e[0] < TOL and e[1] < TOL, (sigma, e)
    mean = float(acc.item()) / off.numel()
    assert abs(mean - float(off.abs().mean())) < 1e-2 * float(off.abs().mean())


def test_offset_warning_is_deferred_not_dropped(ops, caplog):
    """arch_util.py:249-253: `Offset abs mean is X, larger than 50.` - here accumulated on the device and emitted by the next
    DCNv2Pack call (no host sync inside the call), in the reference's wording, on the reference's logger."""
    import logging
    from edvr_b200.dcn import DCNv2Pack, _MONITOR
    m = DCNv2Pack(64, 64, 3, padding=1, deformable_groups=8).cuda().eval()
    with torch.no_grad():
        m.conv_offset.bias[:144] = 80.0                  # every offset = 80 px
    x = torch.randn(1, 64, 12, 16, device="cuda")
    with torch.no_grad(), caplog.at_level(logging.WARNING, logger="basicsr"):
        m(x, x)
        assert not [r for r in caplog.records if "larger than 50" in r.message]     # nothing read back yet, no sync
        torch.cuda.synchronize()
        m(x, x)                                          # the next call reports the previous one
    msgs = [r.message for r in caplog.records if "larger than 50" in r.message]
    assert msgs and msgs[0].startswith("Offset abs mean is 80.0")
    assert abs(m.check_offset_absmean() - 80.0) < 1e-3
    _MONITOR.poll(block=True)


def test_elementwise_stages_vs_torch(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 64, 10, 14, device="cuda", generator=g).half().float()
    v = ops.nchw_to_nhwc(x)
    # bilinear x2 (* 2) into a channel slice
    dst = ops.new_act(2, 20, 28, 128)
    dst.t.zero_()
    ops.upsample2x(v, dst.slice(64, 64), mul=2.0)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) * 2
    assert rel_err(ops.nhwc_to_nchw(dst.slice(64, 64)).cpu(), ref.cpu())[0] < TOL
    assert float(dst.t[..., :64].abs().max()) == 0.0
    # max + avg pool (count_include_pad), odd sizes
    xo = torch.randn(1, 64, 9, 13, device="cuda", generator=g).half().float()
    pd = ops.new_act(1, 5, 7, 128)
    ops.pool_max_avg(ops.nchw_to_nhwc(xo), pd)
    got = ops.nhwc_to_nchw(pd)
    assert rel_err(got[:, :64].cpu(), F.max_pool2d(xo, 3, 2, 1).cpu())[0] < 1e-6
    assert rel_err(got[:, 64:].cpu(), F.avg_pool2d(xo, 3, 2, 1).cpu())[0] < TOL


def test_tsa_temporal_and_modulate_vs_torch(ops):
    B, T, C, H, W = 2, 3, 64, 6, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    emb = torch.randn(B * T, C, H, W, device="cuda", generator=g).half().float() * 0.3
    ref_e = torch.randn(B, C, H, W, device="cuda", generator=g).half().float() * 0.3
    al = torch.randn(B * T, C, H, W, device="cuda", generator=g).half().float()
    dst = ops.new_act(B, H, W, T * C)
    ops.tsa_temporal(ops.nchw_to_nhwc(emb), ops.nchw_to_nhwc(ref_e), ops.nchw_to_nhwc(al), dst, B, T)
    prob = torch.sigmoid((emb.view(B, T, C, H, W) * ref_e.unsqueeze(1)).sum(2))
    want = (al.view(B, T, C, H, W) * prob.unsqueeze(2)).reshape(B, T * C, H, W)
    assert rel_err(ops.nhwc_to_nchw(dst).cpu(), want.cpu())[0] < TOL
    feat, attn, add = (torch.randn(B, C, H, W, device="cuda", generator=g).half().float() for _ in range(3))
    o16 = ops.new_act(B, H, W, C)
    o32 = torch.empty(B, H, W, C, device="cuda")
    ops.tsa_modulate(ops.nchw_to_nhwc(feat), ops.nchw_to_nhwc(attn), ops.nchw_to_nhwc(add), out16=o16, out32=o32)
    want = feat * torch.sigmoid(attn) * 2 + add
    assert rel_err(o32.permute(0, 3, 1, 2).cpu(), want.cpu())[0] < 1e-5
    assert rel_err(ops.nhwc_to_nchw(o16).cpu(), want.cpu())[0] < TOL


def test_conv_first_and_last_vs_torch(ops):
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.rand(2, 3, 12, 16, device="cuda", generator=g)
    w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    out = ops.new_act(2, 12, 16, 64)
    ops.conv_first(x, w, b, out)
    torch.backends.cudnn.allow_tf32 = False
    want = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.1)
    assert rel_err(ops.nhwc_to_nchw(out).cpu(), want.cpu())[0] < TOL
    # conv_last + bilinear x4 base
    hr = torch.randn(2, 64, 48, 64, device="cuda", generator=g).half().float()
    wl = torch.randn(3, 64, 3, 3, device="cuda", generator=g) * 0.05
    bl = torch.randn(3, device="cuda", generator=g) * 0.1
    got = torch.empty(2, 3, 48, 64, device="cuda")
    ops.conv_last(ops.nchw_to_nhwc(hr), wl, bl, x, 3 * 12 * 16, 4, got)
    want = F.conv2d(hr, wl, bl, padding=1) + F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)
    assert rel_err(got.cpu(), want.cpu())[0] < 1e-4


def test_b1_extension_shim_signature_and_inplace_semantics(ops, golden_dir):
    """edvr_b200.deform_conv_ext mirrors deform_conv_ext.cpp:106-146: caller-allocated outputs written in place,
    grad_weight / grad_bias accumulated into."""
    from edvr_b200 import deform_conv_ext as ext
    z = np.load(os.path.join(golden_dir, "dcn_ref_cuda_g_c64_dg8.npz"))
    N, C, H, W, Cout, dg, stride, pad, dil, groups = (int(v) for v in z["meta"])
    t = {k: torch.from_numpy(z[k]).cuda() for k in ("x", "offset", "mask", "weight", "bias", "grad_out")}
    out = torch.empty(N, Cout, H, W, device="cuda")
    e = torch.empty(0, device="cuda")
    ext.modulated_deform_conv_forward(t["x"], t["weight"], t["bias"], e, t["offset"], t["mask"], out, e, 3, 3, stride,
                                      stride, pad, pad, dil, dil, groups, dg, True)
    assert rel_err(out.cpu(), z["out"])[0] < TOL
    gx, goff, gm = torch.zeros_like(t["x"]), torch.zeros_like(t["offset"]), torch.zeros_like(t["mask"])
    gw, gb = torch.ones_like(t["weight"]), torch.ones_like(t["bias"])          # pre-filled: must be accumulated into
    ext.modulated_deform_conv_backward(t["x"], t["weight"], t["bias"], e, t["offset"], t["mask"], e, gx, gw, gb, goff,
                                       gm, t["grad_out"], 3, 3, stride, stride, pad, pad, dil, dil, groups, dg, True)
    assert rel_err(gx.cpu(), z["grad_x"])[0] < TOL
    assert rel_err(goff.cpu(), z["grad_offset"])[0] < TOL
    assert rel_err((gw - 1).cpu(), z["grad_weight"])[0] < TOL
    assert rel_err((gb - 1).cpu(), z["grad_bias"])[0] < TOL
    with pytest.raises(RuntimeError, match="not implemented on CPU"):
        ext.modulated_deform_conv_forward(t["x"].cpu(), t["weight"].cpu(), t["bias"].cpu(), e, t["offset"].cpu(),
                                          t["mask"].cpu(), out.cpu(), e, 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
    assert all(callable(getattr(ext, n)) for n in ("deform_conv_forward", "deform_conv_backward_input",
                                                    "deform_conv_backward_parameters"))   # deform_conv_ext.cpp:149-163


# ---------------------------------------------------------------- full-size, size-independent properties
def test_dcn_full_size_zero_offset_unit_mask_equals_dense_conv(ops):
    """BASELINE L1 size (1x128x180x320): with zero offsets and mask == 1 the DCN is a plain 3x3 convolution
    (deform_conv_cuda_kernel.cu:608-628 with dh = dw = 0) - checked against fp32 cuDNN and against our own dense kernel."""
    g = torch.Generator(device="cuda").manual_seed(11)
    N, C, H, W, dg = 1, 128, 180, 320, 8
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 34
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    off = torch.zeros(N, dg * 18, H, W, device="cuda")
    mask = torch.ones(N, dg * 9, H, W, device="cuda")
    y = ops.mdcn_forward(x, off, mask, w, b, 1, 1, 1, 1, dg)
    torch.backends.cudnn.allow_tf32 = False
    want = F.conv2d(x.half().float(), w.half().float(), b, padding=1)
    e = rel_err(y.cpu(), want.cpu())
    assert e[0] < TOL and e[1] < TOL, e
    dense = ops.new_act(N, H, W, C)
    ops.conv2d(ops.pack_conv(w, b), [ops.nchw_to_nhwc(x)], out16=dense)
    assert rel_err(ops.nhwc_to_nchw(dense).cpu(), y.cpu())[0] < TOL


def test_dcn_full_size_linearity_and_determinism(ops):
    """At the L1 size: linear in x (and in the weights) for fixed offsets/mask, and bit-identical across runs
    (no atomics in the forward path)."""
    g = torch.Generator(device="cuda").manual_seed(12)
    N, C, H, W, dg = 1, 128, 180, 320, 8
    x1, x2 = (torch.randn(N, C, H, W, device="cuda", generator=g).half().float() for _ in range(2))
    off = torch.randn(N, dg * 18, H, W, device="cuda", generator=g) * 3
    mask = torch.sigmoid(torch.randn(N, dg * 9, H, W, device="cuda", generator=g))
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 34
    f = lambda xx, ww: ops.mdcn_forward(xx, off, mask, ww, None, 1, 1, 1, 1, dg)
    y1, y2, y12 = f(x1, w), f(x2, w), f(x1 + x2, w)
    assert rel_err((y1 + y2).cpu(), y12.cpu())[0] < 2e-3
    assert rel_err((2 * y1).cpu(), f(x1, 2 * w).cpu())[0] < 1e-6        # exact power-of-two scaling
    assert torch.equal(f(x1, w), y1)


def test_edvr_full_size_shape_determinism_and_batch_consistency():
    """cfg 3 at full size (7x3x180x320 -> 720x1280): output shape, run-to-run bit equality, and every clip of a
    batch equals the same clip run alone (clips are independent units)."""
    from edvr_b200.engine import EDVREngine
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=128, num_frame=7, num_reconstruct_block=40, seed=0)
    eng = EDVREngine(sd, num_frame=7)
    x = torch.rand(2, 7, 3, 180, 320, generator=torch.Generator().manual_seed(0)).cuda()
    y = eng.forward(x).clone()
    assert y.shape == (2, 3, 720, 1280) and torch.isfinite(y).all()
    assert torch.equal(eng.forward(x), y)
    y0 = eng.forward(x[:1].contiguous())
    assert torch.equal(y0[0], y[0])
    assert float(eng.offset_absmeans().max()) < 50


# ---- frame staging either side of the network: edvr_b200/img.py, bit-exact ------------------------------
def test_frame_staging_matches_reference_golden_bit_exact(ops):
    """eb_frames_u8_to_f32 / eb_tensor2img_u8 against the fixture recorded from the unmodified reference helpers
    (tests/golden/img_ref_import.npz, oracle/make_golden_img.py) and against oracle/img_ref.py: byte / fp32-bit equality,
    incl. values outside [0, 1], products on k + 0.5 (round half to even), gray images, min_max = (-1, 1), RGB kept."""
    from edvr_b200 import frames_to_tensor, tensor2img
    from oracle import img_ref
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "img_ref_import.npz"))
    x = frames_to_tensor(torch.from_numpy(z["frames"]).cuda())
    assert np.array_equal(x.cpu().numpy().view(np.uint32), z["x"].view(np.uint32))
    out = torch.from_numpy(z["out"]).cuda()
    assert np.array_equal(tensor2img([out]), z["y"])
    assert np.array_equal(tensor2img(out, rgb2bgr=False), z["y_rgb"])
    g = tensor2img(out[:, :1])
    assert g.shape == z["gray"].shape and np.array_equal(g, z["gray"])
    assert np.array_equal(tensor2img(out * 2 - 1, min_max=(-1, 1)), z["y_pm1"])
    assert np.array_equal(tensor2img(out[0, 0]), img_ref.tensor2img(z["out"][0, 0]))
    both = tensor2img([out, out[:, :1]])
    assert isinstance(both, list) and np.array_equal(both[0], z["y"]) and np.array_equal(both[1], z["gray"])
    f = tensor2img(out, out_type=np.float32)
    assert f.dtype == np.float32 and f.shape == (12, 18, 3) and float(f.min()) >= 0.0 and float(f.max()) <= 1.0
    with pytest.raises(NotImplementedError):
        tensor2img(out.cpu())
    with pytest.raises(NotImplementedError):
        tensor2img(torch.zeros(2, 3, 4, 4, device="cuda"))
    with pytest.raises(TypeError):
        tensor2img("not a tensor")


def test_frame_staging_full_size_round_trip_and_files(ops, tmp_path):
    """Size-independent properties at the BASELINE frame sizes: every byte survives frames -> [0, 1] floats -> bytes
    (7 LR frames 180x320 and one HR frame 720x1280, ragged sizes too), the kernels agree with the oracle on them, and
    read_img_seq on PNG files equals the oracle applied to the decoded frames."""
    import cv2
    from edvr_b200 import frames_to_tensor, read_img_seq, tensor2img
    from oracle import img_ref
    rng = np.random.default_rng(1)
    for shape in ((7, 180, 320, 3), (1, 720, 1280, 3), (2, 37, 53, 3)):
        frames = rng.integers(0, 256, size=shape, dtype=np.uint8)
        x = frames_to_tensor(torch.from_numpy(frames).cuda())
        assert np.array_equal(x.cpu().numpy().view(np.uint32), img_ref.frames_to_tensor(frames).view(np.uint32))
        for i in range(shape[0]):
            assert np.array_equal(tensor2img(x[i]), frames[i])
    hr = torch.randn(1, 3, 720, 1280, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) * 0.4 + 0.5
    assert np.array_equal(tensor2img(hr), img_ref.tensor2img(hr.cpu().numpy()))
    frames = rng.integers(0, 256, size=(3, 21, 34, 3), dtype=np.uint8)
    paths = []
    for i, f in enumerate(frames):
        paths.append(str(tmp_path / f"{i:08d}.png"))
        assert cv2.imwrite(paths[-1], f)
    want = img_ref.frames_to_tensor(frames)
    for arg in (paths, str(tmp_path)):
        got = read_img_seq(arg)
        assert got.is_cuda and np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    got = read_img_seq(paths, require_mod_crop=True, scale=4)
    assert np.array_equal(got.cpu().numpy(), want[:, :, :20, :32])
