"""GPU end-to-end parity (-m gpu): the fused EDVR executor vs the oracle graph (oracle/edvr_ref.py, the
bit-exact port of the reference's edvr_arch.py) on the same weights and inputs, plus the golden outputs
recorded from the imported reference itself.

Metrics and bars.  (1) max|out - ref| / max|ref| < 1e-3: north_star's bound, on the tensor the user receives.
(2) rel-L2 of the network's own contribution r = out - base (base = bilinear x4 of the centre frame, identical in both
paths; |base| ~ 1 would otherwise hide everything) < 5e-3: ~100 layers of fp16-operand / fp32-accumulate convolutions
each add ~2-4e-4 of independent rounding noise (the same 10-bit operand mantissa as the TF32 path of the reference's
cuDNN convs), which accumulates to a few 1e-3 on the residual alone; a single wrong tap, mask or offset shows up at
>= 1e-1 there.  Measured values are printed and recorded in DESIGN.md (1e-5 .. 1e-4 on (1), 1e-4 .. 2e-3 on (2)).
Offsets: `offset_std` scales the random conv_offset init, so the sampling offsets range from ~0.02 px (fresh model) to
several pixels (trained model; the reference warns at a mean of 50 px, arch_util.py:249-253).
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _compare(sd, x, hr_in=False, tol_l2=5e-3, tol_max=1e-3, center=None, dcn=None):
    from edvr_b200.engine import EDVREngine
    from oracle import edvr_ref
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    kw = {} if dcn is None else {"dcn": dcn}
    ref = edvr_ref.edvr_forward(sdc, x.cuda(), hr_in=hr_in, **kw)     # fp32 oracle graph (torchvision DCN by default)
    eng = EDVREngine(sd, num_frame=x.shape[1], hr_in=hr_in)
    out = eng.forward(x.cuda())
    torch.cuda.synchronize()
    xc = x[:, x.shape[1] // 2].cuda()
    base = xc if hr_in else F.interpolate(xc, scale_factor=4, mode="bilinear", align_corners=False)
    r_ref, r_out = (ref - base).cpu(), (out - base).cpu()
    l2 = float((r_out - r_ref).norm() / r_ref.norm())
    mx = float((out - ref).abs().max() / ref.abs().max())
    print(f"edvr parity: rel-L2(residual)={l2:.3e}  max|d|/max|out|={mx:.3e}  absmeans={eng.offset_absmeans()}")
    assert out.shape == ref.shape
    assert l2 < tol_l2 and mx < tol_max, (l2, mx)
    return out, ref


def test_edvr_m_small_vs_oracle_graph():
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=64, num_frame=5, num_extract_block=5, num_reconstruct_block=10, seed=0)
    x = torch.rand(1, 5, 3, 32, 48, generator=torch.Generator().manual_seed(0))
    _compare(sd, x)


def test_edvr_batch2_ragged_tiles_vs_oracle_graph():
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=64, num_frame=3, num_extract_block=2, num_reconstruct_block=3, seed=1)
    x = torch.rand(2, 3, 3, 20, 36, generator=torch.Generator().manual_seed(1))
    _compare(sd, x)


def test_edvr_l_width_vs_oracle_graph():
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=128, num_frame=7, num_extract_block=5, num_reconstruct_block=40, seed=2)
    x = torch.rand(1, 7, 3, 36, 64, generator=torch.Generator().manual_seed(2))
    _compare(sd, x)


@pytest.mark.parametrize("offset_std", [0.5, 3.0])
def test_edvr_multi_pixel_offsets_vs_oracle_graph(offset_std):
    """Sampling offsets of a trained model's magnitude (mean |offset| printed: ~1 px and ~5 px at L1)."""
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=64, num_frame=5, num_extract_block=2, num_reconstruct_block=4, seed=6,
                                  offset_std=offset_std)
    x = torch.rand(1, 5, 3, 40, 56, generator=torch.Generator().manual_seed(6))
    _compare(sd, x)


def _reference_ext_dcn():
    """B2-signature DCN running the UNMODIFIED reference CUDA extension (oracle/_ref), or None if it was not built."""
    from oracle import build_ref
    if not os.path.exists(build_ref.so_path()):
        return None
    ext = build_ref.load_ref()

    def dcn(x, off, mask, w, b, s, p, d, g, dg):
        x = x.contiguous()
        out = x.new_empty(x.shape[0], w.shape[0], x.shape[2], x.shape[3])
        ext.modulated_deform_conv_forward(x, w, b, x.new_empty(0), off, mask, out, x.new_empty(0), 3, 3, s, s, p, p,
                                          d, d, g, dg, True)
        return out
    return dcn


@pytest.mark.parametrize("offset_std", [0.02, 1.0])
def test_edvr_l_full_size_vs_reference_cuda_ext_live(offset_std):
    """BASELINE cfg 3 at FULL size (EDVR-L, 2 clips of 7x3x180x320 -> 3x720x1280): the fused executor vs the oracle graph
    with the reference's own CUDA dcn extension (compiled unmodified into oracle/_ref) running live on the same GPU, fp32,
    TF32 off."""
    from oracle import edvr_ref
    dcn = _reference_ext_dcn()
    if dcn is None:
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    sd = edvr_ref.make_state_dict(num_feat=128, num_frame=7, num_extract_block=5, num_reconstruct_block=40, seed=9,
                                  offset_std=offset_std)
    x = torch.rand(2, 7, 3, 180, 320, generator=torch.Generator().manual_seed(9))
    _compare(sd, x, dcn=dcn)


def test_edvr_no_tsa_vs_oracle_graph():
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=64, num_frame=3, num_extract_block=1, num_reconstruct_block=2,
                                  with_tsa=False, seed=3)
    x = torch.rand(1, 3, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    _compare(sd, x)


def test_edvr_predeblur_hr_in_vs_oracle_graph():
    from oracle import edvr_ref
    sd = edvr_ref.make_state_dict(num_feat=64, num_frame=3, num_extract_block=1, num_reconstruct_block=2,
                                  with_predeblur=True, hr_in=True, seed=4)
    x = torch.rand(1, 3, 3, 64, 96, generator=torch.Generator().manual_seed(4))
    _compare(sd, x, hr_in=True)


def test_edvr_vs_reference_import_golden(golden_dir):
    """Same weights/input as the fixture recorded from the imported reference graph (nf=64 case)."""
    from oracle import edvr_ref
    z = np.load(os.path.join(golden_dir, "edvr_ref_import_nf64.npz"), allow_pickle=True)
    kw = z["kwargs"].item()
    sd = edvr_ref.make_state_dict(**kw, seed=int(z["seed"]))
    x = torch.from_numpy(z["x"])
    out, _ = _compare(sd, x)
    want = torch.from_numpy(z["y"])
    assert float((out.cpu() - want).abs().max() / want.abs().max()) < 1e-3


def test_drop_in_modules_state_dict_and_forward():
    """B3 boundary: same class names / ctor / state_dict keys as the reference; forward through the engine."""
    from edvr_b200.edvr import EDVR
    from oracle import edvr_ref
    kw = dict(num_feat=64, num_frame=3, deformable_groups=8, num_extract_block=1, num_reconstruct_block=2)
    sd = edvr_ref.make_state_dict(**kw, seed=5)
    net = EDVR(center_frame_idx=None, **kw).cuda().eval()
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd, strict=True)
    x = torch.rand(1, 3, 3, 16, 24, generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        y = net(x)
    ref = edvr_ref.edvr_forward({k: v.cuda() for k, v in sd.items()}, x)
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-3
    with pytest.raises(AssertionError, match="multiple of 4"):
        net(torch.rand(1, 3, 3, 18, 24).cuda())


def test_training_step_gradients_vs_oracle_graph():
    """BASELINE cfg 5 (reduced): one Charbonnier-loss training step through the drop-in modules with autograd ON.
    The graph runs differentiable PyTorch ops around OUR DCN forward + backward kernels (edvr_b200.dcn); gradients
    are compared with the oracle graph differentiated through torchvision's deform_conv2d (fp32, TF32 off).
    Offsets are kept away from the -1 edge by the random conv_offset init."""
    from edvr_b200.edvr import EDVR
    from oracle import edvr_ref
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    kw = dict(num_feat=64, num_frame=3, deformable_groups=8, num_extract_block=1, num_reconstruct_block=2)
    sd = edvr_ref.make_state_dict(**kw, seed=7)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 3, 16, 16, generator=g).cuda()
    gt = torch.rand(2, 3, 64, 64, generator=g).cuda()

    def charbonnier(p, t):       # losses.py:24-25 with reduction='sum' (train_EDVR_L_x4_SR_REDS.yml:90-93)
        return torch.sqrt((p - t) ** 2 + 1e-12).sum()

    net = EDVR(center_frame_idx=None, **kw).cuda().train()
    net.load_state_dict(sd, strict=True)
    loss = charbonnier(net(x), gt)
    loss.backward()

    ref_params = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out_ref = edvr_ref.edvr_forward.__wrapped__(ref_params, x)      # un-decorated (no_grad) forward
        loss_ref = charbonnier(out_ref, gt)
    loss_ref.backward()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 1e-3
    named = dict(net.named_parameters())
    for key in ("conv_first.weight", "pcd_align.dcn_pack.l1.weight", "pcd_align.dcn_pack.l1.conv_offset.weight",
                "pcd_align.cas_dcnpack.bias", "fusion.feat_fusion.weight", "reconstruction.1.conv2.weight",
                "conv_last.weight"):
        a, b = named[key].grad, ref_params[key].grad
        err = float((a - b).norm() / b.norm().clamp_min(1e-20))
        print(f"grad {key}: rel-L2 {err:.2e}")
        assert err < 2e-2, (key, err)


def test_sliding_window_video_inference_is_bit_identical_to_per_clip():
    """forward_video shares the per-frame pyramid between windows; every output frame must equal forward()
    on the explicitly gathered window (same kernels on the same values => torch.equal), for two padding modes."""
    from edvr_b200.engine import EDVREngine, frame_window_indices
    from oracle import edvr_ref
    kw = dict(num_feat=64, num_frame=5, deformable_groups=8, num_extract_block=2, num_reconstruct_block=3)
    eng = EDVREngine(edvr_ref.make_state_dict(**kw, seed=11), num_frame=5)
    frames = torch.rand(9, 3, 24, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    for padding in ("reflection_circle", "replicate"):
        video = eng.forward_video(frames, clips_per_step=4, padding=padding)
        assert video.shape == (9, 3, 96, 128)
        for c in (0, 1, 4, 8):
            clip = frames[frame_window_indices(c, 9, 5, padding)].unsqueeze(0)
            assert torch.equal(eng.forward(clip)[0], video[c]), (padding, c)
