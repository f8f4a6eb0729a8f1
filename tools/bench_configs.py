"""Measurements for the BASELINE.json configurations other than the headline one (which bench.py owns):
cfg 1 single DCN layer (fwd and fwd+bwd), cfg 2 EDVR-M, cfg 4 EDVR-L deblur, plus the DCN operator at the EDVR-L
pyramid sizes — each next to the unmodified reference CUDA path on the same GPU.  Device-side CUDA-event timing,
3 warm-ups, inputs resident in HBM.  Writes gpurun_out/bench_configs.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import ops  # noqa: E402
from edvr_b200.dcn import mdcn_backward  # noqa: E402
from edvr_b200.engine import EDVREngine  # noqa: E402
from oracle import build_ref, edvr_ref  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ref_fns():
    ext = build_ref.load_ref()

    def fwd(x, off, mask, w, b, dg):
        out = x.new_empty(x.shape[0], w.shape[0], x.shape[2], x.shape[3])
        ext.modulated_deform_conv_forward(x, w, b, x.new_empty(0), off, mask, out, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
        return out

    def bwd(x, off, mask, w, b, go, dg):
        gx, goff, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, mask, w, b))
        ext.modulated_deform_conv_backward(x, w, b, x.new_empty(0), off, mask, x.new_empty(0), gx, gw, gb, goff, gm, go, 3, 3,
                                           1, 1, 1, 1, 1, 1, 1, dg, True)
        return gx, goff, gm, gw, gb

    return fwd, bwd


def dcn_op(N, C, H, W, dg=8):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    off = torch.randn(N, dg * 18, H, W, device="cuda", generator=g) * 2
    mask = torch.sigmoid(torch.randn(N, dg * 9, H, W, device="cuda", generator=g))
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    b = torch.zeros(C, device="cuda")
    go = torch.randn(N, C, H, W, device="cuda", generator=g)
    rf, rb = ref_fns()
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    r = {"shape": [N, C, H, W], "dg": dg}
    r["ours_fwd_us"] = 1e3 * timeit(lambda: ops.mdcn_forward(x, off, mask, w, b, 1, 1, 1, 1, dg))
    r["ref_cuda_fwd_us"] = 1e3 * timeit(lambda: rf(x, off, mask, w, b, dg))
    r["ours_fwd_bwd_us"] = 1e3 * timeit(lambda: (ops.mdcn_forward(x, off, mask, w, b, 1, 1, 1, 1, dg),
                                                  mdcn_backward(x, off, mask, w, go, True, 1, 1, 1, 1, dg)))
    r["ref_cuda_fwd_bwd_us"] = 1e3 * timeit(lambda: (rf(x, off, mask, w, b, dg), rb(x, off, mask, w, b, go, dg)))
    alg_bytes = 4 * (x.numel() + off.numel() + mask.numel() + w.numel() + x.numel())       # SURVEY §8d, fp32 I/O at the operator boundary
    r["algorithmic_MB"] = alg_bytes / 1e6
    r["ours_fwd_GBps_of_algorithmic"] = alg_bytes / r["ours_fwd_us"] / 1e3
    r["gemm_GFLOP"] = 2.0 * N * H * W * C * C * 9 / 1e9
    return r


def edvr_cfg(name, kw, shape, hr_in=False, B=1):
    sd = edvr_ref.make_state_dict(**kw)
    eng = EDVREngine(sd, num_frame=kw["num_frame"], hr_in=hr_in)
    x = torch.rand(B, *shape, device="cuda")
    ms = timeit(lambda: eng.forward(x), iters=5)
    rf, _ = ref_fns()
    sdc = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cudnn.benchmark = True
    x1 = x[:1]
    dcn = lambda xx, off, mask, w, b, s, p, d, g, dg: rf(xx.contiguous(), off, mask, w, b, dg)
    ms_ref = timeit(lambda: edvr_ref.edvr_forward(sdc, x1, hr_in=hr_in, dcn=dcn), iters=3, warm=2)
    return {"config": name, "clips_per_step": B, "ours_ms_per_step": ms, "ours_frames_per_s": 1e3 * B / ms,
            "ref_cuda_ms_per_clip": ms_ref, "ref_cuda_frames_per_s": 1e3 / ms_ref}


def train_cfg5(B=4):
    """BASELINE cfg 5 (per-GPU slice): EDVR-L t=5, batch 4, LR 64x64 -> GT 256x256, Charbonnier(sum), Adam(4e-4, (0.9, 0.99)):
    forward + backward + optimizer step.  Ours: edvr_b200.edvr.EDVR under autograd (PyTorch convs around OUR DCN forward and
    backward kernels).  Reference arm: the same functional graph with the UNMODIFIED reference DCN extension wrapped in an
    autograd Function.  fp32 parameters, cuDNN TF32 allowed (torch default), no autocast."""
    from edvr_b200.edvr import EDVR
    kw = dict(num_feat=128, num_frame=5, num_reconstruct_block=40)
    sd = edvr_ref.make_state_dict(**kw)
    x = torch.rand(B, 5, 3, 64, 64, device="cuda")
    gt = torch.rand(B, 3, 256, 256, device="cuda")
    charb = lambda p, t: torch.sqrt((p - t) ** 2 + 1e-12).sum()
    torch.backends.cudnn.benchmark = True

    net = EDVR(center_frame_idx=None, **kw).cuda().train()
    net.load_state_dict(sd, strict=True)
    opt = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.9, 0.99))

    def step_ours():
        opt.zero_grad(set_to_none=True)
        loss = charb(net(x), gt)
        loss.backward()
        opt.step()
        return loss

    ms_ours = timeit(step_ours, iters=5, warm=2)

    rf, rb = ref_fns()

    class RefDCN(torch.autograd.Function):
        @staticmethod
        def forward(ctx, xx, off, mask, w, b, dg):
            ctx.save_for_backward(xx, off, mask, w, b)
            ctx.dg = dg
            return rf(xx.contiguous(), off.contiguous(), mask.contiguous(), w, b, dg)

        @staticmethod
        def backward(ctx, go):
            xx, off, mask, w, b = ctx.saved_tensors
            gx, goff, gm, gw, gb = rb(xx.contiguous(), off.contiguous(), mask.contiguous(), w, b, go.contiguous(), ctx.dg)
            return gx, goff, gm, gw, gb, None

    params = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    opt_ref = torch.optim.Adam(list(params.values()), lr=4e-4, betas=(0.9, 0.99))
    dcn = lambda xx, off, mask, w, b, s, p, d, g, dg: RefDCN.apply(xx, off, mask, w, b, dg)

    def step_ref():
        opt_ref.zero_grad(set_to_none=True)
        with torch.enable_grad():
            loss = charb(edvr_ref.edvr_forward.__wrapped__(params, x, dcn=dcn), gt)
        loss.backward()
        opt_ref.step()
        return loss

    ms_ref = timeit(step_ref, iters=3, warm=2)
    return {"config": "cfg5 EDVR-L t=5 train step, batch 4 x 5x3x64x64 -> 256x256 per GPU", "ours_ms_per_step": ms_ours,
            "ours_samples_per_s": 1e3 * B / ms_ours, "ref_cuda_ms_per_step": ms_ref, "ref_cuda_samples_per_s": 1e3 * B / ms_ref,
            "note": "ours = autograd path (PyTorch/cuDNN convs + our DCN fwd/bwd kernels); conv dgrad/wgrad kernels are not built"}


def video_cfg3(n_frames=32, B=4):
    """Application-level mode: one 32-frame 180x320 sequence restored with sliding 7-frame windows
    (EDVREngine.forward_video: per-frame features computed once) vs the same windows as independent clips."""
    kw = dict(num_feat=128, num_frame=7, num_reconstruct_block=40)
    eng = EDVREngine(edvr_ref.make_state_dict(**kw), num_frame=7)
    frames = torch.rand(n_frames, 3, 180, 320, device="cuda")
    ms_video = timeit(lambda: eng.forward_video(frames, clips_per_step=B), iters=3, warm=1)
    x = torch.rand(B, 7, 3, 180, 320, device="cuda")
    ms_clip = timeit(lambda: eng.forward(x), iters=5)
    return {"config": f"cfg3 network, {n_frames}-frame sequence, sliding windows, {B} windows per step",
            "video_ms_per_sequence": ms_video, "video_frames_per_s": 1e3 * n_frames / ms_video,
            "independent_clips_frames_per_s": 1e3 * B / ms_clip}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "video":
        r = video_cfg3()
        print("cfg3_video", json.dumps(r), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(r, open(os.path.join(ROOT, "gpurun_out", "bench_cfg3_video.json"), "w"), indent=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        r = train_cfg5()
        print("cfg5_train_step", json.dumps(r), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(r, open(os.path.join(ROOT, "gpurun_out", "bench_cfg5_train.json"), "w"), indent=1)
        return
    res = {"cfg1_single_dcn_layer_1x64x64x64": dcn_op(1, 64, 64, 64),
           "dcn_L1_1x128x180x320": dcn_op(1, 128, 180, 320),
           "dcn_L2_1x128x90x160": dcn_op(1, 128, 90, 160),
           "dcn_L3_1x128x45x80": dcn_op(1, 128, 45, 80),
           "dcn_train_20x128x64x64": dcn_op(20, 128, 64, 64)}
    for k, v in res.items():
        print(k, json.dumps(v), flush=True)
    res["cfg2_edvr_m_B1"] = edvr_cfg("cfg2 EDVR-M 5x3x128x128", dict(num_feat=64, num_frame=5, num_reconstruct_block=10), (5, 3, 128, 128), B=1)
    res["cfg2_edvr_m_B16"] = edvr_cfg("cfg2 EDVR-M 5x3x128x128", dict(num_feat=64, num_frame=5, num_reconstruct_block=10), (5, 3, 128, 128), B=16)
    res["cfg3_t5_B4"] = edvr_cfg("cfg3 variant t=5 (shipped REDS yml)", dict(num_feat=128, num_frame=5, num_reconstruct_block=40), (5, 3, 180, 320), B=4)
    res["cfg4_deblur_B1"] = edvr_cfg("cfg4 EDVR-L deblur 5x3x720x1280", dict(num_feat=128, num_frame=5, num_reconstruct_block=40, with_predeblur=True, hr_in=True),
                                    (5, 3, 720, 1280), hr_in=True, B=1)
    for k in ("cfg2_edvr_m_B1", "cfg2_edvr_m_B16", "cfg3_t5_B4", "cfg4_deblur_B1"):
        print(k, json.dumps(res[k]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
