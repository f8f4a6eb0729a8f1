"""bench.py --mode train: BASELINE cfg 5 - one EDVR-L 4x SR training step per "step" (Charbonnier loss, DCNv2 backward,
Adam, DistributedDataParallel gradient all-reduce over NCCL), weak scaling at 4 clips per GPU.

Configuration of options/train/EDVR/train_EDVR_L_x4_SR_REDS.yml: num_feat 128, num_frame 5, 40 reconstruction blocks,
batch_size_per_gpu 4, GT 256x256 -> LR 64x64, Adam lr 4e-4 betas (0.9, 0.99), CharbonnierLoss reduction sum; the net is
wrapped in DDP like basicsr/models/base_model.py:62-69.  Synthetic clips, random-init weights.

value = clips/s over all ranks, device-timed (max over ranks); `e2e` includes the H2D copy of each step's clips and
targets from pinned host memory and a D2H read of the loss.  `ddp` reports the exposed gradient-exchange time:
step time with the all-reduce minus step time under no_sync() (same kernels, no exchange).  `ref_cuda` = the UNMODIFIED
reference (basicsr EDVR + its dcn extension from baseline/_ref, fp32, cuDNN) doing the same step on the same GPU."""
import json
import os
import sys
import time

CFG5 = dict(num_feat=128, num_frame=5, deformable_groups=8, num_extract_block=5, num_reconstruct_block=40)
LR = 64
METRIC = "training clips/sec EDVR-L 4xSR 5f 64x64->256x256 (cfg 5: Charbonnier, DCNv2 backward, Adam, DDP)"


def _step_fn(net, opt, loss_fn):
    def step(x, gt):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(x), gt)
        loss.backward()
        opt.step()
        return loss
    return step


def _time_steps(step, x, gt, steps, warmup, barrier):
    import torch
    for _ in range(warmup):
        step(x, gt)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(x, gt)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps, float(loss.detach())


def reference_train_rate(root, sd, B, steps=5, warmup=2):
    import torch
    ref_root = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "basicsr", "models", "archs")):
        return {"unavailable": "baseline/_ref (pip install of /root/reference) not present"}
    sys.path.insert(0, ref_root)
    from basicsr.models.archs.edvr_arch import EDVR as RefEDVR
    from edvr_b200.train import charbonnier_loss
    torch.backends.cudnn.benchmark = True
    net = RefEDVR(num_in_ch=3, num_out_ch=3, center_frame_idx=None, hr_in=False, with_predeblur=False, with_tsa=True,
                  **CFG5).cuda().train()
    net.load_state_dict(sd, strict=True)
    opt = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.9, 0.99))
    x = torch.rand(B, 5, 3, LR, LR, device="cuda")
    gt = torch.rand(B, 3, 4 * LR, 4 * LR, device="cuda")
    ms, loss = _time_steps(_step_fn(net, opt, charbonnier_loss), x, gt, steps, warmup, torch.cuda.synchronize)
    return {"value": 1000.0 * B / ms, "unit": "clips/s", "ms_per_step": ms, "clips_per_step": B, "loss": loss,
            "how": "UNMODIFIED reference: basicsr EDVR + its compiled dcn extension (baseline/_ref), fp32 (its only dtype; "
                   f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}), cuDNN fwd/dgrad/wgrad, torch Adam, 1 GPU"}


def main(args):
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from edvr_b200 import ops, synth
    from edvr_b200.edvr import EDVR
    from edvr_b200.shard import max_over_ranks
    from edvr_b200.train import charbonnier_loss

    B = args.clips
    steps, warmup = args.steps, max(args.warmup, 3)
    sd = synth.make_state_dict(**CFG5, seed=0, offset_std=args.offset_std)
    net = EDVR(center_frame_idx=None, **CFG5).cuda().train()
    net.load_state_dict(sd, strict=True)
    net.train_dtype = torch.bfloat16
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], bucket_cap_mb=25) if world > 1 else net
    opt = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.9, 0.99), fused=True)
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.rand(B, 5, 3, LR, LR, device="cuda", generator=g)
    gt = torch.rand(B, 3, 4 * LR, 4 * LR, device="cuda", generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = _step_fn(model, opt, charbonnier_loss)
    c0 = ops.LAUNCHES[0]
    ms, loss = _time_steps(step, x, gt, steps, warmup, barrier)
    launches = (ops.LAUNCHES[0] - c0) * steps // (steps + warmup)
    ms = max_over_ranks(ms, world)

    # e2e: this step's clips and targets come from pinned host memory, the loss is read back
    xh, gh = x.cpu().pin_memory(), gt.cpu().pin_memory()
    lh = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e(_x, _gt):
        loss_ = step(xh.cuda(non_blocking=True), gh.cuda(non_blocking=True))
        lh.copy_(loss_.detach(), non_blocking=True)
        return loss_
    ms_e2e, _ = _time_steps(step_e2e, x, gt, steps, 1, barrier)
    ms_e2e = max_over_ranks(ms_e2e, world)

    # one-launch variant: the whole step as a CUDA graph (single GPU; DDP's bucketed all-reduce stays eager)
    graphed = None
    if world == 1:
        try:
            from edvr_b200.train import GraphedTrainStep
            opt_g = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.9, 0.99), capturable=True)
            gstep = GraphedTrainStep(net, opt_g, charbonnier_loss, x, gt)
            ms_g, loss_g = _time_steps(lambda a, b: gstep(), x, gt, steps, 2, barrier)
            ms_ge, _ = _time_steps(lambda a, b: (gstep(xh, gh), lh.copy_(gstep.loss.detach(), non_blocking=True))[0], x, gt,
                                   steps, 1, barrier)
            graphed = {"ms_per_step": ms_g, "value": B * 1000.0 / ms_g, "unit": "clips/s", "e2e_ms_per_step": ms_ge,
                       "e2e_value": B * 1000.0 / ms_ge, "last_loss": loss_g,
                       "how": "edvr_b200.train.GraphedTrainStep: forward + loss + backward + Adam captured once, replayed per step"}
        except Exception as e:      # noqa: BLE001
            graphed = {"unavailable": repr(e)[:300]}

    ddp = None
    if world > 1:
        def step_nosync(x_, gt_):
            with model.no_sync():
                return step(x_, gt_)
        ms_ns, _ = _time_steps(step_nosync, x, gt, max(5, steps // 2), 1, barrier)
        ms_ns = max_over_ranks(ms_ns, world)
        nbytes = sum(p.numel() for p in net.parameters()) * 4
        ddp = {"grad_bytes": nbytes, "ms_per_step_no_sync": ms_ns, "exposed_allreduce_ms": ms - ms_ns,
               "bucket_cap_mb": 25, "note": "NCCL all-reduce of fp32 gradients in 25 MB buckets, overlapped with the backward "
                                            "pass by DDP; exposed = step - step under no_sync()"}
    if rank == 0:
        line = {"metric": METRIC, "value": world * B * 1000.0 / ms, "unit": "clips/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 activations and gradients (tcgen05 kind::f16, bf16 operands) / f32 accumulate, f32 master weights",
                "data": "synthetic",
                "config": {"workload": "EDVR-L 4xSR training step cfg 5: nf=128, 5 frames, 40 recon blocks, LR 64x64 -> GT 256x256",
                           "clips_per_gpu_per_step": B, "global_clips_per_step": world * B, "parallelism": f"ddp{world}",
                           "optimizer": "Adam lr 4e-4 betas (0.9, 0.99), fused", "loss": "Charbonnier, reduction sum",
                           "last_loss": loss},
                "e2e": {"value": world * B * 1000.0 / ms_e2e, "unit": "clips/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": (xh.numel() + gh.numel()) * 4, "d2h_bytes_per_step": 4,
                        "api": "edvr_b200.edvr.EDVR (train mode) under torch DDP + torch.optim.Adam"},
                "gpu_launches": launches, "graphed_step": graphed, "ddp": ddp}
        if os.environ.get("EDVR_BENCH_PROFILING") != "1":
            try:
                line["ref_cuda"] = reference_train_rate(root, sd, B)
            except Exception as e:      # noqa: BLE001
                line["ref_cuda"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
