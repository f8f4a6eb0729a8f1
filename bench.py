#!/usr/bin/env python
"""bench.py — HR frames/s of the EDVR-L 4x SR hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (oracle port)

A "step" is one forward pass of EDVR-L (num_feat 128, 7 frames, 40 reconstruction blocks) over a batch
of B synthetic REDS-shaped clips [B,7,3,180,320] -> B HR frames [B,3,720,1280] per GPU (weak scaling:
clips shard batch-parallel, no data-path collective).  `value` is device-timed with the
inputs resident in HBM; `e2e` goes through the public drop-in module (edvr_b200.edvr.EDVR.forward) with
pinned HOST buffers, H2D and D2H copies inside the timed region.  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG3 = dict(num_feat=128, num_frame=7, deformable_groups=8, num_extract_block=5, num_reconstruct_block=40)
LR_H, LR_W = 180, 320
TFLOP_PER_CLIP = 5.251          # SURVEY §8(d) / BASELINE.md §2: 2*MACs of every conv + DCN GEMM, cfg 3
METRIC = "HR frames/sec EDVR-L 4xSR 7f 180x320->1280x720"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        for line in self.f.read().strip().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(threads=None, crop=(96, 160), steps=1, warmup=1):
    """Reference graph on the host cores: oracle port of edvr_arch.py (bit-exact, tests/test_oracle.py) with
    the DCN via torchvision's CPU deform_conv2d.  Bounded sample: one EDVR-L clip cropped to
    `crop` LR pixels; conv work is linear in pixels, so clips/s = (crop / full pixels) / seconds."""
    import torch
    from oracle import edvr_ref
    if threads:
        torch.set_num_threads(threads)
    threads = torch.get_num_threads()          # PyTorch's default = all the cores it will use
    sd = edvr_ref.make_state_dict(**CFG3, seed=0)
    h, w = crop
    x = torch.rand(1, 7, 3, h, w, generator=torch.Generator().manual_seed(0))
    frac = (h * w) / float(LR_H * LR_W)
    for _ in range(warmup):
        edvr_ref.edvr_forward(sd, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        edvr_ref.edvr_forward(sd, x)
    dt = (time.perf_counter() - t0) / steps
    return {"value": frac / dt, "unit": "HR frames/s", "cores": threads, "kind": "port",
            "sample": f"1 EDVR-L clip cropped to 7x3x{h}x{w} LR ({frac:.4f} of 180x320) per step, {steps} step(s), "
                      f"{dt:.2f} s/step; DCN = torchvision CPU deform_conv2d",
            "sec_per_step": dt}


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cb = cpu_reference_rate(steps=max(args.steps, 1), warmup=max(args.warmup, 0))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "HR frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "EDVR-L 4xSR inference, 7x3x180x320 -> 3x720x1280 (cfg 3), CPU reference graph",
                       "clips_per_step": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "HR frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def reference_cuda_rate(sd, B, iters=5):
    """The reference's own CUDA dcn extension (oracle/_ref, unmodified) + stock PyTorch/cuDNN graph, same box."""
    import torch
    from oracle import build_ref, edvr_ref
    if not os.path.exists(build_ref.so_path()):
        return {"unavailable": "oracle/_ref/deform_conv_ext_ref.so not built"}
    ext = build_ref.load_ref()
    torch.backends.cudnn.benchmark = True          # as basicsr/test.py:17

    def dcn(x, off, mask, w, b, s, p, d, g, dg):
        x = x.contiguous()
        out = x.new_empty(x.shape[0], w.shape[0], x.shape[2], x.shape[3])
        ext.modulated_deform_conv_forward(x, w, b, x.new_empty(0), off, mask, out, x.new_empty(0), 3, 3, s, s, p, p,
                                          d, d, g, dg, True)
        return out

    sdc = {k: v.cuda() for k, v in sd.items()}
    x = torch.rand(B, 7, 3, LR_H, LR_W, device="cuda")
    for _ in range(2):
        edvr_ref.edvr_forward(sdc, x, dcn=dcn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        edvr_ref.edvr_forward(sdc, x, dcn=dcn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": 1000.0 * B / ms, "unit": "HR frames/s", "ms_per_step": ms, "clips_per_step": B,
            "how": "unmodified reference dcn CUDA ext (oracle/_ref) + oracle/edvr_ref.py graph on cuDNN, fp32, "
                   f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}, cudnn.benchmark=True, 1 GPU"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from edvr_b200 import ops
    from edvr_b200.edvr import EDVR
    from oracle import edvr_ref      # weight generator only (synthetic reference-format state_dict)

    B = args.clips
    sd = edvr_ref.make_state_dict(**CFG3, seed=0)
    net = EDVR(center_frame_idx=None, **CFG3).cuda().eval()
    net.load_state_dict(sd, strict=True)
    eng = net.engine()
    g = torch.Generator(device="cuda").manual_seed(rank)          # seed + rank, like basicsr/train.py:53
    x_dev = torch.rand(B, 7, 3, LR_H, LR_W, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value)
    counter0 = ops_launch_counter()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        y = eng.forward(x_dev)
    per_step_launches = None
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c_before = ops_launch_counter()
    e0.record()
    for _ in range(args.steps):
        flush.zero_()
        y = eng.forward(x_dev)
    e1.record()
    barrier()
    c_after = ops_launch_counter()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    clocks = sampler.stop() if sampler else None
    launches = c_after - c_before

    # ---------------- end-to-end through the public module API with host buffers (e2e)
    x_host = torch.rand(B, 7, 3, LR_H, LR_W).pin_memory()
    y_host = torch.empty(B, 3, 4 * LR_H, 4 * LR_W).pin_memory()
    with torch.no_grad():
        for _ in range(2):
            y_host.copy_(net(x_host.cuda(non_blocking=True)), non_blocking=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            xd = x_host.cuda(non_blocking=True)            # H2D of this step's clips
            yd = net(xd)                                   # public drop-in module call
            y_host.copy_(yd, non_blocking=True)            # D2H of the HR frames
        e1.record()
        barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e_step = float(ms_e2e.item()) / args.steps

    if rank == 0:
        peak_tf, peak_hbm, peak_src = measured_peaks()
        prof = profile_step(eng, x_dev)
        dom = prof["dominant"]
        value = world * B * 1000.0 / ms_step
        line = {
            "metric": METRIC, "value": value, "unit": "HR frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
            "config": {"workload": "EDVR-L 4xSR inference cfg 3: nf=128, 7 frames, 40 recon blocks, "
                                   "7x3x180x320 -> 3x720x1280", "clips_per_gpu_per_step": B,
                       "global_clips_per_step": world * B, "parallelism": f"dp{world} (batch-sharded, no collective)",
                       "weights": "random init, reference initialisers, conv_offset ~ N(0, 0.02^2)",
                       "l2": "256 MiB memset between steps inside the timed region; per-step activation working set >> 126 MB L2",
                       "achieved_tflops": TFLOP_PER_CLIP * value, "tflop_per_clip": TFLOP_PER_CLIP},
            "clocks": clocks,
            "e2e": {"value": world * B * 1000.0 / ms_e2e_step, "unit": "HR frames/s", "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4,
                    "api": "edvr_b200.edvr.EDVR.forward (drop-in for basicsr.models.archs.edvr_arch.EDVR), pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"kernel": dom["name"], "bound": "tensor", "achieved": dom["tflops"], "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": dom["tflops"] / peak_tf, "traffic": 74.1e6,
                         "traffic_note": "dram read 59.3 MB + write 14.8 MB per launch of the 128->128 3x3 trunk conv at "
                                         "4x180x320 on the CTA-pair kernel (ncu --set full, "
                                         "profiles/r01_ncu_conv_pair_trunk_4x180x320_summary.csv); algorithmic I/O of that "
                                         "launch 118 MB fp16 - most of the output stays in the 126 MB L2",
                         "peak_source": peak_src + ", bf16 sustained (fp16 runs at the same tensor rate)",
                         "launches_per_step": dom["launches"], "avg_launch_ms": dom["avg_ms"],
                         "share_of_step": dom["share"],
                         "how": "algorithmic FLOPs (2*N*H*W*Cout*Cin*k*k per launch) / CUDA-event time of each launch, "
                                "one instrumented step after the timed region"},
            "kernel_shares": prof["shares"],
        }
        if world == 1 and os.environ.get("EDVR_BENCH_PROFILING") == "1":
            # launch-list runs under ncu (profiles/): the baseline arms would only add minutes of serialised replays
            line["cpu_baseline"] = {"skipped": "EDVR_BENCH_PROFILING=1"}
        elif world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate().items() if k != "sec_per_step"}
            try:
                line["ref_cuda"] = reference_cuda_rate(sd, 1)
            except Exception as e:       # baseline leg only; never hides the product number
                line["ref_cuda"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ops_launch_counter():
    from edvr_b200 import ops
    return ops.LAUNCHES[0]


def profile_step(eng, x):
    """One instrumented forward: CUDA events around every launch class (on the launching stream)."""
    import torch
    from edvr_b200 import ops
    ops.PROFILE = []
    eng.forward(x)
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    agg, total = {}, 0.0
    for name, flops, e0, e1, _detail in recs:
        ms = e0.elapsed_time(e1)
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += ms; a[1] += flops; a[2] += 1
        total += ms
    shares = {k: {"ms": round(v[0], 4), "share": round(v[0] / total, 4), "launches": v[2],
                  "tflops": round(v[1] / (v[0] * 1e9), 1) if v[0] > 0 and v[1] > 0 else None} for k, v in agg.items()}
    name = max(agg, key=lambda k: agg[k][0])
    v = agg[name]
    return {"shares": shares, "dominant": {"name": name, "tflops": v[1] / (v[0] * 1e9), "launches": v[2],
                                           "avg_ms": v[0] / v[2], "share": v[0] / total}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=4, help="clips per GPU per step (8 runs into the 1 kW power cap)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
