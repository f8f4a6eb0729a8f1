"""Time one conv shape under the profiling switches of the pixel-major kernel (EDVR_B200_DBG bits, see ConvParams::dbg):
which role bounds the tile time?  Results of dbg != 0 runs are wrong by construction."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import ops  # noqa: E402

os.environ["EDVR_B200_CONV_V1"] = "1"
os.environ["EDVR_B200_CONV_PAIR"] = "0"          # the switches live in the single-CTA pixel-major kernel


def run(N, H, W, cin, cout, k, dbgs, f32=False, pair_dbgs=(0,)):
    x = ops.nchw_to_nhwc(torch.randn(N, cin, H, W, device="cuda"))
    w = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
    pc = ops.pack_conv(w, torch.zeros(cout, device="cuda"))
    out = ops.new_act(N, H, W, cout)
    stream = ops.Blocked32(N, H, W, cout) if f32 else None
    fl = 2.0 * N * H * W * cout * cin * k * k
    tiles = N * ((H + 15) // 16) * ((W + 15) // 16) * pc.n_tiles
    rounds = -(-tiles // 148)
    for dbg in dbgs:
        os.environ["EDVR_B200_DBG"] = str(dbg)
        fn = lambda: ops.conv2d(pc, [x], out16=out, act=ops.ACT_RELU if not f32 else ops.ACT_NONE, res32=stream, out32=stream)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"N={N} {H}x{W} {cin}->{cout} k{k} f32={int(f32)} dbg={dbg:2d}: {us:8.1f} us  {fl / us / 1e6:6.0f} TF/s  "
              f"{us / rounds * 1.9e3:7.0f} clk/tile-round ({rounds} rounds)", flush=True)
    os.environ["EDVR_B200_DBG"] = "0"
    if pc.wpair is not None:
        os.environ["EDVR_B200_CONV_PAIR"] = "1"
        prounds = -(-(N * ((H + 15) // 16) * (((W + 15) // 16 + 1) // 2) * pc.n_tiles) // 74)
        for dbg in pair_dbgs:
            os.environ["EDVR_B200_DBG"] = str(dbg)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print(f"N={N} {H}x{W} {cin}->{cout} k{k} f32={int(f32)} CTA-pair dbg={dbg:2d}: {us:8.1f} us  {fl / us / 1e6:6.0f} TF/s  "
                  f"{us / prounds * 1.9e3:7.0f} clk/pair-tile-round ({prounds} rounds)", flush=True)
        os.environ["EDVR_B200_DBG"] = "0"
        os.environ["EDVR_B200_CONV_PAIR"] = "0"


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":       # ncu target: the trunk convolution on the CTA-pair kernel only
        os.environ["EDVR_B200_CONV_PAIR"] = "1"
        x = ops.nchw_to_nhwc(torch.randn(4, 128, 180, 320, device="cuda"))
        pc = ops.pack_conv(torch.randn(128, 128, 3, 3, device="cuda") / 34, torch.zeros(128, device="cuda"))
        out = ops.new_act(4, 180, 320, 128)
        for _ in range(4):
            ops.conv2d(pc, [x], out16=out, act=ops.ACT_RELU)
        stream = ops.Blocked32(4, 180, 320, 128)             # the trunk's second conv: fp32 residual stream in and out
        for _ in range(4):
            ops.conv2d(pc, [x], out16=out, act=ops.ACT_NONE, res32=stream, out32=stream)
        torch.cuda.synchronize()
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    DBGS = [0] if quick else [0, 1, 4, 8, 16, 24, 28, 60, 32]
    run(28, 180, 320, 128, 128, 3, DBGS, pair_dbgs=(0, 1, 4, 16, 20, 32, 52))
    run(4, 180, 320, 128, 128, 3, DBGS[:1])
    run(4, 180, 320, 128, 128, 3, [0] if quick else [0, 4, 8, 28], f32=True, pair_dbgs=(0, 1, 4))
    run(28, 180, 320, 128, 256, 3, [0])
    run(4, 720, 1280, 64, 64, 3, [0])
    run(28, 180, 320, 256, 128, 3, [0] if quick else [0, 8, 28])
    run(4, 180, 320, 896, 256, 1, [0] if quick else [0, 8, 16, 28])
