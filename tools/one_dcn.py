"""Launch the fused DCN (packed-offset pipeline entry) on one shape a few times: ncu target / quick timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import ops  # noqa: E402

N, H, W, C, dg = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (4, 180, 320, 128, 8)))
g = torch.Generator(device="cuda").manual_seed(0)
x = ops.nchw_to_nhwc(torch.randn(N, C, H, W, device="cuda", generator=g))
feat = ops.nchw_to_nhwc(torch.randn(N, C, H, W, device="cuda", generator=g))
wo = torch.randn(dg * 27, C, 3, 3, device="cuda", generator=g) * 0.02
bo = torch.randn(dg * 27, device="cuda", generator=g) * 0.5
w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
po = ops.pack_conv(wo, bo, row_map=ops.dcn_offset_row_map(dg))
pw = ops.pack_conv(w, torch.zeros(C, device="cuda"), tap_major=False)
offp = ops.new_act(N, H, W, dg * 32)
ops.conv2d(po, [feat], out16=offp, act=ops.ACT_DCN_PACK)
out = ops.new_act(N, H, W, C)
for _ in range(3):
    ops.dcn_nhwc(pw, x, offp, dg, out16=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.dcn_nhwc(pw, x, offp, dg, out16=out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"dcn N={N} {H}x{W} C={C}: {us:.1f} us  {2.0 * N * H * W * C * C * 9 / us / 1e6:.0f} TFLOP/s  checksum {float(out.t.float().abs().mean()):.6f}")
