"""Launch one trunk-shaped conv a few times (ncu target)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import ops
N, H, W, C = 4, 180, 320, 128
x = ops.nchw_to_nhwc(torch.randn(N, C, H, W, device="cuda"))
w = torch.randn(C, C, 3, 3, device="cuda") / 34
pc = ops.pack_conv(w, torch.zeros(C, device="cuda"))
out = ops.new_act(N, H, W, C)
for _ in range(4):
    ops.conv2d(pc, [x], out16=out, act=ops.ACT_RELU)
torch.cuda.synchronize()
print("done")
