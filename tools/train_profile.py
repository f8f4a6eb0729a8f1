"""Where the GPU time of one cfg-5 training step goes: torch.profiler kernel table (top kernels by device time)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import synth  # noqa: E402
from edvr_b200.edvr import EDVR  # noqa: E402
from edvr_b200.train import charbonnier_loss  # noqa: E402
from edvr_b200.train_bench import CFG5, LR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = EDVR(center_frame_idx=None, **CFG5).cuda().train()
net.load_state_dict(synth.make_state_dict(**CFG5, seed=0), strict=True)
opt = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.9, 0.99), fused=True)
x = torch.rand(B, 5, 3, LR, LR, device="cuda")
gt = torch.rand(B, 3, 4 * LR, 4 * LR, device="cuda")


def step():
    opt.zero_grad(set_to_none=True)
    loss = charbonnier_loss(net(x), gt)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted((k for k in ka if k.device_time_total > 0 and k.device_type.name == "CUDA"), key=lambda k: -k.device_time_total)
tot = sum(k.device_time_total for k in rows)
print(f"total device time of one step: {tot / 1e3:.2f} ms over {sum(k.count for k in rows)} kernels")
for k in rows[:28]:
    print(f"{k.device_time_total / 1e3:8.3f} ms {100 * k.device_time_total / tot:5.1f}%  x{k.count:5d}  {k.key[:110]}")
