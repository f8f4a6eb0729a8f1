"""Per-launch-shape breakdown of one EDVR executor step (CUDA events around every call; cfg 3 by default)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import ops  # noqa: E402
from edvr_b200.engine import EDVREngine  # noqa: E402
from oracle import edvr_ref  # noqa: E402  (weights generator only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    kw = dict(num_feat=128, num_frame=a.frames, num_reconstruct_block=40)
    eng = EDVREngine(edvr_ref.make_state_dict(**kw), num_frame=a.frames)
    x = torch.rand(a.clips, a.frames, 3, 180, 320, device="cuda")
    for _ in range(3):
        eng.forward(x)
    torch.cuda.synchronize()
    agg = {}
    for _ in range(a.reps):
        ops.PROFILE = []
        eng.forward(x)
        torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
        for name, flops, e0, e1, detail in recs:
            v = agg.setdefault((name, detail), [0.0, 0.0, 0])
            v[0] += e0.elapsed_time(e1) / a.reps
            v[1] += flops / a.reps
            v[2] += 1
    total = sum(v[0] for v in agg.values())
    print(f"step total (sum of per-call event times) {total:.3f} ms, B={a.clips}")
    for (name, detail), v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        n = v[2] // a.reps
        tf = f"{v[1] / (v[0] * 1e9):7.0f} TF/s" if v[1] > 0 else " " * 12
        print(f"{v[0]:8.3f} ms {100 * v[0] / total:5.1f}%  x{n:3d}  {1e3 * v[0] / max(n, 1):8.1f} us  {tf}  {name} {detail}")


if __name__ == "__main__":
    main()
