"""Quick device-side timing of the EDVR executor (not the bench contract; see bench.py)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200.engine import EDVREngine  # noqa: E402
from oracle import edvr_ref  # noqa: E402  (weights generator only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="3")
    ap.add_argument("--batches", default="1,2,4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--graph", type=int, default=1)
    a = ap.parse_args()
    cfgs = {"2": (dict(num_feat=64, num_frame=5, num_reconstruct_block=10), (5, 3, 128, 128)),
            "3": (dict(num_feat=128, num_frame=7, num_reconstruct_block=40), (7, 3, 180, 320)),
            "3t5": (dict(num_feat=128, num_frame=5, num_reconstruct_block=40), (5, 3, 180, 320))}
    kw, shp = cfgs[a.cfg]
    sd = edvr_ref.make_state_dict(**kw)
    eng = EDVREngine(sd, num_frame=kw["num_frame"])
    res = {}
    for B in [int(b) for b in a.batches.split(",")]:
        x = torch.rand(B, *shp, device="cuda")
        for _ in range(3):
            y = eng.forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            y = eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        r = {"ms_per_step": ms, "fps": 1000.0 * B / ms}
        if a.graph:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                eng.forward(x)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    yg = eng.forward(x)
            torch.cuda.synchronize()
            for _ in range(2):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            msg = e0.elapsed_time(e1) / a.iters
            r["graph_ms_per_step"] = msg
            r["graph_fps"] = 1000.0 * B / msg
            r["graph_matches"] = bool(torch.equal(yg, y))
        res[f"B{B}"] = r
        print(f"cfg {a.cfg} B={B}: {r}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"time_engine_cfg{a.cfg}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
