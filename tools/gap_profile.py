"""GPU timeline of one cfg-3 inference step (torch.profiler / CUPTI kernel records): busy time, idle gaps between consecutive
kernels, and which kernels the largest gaps follow.  Answers "step time - sum of kernel times = ?".
    python tools/gap_profile.py [clips]"""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edvr_b200 import synth  # noqa: E402
from edvr_b200.engine import EDVREngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kw = dict(num_feat=128, num_frame=7, deformable_groups=8, num_extract_block=5, num_reconstruct_block=40)
eng = EDVREngine(synth.make_state_dict(**kw, seed=0), num_frame=7)
x = torch.rand(B, 7, 3, 180, 320, device="cuda")
for _ in range(3):
    eng.forward(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        eng.forward(x)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
ev = ev[len(ev) // 2:]                      # the second forward
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy = sum(e["dur"] for e in ev)
gaps = []
end = ev[0]["ts"] + ev[0]["dur"]
for prev, e in zip(ev, ev[1:]):
    g = e["ts"] - end
    gaps.append((g, prev["name"][:60], e["name"][:60]))
    end = max(end, e["ts"] + e["dur"])
pos = [g for g in gaps if g[0] > 0]
neg = [g for g in gaps if g[0] < 0]
print(f"B={B}: span {(t1 - t0) / 1e3:.3f} ms, sum of kernel durations {busy / 1e3:.3f} ms over {len(ev)} records, "
      f"idle gaps {sum(g[0] for g in pos) / 1e3:.3f} ms ({len(pos)}), overlaps (PDL) {-sum(g[0] for g in neg) / 1e3:.3f} ms ({len(neg)})")
agg = {}
for g, a, b in pos:
    k = (a.split("<")[0].split("(")[0][-40:], b.split("<")[0].split("(")[0][-40:])
    v = agg.setdefault(k, [0.0, 0])
    v[0] += g
    v[1] += 1
for (a, b), (tot, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {tot / 1e3:7.3f} ms  x{n:3d}  avg {tot / n:6.1f} us   {a}  ->  {b}")
