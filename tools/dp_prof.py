"""Where does the CTA-pair DCN site kernel (dcn_pair.cuh) wait?  `--build` compiles a -DDP_PROF copy of the library into
edvr_b200/libedvr_b200_prof.so (in-tree, so that it travels to the GPU box; git-ignored like every .so); the run loads it
through EDVR_B200_LIB, launches the site once per EDVR_B200_DP_DBG setting and prints, per role, the mean cycles per tile
spent inside each kind of wait (profiles/r02_dcn_pair_role_timing.txt).

    python tools/dp_prof.py --build                               # where nvcc is
    python tools/dp_prof.py --build-variant NAME -DDP_CFG_...     # A/B builds of the ring depths / window margin
    python tools/dp_prof.py [N] [sigma] [dbg,dbg,...]             # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
PROF_LIB = os.path.join(ROOT, "edvr_b200", "libedvr_b200_prof.so")

if "--build" in sys.argv:
    from edvr_b200 import build as b
    cmd = [b.NVCC] + b.FLAGS + ["-DDP_PROF", "-o", PROF_LIB, os.path.join(b.CSRC, "capi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stderr[-2000:] if r.returncode else "built " + PROF_LIB)
    sys.exit(r.returncode)
if "--build-variant" in sys.argv:        # python tools/dp_prof.py --build-variant NAME -DDP_CFG_A_STAGES=4 ...
    from edvr_b200 import build as b
    i = sys.argv.index("--build-variant")
    out = os.path.join(ROOT, "edvr_b200", f"libedvr_b200_{sys.argv[i + 1]}.so")
    cmd = [b.NVCC] + b.FLAGS + sys.argv[i + 2:] + ["-o", out, os.path.join(b.CSRC, "capi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stderr[-2000:] if r.returncode else "built " + out)
    sys.exit(r.returncode)

os.environ["EDVR_B200_LIB"] = PROF_LIB
os.environ["EDVR_B200_DCN_SITE"] = "pair"
import torch  # noqa: E402
from edvr_b200 import _lib as L, ops  # noqa: E402
from dcn_sweep import site_weights  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 28
sigma = float(args[1]) if len(args) > 1 else 0.02
dbgs = [int(v) for v in (args[2] if len(args) > 2 else "0").split(",")]
H, W, C, dg = 180, 320, 128, 8
wo, bo, w, b = site_weights(C, dg, sigma, torch.Generator().manual_seed(1))
g = torch.Generator(device="cuda").manual_seed(0)
x = ops.nchw_to_nhwc(torch.randn(N, C, H, W, device="cuda", generator=g))
feat = ops.nchw_to_nhwc(torch.randn(N, C, H, W, device="cuda", generator=g))
out = ops.new_act(N, H, W, C)
site = ops.DcnSite(wo, bo, w, b, dg)
assert site.mode == "pair"
tiles = N * ((H + 15) // 16) * ((W + 7) // 8)
names = {0: ("gather total", 16), 1: ("gather wait off_full", 16), 2: ("gather wait win_full", 16), 3: ("gather wait empty", 16),
         4: ("gather tmem fetch", 16), 5: ("fwd wait gathered", 1), 6: ("fwd total", 1), 7: ("B-issuer wait acc_empty", 0.5),
         8: ("B-issuer wait wfull", 0.5), 9: ("B-issuer wait full", 0.5), 10: ("B-issuer total", 0.5),
         11: ("A-issuer wait off_empty", 0.5), 12: ("A-issuer wait f_full", 0.5), 13: ("A-issuer wait wo_full", 0.5),
         14: ("A-issuer total", 0.5), 15: ("B-prod wait wempty", 1), 16: ("B-prod wait win_empty", 1), 17: ("B-prod total", 1),
         18: ("A-prod wait wo_empty", 1), 19: ("A-prod wait f_empty", 1), 20: ("A-prod total", 1), 21: ("epi wait acc_full", 4),
         22: ("epi total", 4)}
buf = (ctypes.c_ulonglong * 32)()
for d in dbgs:
    os.environ["EDVR_B200_DP_DBG"] = str(d)
    for _ in range(2):
        site(x, feat, out)
    torch.cuda.synchronize()
    L.check(L.lib().eb_dcn_pair_prof_read(buf))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    site(x, feat, out)
    e1.record()
    torch.cuda.synchronize()
    L.check(L.lib().eb_dcn_pair_prof_read(buf))
    print(f"dbg={d} N={N} sigma={sigma}: {e0.elapsed_time(e1) * 1e3:.1f} us; cycles per tile (mean over the role's warps):")
    line = []
    for i, (nm, per_cta_warps) in names.items():
        line.append(f"{nm} {buf[i] / (tiles * per_cta_warps):.0f}")
    print("   " + " | ".join(line[:5]))
    print("   " + " | ".join(line[5:11]))
    print("   " + " | ".join(line[11:17]))
    print("   " + " | ".join(line[17:]))
