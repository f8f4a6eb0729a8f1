"""tcgen05.mma issue-rate probe (eb_selftest_mma_rate): cycles per MMA for tile shapes, shared-memory layouts and
1-CTA vs CTA-pair issue.  Each variant runs in its own process so that a faulting descriptor does not take the rest down."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (cta_group, M, N, layout, a_lbo, a_sbo, b_lbo, b_sbo, kstep_bytes)
VARIANTS = {
    "cg1 nosw 128x64": (1, 128, 64, 0, 2048, 128, 1024, 128, 4096),
    "cg1 nosw 128x128": (1, 128, 128, 0, 2048, 128, 2048, 128, 4096),
    "cg1 nosw 128x256": (1, 128, 256, 0, 2048, 128, 4096, 128, 4096),
    "cg1 nosw 128x128 halo-pitch A": (1, 128, 128, 0, 5200, 288, 2048, 128, 10400),
    "cg1 sw128 128x64": (1, 128, 64, 2, 16, 1024, 16, 1024, 32),
    "cg1 sw128 128x128": (1, 128, 128, 2, 16, 1024, 16, 1024, 32),
    "cg1 sw128 128x256": (1, 128, 256, 2, 16, 1024, 16, 1024, 32),
    "cg1 sw64 128x128": (1, 128, 128, 4, 16, 512, 16, 512, 32),
    "cg1 sw32 128x128": (1, 128, 128, 6, 16, 256, 16, 256, 4096),
    "cg2 nosw 256x128": (2, 256, 128, 0, 2048, 128, 1024, 128, 4096),
    "cg2 nosw 256x256": (2, 256, 256, 0, 2048, 128, 2048, 128, 4096),
    "cg2 sw128 256x128": (2, 256, 128, 2, 16, 1024, 16, 1024, 32),
    "cg2 sw128 256x256": (2, 256, 256, 2, 16, 1024, 16, 1024, 32),
    "cg2 nosw 256x128 halo-pitch A": (2, 256, 128, 0, 5200, 288, 1024, 128, 10400),
}


def one(name):
    import torch
    from edvr_b200 import _lib as L
    cg, M, N, layout, al, asb, bl, bsb, ks = VARIANTS[name]
    reps = 4096
    cyc = torch.zeros(160, dtype=torch.int64, device="cuda")
    n = ctypes.c_int(0)
    for _ in range(2):
        L.check(L.lib().eb_selftest_mma_rate(cg, M, N, layout, al, asb, bl, bsb, ks, reps, L.ptr(cyc), ctypes.byref(n),
                                             L.stream_ptr()), "mma_rate")
        torch.cuda.synchronize()
    c = cyc[:n.value].double()
    c = c[c > 0] / reps
    flop = 2.0 * M * N * 16
    print(json.dumps({"variant": name, "ctas": n.value, "cyc_per_mma_min": float(c.min()), "avg": float(c.mean()),
                      "max": float(c.max()), "flop_per_clk_per_sm": flop / float(c.mean()) / cg,
                      "smem_B_per_clk_per_sm": (128 * 32 + N // cg * 32) / float(c.mean())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        for name in VARIANTS:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=120)
            out = p.stdout.strip().splitlines()
            print(out[-1] if out and p.returncode == 0 else f"{name}: FAILED rc={p.returncode} {p.stderr.strip()[-300:]}", flush=True)
