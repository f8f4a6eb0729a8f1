"""Static SASS instruction counts per kernel of libedvr_b200.so (cuobjdump -sass): the evidence table under profiles/.
    python tools/sass_table.py > profiles/r02_sass_mnemonics_v2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edvr_b200", "libedvr_b200.so")
COLS = [("UTCHMMA(1cta)", r"\bUTCHMMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
        ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HFMA2", r"\bHFMA2"),
        ("LDS", r"\bLDS\b"), ("LDG", r"\bLDG\b"), ("HMMA", r"\bHMMA")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order = collections.OrderedDict(), []
cur = None
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"^void ", "", cur).replace("eb::", "")
        cur = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", cur)
        cur = re.sub(r"\((?:int|bool)\)", "", cur)
        counts[cur] = [0] * len(COLS)
        continue
    if cur is None or "/*" not in line:
        continue
    for i, (_n, pat) in enumerate(COLS):
        if re.search(pat, line):
            counts[cur][i] += 1
print("# SASS evidence for edvr_b200/libedvr_b200.so (cuobjdump -sass, static instruction counts per kernel; end of round 2).")
print("# UTCHMMA = tcgen05.mma (\".2CTA\" = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores,")
print("# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none).")
print()
print(f"{'kernel':60s}" + "".join(f"{n:>14s}" for n, _ in COLS))
tot = [0] * len(COLS)
for k, v in counts.items():
    if sum(v[:8]) == 0:
        continue
    print(f"{k[:60]:60s}" + "".join(f"{c:14d}" for c in v))
    tot = [a + b for a, b in zip(tot, v)]
print(f"{'TOTAL':60s}" + "".join(f"{c:14d}" for c in tot))
