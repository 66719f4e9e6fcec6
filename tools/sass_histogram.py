"""SASS opcode histogram per kernel of libdeepsc_b200.so (cuobjdump -sass): the tcgen05 / TMEM / bulk-copy mnemonics that
show which kernels are tensor-core kernels (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops).   python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deepsc-gan_b200", "csrc", "libdeepsc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "FFMA", "HMMA", "SHFL", "LDG", "STG", "LDS", "STS", "LDL", "STL", "MUFU", "BAR")
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: static SASS instruction counts per kernel (sm_100a)")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>8s}" for k in KEY))
tot = collections.Counter()
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    print(f"{k[:58]:58s} {sum(h.values()):6d} " + " ".join(f"{h.get(x, 0):8d}" for x in KEY))
    tot.update(h)
print(f"{'ALL KERNELS':58s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(x, 0):8d}" for x in KEY))
