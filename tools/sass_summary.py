"""Per-kernel SASS evidence for the tensor-core / TMA claims: counts of the Blackwell instructions in every kernel of
liblivae_sm100.so.  UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM =
tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync, REDG / ATOMG =
global reductions / atomics, ATOMS = shared-memory atomics.
usage: python tools/sass_summary.py > profiles/r02_sass_tc_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "li-vae_b200", "livae", "liblivae_sm100.so")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "REDG", "ATOMG", "ATOMS", "LDGSTS", "FFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, counts, total = None, {}, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                counts[kern][o] += 1
print(f"SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); {len(counts)} kernels")
print(f"{'kernel':86s} {'instr':>6s} " + " ".join(f"{o:>7s}" for o in OPS))
agg = collections.Counter()
for k in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"] * 100000 - total[k]):
    name = re.sub(r"\(.*", "", demangle(k))[:86]
    print(f"{name:86s} {total[k]:6d} " + " ".join(f"{counts[k][o]:7d}" for o in OPS))
    agg.update(counts[k])
print(f"{'TOTAL':86s} {sum(total.values()):6d} " + " ".join(f"{agg[o]:7d}" for o in OPS))
