"""Per-kernel SASS evidence of the Blackwell paths: counts of tcgen05 MMA (UTCHMMA), TMEM loads (LDTM), TMA loads (UTMALDG),
TMEM alloc (UTCATOMSWS / UTCBAR), cluster barriers (UCGABAR) and distributed-shared-memory loads (LD...shared::cluster show
as LDS with cluster addressing is not distinguishable; MAPA is counted) in avsr_b200/libavsr_b200.so.

    python tools/sass_summary.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "avsr_b200", "libavsr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = {"UTCHMMA": r"\bUTCHMMA", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG", "UBLKCP(bulk copy)": r"\bUBLKCP", "UTMAPF(L2 prefetch)": r"\bUBLKPF|\bUTMAPF", "UTCBAR/commit": r"\bUTCBAR",
        "UCGABAR(cluster barrier)": r"\bUCGABAR", "SYNCS(mbarrier)": r"\bSYNCS", "LDGSTS(cp.async)": r"\bLDGSTS",
        "instructions": r"^\s+/\*[0-9a-f]{4,}\*/"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
demangled = {}
try:
    names = list(counts)
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(names, dm))
except OSError:
    pass
print(f"SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a); columns = instruction counts per kernel\n")
keys = list(pats)
print(f"{'kernel':70s} " + " ".join(f"{k.split('(')[0]:>9s}" for k in keys))
for fn, c in counts.items():
    name = demangled.get(fn, fn).replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    if not any(c[k] for k in keys[:-1]) and "--all" not in sys.argv:
        continue
    print(f"{name[:70]:70s} " + " ".join(f"{c[k]:9d}" for k in keys))
