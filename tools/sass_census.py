"""Per-kernel census of the Blackwell-specific SASS in libhmgan.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA), TMA loads /
stores (UTMALDG / UTMASTG), TMEM loads (LDTM), tcgen05 commits (UTCBAR), bulk reductions (REDG / RED) and -- as the thing
that must NOT be there -- legacy warp-level MMAs (HMMA).   usage: python tools/sass_census.py > profiles/rN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gan-heightmaps_b200", "libhmgan.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "RED", "ATOMG", "SYNCS"]
cnt = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cnt[cur]["_total"] += 1
        for p in pats:
            if op.startswith(p):
                cnt[cur][p] += 1
print("%-58s %7s " % ("kernel", "instrs") + " ".join("%8s" % p for p in pats))
tot = collections.Counter()
for k, c in cnt.items():
    if any(c[p] for p in pats[:6]):
        print("%-58s %7d " % (k[:58], c["_total"]) + " ".join("%8d" % c[p] for p in pats))
    tot.update(c)
print("%-58s %7d " % ("ALL %d kernels" % len(cnt), tot["_total"]) + " ".join("%8d" % tot[p] for p in pats))
