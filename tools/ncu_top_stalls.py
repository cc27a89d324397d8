"""Top stall locations of one kernel from an ncu report: `ncu -i rep --page source --csv` parsed, rows sorted by warp-stall
samples.  usage: python tools/ncu_top_stalls.py report.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
for view in ("cuda", "sass"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next((r for r in rows if any("Sampl" in c for c in r)), None)
    if hdr is None:
        print("[%s] no sampling columns; first lines:" % view, out[:400])
        continue
    k = rows.index(hdr)
    si = next(i for i, c in enumerate(hdr) if "Sampl" in c and "All" in c) if any(
        "Sampl" in c and "All" in c for c in hdr) else next(i for i, c in enumerate(hdr) if "Sampl" in c)
    src = next((i for i, c in enumerate(hdr) if c.strip() in ("Source", "Source (SASS)", "Source (CUDA-C)")), 1)
    body = []
    for r in rows[k + 1:]:
        try:
            body.append((float(r[si].replace(",", "") or 0), r))
        except (ValueError, IndexError):
            pass
    tot = sum(v for v, _ in body) or 1.0
    print("[%s] column %r, total samples %d" % (view, hdr[si], tot))
    for v, r in sorted(body, key=lambda t: -t[0])[:n]:
        print("%6.2f%%  #%-5s %s" % (100 * v / tot, r[0], r[src][:150]))
