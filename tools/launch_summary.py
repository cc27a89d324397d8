"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family and the
slowest individual launches (of the last complete training step when the capture holds several).  usage: python tools/launch_summary.py gpurun_out/launches.csv [min_ms] [top]"""
import collections
import csv
import re
import sys


def main(path, min_ms=1.0, top=20):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size'))
    rows = [row for row in r if len(row) > vi]
    # a training step ends with the optimiser launches: keep only the last complete step when there are several
    marks = [i for i, row in enumerate(rows) if 'rmsprop' in row[ki] or 'adam_kernel' in row[ki]]
    # a step starts with the cast of the latent batch into the generator's input buffer (one cast_kernel<float, __half>
    # per step; the optimiser launches are spread over two streams and no longer mark a boundary)
    starts = [i for i, row in enumerate(rows) if 'cast_kernel<float' in row[ki]]
    if len(starts) < 2:
        starts = [i for k, i in enumerate(marks) if k == 0 or i - marks[k - 1] > 60]
    if len(starts) >= 2:
        rows = rows[starts[-2]:starts[-1]]
        print('last full training step period (%d of %d launches)' % (len(rows), starts[-1]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, big = 0.0, []
    for row in rows:
        v = float(row[vi].replace(',', ''))
        v = v / 1e6 if row[ui] in ('nsecond', 'ns') else (v / 1e3 if row[ui] in ('usecond', 'us') else v)
        name = re.sub(r'\(.*', '', row[ki]).replace('void ', '')
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        if v >= min_ms:
            big.append((v, row[gi], name))
    print('total %.2f ms over %d launches' % (tot, sum(n for n, _ in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%8.2f ms %5.1f%% n=%4d  %s' % (t, 100 * t / tot, n, k))
    print('launches >= %.1f ms:' % min_ms)
    for b in big:
        print('  %8.2f ms  grid=%-16s %s' % b)


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0, int(sys.argv[3]) if len(sys.argv) > 3 else 20)
