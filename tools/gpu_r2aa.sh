#!/bin/bash
# launch list (ncu, serialised) of one replayed pix2pix step and one joint step
out=gpurun_out; mkdir -p $out
for wl in p2p both; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/launches_$wl.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --workload $wl > /tmp/launches_$wl.log 2>&1
  python tools/launch_summary.py /tmp/launches_$wl.csv 0.15 32 > $out/r2aa_launches_$wl.txt 2>&1; head -45 $out/r2aa_launches_$wl.txt
done
echo "[r2aa] done"
