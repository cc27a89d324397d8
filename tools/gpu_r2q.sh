#!/bin/bash
# visit Q: same-box A/B of HMGAN_SIDE_EXTRA (deferred weighted max-pool copy + D layer 1 dW on the side stream), then the
# launch list of one step period with the final kernel set.
out=gpurun_out; mkdir -p $out
for v in "HMGAN_SIDE_EXTRA=0 --workload dcgan" "HMGAN_SIDE_EXTRA=1 --workload dcgan" "HMGAN_SIDE_EXTRA=0 --workload dcgan" "HMGAN_SIDE_EXTRA=1 --workload dcgan" "HMGAN_SIDE_EXTRA=0 --workload both" "HMGAN_SIDE_EXTRA=1 --workload both"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2q] bench $v"
  env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2q_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'])" || tail -5 $out/r2q_bench.err
done > $out/r2q_bench_variants.txt 2>&1; cat $out/r2q_bench_variants.txt | cut -c1-300
echo "[r2q] launch list of one step period"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2q_launches_step.txt 2>&1; head -30 $out/r2q_launches_step.txt
echo "[r2q] done"
