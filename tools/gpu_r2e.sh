#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2e] determinism of the variants (off vs off should be ~1e-7 everywhere)"
( echo "##### HOIST only"; HMGAN_FORK=0 HMGAN_EW_HOIST=1 timeout 300 python tools/dbg_side_stream.py wide64 fast
  echo "##### SPLITK only"; HMGAN_FORK=0 HMGAN_TC_SPLITK=1 timeout 300 python tools/dbg_side_stream.py wide64 fast ) > $out/r2e_determinism.log 2>&1
grep "#####\|==\|param" $out/r2e_determinism.log | awk '{ if ($0 ~ /#####|==/) print; else if ($NF+0 > 1e-4) print }' | cut -c1-200
echo "[r2e] bench default (live head, secondary workloads)"
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2e_bench.json 2> $out/r2e_bench.err; tail -c 1500 $out/r2e_bench.json; tail -3 $out/r2e_bench.err
for v in "HMGAN_X=1 --head-bias 0" "HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 --head-bias 0" "HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 --head-bias 0.6" "HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 --head-bias 0.6"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2e] bench $v"
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary $args 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'])"
done > $out/r2e_bench_variants.txt 2>&1; cat $out/r2e_bench_variants.txt
echo "[r2e] reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | cut -c1-900
echo "[r2e] done"
