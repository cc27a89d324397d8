#!/bin/bash
# visit Y: c1s2_bwd builders with the next tile's loads issued pass by pass during the build
out=gpurun_out; mkdir -p $out
echo "[r2y] c1 kernel tests"
timeout 400 python -m pytest tests/test_tc_gpu.py -q -x --timeout 200 -k "c1s2" 2>&1 | tail -4 | cut -c1-300
echo "[r2y] c1 kernels timed alone"
timeout 200 python tools/tc_probe.py c1bwd 2>&1 | tail -9 | cut -c1-200
echo "[r2y] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2y_pytest.log 2>&1; tail -6 $out/r2y_pytest.log | cut -c1-300
for v in "HMGAN_X=0 --workload dcgan" "HMGAN_X=0 --workload dcgan"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2y] bench $v"
  env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2y_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'], d.get('head_alive_frac'))" || tail -5 $out/r2y_bench.err
done > $out/r2y_bench_variants.txt 2>&1; cat $out/r2y_bench_variants.txt | cut -c1-300
echo "[r2y] done"
