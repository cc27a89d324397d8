#!/bin/bash
# visit R: hm_c1s2_wgrad (generator output layer weight gradient), the reordered host path; same-box A/B of both knobs
out=gpurun_out; mkdir -p $out
echo "[r2r] c1wg kernel test"
timeout 300 python -m pytest tests/test_tc_gpu.py -q -x --timeout 200 -k "c1s2" 2>&1 | tail -4 | cut -c1-300
echo "[r2r] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2r_pytest.log 2>&1; tail -8 $out/r2r_pytest.log | cut -c1-300
for v in "HMGAN_C1WG=0 --workload dcgan" "HMGAN_C1WG=1 --workload dcgan" "HMGAN_C1WG=0 --workload dcgan" "HMGAN_C1WG=1 --workload dcgan" "HMGAN_SIDE_D1DW=1 --workload dcgan" "HMGAN_X=0 --workload both" "HMGAN_X=0 --workload p2p"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2r] bench $v"
  env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2r_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'])" || tail -5 $out/r2r_bench.err
done > $out/r2r_bench_variants.txt 2>&1; cat $out/r2r_bench_variants.txt | cut -c1-300
echo "[r2r] done"
