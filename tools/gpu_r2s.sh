#!/bin/bash
# visit S: the U-Net bottleneck deconvolution as tensor-core GEMMs (HMGAN_DC1), full GPU suite, same-box A/B
out=gpurun_out; mkdir -p $out
echo "[r2s] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2s_pytest.log 2>&1; tail -8 $out/r2s_pytest.log | cut -c1-300
for v in "HMGAN_DC1=0 --workload p2p" "HMGAN_DC1=1 --workload p2p" "HMGAN_DC1=0 --workload p2p" "HMGAN_DC1=1 --workload p2p" "HMGAN_DC1=0 --workload both" "HMGAN_DC1=1 --workload both"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2s] bench $v"
  env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2s_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'], d.get('head_alive_frac'))" || tail -5 $out/r2s_bench.err
done > $out/r2s_bench_variants.txt 2>&1; cat $out/r2s_bench_variants.txt | cut -c1-300
echo "[r2s] done"
