#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2p] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2p_pytest.log 2>&1; tail -8 $out/r2p_pytest.log | cut -c1-300
for v in "HMGAN_PACK_TILED=0 --workload dcgan" "HMGAN_PACK_TILED=1 --workload dcgan" "HMGAN_PACK_TILED=0 HMGAN_FORK_P2P=0 --workload both" "HMGAN_PACK_TILED=1 HMGAN_FORK_P2P=0 --workload both" "HMGAN_PACK_TILED=1 HMGAN_FORK_P2P=1 --workload both" "HMGAN_PACK_TILED=0 --workload p2p" "HMGAN_PACK_TILED=1 --workload p2p"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2p] bench $v"
  env $envs timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2p_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'])" || tail -5 $out/r2p_bench.err
done > $out/r2p_bench_variants.txt 2>&1; cat $out/r2p_bench_variants.txt | cut -c1-300
echo "[r2p] done"
