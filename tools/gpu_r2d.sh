#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2d] side stream diagnostics"
( HMGAN_FORK=0 timeout 300 python tools/dbg_side_stream.py wide64 fast
  HMGAN_FORK=0 HMGAN_CUDA_GRAPHS=0 CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/dbg_side_stream.py wide64 fast ) > $out/r2d_side.log 2>&1
grep -c "val\|param" $out/r2d_side.log
echo "[r2d] update-direction measurement"
timeout 900 python tools/dbg_tc32.py precision=tc32,fast dcgan512 B=8 bias=0.6 steps=3 lr=1e-4 > $out/r2d_steps.log 2>&1; grep "params\|losses ours" $out/r2d_steps.log
echo "[r2d] pytest -m gpu (defaults: fork on, deferred wgrad join is moot without the side stream)"
timeout 900 python -m pytest tests -m gpu -q -rf > $out/r2d_pytest_default.log 2>&1; tail -8 $out/r2d_pytest_default.log | cut -c1-300
echo "[r2d] pytest -m gpu, variants on"
HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 HMGAN_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -rf > $out/r2d_pytest_all.log 2>&1
tail -8 $out/r2d_pytest_all.log | cut -c1-300
for v in "HMGAN_FORK=0" "HMGAN_FORK=1" "HMGAN_FORK=0 HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1" "HMGAN_FORK=1 HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1"; do
  echo "[r2d] bench with $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2])"
done > $out/r2d_bench_variants.txt 2>&1; cat $out/r2d_bench_variants.txt
echo "[r2d] done"
