#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2c] step diagnostics"
( timeout 600 python tools/dbg_tc32.py precision=parity,tc32 wide64 B=2
  timeout 600 python tools/dbg_tc32.py precision=parity,tc32 wide64 B=8
  timeout 900 python tools/dbg_tc32.py precision=parity,tc32,fast dcgan512 B=2 bias=0.6
  timeout 900 python tools/dbg_tc32.py precision=tc32,fast dcgan512 B=8 bias=0.6 steps=3
  timeout 900 python tools/dbg_tc32.py precision=parity,tc32 joint512 B=2 bias=0.6
  timeout 900 python tools/dbg_tc32.py precision=tc32,fast joint512 B=6 bias=0.6 ) > $out/r2c_steps.log 2>&1
grep -c "==" $out/r2c_steps.log
echo "[r2c] side-stream test, all variants on"
HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 HMGAN_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -rf > $out/r2c_pytest_all.log 2>&1
tail -12 $out/r2c_pytest_all.log | cut -c1-300
echo "[r2c] default pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf > $out/r2c_pytest_default.log 2>&1
tail -8 $out/r2c_pytest_default.log | cut -c1-300
echo "[r2c] done"
