#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2g] up2 backward probes"
timeout 600 python tools/tc_probe.py up2bwd > $out/r2g_up2bwd.log 2>&1; cat $out/r2g_up2bwd.log | cut -c1-200
echo "[r2g] pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/r2g_pytest.log 2>&1; tail -10 $out/r2g_pytest.log | cut -c1-300
for v in "HMGAN_DG6=0 HMGAN_WG8=0" "HMGAN_DG6=1 HMGAN_WG8=0" "HMGAN_DG6=1 HMGAN_WG8=1"; do
  echo "[r2g] bench $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'])"
done > $out/r2g_bench_variants.txt 2>&1; cat $out/r2g_bench_variants.txt
echo "[r2g] launch list of one step period (serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2g_launches_step.txt 2>&1; head -60 $out/r2g_launches_step.txt
echo "[r2g] per-op times of an eager step"
timeout 600 python tools/op_times.py > $out/r2g_op_times.txt 2>&1; head -50 $out/r2g_op_times.txt
echo "[r2g] done"
