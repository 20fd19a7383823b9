#!/bin/bash
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 300 python tools/bench_f16acc.py r2ab_f16acc 2>&1 | grep -v Warning | cut -c1-300
run 400 python -m pytest tests/test_msda_gpu.py tests/test_module_gpu.py -m gpu -q -k "f16acc or scaled_fp16" -s 2>&1 | grep -E "f16acc|scaled|worst|passed|failed|Error|assert" | cut -c1-220 | head -40
for c in 2 4; do
  for m in 0 1; do
    MSDA_B200_F16ACC=$m run 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2ab_bench_c${c}_f16acc$m.json 2> gpurun_out/r2ab_bench_c${c}_f16acc$m.err
    python -c "import json,sys; d=json.loads(open('gpurun_out/r2ab_bench_c${c}_f16acc$m.json').read().strip().splitlines()[-1]); print('config $c f16acc=$m', d['value'], d['ms_per_step'])"
  done
done
