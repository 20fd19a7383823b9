#!/bin/bash
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 600 python -m pytest tests/test_layers_gpu.py tests/test_gemm_gpu.py tests/test_module_gpu.py -m gpu -q -x -s 2>&1 | grep -E "folds|query projection of|scaled-fp16|passed|failed|Error|assert" | cut -c1-220 | head -40
for c in 2 4; do
  for m in 0 1; do
    MSDA_B200_FOLD_POS=$m MSDA_B200_DGRAD_CAT=$m run 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2ac_bench_c${c}_folds$m.json 2> gpurun_out/r2ac_bench_c${c}_folds$m.err
    python - <<PY
import json
for line in open('gpurun_out/r2ac_bench_c${c}_folds$m.json'):
    if line.startswith('{"metric"'):
        d = json.loads(line); print('config $c folds=$m', d['value'], d['ms_per_step'], d.get('errors'))
PY
  done
done
