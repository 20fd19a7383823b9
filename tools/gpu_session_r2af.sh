#!/bin/bash
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 300 python tools/bench_f16acc.py r2af_f16acc 2>&1 | grep -v Warning | cut -c1-900
run 400 python -m pytest tests/test_msda_gpu.py tests/test_module_gpu.py -m gpu -q -k "f16acc or scaled_fp16" 2>&1 | tail -3
