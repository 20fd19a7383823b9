#!/bin/bash
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 300 python tools/bench_f16acc.py r2aa_f16acc 2>&1 | grep -v Warning | cut -c1-900
run 400 python -m pytest tests/test_msda_gpu.py -m gpu -q -k f16acc -s 2>&1 | grep -E "f16acc|passed|failed|Error|assert" | cut -c1-220 | head -60
