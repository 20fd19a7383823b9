#!/bin/bash
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "folds|query projection of|scaled-fp16|worst case|passed|failed|Error|assert|FAILED" | cut -c1-220 | head -60
