#!/bin/bash
# Session j: bidirectional-attention kernels -- tests, timing at model size, ncu summaries.
mkdir -p gpurun_out
run() { echo "== $*"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 300 python -m pytest tests/test_biattn_gpu.py tests/test_fuse_modules_gpu.py -m gpu -q 2>&1 | tail -15
run 300 python tools/bench_biattn.py r2j 2>&1 | tail -20
for k in biattn_pv_kernel biattn_ds_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o gpurun_out/r2j_$k -f python tools/bench_biattn.py r2j_ncu > gpurun_out/r2j_ncu_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2j_$k.ncu-rep > gpurun_out/r2j_ncu_$k.txt 2>&1
  rm -f gpurun_out/r2j_$k.ncu-rep
done
tail -5 gpurun_out/r2j_ncu_biattn_pv_kernel.txt
