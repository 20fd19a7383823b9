#!/bin/bash
# Session q: bidirectional attention v3 -- full GPU tests, block + kernel timing, default bench line, ncu summaries.
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
run 120 python tools/bench_biattn.py r2q 2>&1 | tail -14 | cut -c1-330

run 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 1500 gpurun_out/r2q_bench.json
for k in biattn_pv_kernel biattn_ds_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 4 -o gpurun_out/r2q_$k -f python tools/bench_biattn.py r2q_ncu > gpurun_out/r2q_ncu_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2q_$k.ncu-rep > gpurun_out/r2q_ncu_$k.txt 2>&1
  rm -f gpurun_out/r2q_$k.ncu-rep
done
grep -c "==" gpurun_out/r2q_ncu_biattn_pv_kernel.txt
