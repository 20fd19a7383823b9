#!/bin/bash
# Final round-2 session (after the operand folds and the opt-in scaled-fp16 scatter): full GPU tests, smoke, default bench +
# reference arm, per-workload in-graph kernel times, ncu summaries of the new kernel variants.
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
run 300 python __graft_entry__.py --smoke 2>&1 | tail -2
run 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; tail -c 500 gpurun_out/r2aj_bench.json; echo
run 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2aj_bench_reference.json 2> gpurun_out/r2aj_bench_reference.err; tail -c 300 gpurun_out/r2aj_bench_reference.json; echo
for c in 2 4; do
  run 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps > gpurun_out/r2aj_gaps_config$c.txt 2>&1
  grep -v Warning gpurun_out/r2aj_gaps_config$c.txt | head -8 | cut -c1-150
done
run 200 python tools/bench_f16acc.py r2aj_f16acc 2>&1 | grep -v Warning | cut -c1-200
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_vec -c 2 -o gpurun_out/r2aj_f16acc -f python tools/prof_f16acc.py > gpurun_out/r2aj_ncu_f16acc.log 2>&1
python tools/ncu_summary.py gpurun_out/r2aj_f16acc.ncu-rep > gpurun_out/r2aj_ncu_f16acc.txt 2>&1
rm -f gpurun_out/r2aj_f16acc.ncu-rep
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -c 4 -o gpurun_out/r2aj_gemm2 -f python tools/prof_f16acc.py > gpurun_out/r2aj_ncu_gemm2.log 2>&1
python tools/ncu_summary.py gpurun_out/r2aj_gemm2.ncu-rep > gpurun_out/r2aj_ncu_gemm2.txt 2>&1
rm -f gpurun_out/r2aj_gemm2.ncu-rep
wc -l gpurun_out/r2aj_ncu_*.txt
