#!/bin/bash
# Final round-2 session: full GPU tests, smoke, default bench + reference arm, per-workload in-graph kernel times,
# fusion-block timings, ncu summaries of the attention / FFN kernels.
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
run 300 python __graft_entry__.py --smoke 2>&1 | tail -2
run 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 600 gpurun_out/r2z_bench.json; echo
run 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; tail -c 400 gpurun_out/r2z_bench_reference.json; echo
for c in 2 4; do
  run 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps > gpurun_out/r2z_gaps_config$c.txt 2>&1
  grep -v Warning gpurun_out/r2z_gaps_config$c.txt | head -8 | cut -c1-150
done
run 200 python tools/bench_biattn.py r2z > gpurun_out/r2z_biattn.log 2>&1; tail -12 gpurun_out/r2z_biattn.log | cut -c1-200
for k in biattn_pv_kernel biattn_ds_kernel biattn_tn_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 4 -o gpurun_out/r2z_$k -f python tools/bench_biattn.py r2z_ncu > gpurun_out/r2z_ncu_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2z_$k.ncu-rep > gpurun_out/r2z_ncu_$k.txt 2>&1
  rm -f gpurun_out/r2z_$k.ncu-rep
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:ffn_chain2_kernel -c 2 -o gpurun_out/r2z_ffn -f python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-config5 --no-fusion --no-graph > gpurun_out/r2z_ncu_ffn.log 2>&1
python tools/ncu_summary.py gpurun_out/r2z_ffn.ncu-rep > gpurun_out/r2z_ncu_ffn_chain.txt 2>&1
rm -f gpurun_out/r2z_ffn.ncu-rep
grep -c "==" gpurun_out/r2z_ncu_*.txt
