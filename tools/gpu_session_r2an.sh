#!/bin/bash
# Final round-2 session (final code: operand folds, LayerNorm in the output projection's epilogue, tiles-per-CTA policy, opt-in
# scaled-fp16 scatter): full GPU tests, smoke, default bench + reference arm, per-workload in-graph kernel times, launch list.
mkdir -p gpurun_out
run() { echo "== ${*:2}"; timeout -s KILL "$1" "${@:2}"; echo "[rc=$?] ${*:2}"; }
run 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
run 300 python __graft_entry__.py --smoke 2>&1 | tail -2
run 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err; tail -c 300 gpurun_out/r2an_bench.json; echo
run 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2an_bench_reference.json 2> gpurun_out/r2an_bench_reference.err; tail -c 200 gpurun_out/r2an_bench_reference.json; echo
for c in 2 4; do
  run 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps > gpurun_out/r2an_gaps_config$c.txt 2>&1
  grep -v Warning gpurun_out/r2an_gaps_config$c.txt | head -6 | cut -c1-150
done
for c in 2 4; do
  timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2an_launches_config$c.csv python bench.py --config $c --profile-step --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2an_launches_config$c.log 2>&1
  python tools/launch_summary.py gpurun_out/r2an_launches_config$c.csv > gpurun_out/r2an_launches_config$c.txt 2>&1
  head -5 gpurun_out/r2an_launches_config$c.txt | cut -c1-150
done
