#!/bin/bash
# One GPU-box session of round 2: parity first, then the bench line, then A/B runs and probes.  Everything is bounded by
# `timeout` (a hung kernel must not hold the box) and writes under gpurun_out/ (merged back by gpurun).
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout $t "$@"; echo "[rc=$?] $*" >> $O/r2_session.log; }
: > $O/r2_session.log
run 420 python -m pytest tests/test_msda_gpu.py -m gpu -q -s -k "mma" > $O/r2_tests_mma.log 2>&1
tail -4 $O/r2_tests_mma.log
run 300 python tools/debug_mma_scatter.py --time > $O/r2_debug_mma.log 2>&1
tail -30 $O/r2_debug_mma.log
run 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench1.json 2> $O/r2_bench1.err
cut -c1-400 $O/r2_bench1.json
run 1500 python -m pytest tests -m gpu -q -s > $O/r2_tests_full.log 2>&1
tail -4 $O/r2_tests_full.log
for lv in 0 2 3 4; do
  MSDA_B200_TUNING=bwd_mma_levels=$lv run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2_bench_c2_mma2_lv$lv.json 2> $O/r2_bench_c2_mma2_lv$lv.err
done
MSDA_B200_TUNING=tap_share=1 run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2_bench_c2_tapshare.json 2> $O/r2_bench_c2_tapshare.err
MSDA_B200_TUNING=tap_share=1,bwd_mma_levels=4 run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2_bench_c2_tapshare_lv4.json 2> $O/r2_bench_c2_tapshare_lv4.err
MSDA_B200_OWN_K_FFN=1 run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2_bench_c2_ownkffn.json 2> $O/r2_bench_c2_ownkffn.err
MSDA_B200_TUNING=bwd_mma=0 run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2_bench_c2_nomma.json 2> $O/r2_bench_c2_nomma.err
run 300 python bench.py --gaps --config 4 2> $O/r2_gaps_c4.txt
run 300 python bench.py --gaps --config 2 2> $O/r2_gaps_c2.txt
MSDA_B200_TUNING=bwd_mma=0 run 300 python bench.py --gaps --config 2 2> $O/r2_gaps_c2_nomma.txt
MSDA_B200_TUNING=bwd_mma_levels=4 run 300 python bench.py --gaps --config 2 2> $O/r2_gaps_c2_mma2_lv4.txt
run 300 python tools/probe_scatter.py > $O/r2_probe_scatter.log 2>&1
run 600 python tools/parity_table.py > $O/r2_parity_table.txt 2>&1
cat $O/r2_session.log
