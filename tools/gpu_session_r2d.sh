#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout $t "$@"; echo "[rc=$?] $*" >> $O/r2d_session.log; }
: > $O/r2d_session.log
run 600 python -m pytest tests -m gpu -q -s > $O/r2d_tests_full.log 2>&1
tail -8 $O/r2d_tests_full.log | cut -c1-200
run 300 python tools/debug_lpg16.py > $O/r2d_debug_lpg16.log 2>&1
cat $O/r2d_debug_lpg16.log | cut -c1-250
run 900 python bench.py --steps 20 --warmup 5 > $O/r2d_bench.json 2> $O/r2d_bench.err
cut -c1-300 $O/r2d_bench.json; tail -3 $O/r2d_bench.err
run 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2d_bench_reference.json 2> $O/r2d_bench_reference.err
# launch list of one eager step (config 2 and config 4), device time per launch
run 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2d_launches_c2.csv python bench.py --profile-step --config 2 > /dev/null 2>&1
run 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2d_launches_c4.csv python bench.py --profile-step --config 4 > /dev/null 2>&1
# full ncu capture of the round-2 kernels
run 900 ncu --set full --clock-control none --import-source on -k regex:'msda_bwd_vec|msda_fwd_vec|ffn_chain|linear_tf32x3|flatten_levels|encoder_proposals' -o $O/r2d_kernels python tools/prof_r2.py > $O/r2d_ncu_kernels.log 2>&1
MSDA_B200_TUNING=bwd_mma=1 run 600 ncu --set full --clock-control none --import-source on -k regex:'msda_bwd_vec|msda_scatter_mma' -c 4 -o $O/r2d_kernels_mma python tools/prof_r2.py > $O/r2d_ncu_kernels_mma.log 2>&1
ls -la $O/*.ncu-rep
cat $O/r2d_session.log
