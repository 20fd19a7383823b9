#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout $t "$@"; echo "[rc=$?] $*" >> $O/r2f_session.log; }
: > $O/r2f_session.log
run 600 python -m pytest tests -m gpu -q > $O/r2f_tests_full.log 2>&1
tail -3 $O/r2f_tests_full.log | cut -c1-200
run 900 python bench.py --steps 20 --warmup 5 > $O/r2f_bench.json 2> $O/r2f_bench.err
run 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err
run 300 python bench.py --gaps --config 2 2> $O/r2f_gaps_c2.txt
run 300 python bench.py --gaps --config 4 2> $O/r2f_gaps_c4.txt
run 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2f_launches_c2.csv python bench.py --profile-step --config 2 > /dev/null 2>&1
run 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2f_launches_c4.csv python bench.py --profile-step --config 4 > /dev/null 2>&1
# full ncu captures, one kernel family per report; summarised HERE (the reports of the big kernels are 25+ MB), then dropped
for k in msda_bwd_vec msda_fwd_vec ffn_chain linear_tf32x3 flatten_levels encoder_proposals; do
  run 400 ncu --set full --clock-control none -k regex:$k -c 2 -s 1 -o $O/r2f_$k python tools/prof_r2.py > $O/r2f_ncu_$k.log 2>&1
  python tools/ncu_summary.py $O/r2f_$k.ncu-rep > $O/r2f_ncu_summary_$k.txt 2>&1
  rm -f $O/r2f_$k.ncu-rep
done
MSDA_B200_TUNING=bwd_mma=1 run 400 ncu --set full --clock-control none -k regex:'msda_bwd_vec|msda_scatter_mma' -c 2 -s 2 -o $O/r2f_scatter_mma python tools/prof_r2.py > $O/r2f_ncu_scatter_mma.log 2>&1
python tools/ncu_summary.py $O/r2f_scatter_mma.ncu-rep > $O/r2f_ncu_summary_scatter_mma.txt 2>&1
rm -f $O/r2f_scatter_mma.ncu-rep
run 300 python tools/parity_table.py > $O/r2f_parity_table.txt 2>&1
run 300 python tools/debug_mma_scatter.py --time > $O/r2f_debug_mma.log 2>&1
du -sh $O
cat $O/r2f_session.log
