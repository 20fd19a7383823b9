#!/bin/bash
# Re-capture of the shipped scatter (two unit tiles per CTA) and gather for the committed ncu summaries / roofline.traffic.
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:'msda_bwd_vec|msda_fwd_vec' -c 4 -o gpurun_out/r2aq_core -f python tools/prof_r2.py > gpurun_out/r2aq_ncu_core.log 2>&1
python tools/ncu_summary.py gpurun_out/r2aq_core.ncu-rep > gpurun_out/r2aq_ncu_core.txt 2>&1
rm -f gpurun_out/r2aq_core.ncu-rep
grep -E "^==|gpu__time_duration|grid_size|dram__bytes|issue_active|l1tex__throughput|lts__throughput" gpurun_out/r2aq_ncu_core.txt | cut -c1-150
