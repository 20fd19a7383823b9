#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -c 6 -o gpurun_out/r2ap_gemm -f python tools/prof_f16acc.py > gpurun_out/r2ap_ncu_gemm.log 2>&1
python tools/ncu_summary.py gpurun_out/r2ap_gemm.ncu-rep > gpurun_out/r2ap_ncu_gemm.txt 2>&1
rm -f gpurun_out/r2ap_gemm.ncu-rep
grep -E "^==|gpu__time_duration|tensor_cycles|dram__bytes|registers_per_thread" gpurun_out/r2ap_ncu_gemm.txt | cut -c1-160
