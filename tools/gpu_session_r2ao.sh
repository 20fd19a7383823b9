#!/bin/bash
# ncu launch lists of ONE eager step per workload (final code); cudaProfilerStart/Stop bracket the step.
mkdir -p gpurun_out
for c in 2 4; do
  timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ao_launches_config$c.csv python bench.py --config $c --profile-step --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2ao_launches_config$c.log 2>&1
  python tools/launch_summary.py gpurun_out/r2ao_launches_config$c.csv > gpurun_out/r2ao_launches_config$c.txt 2>&1
  head -8 gpurun_out/r2ao_launches_config$c.txt | cut -c1-150
  rm -f gpurun_out/r2ao_launches_config$c.csv
done
