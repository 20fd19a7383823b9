#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_layers_gpu.py -m gpu -q -x -s -k "one_launch or in_the_projection" 2>&1 | grep -E "passed|failed|Error|assert|LayerNorm in" | cut -c1-200 | head -20
for m in 0 1; do
MSDA_B200_OUT_LN=$m timeout -s KILL 400 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps > gpurun_out/r2am_gaps_config2_$m.txt 2>&1
grep -E "gaps:|add_ln_fwd|linear_tc_kernel<0" gpurun_out/r2am_gaps_config2_$m.txt | cut -c1-150
done
