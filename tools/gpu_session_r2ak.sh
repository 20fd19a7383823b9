#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_layers_gpu.py tests/test_fuse_modules_gpu.py tests/test_biattn_gpu.py -m gpu -q 2>&1 | tail -3
timeout -s KILL 400 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps > gpurun_out/r2ak_gaps_config2.txt 2>&1
grep -E "gaps:|add_ln|msda_bwd" gpurun_out/r2ak_gaps_config2.txt | cut -c1-150
timeout -s KILL 300 python tools/sweep_fwd_passes.py 2>&1 | grep -v Warn
