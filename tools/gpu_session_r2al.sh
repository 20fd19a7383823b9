#!/bin/bash
# compute-sanitizer memcheck over the tests of the last session's kernels (scaled-fp16 scatter + consumer, two-operand GEMMs,
# two-row add+LayerNorm, tiles-per-CTA policy); the config-2-sized cases are deselected for run time.
mkdir -p gpurun_out
timeout -s KILL 800 compute-sanitizer --tool memcheck --log-file gpurun_out/r2al_sanitizer.log python -m pytest \
  tests/test_msda_gpu.py tests/test_layers_gpu.py tests/test_module_gpu.py -m gpu -q \
  -k "(f16acc and not full_size) or scaled_fp16 or accum2 or query_projection_of or operand_folds or encoder_layer_block_functions or fused_query_backward" \
  > gpurun_out/r2al_sanitizer_pytest.txt 2>&1
tail -3 gpurun_out/r2al_sanitizer_pytest.txt
grep -n "ERROR SUMMARY\|Invalid\|out of bounds\|Program hit" gpurun_out/r2al_sanitizer.log | head -12
