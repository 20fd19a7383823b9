#!/bin/bash
# A/B of the chained FFN: one CTA per row tile (MSDA_B200_FFN_PAIR=0) vs CTA pairs (=1): layer tests + in-graph kernel times.
for pair in 0 1; do
  export MSDA_B200_FFN_PAIR=$pair
  echo "== FFN_PAIR=$pair"
  timeout -s KILL 150 python -m pytest tests/test_layers_gpu.py -m gpu -q 2>&1 | tail -2
  timeout -s KILL 400 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion --gaps 2>&1 | grep -v Warning | grep "ffn_chain\|add_ln_fwd\|gaps:" | cut -c1-130
done
