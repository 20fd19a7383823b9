#!/bin/bash
mkdir -p gpurun_out
for c in 2 4; do
  for t in "bwd_passes=1" "bwd_passes=2" "bwd_passes=4" "bwd_passes=2,fwd_passes=2"; do
    MSDA_B200_TUNING=$t timeout -s KILL 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2ah_tmp.json 2> gpurun_out/r2ah_tmp.err
    python - <<PY
import json
for line in open('gpurun_out/r2ah_tmp.json'):
    if line.startswith('{"metric"'):
        d = json.loads(line); print('config $c $t', round(d['value'],2), round(d['ms_per_step'],3), d['msda_core_us_per_layer']['fwd'], d['msda_core_us_per_layer']['bwd'])
PY
  done
done
MSDA_B200_F16ACC=1 MSDA_B200_TUNING=bwd_passes=4 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 --no-fusion > gpurun_out/r2ah_tmp.json 2> gpurun_out/r2ah_tmp.err
python - <<PY
import json
for line in open('gpurun_out/r2ah_tmp.json'):
    if line.startswith('{"metric"'):
        d = json.loads(line); print('config 4 f16acc passes=4', round(d['value'],2), round(d['ms_per_step'],3))
PY
