#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout $t "$@"; echo "[rc=$?] $*" >> $O/r2c_session.log; }
: > $O/r2c_session.log
run 600 python -m pytest tests -m gpu -q -s > $O/r2c_tests_full.log 2>&1
tail -8 $O/r2c_tests_full.log | cut -c1-200; grep "bwd_dots" $O/r2c_tests_full.log | head -12
MSDA_B200_TUNING=bwd_dots=1 run 600 python -m pytest tests -m gpu -q -x > $O/r2c_tests_dots.log 2>&1
tail -4 $O/r2c_tests_dots.log | cut -c1-200
run 400 python tools/debug_mma_scatter.py --time > $O/r2c_debug_mma.log 2>&1
grep regime $O/r2c_debug_mma.log
for v in "bwd_dots=1" "bwd_dots=1,bwd_mma=0" "bwd_dots=1,bwd_mma_levels=2"; do
  MSDA_B200_TUNING=$v MSDA_B200_FFN_CHAIN=1 run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > "$O/r2c_bench_c2_$v.json" 2> "$O/r2c_bench_c2_$v.err"
done
MSDA_B200_TUNING=bwd_dots=1 MSDA_B200_FFN_CHAIN=1 run 300 python bench.py --gaps --config 2 2> $O/r2c_gaps_c2.txt
MSDA_B200_TUNING=bwd_dots=1 MSDA_B200_FFN_CHAIN=1 run 300 python bench.py --gaps --config 4 2> $O/r2c_gaps_c4.txt
for f in $O/r2c_bench_c2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.1f ms %.2f"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["msda_core_us_per_layer"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
head -14 $O/r2c_gaps_c2.txt | cut -c1-130; head -14 $O/r2c_gaps_c4.txt | cut -c1-130
cat $O/r2c_session.log
