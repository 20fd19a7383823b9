#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout $t "$@"; echo "[rc=$?] $*" >> $O/r2g_session.log; }
: > $O/r2g_session.log
run 600 python -m pytest tests -m gpu -q > $O/r2g_tests_full.log 2>&1
tail -3 $O/r2g_tests_full.log | cut -c1-200
MSDA_B200_TUNING=pk2=1 run 600 python -m pytest tests -m gpu -q > $O/r2g_tests_pk2.log 2>&1
tail -3 $O/r2g_tests_pk2.log | cut -c1-200
for v in "" "pk2=1"; do
  MSDA_B200_TUNING=$v run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > "$O/r2g_bench_c2_$v.json" 2> "$O/r2g_bench_c2_$v.err"
done
MSDA_B200_TUNING=pk2=1 run 300 python bench.py --gaps --config 2 2> $O/r2g_gaps_c2_pk2.txt
for f in $O/r2g_bench_c2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.1f ms %.2f"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["msda_core_us_per_layer"].items() if isinstance(v,float)}, "refAB", d.get("ref_cuda_us_per_layer"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
head -8 $O/r2g_gaps_c2_pk2.txt | cut -c1-130
cat $O/r2g_session.log
