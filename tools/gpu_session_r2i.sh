#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout -k 10 $t "$@"; echo "[rc=$?] $*" >> $O/r2i_session.log; }
: > $O/r2i_session.log
run 600 python -m pytest tests -m gpu -q > $O/r2i_tests_full.log 2>&1
tail -3 $O/r2i_tests_full.log | cut -c1-200
run 600 python tools/bench_configs.py r2 > $O/r2i_bench_configs.log 2>&1
cp $O/configs_r2.jsonl $O/r2i_configs_1_3_5.jsonl 2>/dev/null; cat $O/r2i_configs_1_3_5.jsonl | cut -c1-400
run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2i_bench_c2_minb3.json 2> $O/r2i_bench_c2_minb3.err
run 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2i_bench_c4_minb3.json 2> $O/r2i_bench_c4_minb3.err
# A/B: scatter kernel compiled for 4 resident CTAs per SM (64 registers)
cp ziragroundingdino_b200/_lib/libmsda_b200.so /tmp/libmsda_b200.keep
MSDA_NVCC_EXTRA=-DMSDA_BWD_MINB=4 run 400 bash ziragroundingdino_b200/csrc/build.sh > $O/r2i_build_minb4.log 2>&1
run 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $O/r2i_bench_c2_minb4.json 2> $O/r2i_bench_c2_minb4.err
run 300 python -m pytest tests/test_msda_gpu.py tests/test_gemm_gpu.py -m gpu -q -x > $O/r2i_tests_minb4.log 2>&1
tail -2 $O/r2i_tests_minb4.log | cut -c1-200
cp /tmp/libmsda_b200.keep ziragroundingdino_b200/_lib/libmsda_b200.so
for f in $O/r2i_bench_c*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "value %.1f ms %.2f"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["msda_core_us_per_layer"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
cat $O/r2i_session.log
