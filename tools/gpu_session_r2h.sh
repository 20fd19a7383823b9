#!/bin/bash
# 2-GPU session: default bench at N=2, then the captured all-reduce variant (bounded: a hang must not hold the box)
O=gpurun_out
mkdir -p $O
run() { local t=$1; shift; timeout -k 10 $t "$@"; echo "[rc=$?] $*" >> $O/r2h_session.log; }
: > $O/r2h_session.log
export BENCH_WATCHDOG_S=240
run 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2h_bench_n2.json 2> $O/r2h_bench_n2.err
cut -c1-250 $O/r2h_bench_n2.json
run 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --graph-allreduce --config 4 --no-config5 > $O/r2h_bench_n2_graph_allreduce.json 2> $O/r2h_bench_n2_graph_allreduce.err
cut -c1-250 $O/r2h_bench_n2_graph_allreduce.json; tail -5 $O/r2h_bench_n2_graph_allreduce.err | cut -c1-200
run 200 python -m pytest tests/test_dp_gloo.py -q > $O/r2h_dp_test.log 2>&1
cat $O/r2h_session.log
