#!/bin/bash
# N-GPU session (final code): default bench under torchrun at N = $1.
N=${1:-2}
O=gpurun_out
mkdir -p $O
export BENCH_WATCHDOG_S=300
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/r2n_bench_n$N.json 2> $O/r2n_bench_n$N.err
echo "[rc=$?]"; cut -c1-300 $O/r2n_bench_n$N.json; tail -3 $O/r2n_bench_n$N.err | cut -c1-200
