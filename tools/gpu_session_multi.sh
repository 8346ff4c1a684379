#!/bin/bash
# Multi-GPU check under `gpurun --gpus N`: the rank-equality test, then the bench line at N ranks.
# Usage: bash tools/gpu_session_multi.sh N [notest]
set -u
N=${1:-2}
O=gpurun_out; mkdir -p $O
if [ "${2:-}" != "notest" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $O/r2_mgpu_tests_n$N.log 2>&1; echo "mgpu tests rc=$?"; tail -3 $O/r2_mgpu_tests_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; echo "bench n$N rc=$?"
head -c 500 $O/r2_bench_n$N.json; echo
