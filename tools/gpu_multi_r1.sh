#!/bin/bash
# Multi-GPU pass of round 1 (run under `gpurun --gpus N -- bash tools/gpu_multi_r1.sh N`):
#   host-thread tests (one thread per device), bench.py at N ranks, the C++ drivers sharded over N GPUs
#   (hex nq=8 at the 1 Gi-point size of configs[4]; quad nq=4, a bank-fed kernel, at 2 x 4 Mi elements on 2 GPUs).
set -u
N=${1:-4}
mkdir -p gpurun_out
python -m pytest tests/test_threads_gpu.py -q -m gpu > gpurun_out/threads_${N}gpu.log 2>&1
echo "threads rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$N" --no-sweep > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"
B200FE_NGPUS=$N B200FE_NELMT=2097152 ./benchmark05/build/benchmark05 8 8 8 > gpurun_out/driver_b05_${N}gpu.txt 2>&1
echo "b05 rc=$?"
B200FE_NGPUS=2 B200FE_NELMT=8388608 ./benchmark04/build/benchmark04 4 4 > gpurun_out/driver_b04_2gpu.txt 2>&1
echo "b04 (2 GPUs) rc=$?"
B200FE_NGPUS=$N B200FE_NELMT=$((4194304 * N)) ./benchmark04/build/benchmark04 4 4 > gpurun_out/driver_b04_${N}gpu.txt 2>&1
echo "b04 rc=$?"
tail -2 gpurun_out/threads_${N}gpu.log
