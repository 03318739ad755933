#!/bin/bash
# Multi-GPU pass of round 2 (run under `gpurun --gpus N -- bash tools/gpu_multi_r2.sh N`):
#   host-thread tests, the reference's UNMODIFIED run.sh (it pins device 1: needs >= 2 GPUs) through the driver tests
#   and in place, bench.py at N ranks (weak, strong, reference arm under torchrun), the C++ drivers sharded over N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_threads_gpu.py tests/test_drivers_gpu.py -q -m gpu > gpurun_out/r02_pytest_${N}gpu.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_${N}gpu.txt
$TR --master-port 29517 bench.py --gpus "$N" --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench weak rc=$?"
$TR --master-port 29518 bench.py --gpus "$N" --steps 20 --warmup 5 --scaling strong > gpurun_out/r02_bench_strong_${N}gpu.json 2> gpurun_out/r02_bench_strong_${N}gpu.err
echo "bench strong rc=$?"
$TR --master-port 29519 bench.py --gpus "$N" --steps 10 --warmup 3 --impl reference > gpurun_out/r02_bench_ref_${N}gpu.json 2> gpurun_out/r02_bench_ref_${N}gpu.err
echo "bench reference rc=$?"
B200FE_NGPUS=$N B200FE_NELMT=2097152 ./benchmark05/build/benchmark05 8 8 8 > gpurun_out/r02_driver_b05_${N}gpu.txt 2>&1
echo "b05 rc=$?"
B200FE_NGPUS=$N B200FE_NELMT=$((4194304 * N)) ./benchmark04/build/benchmark04 4 4 > gpurun_out/r02_driver_b04_${N}gpu.txt 2>&1
echo "b04 rc=$?"
if [ "$N" = "2" ] && [ "${RUNSH:-1}" = "1" ]; then
  # the reference's run.sh in place, byte for byte; logs land next to it exactly as its README describes.
  # B200FE_CPU_REPS=1 only shortens the host columns; the script itself is untouched.
  for b in benchmark04 benchmark05; do
    ( cd $b && B200FE_CPU_REPS=1 timeout 1200 bash run.sh > ../gpurun_out/r02_runsh_$b.out 2>&1; echo "$b run.sh rc=$?" )
    mkdir -p gpurun_out/r02_runsh_logs/$b && cp $b/nq*.log gpurun_out/r02_runsh_logs/$b/ 2>/dev/null
  done
fi
