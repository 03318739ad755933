#!/bin/bash
# round-1 profiling pass (run under gpurun): launch lists + full ncu captures of the dominant kernels
set -u
mkdir -p gpurun_out
export B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3
prof() { # name, kernel regex, env..., -- command
  name=$1; regex=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed: $name"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 6 -c 1 -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "profiled $name rc=$?"
}
# bench.py launch list (same command, without and then with ncu)
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
B200FE_NELMT=262144 prof hex8_f64_mma bwdtrans_hex_ benchmark05/build/benchmark05 8 8 8
B200FE_NELMT=262144 prof quad16_f64_mma bwdtrans_quad_ benchmark04/build/benchmark04 16 16
B200FE_NELMT=65536 prof quad32_f64_mma bwdtrans_quad_ benchmark04/build/benchmark04 32 32
B200FE_NELMT=65536 B200FE_DTYPE=float prof quad32_f32_mma bwdtrans_quad_ benchmark04/build/benchmark04 32 32
B200FE_NELMT=67104 prof hex10_f64 bwdtrans_hex_ benchmark05/build/benchmark05 10 10 10
B200FE_NELMT=67104 B200FE_DTYPE=float prof hex10_f32 bwdtrans_hex_ benchmark05/build/benchmark05 10 10 10
ls -la gpurun_out
