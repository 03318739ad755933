#!/bin/bash
# round-1 profiling pass for the lanes kernels (run under gpurun): one full ncu capture per kernel, each only
# after the same command has exited 0 without ncu
set -u
mkdir -p gpurun_out
export B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3
prof() { # name, kernel regex, skip, env..., -- command
  name=$1; regex=$2; skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed: $name"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "profiled $name rc=$?"
}
B200FE_NELMT=262144 prof quad16_f64_lanesem lanesem_kernel 6 benchmark04/build/benchmark04 16 16
B200FE_NELMT=342368 B200FE_DTYPE=float prof quad14_f32_lanesem lanesem_kernel 6 benchmark04/build/benchmark04 14 14
B200FE_NELMT=262144 prof quad16_f64_lanes_coa quad_lanes_kernel 2 benchmark04/build/benchmark04 16 16
B200FE_NELMT=131072 B200FE_DTYPE=float prof hex8_f32_lanesem lanesem_kernel 6 benchmark05/build/benchmark05 8 8 8
B200FE_NELMT=131072 prof hex8_f64_lanes_coa hex_lanes_kernel 2 benchmark05/build/benchmark05 8 8 8
B200FE_NELMT=67104 B200FE_DTYPE=float prof hex10_f32_lanesem lanesem_kernel 6 benchmark05/build/benchmark05 10 10 10
ls -la gpurun_out | head -40
