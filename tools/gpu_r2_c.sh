# round 2, GPU pass C: ncu --set full of the interleaved (_Coa) kernels still below the 75 % target
set -u
mkdir -p gpurun_out
export B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3 B200FE_PLAN=0
prof() { # name, kernel regex, skip, -- command
  name=$1; regex=$2; skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed: $name"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "profiled $name rc=$?"
}
B200FE_NELMT=67104 prof hex10_f64_coa lanesq_kernel 1 benchmark05/build/benchmark05 10 10 10
B200FE_NELMT=67104 B200FE_DTYPE=float prof hex10_f32_coa hex_lanes_kernel 1 benchmark05/build/benchmark05 10 10 10
B200FE_NELMT=65536 prof quad32_f64_coa quad_lanes_kernel 1 benchmark04/build/benchmark04 32 32
B200FE_NELMT=65536 B200FE_DTYPE=float prof quad32_f32_coa quad_lanes_kernel 1 benchmark04/build/benchmark04 32 32
