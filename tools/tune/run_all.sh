#!/bin/bash
# run every tuner binary; CSV on stdout
cd "$(dirname "$0")/build" || exit 1
echo dim,dtype,nq,backend,E,threads,R,V,smem,ctas_per_sm,regs,ms_min,ms_med,GBs,hbm_frac,ok
timeout 120 ./tune_2_f32_6 "$@" || echo "# tune_2_f32_6 exited $?"
timeout 120 ./tune_2_f32_10 "$@" || echo "# tune_2_f32_10 exited $?"
timeout 120 ./tune_2_f32_14 "$@" || echo "# tune_2_f32_14 exited $?"
timeout 120 ./tune_3_f32_6 "$@" || echo "# tune_3_f32_6 exited $?"
timeout 120 ./tune_3_f32_10 "$@" || echo "# tune_3_f32_10 exited $?"
timeout 120 ./tune_2_f64_6 "$@" || echo "# tune_2_f64_6 exited $?"
timeout 120 ./tune_3_f64_6 "$@" || echo "# tune_3_f64_6 exited $?"
timeout 120 ./tune_2_f32_2 "$@" || echo "# tune_2_f32_2 exited $?"
timeout 120 ./tune_3_f32_2 "$@" || echo "# tune_3_f32_2 exited $?"
