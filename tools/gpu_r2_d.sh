# round 2, GPU pass D (1 GPU), final build: the default bench.py line with all its legs, the reference arm, the launch
# list of the bench command and ncu --set full of the headline kernel (re-tuned shape: G = 1, 16 warps).
#   gpurun --timeout 1500 -- bash tools/gpu_r2_d.sh
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 700 python bench.py --steps 20 --warmup 5 --detail gpurun_out/bench_detail_n1.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench_n1.err
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-sustained > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-sustained > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
B200FE_NELMT=262144 B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3 benchmark05/build/benchmark05 8 8 8 > gpurun_out/plain_hex8.log 2>&1 && \
B200FE_NELMT=262144 B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3 ncu --set full --clock-control none --import-source on \
    -k regex:hex_mma -s 6 -c 1 -o gpurun_out/r02_hex8_f64_mma benchmark05/build/benchmark05 8 8 8 > gpurun_out/ncu_hex8.log 2>&1
echo "ncu hex8 rc=$?"
