# round 2, GPU pass B (1 GPU): full -m gpu suite, the default bench.py line, launch list of the bench command,
# ncu --set full of the tcgen05 kernel and of the headline kernel.   gpurun --timeout 1500 -- bash tools/gpu_r2_b.sh
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2>&1; echo "ref rc=$?"
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-sustained > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-sustained > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/ncu_umma.py > gpurun_out/umma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma -s 1 -c 1 -o gpurun_out/r02_umma \
    python tools/ncu_umma.py > gpurun_out/ncu_umma.log 2>&1
echo "ncu umma rc=$?"
B200FE_NELMT=262144 B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3 benchmark05/build/benchmark05 8 8 8 > gpurun_out/plain_hex8.log 2>&1 && \
B200FE_NELMT=262144 B200FE_SKIP_CPU=1 B200FE_SKIP_CUBLAS=1 B200FE_REPS=3 ncu --set full --clock-control none --import-source on \
    -k regex:hex_mma -s 6 -c 1 -o gpurun_out/r02_hex8_f64_mma benchmark05/build/benchmark05 8 8 8 > gpurun_out/ncu_hex8.log 2>&1
echo "ncu hex8 rc=$?"
