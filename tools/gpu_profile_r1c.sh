set -u
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
bash tools/gpu_profile_r1b.sh 2>&1 | grep profiled
