# round 2, GPU pass A (1 GPU): full -m gpu suite, the default bench.py line with all its legs, the reference arm,
# ncu --set full of the benchmark01-03 kernels.   gpurun --timeout 1500 -- bash tools/gpu_r2_a.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_a_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2>&1; echo "ref rc=$?"
kill $SMI
python tools/ncu_vec.py > gpurun_out/vec_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"reduce_partials|add_vector|matvec" -c 8 \
    -o gpurun_out/r02_vec python tools/ncu_vec.py > gpurun_out/ncu_vec.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_vec.log
