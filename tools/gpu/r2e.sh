set -x
mkdir -p gpurun_out/r2e
for m in 1 0; do EMB_SPMV_TMA=$m timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | sed "s/^/TMA=$m /"; done > gpurun_out/r2e/spmv_tma.txt 2>&1
cat gpurun_out/r2e/spmv_tma.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_dropin.py tests/test_gpu_sharded.py -m gpu -x -q --durations=5 > gpurun_out/r2e/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2e/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps -1 --no-full-sweep > gpurun_out/r2e/bench_20.json 2> gpurun_out/r2e/bench_20.err
tail -c 1800 gpurun_out/r2e/bench_20.json; tail -5 gpurun_out/r2e/bench_20.err
