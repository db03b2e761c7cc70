set -x
mkdir -p gpurun_out/r2c
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c/pytest_gpu.txt 2>&1
tail -25 gpurun_out/r2c/pytest_gpu.txt
timeout 600 python tools/asm_bench.py 44,20,190 3 > gpurun_out/r2c/asm_bench.json 2> gpurun_out/r2c/asm_bench.err
cat gpurun_out/r2c/asm_bench.json; tail -3 gpurun_out/r2c/asm_bench.err
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c/bench_20.json 2> gpurun_out/r2c/bench_20.err
tail -c 1500 gpurun_out/r2c/bench_20.json; tail -5 gpurun_out/r2c/bench_20.err
