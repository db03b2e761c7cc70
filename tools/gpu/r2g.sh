set -x
mkdir -p gpurun_out/r2g
timeout 2400 python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/r2g/pytest_gpu.txt 2>&1
grep -E "locate 20k|topology device|passed|failed|Error" gpurun_out/r2g/pytest_gpu.txt | tail -12
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g/bench_20.json 2> gpurun_out/r2g/bench_20.err
tail -c 600 gpurun_out/r2g/bench_20.json; tail -3 gpurun_out/r2g/bench_20.err
