set -x
mkdir -p gpurun_out/r2q
timeout 300 python tools/malloc_probe.py 48 > gpurun_out/r2q/malloc.txt 2>&1
cat gpurun_out/r2q/malloc.txt
timeout 600 python tools/e2e_trace.py 20 > gpurun_out/r2q/trace.txt 2>&1
cat gpurun_out/r2q/trace.txt
