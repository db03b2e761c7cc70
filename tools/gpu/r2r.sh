set -x
mkdir -p gpurun_out/r2r
timeout 600 python tools/sell_variants.py 44,20,190 > gpurun_out/r2r/variants.txt 2>&1
cat gpurun_out/r2r/variants.txt
timeout 600 python tools/e2e_trace.py 20 > gpurun_out/r2r/trace.txt 2>&1
cat gpurun_out/r2r/trace.txt
