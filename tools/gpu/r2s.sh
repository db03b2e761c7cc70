set -x
mkdir -p gpurun_out/r2s
timeout 900 python tools/sell_variants.py 44,20,190 > gpurun_out/r2s/variants.txt 2>&1
cat gpurun_out/r2s/variants.txt
