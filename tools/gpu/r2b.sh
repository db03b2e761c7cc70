set -x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2b/pytest_gpu.txt
timeout 600 python tools/asm_bench.py 44,20,190 3 > gpurun_out/r2b/asm_bench.json 2> gpurun_out/r2b/asm_bench.err
cat gpurun_out/r2b/asm_bench.json; tail -3 gpurun_out/r2b/asm_bench.err
EMB_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name regex:"k_asm_rows|k_tet_records" -c 6 -o gpurun_out/r2b/asm_fused_full python tools/asm_bench.py 44,20,190 2 > gpurun_out/r2b/ncu_asm.log 2>&1
tail -3 gpurun_out/r2b/ncu_asm.log
