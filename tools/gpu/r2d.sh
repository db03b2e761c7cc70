set -x
mkdir -p gpurun_out/r2d
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_dropin.py -m gpu -x -q > gpurun_out/r2d/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2d/pytest_gpu.txt
for w in 16 20 24; do EMB_ASM_WARPS=$w timeout 600 python tools/asm_bench.py 44,20,190 3 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('warps $w', j['fused'], j['fused_vs_coo_rel_diff'])" ; done > gpurun_out/r2d/asm_warps.txt 2>&1
cat gpurun_out/r2d/asm_warps.txt
for s in 0 1; do EMB_SPMV_SELL=$s timeout 600 python tools/spmv_tune.py 44,20,190 2>&1 | grep nv= ; done > gpurun_out/r2d/spmv_sell.txt 2>&1
cat gpurun_out/r2d/spmv_sell.txt
EMB_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name regex:"k_asm_rows" -c 1 -o gpurun_out/r2d/asm_v3_full python tools/asm_bench.py 44,20,190 2 > gpurun_out/r2d/ncu_asm.log 2>&1
tail -2 gpurun_out/r2d/ncu_asm.log
