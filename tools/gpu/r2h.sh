set -x
mkdir -p gpurun_out/r2h
free -g | head -2; nproc
timeout 600 python -m pytest tests/test_gpu_postproc.py -x -q 2>&1 | tail -5
for ord in lex morton; do for blk in 0 1; do
EMB_MESH_ORDER=$ord EMB_SELL_BLOCKED=$blk SPMV_TUNE_ONLY=2,1 timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | sed "s/^/order=$ord blocked=$blk /" | tee -a gpurun_out/r2h/spmv_order.txt
done; done
EMB_MESH_ORDER=morton timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | sed "s/^/order=morton all /" | tee -a gpurun_out/r2h/spmv_order.txt
timeout 1700 python bench.py --workload slabs --cells 76,34,323 --points 401 --steps 6 --warmup 3 --e2e-steps -1 --no-full-sweep --no-cpu-baseline --recycle 16 > gpurun_out/r2h/slabs_1gpu.json 2> gpurun_out/r2h/slabs_1gpu.err
tail -c 3000 gpurun_out/r2h/slabs_1gpu.json; tail -15 gpurun_out/r2h/slabs_1gpu.err
