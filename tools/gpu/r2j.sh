set -x
mkdir -p gpurun_out/r2j
SPMV_TUNE_ONLY=2,1 timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | tee gpurun_out/r2j/spmv_v3.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2j/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2j/pytest_gpu.txt
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j/bench_20.json 2> gpurun_out/r2j/bench_20.err
tail -3 gpurun_out/r2j/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2j/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'])
print(d['e2e_setup'], d['full_sweep']['per_rank'])
PY
SPMV_TUNE_ONLY=2,1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_bsell_tma" --launch-skip 5 -c 1 \
   -o gpurun_out/r2j/bsell_tma_v3_full python tools/spmv_tune.py 44,20,190 > gpurun_out/r2j/ncu_tma.log 2>&1
ncu -i gpurun_out/r2j/bsell_tma_v3_full.ncu-rep --page raw --csv > gpurun_out/r2j/bsell_v3_raw.csv 2>/dev/null
