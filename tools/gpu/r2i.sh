set -x
mkdir -p gpurun_out/r2i
SPMV_TUNE_ONLY=2,1 timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | tee gpurun_out/r2i/spmv_v2.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2i/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2i/pytest_gpu.txt
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i/bench_20.json 2> gpurun_out/r2i/bench_20.err
tail -c 400 gpurun_out/r2i/bench_20.json; tail -3 gpurun_out/r2i/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2i/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['full_sweep']['value'], d['full_sweep']['e2e_value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'])
PY
SPMV_TUNE_ONLY=2,1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_bsell_tma" --launch-skip 5 -c 1 \
   -o gpurun_out/r2i/bsell_tma_v2_full python tools/spmv_tune.py 44,20,190 > gpurun_out/r2i/ncu_tma.log 2>&1
tail -2 gpurun_out/r2i/ncu_tma.log
ncu -i gpurun_out/r2i/bsell_tma_v2_full.ncu-rep --page raw --csv > gpurun_out/r2i/bsell_v2_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2i/iter_launches.csv python tools/profile_iter.py 44,20,190 14 > gpurun_out/r2i/prof_iter.log 2>&1
tail -2 gpurun_out/r2i/prof_iter.log
