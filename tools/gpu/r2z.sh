set -x
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sell.py -x -q -m gpu > gpurun_out/r2z/tests.txt 2>&1
tail -3 gpurun_out/r2z/tests.txt
( time timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z/bench_20.json 2> gpurun_out/r2z/bench_20.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2z/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2z/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'])
print(d.get('e2e_split'), d['e2e_setup'], d['hbm_after_sweep'])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2z/iter_launches.csv python tools/profile_iter.py 44,20,190 14 > gpurun_out/r2z/prof_iter.log 2>&1
tail -1 gpurun_out/r2z/prof_iter.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_bsell_tma" --launch-skip 5 -c 1 \
   -o gpurun_out/r2z/bsell_tma2_full python tools/sell_variants.py 44,20,190 default > gpurun_out/r2z/ncu_tma.log 2>&1
tail -2 gpurun_out/r2z/ncu_tma.log
ls -la gpurun_out/r2z
