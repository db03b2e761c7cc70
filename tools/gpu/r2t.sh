set -x
mkdir -p gpurun_out/r2t
timeout 900 python -m pytest tests/test_gpu_sell.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2t/tests.txt 2>&1
tail -15 gpurun_out/r2t/tests.txt
timeout 600 python tools/sell_variants.py 44,20,190 > gpurun_out/r2t/variants.txt 2>&1
cat gpurun_out/r2t/variants.txt
( time timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2t/bench_20.json 2> gpurun_out/r2t/bench_20.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2t/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2t/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'])
print(d.get('e2e_split'), d['e2e_setup'], d['hbm_after_sweep'])
PY
