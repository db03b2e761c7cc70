set -x
mkdir -p gpurun_out/r3e
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sell.py tests/test_gpu_limits.py -x -q -m gpu > gpurun_out/r3e/tests.txt 2>&1
tail -3 gpurun_out/r3e/tests.txt
( time timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r3e/bench_20.json 2> gpurun_out/r3e/bench_20.err ) 2>&1 | tail -4
tail -3 gpurun_out/r3e/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3e/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['ms'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['sampled_launches'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'], 'relres', d['solver']['max_relres'], d['clocks'])
print(d['e2e_setup'])
PY
