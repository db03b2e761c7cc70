set -x
mkdir -p gpurun_out/r3a
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sell.py -x -q -m gpu > gpurun_out/r3a/tests.txt 2>&1
tail -5 gpurun_out/r3a/tests.txt
for lks in 1 0; do
EMB_LKS=$lks timeout 900 python bench.py --steps 20 --warmup 3 --e2e-steps -1 --no-cpu-baseline > gpurun_out/r3a/bench_lks$lks.json 2> gpurun_out/r3a/bench_lks$lks.err
tail -3 gpurun_out/r3a/bench_lks$lks.err
python - $lks <<'PY'
import json, sys
l=[x for x in open(f'gpurun_out/r3a/bench_lks{sys.argv[1]}.json') if x.startswith('{')]
d=json.loads(l[-1])
print('LKS', sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, 'full', d['full_sweep']['value'], d['full_sweep']['ms'], 'spmm', d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], 'iters(20 pts)', d['solver']['lockstep_iterations_total'], 'full iters', d['full_sweep']['per_rank'][0]['iterations'], 'iterating', d['full_sweep']['per_rank'][0]['iterating_points'], 'max relres', d['solver']['max_relres'])
PY
done
