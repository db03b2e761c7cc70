set -x
mkdir -p gpurun_out/r2k
for x in 0 1; do
EMB_SELL_X32=$x SPMV_TUNE_ONLY=2,1 timeout 300 python tools/spmv_tune.py 44,20,190 2>&1 | grep -E "nv=|rror" | sed "s/^/x32=$x /" | tee -a gpurun_out/r2k/spmv_x32.txt
done
EMB_SELL_X32=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2k/pytest_x32.txt 2>&1
tail -5 gpurun_out/r2k/pytest_x32.txt
EMB_SELL_X32=1 timeout 900 python bench.py --steps 20 --warmup 3 --e2e-steps -1 --no-full-sweep --no-cpu-baseline > gpurun_out/r2k/bench_x32.json 2> gpurun_out/r2k/bench_x32.err
tail -3 gpurun_out/r2k/bench_x32.err
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2k/bench_20.json 2> gpurun_out/r2k/bench_20.err
tail -3 gpurun_out/r2k/bench_20.err
python - <<'PY'
import json
for f in ('bench_x32','bench_20'):
    l=[x for x in open(f'gpurun_out/r2k/{f}.json') if x.startswith('{')]
    d=json.loads(l[-1])
    print(f, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'], d['solver']['max_relres'])
    print(d.get('e2e_split'), d['e2e_setup'])
PY
