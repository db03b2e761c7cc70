set -x
mkdir -p gpurun_out/r2n
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2n/pytest_gpu.txt 2>&1
tail -12 gpurun_out/r2n/pytest_gpu.txt
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n/bench_20.json 2> gpurun_out/r2n/bench_20.err
tail -3 gpurun_out/r2n/bench_20.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2n/bench_20.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'prec', d['solver']['precond_apply_ms'], d['solver']['lockstep_iterations_total'])
print(d.get('e2e_split'), d['e2e_setup'])
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2n/iter_launches.csv python tools/profile_iter.py 44,20,190 14 > gpurun_out/r2n/prof_iter.log 2>&1
tail -1 gpurun_out/r2n/prof_iter.log | cut -c1-200
for tool in memcheck racecheck initcheck; do
timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2n/$tool.txt python __graft_entry__.py --smoke > gpurun_out/r2n/${tool}_stdout.txt 2>&1
tail -2 gpurun_out/r2n/$tool.txt
done
