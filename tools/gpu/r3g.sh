set -x
mkdir -p gpurun_out/r3g
( time timeout 300 python bench.py --cells 12,6,24 --steps 5 --warmup 3 > gpurun_out/r3g/bench_small.json 2> gpurun_out/r3g/bench_small.err ) 2>&1 | tail -4
tail -3 gpurun_out/r3g/bench_small.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3g/bench_small.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_comparable_across_n','strong_scaling_value','strong_scaling_e2e','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'], d['same_size_as_cpu_sample']['same_points'])
PY
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
