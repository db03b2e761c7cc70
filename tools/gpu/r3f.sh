set -x
mkdir -p gpurun_out/r3f
( time timeout 420 python bench.py --workload slabs --cells 76,34,323 --points 401 --steps 2 --warmup 0 --e2e-steps -1 --no-full-sweep --no-cpu-baseline --recycle 12 > gpurun_out/r3f/slabs_1gpu.json 2> gpurun_out/r3f/slabs_1gpu.err ) 2>&1 | tail -4
tail -5 gpurun_out/r3f/slabs_1gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3f/slabs_1gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['solver'], d['hbm_after_sweep'], d['roofline'], d['roofline_assembly']['Mtet_per_s'], d['sizes'])
PY
