set -x
mkdir -p gpurun_out/r2m
nvidia-smi -L | wc -l; free -g | head -2; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2m/bench_8gpu.json 2> gpurun_out/r2m/bench_8gpu.err
tail -3 gpurun_out/r2m/bench_8gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2m/bench_8gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'])
for r in d['full_sweep']['per_rank']: print(r)
print(d.get('e2e_split'))
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload slabs --cells 76,34,323 --points 401 --steps 4 --warmup 3 --e2e-steps -1 --no-full-sweep --no-cpu-baseline --recycle 12 > gpurun_out/r2m/slabs_8gpu.json 2> gpurun_out/r2m/slabs_8gpu.err
tail -5 gpurun_out/r2m/slabs_8gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2m/slabs_8gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['solver'], d['hbm_after_sweep'], d['roofline_assembly'])
for r in d['full_sweep']['per_rank']: print(r)
PY
