set -x
mkdir -p gpurun_out/r3c
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r3c/bench_4gpu.json 2> gpurun_out/r3c/bench_4gpu.err ) 2>&1 | tail -4
tail -3 gpurun_out/r3c/bench_4gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3c/bench_4gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'))
fs=d['full_sweep']; print({k:fs[k] for k in fs if k not in ('per_rank','what')})
for r in fs['per_rank']: print(r)
print(d['e2e_split'])
PY
