set -x
mkdir -p gpurun_out/r3h
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3h/bench_2gpu.json 2> gpurun_out/r3h/bench_2gpu.err ) 2>&1 | tail -4
tail -3 gpurun_out/r3h/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3h/bench_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','value_comparable_across_n','strong_scaling_value','strong_scaling_e2e')}, 'e2e', d['e2e'].get('value'))
fs=d['full_sweep']; print({k:fs[k] for k in fs if k not in ('per_rank','what')})
for r in fs['per_rank']: print(r)
PY
