set -x
mkdir -p gpurun_out/r2l
nvidia-smi -L | head -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2l/bench_2gpu.json 2> gpurun_out/r2l/bench_2gpu.err
tail -5 gpurun_out/r2l/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2l/bench_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'].get('value'), 'full', d['full_sweep']['value'], d['full_sweep']['e2e_value'])
for r in d['full_sweep']['per_rank']: print(r)
print(d.get('e2e_split'))
PY
