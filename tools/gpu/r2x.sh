set -x
mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r2x/tests.txt 2>&1
tail -4 gpurun_out/r2x/tests.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2x/bench_2gpu.json 2> gpurun_out/r2x/bench_2gpu.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2x/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2x/bench_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'].get('value'))
fs=d['full_sweep']; print({k:fs[k] for k in fs if k not in ('per_rank','what')})
for r in fs['per_rank']: print(r)
PY
