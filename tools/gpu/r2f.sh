set -x
mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_topology.py -m gpu -x -q -s > gpurun_out/r2f/pytest_topo.txt 2>&1
tail -8 gpurun_out/r2f/pytest_topo.txt
cat > /tmp/sweep_sigma.py <<'PY'
import os, subprocess, sys
for sigma in (8, 32, 64, 256):
    for tma in (0, 1):
        env = dict(os.environ, EMB_SELL_SIGMA=str(sigma), EMB_SPMV_TMA=str(tma), SPMV_TUNE_ONLY="2,1")
        out = subprocess.run([sys.executable, "tools/spmv_tune.py", "44,20,190"], env=env, capture_output=True, text=True, timeout=300)
        print([l for l in out.stdout.splitlines() if "nv=" in l] or out.stderr[-300:], flush=True)
PY
timeout 2400 python /tmp/sweep_sigma.py > gpurun_out/r2f/sigma_sweep.txt 2>&1
cat gpurun_out/r2f/sigma_sweep.txt
SPMV_TUNE_ONLY=2,1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_bsell_tma" --launch-skip 5 -c 1 \
   -o gpurun_out/r2f/bsell_tma_full python tools/spmv_tune.py 44,20,190 > gpurun_out/r2f/ncu_tma.log 2>&1
tail -2 gpurun_out/r2f/ncu_tma.log
