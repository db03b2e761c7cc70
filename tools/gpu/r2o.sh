set -x
mkdir -p gpurun_out/r2o
timeout 600 python tools/e2e_probe.py 12 > gpurun_out/r2o/probe_lazy.txt 2>&1
cat gpurun_out/r2o/probe_lazy.txt | cut -c1-1500
CUDA_MODULE_LOADING=EAGER timeout 600 python tools/e2e_probe.py 12 > gpurun_out/r2o/probe_eager.txt 2>&1
cat gpurun_out/r2o/probe_eager.txt | cut -c1-1500
