set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a/memcheck.txt python __graft_entry__.py --smoke > gpurun_out/r2a/memcheck_stdout.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a/racecheck.txt python __graft_entry__.py --smoke > gpurun_out/r2a/racecheck_stdout.txt 2>&1
timeout 900 compute-sanitizer --tool initcheck --log-file gpurun_out/r2a/initcheck.txt python __graft_entry__.py --smoke > gpurun_out/r2a/initcheck_stdout.txt 2>&1
timeout 600 python bench.py --steps 20 --e2e-steps -1 --no-cpu-baseline > gpurun_out/r2a/bench_base.json 2> gpurun_out/r2a/bench_base.err
timeout 600 python bench.py --steps 20 --e2e-steps -1 --no-cpu-baseline --coarse-basis > gpurun_out/r2a/bench_coarse.json 2> gpurun_out/r2a/bench_coarse.err
tail -c 600 gpurun_out/r2a/bench_base.json; tail -c 600 gpurun_out/r2a/bench_coarse.json
