set -x
mkdir -p gpurun_out/r3b
for tool in memcheck racecheck initcheck; do
timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r3b/$tool.txt python __graft_entry__.py --smoke > gpurun_out/r3b/${tool}_stdout.txt 2>&1
tail -2 gpurun_out/r3b/$tool.txt
done
