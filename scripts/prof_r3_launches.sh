# ncu launch list (gpu__time_duration.sum per launch) of one bench step at the default workload, after the same command
# exited 0 plainly.  One ncu invocation per gpurun call.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --no-parity"
timeout 300 $CMD > gpurun_out/r3l_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3l_launches.csv $CMD > gpurun_out/r3l_ncu.log 2>&1
tail -c 400 gpurun_out/r3l_plain.log; tail -c 300 gpurun_out/r3l_ncu.log; wc -l gpurun_out/r3l_launches.csv
