# ncu --set full capture of the hot kernels (one ncu invocation per gpurun call), after the same command exited 0 plainly.
set -x
mkdir -p gpurun_out
CMD="python scripts/ncu_targets.py"
timeout 200 $CMD > gpurun_out/r1t_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_gram_i8_persist|k_compact_sym|k_scan" -c 9 -o gpurun_out/r1t_targets $CMD > gpurun_out/r1t_ncu.log 2>&1
tail -3 gpurun_out/r1t_plain.log; tail -5 gpurun_out/r1t_ncu.log; ls -la gpurun_out/r1t_targets.ncu-rep
