# ncu --set full capture of the hot kernels (one ncu invocation per gpurun call), after the same command exited 0 plainly.
set -x
mkdir -p gpurun_out
CMD="python scripts/ncu_targets_r2.py"
timeout 200 $CMD > gpurun_out/r2t_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:k_gram_i8_persist|k_compact_sym|k_colsum_batch|k_scan|k_wk_batch" -c 48 -o gpurun_out/r2t_targets $CMD > gpurun_out/r2t_ncu.log 2>&1
tail -4 gpurun_out/r2t_plain.log; tail -5 gpurun_out/r2t_ncu.log; ls -la gpurun_out/r2t_targets.ncu-rep
