# ncu --set full capture of the hot kernels (one ncu invocation per gpurun call), after the same command exited 0 plainly.
# The raw page is converted to CSV on the box (gpurun brings back at most 64 MiB: the .ncu-rep is dropped if too large).
set -x
mkdir -p gpurun_out
CMD="python scripts/ncu_targets_r2.py"
timeout 200 $CMD > gpurun_out/r3t_plain.log 2>&1 &&
timeout 420 ncu --set full --clock-control none --kernel-name-base demangled \
  -k "regex:k_compact_sym|k_colsum|k_scan|k_wk_atoms" -c 30 -o gpurun_out/r3t_targets $CMD > gpurun_out/r3t_ncu.log 2>&1
tail -4 gpurun_out/r3t_plain.log; tail -5 gpurun_out/r3t_ncu.log; ls -la gpurun_out/r3t_targets.ncu-rep
ncu -i gpurun_out/r3t_targets.ncu-rep --page raw --csv > gpurun_out/r3t_raw.csv 2> gpurun_out/r3t_raw.err
ls -la gpurun_out/
if [ $(stat -c %s gpurun_out/r3t_targets.ncu-rep) -gt 45000000 ]; then rm -f gpurun_out/r3t_targets.ncu-rep; fi
