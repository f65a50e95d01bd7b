# Round-1 ncu evidence: launch list (shares) + one --set full capture per hot kernel, each after the same command has
# exited 0 without ncu.  Run under gpurun (1 GPU); outputs in gpurun_out/, summaries made by scripts/ncu_summary.py.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --bootnr 2 --no-cpu --no-e2e"
$CMD > gpurun_out/r1f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1f_launches.csv $CMD > gpurun_out/r1f_ncu_launches.log 2>&1
$CMD > gpurun_out/r1f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gram_i8 -c 1 -o gpurun_out/r1f_gram_i8 $CMD > gpurun_out/r1f_ncu_gram.log 2>&1
$CMD > gpurun_out/r1f_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_scan<1" -c 2 -o gpurun_out/r1f_scan_swap $CMD > gpurun_out/r1f_ncu_scan.log 2>&1
$CMD > gpurun_out/r1f_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_compact|k_scan_delta" -c 3 -o gpurun_out/r1f_compact_delta $CMD > gpurun_out/r1f_ncu_compact.log 2>&1
tail -c 300 gpurun_out/r1f_plain.log
ls -la gpurun_out | head -30
