set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --bootnr 2 --no-cpu --no-e2e"
$CMD > gpurun_out/r1g_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_scan<1" -c 2 -o gpurun_out/r1g_scan_swap $CMD > gpurun_out/r1g_ncu_scan.log 2>&1
$CMD > gpurun_out/r1g_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gram_i8 -c 1 -o gpurun_out/r1g_gram_i8 $CMD > gpurun_out/r1g_ncu_gram.log 2>&1
$CMD > gpurun_out/r1g_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_compact -c 1 -o gpurun_out/r1g_compact $CMD > gpurun_out/r1g_ncu_compact.log 2>&1
ls -la gpurun_out | grep r1g
