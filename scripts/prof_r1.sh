set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --bootnr 1 --no-cpu --no-e2e"
$CMD > gpurun_out/r1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu_launches.log 2>&1
$CMD > gpurun_out/r1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 40 -c 3 -o gpurun_out/r1_scan $CMD > gpurun_out/r1_ncu_scan.log 2>&1
$CMD > gpurun_out/r1_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gram -c 1 -o gpurun_out/r1_gram $CMD > gpurun_out/r1_ncu_gram.log 2>&1
tail -2 gpurun_out/r1_plain.log | cut -c1-400
ls -la gpurun_out
