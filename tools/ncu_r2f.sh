set -x
B="python bench.py --steps 2 --warmup 1 --ransac none --no-cpu --no-cull --no-e2e"
$B > gpurun_out/r2f_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv $B > gpurun_out/r2f_ncu_bench.log 2>&1
B2="python bench.py --steps 2 --warmup 1 --ransac none --no-cpu --no-cull --no-e2e --points 2097152"
$B2 > gpurun_out/r2f_plain_bench2m.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench_2Mi.csv $B2 > gpurun_out/r2f_ncu_bench2m.log 2>&1
R="python tools/ransac_e2e.py --scene c4 --no-cpu --no-check"
$R > gpurun_out/r2f_plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2f_launches_ransac_c4.csv $R > gpurun_out/r2f_ncu_c4.log 2>&1
B3="python bench.py --steps 1 --warmup 1 --ransac none --no-cpu --no-cull --no-e2e"
$B3 > gpurun_out/r2f_plain_bench3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 5 -c 5 -o gpurun_out/r2f_score_full $B3 > gpurun_out/r2f_ncu_full.log 2>&1
ls -la gpurun_out/r2f_*
