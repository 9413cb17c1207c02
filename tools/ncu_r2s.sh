set -x
R="python tools/ransac_e2e.py --scene c4 --no-cpu --no-check"
$R > gpurun_out/r2s_plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2s_launches_ransac_c4.csv $R > gpurun_out/r2s_ncu_c4.log 2>&1
ls -la gpurun_out/r2s_*
