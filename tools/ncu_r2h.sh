set -x
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err; tail -c 600 gpurun_out/r2h_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.json 2>&1; tail -c 300 gpurun_out/r2h_bench_ref.json
R="python tools/ransac_e2e.py --scene c4 --no-cpu --no-check"
$R > gpurun_out/r2h_plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2h_launches_ransac_c4.csv $R > gpurun_out/r2h_ncu_c4.log 2>&1
ls -la gpurun_out/r2h_*
