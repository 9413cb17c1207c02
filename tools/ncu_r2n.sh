set -x
R="python tools/ransac_multi.py --scene c4 --sampler octree --lw-period 16"
$R > gpurun_out/r2n_plain_octree.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r2n_launches_octree_c4.csv $R > gpurun_out/r2n_ncu_octree.log 2>&1
tail -c 300 gpurun_out/r2n_plain_octree.log; ls -la gpurun_out/r2n_*
