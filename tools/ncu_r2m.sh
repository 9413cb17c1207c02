set -x
C="python tools/cull_bench.py --points 4194304 --cands 4096 --reps 2"
$C > gpurun_out/r2m_cull_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cull_score -s 1 -c 1 -o gpurun_out/r2m_cull_full $C > gpurun_out/r2m_ncu_cull.log 2>&1
tail -2 gpurun_out/r2m_cull_plain.log; ls -la gpurun_out/r2m_*
