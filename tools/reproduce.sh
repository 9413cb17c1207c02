#!/usr/bin/env bash
# The commands behind profiles/ (run from the repo root on a B200 box; each block is independent).
# Nothing here is needed to use the library -- it documents how every committed number was produced.
set -euo pipefail
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

# build + parity
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests -q -m "not gpu"
python -m pytest tests -q -m gpu
python -c "import __graft_entry__ as g; g.smoke()"

# bench lines (profiles/r1f_bench_*.json)
python bench.py                                             > gpurun_out/bench_1gpu.json
python bench.py --impl reference                            > gpurun_out/bench_reference_arm.json
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${n}gpu.json
done

# ransac() on c4 / c5, sharded, with the sharded == unsharded check (profiles/r1f_ransac_*.json)
$TR --nproc-per-node 8 --master-port 29511 tools/ransac_multi.py --scene c4 > gpurun_out/ransac_c4_8gpu.json
$TR --nproc-per-node 8 --master-port 29512 tools/ransac_multi.py --scene c5 > gpurun_out/ransac_c5_8gpu.json
python tools/ransac_multi.py --scene c4 --sampler octree --levels 9          > gpurun_out/ransac_c4_octree.json
python tools/ransac_probe.py                                                  # where ransac() wall time goes

# ncu: launch list of the bench, full capture of the score kernels and of the refit kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --ransac none
ncu --set full --import-source on --clock-control none -k regex:score_kernel -s 10 -c 5 -o gpurun_out/prof_score \
    python bench.py --no-e2e --no-cpu --ransac none --steps 1 --warmup 1
ncu --set full --import-source on --clock-control none -k regex:extract_mask -c 1 -o gpurun_out/prof_refit \
    python bench.py --no-e2e --no-cpu --ransac none --steps 1 --warmup 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ransac_c4.csv \
    python tools/ransac_multi.py --scene c4

# tiling sweep and microbenchmarks (profiles/*tiling_sweep*, r1_pipe_mix2*, r1_loop_bench*)
python tools/tune_score.py > gpurun_out/tune.jsonl
for t in fp32_peak pipe_mix2 loop_bench loop_bench2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Iransac.jl_b200/csrc -o tools/$t tools/$t.cu
  tools/$t > gpurun_out/$t.jsonl
done

# extensions (profiles/r1g_*): least-squares refit on c4, progressive + lsq sharded over 2 GPUs, culled scoring
for l in 0 1; do RSC_TRACE=1 python tools/ransac_multi.py --scene c4 --lsq $l > gpurun_out/c4_lsq$l.json; done
$TR --nproc-per-node 2 --master-port 29517 tools/ransac_multi.py --scene c2 --progressive 1 --lsq 1 > gpurun_out/c2_2gpu_prog_lsq.json
python tools/cull_estimate.py --points 4194304 --tile 512 --per-type 32 --levels 12     # CPU only
python -m pytest tests/test_cull_gpu.py -q && python tools/cull_bench.py > gpurun_out/cull_bench.json
