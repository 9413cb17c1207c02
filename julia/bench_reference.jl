# Times the real RANSAC.jl package (CPU, single thread) on a scene file written by
# `python -m tools.dump_scene`; for anyone who has Julia -- this image does not.
#   julia --project=. julia/bench_reference.jl scene_c1.bin
using RANSAC, StaticArrays
function main(path)
    raw = reinterpret(Float32, read(path))
    n = length(raw) ÷ 6
    v = [SVector{3,Float64}(raw[3i-2], raw[3i-1], raw[3i]) for i in 1:n]
    nr = [SVector{3,Float64}(raw[3n+3i-2], raw[3n+3i-1], raw[3n+3i]) for i in 1:n]
    pc = RANSACCloud(v, nr, 2)
    p = ransacparameters()
    ransac(pc, p, true; reset_rand=true)            # compile
    t = @elapsed ex, _ = ransac(pc, p, true; reset_rand=true)
    println("ransac(): $(t) s, $(length(ex)) shapes, $(Threads.nthreads()) thread(s)")
    # scorecandidate throughput on subset 1
    fp = FittedPlane(SVector(0.0, 0, 0), SVector(0.0, 0, 1))
    t = @elapsed for _ in 1:100 RANSAC.scorecandidate(pc, fp, 1, p) end
    println("scorecandidate(plane): $(100 * length(pc.subsets[1]) / t / 1e9) G evals/s")
end
main(ARGS[1])
