# runtests.jl of the Julia host binding (julia/RANSACB200): RANSAC.jl's own results against the CUDA
# library behind the ccall boundary, on the exchange scene of tests/golden/julia_inputs.json.
#
# NOT RUN in this repository's image (no Julia toolchain, no GPU in the build container).  Needs: Julia with
# RANSAC v0.6.0, JSON, StaticArrays; a B200 and the built library (ENV["RANSAC_B200_LIB"] =
# ".../ransac.jl_b200/libransac_b200.so").  The Python test-suite exercises the same C ABI calls
# (tests/test_*_gpu.py); this file is their Julia twin, in the style of the reference's test/runtests.jl.
using Test
using RANSAC
using RANSACB200
using JSON
using StaticArrays: SVector
using RANSAC: FittedPlane, FittedSphere, FittedCylinder, FittedCone, FittedShape, ExtractedShape, IterationCandidates,
              RANSACCloud, ransacparameters      # everything make_golden_helpers.jl expects in scope

const GOLDEN = joinpath(@__DIR__, "..", "..", "..", "tests", "golden")
include(joinpath(@__DIR__, "..", "..", "make_golden_helpers.jl"))

inp = JSON.parsefile(joinpath(GOLDEN, "julia_inputs.json"))
params = params_from(inp["params"])
P = [sv(v) for v in inp["points"]]
N = [sv(v) for v in inp["normals"]]
subsets = [Int.(s) .+ 1 for s in inp["subsets"]]
cands = [shape_from(d) for d in inp["candidates"]]

@testset "struct layouts (include/rsc.h)" begin
    @test sizeof(RANSACB200.RscCand) == 64
    @test sizeof(RANSACB200.RscParams) == 160
    @test RANSACB200.version() == 100
end

@testset "scorecandidate / compatibles*: device == RANSAC.jl" begin
    pc = RANSACCloud(P, N, subsets)
    pc.isenabled[Int.(inp["disabled"]) .+ 1] .= false
    dc = RANSACB200.DeviceCloud(pc)
    dev = RANSACB200.scorecandidates(dc, cands, 1, params)    # one launch for all candidates
    for (s, (sc_dev, ip_dev)) in zip(cands, dev)
        sc_ref, ip_ref = RANSAC.scorecandidate(pc, s, 1, params)
        @test ip_dev == ip_ref                                # bit-exact inlier list, subset order
        @test RANSAC.E(sc_dev) ≈ RANSAC.E(sc_ref) rtol = 1e-12
    end
end

@testset "fit x4: device == RANSAC.jl" begin
    pc = RANSACCloud(P, N, subsets)
    dc = RANSACB200.DeviceCloud(pc)
    sets = [Int.(sd) .+ 1 for sd in inp["minimal_sets"]]
    shapes, srcset = RANSACB200.fit_batch(dc, hcat(sets...), params)   # (set, shape_types) order + source set
    k = 1
    for (si, sd) in enumerate(sets), T in params.iteration.shape_types
        f = RANSAC.fit(T, pc.vertices[sd], pc.normals[sd], pc, params)
        f === nothing && continue
        @test k <= length(shapes) && srcset[k] == si
        @test typeof(shapes[k]) <: T
        @test collect(RANSACB200.tocand(shapes[k]).p) ≈ collect(RANSACB200.tocand(f).p) rtol = 1e-5
        k += 1
    end
    @test k - 1 == length(shapes)
end

@testset "refit + invalidate_indexes!: device == RANSAC.jl" begin
    pc = RANSACCloud(P, N, subsets)
    pc.isenabled[Int.(inp["disabled"]) .+ 1] .= false
    dc = RANSACB200.DeviceCloud(pc)
    for s in cands[1:6]
        @test RANSACB200.refit_extract!(dc, s, params; disable = false).inpoints == RANSAC.refit(s, pc, params).inpoints
    end
end

@testset "the whole loop on explicit minimal sets" begin
    # the library's own loop draws Philox sets; the comparable run is RANSAC.jl's control flow with the
    # device doing fit / score / refit through the per-call ABI, fed with the file's index triples
    pc_ref = RANSACCloud(P, N, subsets)
    ex_ref, at_ref, it_ref = loop_with_sets(pc_ref, params, inp["loop"]["sets"])
    pc_dev = RANSACCloud(P, N, subsets)
    ex_dev, at_dev, it_dev = RANSACB200.loop_with_sets(pc_dev, params, inp["loop"]["sets"])
    @test at_dev == at_ref && it_dev == it_ref
    @test [e.inpoints for e in ex_dev] == [e.inpoints for e in ex_ref]
    @test pc_dev.isenabled == pc_ref.isenabled
end

@testset "ransac(pc, params, true) drop-in" begin
    pc = RANSACCloud(P, N, subsets)
    extracted, secs = RANSACB200.ransac(pc, params, true; seed = 777)
    @test length(extracted) >= 3
    taken = vcat([e.inpoints for e in extracted]...)
    @test allunique(taken)
    @test count(pc.isenabled) == pc.size - length(taken)
end
