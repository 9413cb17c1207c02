# make_golden.jl -- run the REAL RANSAC.jl (cserteGT3/RANSAC.jl v0.6.0) on tests/golden/julia_inputs.json
# and write tests/golden/julia_reference.json in the format of julia_expected_by_oracle.json.
#
#   julia --project=<an environment with RANSAC v0.6.0, JSON, StaticArrays> julia/make_golden.jl [tests/golden]
#
# tests/test_julia_golden.py then compares the file with the repo's oracles (NumPy and C) and, on a GPU
# box, with the CUDA library: compatibles* index lists and loop results must be identical, E / fitted
# parameters agree to 1e-9 relative.  This is how the "parity unpinned" rows of DESIGN.md section 5
# (compatibles*, scorecandidate, refit, cylinder / cone fit, estimatescore, the loop) get pinned by
# anyone who has Julia.  NOT RUN in this repository's image (no Julia toolchain): written against the
# reference's source, every call cites the line it relies on.
#
# Indices in the JSON files are 0-based; Julia's are 1-based (converted here).
using RANSAC
using JSON
using StaticArrays: SVector
using RANSAC: FittedPlane, FittedSphere, FittedCylinder, FittedCone, FittedShape, ExtractedShape,
              IterationCandidates, RANSACCloud, ransacparameters

const DIR = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
include(joinpath(@__DIR__, "make_golden_helpers.jl"))

function main()
    inp = JSON.parsefile(joinpath(DIR, "julia_inputs.json"))
    params = params_from(inp["params"])
    P = [sv(v) for v in inp["points"]]
    N = [sv(v) for v in inp["normals"]]
    subsets = [Int.(s) .+ 1 for s in inp["subsets"]]
    cands = [shape_from(d) for d in inp["candidates"]]
    out = Dict{String,Any}("format" => 1, "producer" => "RANSAC.jl $(isdefined(Base, :pkgversion) ? pkgversion(RANSAC) : "(version unknown: Julia < 1.9)") on Julia $(VERSION)")

    pc = RANSACCloud(P, N, subsets)                                                         # octree.jl:96-103
    pc.isenabled[Int.(inp["disabled"]) .+ 1] .= false
    out["compatibles"] = [findall(compat(s, pc.vertices, pc.normals, params)) .- 1 for s in cands]
    out["scorecandidate"] = map(cands) do s
        sc, ip = RANSAC.scorecandidate(pc, s, 1, params)
        Dict("E" => RANSAC.E(sc), "inpoints" => ip .- 1)
    end
    out["refit"] = [RANSAC.refit(s, pc, params).inpoints .- 1 for s in cands]
    out["fits"] = map(inp["minimal_sets"]) do sd0
        sd = Int.(sd0) .+ 1
        map(params.iteration.shape_types) do T
            f = RANSAC.fit(T, pc.vertices[sd], pc.normals[sd], pc, params)
            f === nothing ? nothing : shape_json(f)
        end
    end
    out["estimatescore"] = map(inp["estimatescore"]) do a
        ci = RANSAC.estimatescore(Int(a[1]), Int(a[2]), Int(a[3]))                           # confidenceintervals.jl:71-74
        [ci.min, ci.max, ci.E]
    end

    pc2 = RANSACCloud(P, N, subsets)
    extracted, extracted_at, iterations = loop_with_sets(pc2, params, inp["loop"]["sets"])
    out["loop"] = Dict("iterations" => iterations, "extracted_at" => extracted_at,
                       "extracted" => [merge(shape_json(e.shape), Dict("inpoints" => e.inpoints .- 1)) for e in extracted],
                       "isenabled" => Int.(collect(pc2.isenabled)))
    open(joinpath(DIR, "julia_reference.json"), "w") do io
        JSON.print(io, out)
    end
    println("wrote ", joinpath(DIR, "julia_reference.json"), ": ", length(cands), " candidates, ", length(extracted), " extracted shapes")
end

main()
