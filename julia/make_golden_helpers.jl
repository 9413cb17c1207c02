# make_golden_helpers.jl -- shared by julia/make_golden.jl and julia/RANSACB200/test/runtests.jl:
# exchange-file <-> RANSAC.jl conversions and the reference loop on explicit minimal sets.
# (expects `using RANSAC, JSON`, `using StaticArrays: SVector` and the Fitted* / RANSACCloud names in scope)
const TYPES = Dict("plane" => FittedPlane, "sphere" => FittedSphere, "cylinder" => FittedCylinder, "cone" => FittedCone)

sv(v) = SVector{3,Float64}(v[1], v[2], v[3])

"the nested NamedTuple of ransacparameters (utilities.jl:332-433) from the exchange file"
function params_from(p)
    it = p["iteration"]
    shape_types = [TYPES[t] for t in it["shape_types"]]
    ransacparameters(shape_types;                                            # utilities.jl:463-466
        iteration = (drawN = Int(it["drawN"]), minsubsetN = Int(it["minsubsetN"]), prob_det = Float64(it["prob_det"]),
                     τ = Int(it["tau"]), itermax = Int(it["itermax"]), extract_s = Symbol(it["extract_s"]),
                     terminate_s = Symbol(it["terminate_s"])),
        common = (collin_threshold = Float64(p["common"]["collin_threshold"]), parallelthrdeg = Float64(p["common"]["parallelthrdeg"])),
        plane = (ϵ = Float64(p["plane"]["eps"]), α = Float64(p["plane"]["alpha"])),
        sphere = (ϵ = Float64(p["sphere"]["eps"]), α = Float64(p["sphere"]["alpha"]), sphere_par = Float64(p["sphere"]["sphere_par"])),
        cylinder = (ϵ = Float64(p["cylinder"]["eps"]), α = Float64(p["cylinder"]["alpha"])),
        cone = (ϵ = Float64(p["cone"]["eps"]), α = Float64(p["cone"]["alpha"]), minconeopang = Float64(p["cone"]["minconeopang"])))
end

function shape_from(d)
    t = d["type"]
    t == "plane" && return FittedPlane(sv(d["point"]), sv(d["normal"]))                                     # plane.jl:8-11
    t == "sphere" && return FittedSphere(sv(d["center"]), Float64(d["radius"]), Bool(d["outwards"]))        # sphere.jl:9-13
    t == "cylinder" && return FittedCylinder(sv(d["axis"]), sv(d["center"]), Float64(d["radius"]), Bool(d["outwards"]))  # cylinder.jl:11-16
    return FittedCone(sv(d["apex"]), sv(d["axis"]), Float64(d["opang"]), Bool(d["outwards"]))                # cone.jl:11-19
end

shape_json(s::FittedPlane) = Dict("type" => "plane", "point" => collect(s.point), "normal" => collect(s.normal))
shape_json(s::FittedSphere) = Dict("type" => "sphere", "center" => collect(s.center), "radius" => s.radius, "outwards" => s.outwards)
shape_json(s::FittedCylinder) = Dict("type" => "cylinder", "axis" => collect(s.axis), "center" => collect(s.center),
                                     "radius" => s.radius, "outwards" => s.outwards)
shape_json(s::FittedCone) = Dict("type" => "cone", "apex" => collect(s.apex), "axis" => collect(s.axis), "opang" => s.opang,
                                 "outwards" => s.outwards)

compat(s::FittedPlane, P, N, prm) = RANSAC.compatiblesPlane(s, P, N, prm)        # plane.jl:114-130
compat(s::FittedSphere, P, N, prm) = RANSAC.compatiblesSphere(s, P, N, prm)      # sphere.jl:144-172
compat(s::FittedCylinder, P, N, prm) = RANSAC.compatiblesCylinder(s, P, N, prm)  # cylinder.jl:194-221
compat(s::FittedCone, P, N, prm) = RANSAC.compatiblesCone(s, P, N, prm)          # cone.jl:132-153

"iterations.jl:35-162 with the minimal sets read from the file instead of samplepointcloud4! (Julia's
random stream is not part of the contract; everything else is the package's own code, call for call)"
function loop_with_sets(pc, params, sets)
    it = params.iteration
    drawN, minsubsetN, prob_det, τ, itermax = it.drawN, it.minsubsetN, it.prob_det, it.τ, it.itermax
    candidates = FittedShape[]
    scoredshapes = IterationCandidates()
    extracted = ExtractedShape[]
    shape_octree_level = Int[]
    countcandidates = [0, 0, 0]
    extracted_at = Int[]
    iterations = 0
    for k in 1:itermax
        count(pc.isenabled) < τ && break                                                    # iterations.jl:75
        iterations = k
        for i in 1:minsubsetN
            (k > length(sets) || sets[k][i] === nothing) && continue
            sd = Int.(sets[k][i]) .+ 1
            f_v = @view pc.vertices[sd]
            f_n = @view pc.normals[sd]
            RANSAC.forcefitshapes!(f_v, f_n, params, candidates, shape_octree_level, 1, pc)  # iterations.jl:89
        end
        countcandidates[2] += size(candidates, 1)                                           # iterations.jl:94
        RANSAC.scorecandidates!(pc, scoredshapes, candidates, 1, params, shape_octree_level) # iterations.jl:95
        countcandidates[3] = k * minsubsetN
        countcandidates[1] = length(scoredshapes)
        if !(length(scoredshapes) < 1)
            best = RANSAC.findhighestscore(scoredshapes)
            bestshape = scoredshapes.shapes[best.index]
            scr = RANSAC.E(scoredshapes.scores[best.index])
            s = RANSAC.chooseS(countcandidates, it.extract_s)
            if RANSAC.prob(scr, s, pc.size, drawN) > prob_det                               # iterations.jl:113
                extr_shape = RANSAC.refit(bestshape, pc, params)
                if !(extr_shape === nothing)
                    RANSAC.invalidate_indexes!(pc, extr_shape.inpoints)
                    push!(extracted, extr_shape)
                    push!(extracted_at, k)
                    deleteat!(scoredshapes, best.index)
                    RANSAC.removeinvalidshapes!(pc, scoredshapes)
                end
            end
        end
        RANSAC.updatelevelweight(pc)                                                        # iterations.jl:148
        s = RANSAC.chooseS(countcandidates, it.terminate_s)
        RANSAC.prob(τ, s, pc.size, drawN) > prob_det && break                               # iterations.jl:151-156
    end
    return extracted, extracted_at, iterations
end

