# make_golden.jl - pins the oracle (oracle/gsp_oracle.py) to the REAL GeoStatsProcesses.jl v0.13.0.
#
# This image has no Julia, so the fixtures this script writes do not exist in the repository yet;
# tests/test_reference_fixtures.py is xfail(strict) until they do and then asserts oracle == fixture at 1e-12
# (and, under -m gpu, CUDA path == fixture at 1e-9).  Anyone with Julia >= 1.10 runs, once:
#
#     julia --project=@gsp -e 'using Pkg; Pkg.add(name="GeoStatsProcesses", version="0.13.0");
#                              Pkg.add(["GeoStatsFunctions", "GeoStatsModels", "GeoTables", "Meshes", "Unitful"])'
#     julia --project=@gsp tests/golden/make_golden.jl            # writes tests/golden/julia/
#
# What it settles (DESIGN.md "parity unpinned"): GeoStatsFunctions' model formulas and practical-range scaling, Meshes'
# centroid / linear-index convention, lusim.jl's block algebra and rho-mixing, fftsim.jl's centre index, DC handling and
# variance scaling, and - for conditional FFTSIM - GeoStatsModels' neighbour search (tie-breaking at equal distances) and
# the point/element support of the second Kriging (fftsim.jl:143-149).
#
# Noise capture: the reference draws inside randsingle (lusim.jl:160 `randn(rng, n)`, fftsim.jl:124 `rand(rng, T, dims)`).
# preprocess consumes no random numbers for LUSIM / FFTSIM, so a copy of the RNG taken before `rand(rng, ...)` replays the
# identical draws in the identical order (w1 then w2 per realization, lusim.jl:114-119); the replayed arrays are stored.
#
# Output format: raw little-endian Float64 / Int64 files `<case>.<name>.bin` + `manifest.toml` (stdlib TOML, no extra packages).
using GeoStatsProcesses, GeoStatsFunctions, GeoTables, Meshes, Unitful
using GeoStatsFunctions: pairwise
using Random, TOML, LinearAlgebra

const OUT = joinpath(@__DIR__, "julia")
mkpath(OUT)
const MANIFEST = Dict{String,Any}("generator" => "GeoStatsProcesses $(pkgversion(GeoStatsProcesses))", "cases" => Dict{String,Any}())

function dump(case, name, a::AbstractArray{T}) where {T<:Union{Float64,Int64}}
  open(joinpath(OUT, "$case.$name.bin"), "w") do io
    write(io, htol.(vec(collect(a))))
  end
  Dict("dtype" => string(T), "shape" => collect(size(a)))   # column-major shape as Julia sees it
end
field(t, v) = Float64.(ustrip.(getproperty(t, v)))

# ---- one case per covariance model: pairwise over a small anisotropic 3-D point set (utils.jl:50-62)
let rng = MersenneTwister(1), X = rand(rng, 3, 40) .* 12
  pts = PointSet([Point(X[:, i]...) for i in 1:40])
  ball = MetricBall((9.0, 4.0, 2.0), RotZ(deg2rad(30)))         # helpers.aniso3(kind, sill, (9, 4, 2), 30 deg)
  models = ["spherical" => SphericalCovariance, "exponential" => ExponentialCovariance, "gaussian" => GaussianCovariance,
            "cubic" => CubicCovariance, "pentaspherical" => PentasphericalCovariance, "sinehole" => SineHoleCovariance,
            "circular" => CircularCovariance]
  for (nm, M) in models
    f = M(ball, sill=0.9, nugget=0.1)
    case = "pairwise_$nm"
    MANIFEST["cases"][case] = Dict("kind" => nm, "sill" => 0.9, "nugget" => 0.1, "ranges" => [9.0, 4.0, 2.0], "angle_deg" => 30.0,
                                   "X" => dump(case, "X", X), "C" => dump(case, "C", Float64.(ustrip.(GeoStatsProcesses._pairwise(f, pts)))))
  end
  for ν in (0.5, 1.5, 2.5)
    f = MaternCovariance(ball, sill=0.9, nugget=0.1, order=ν)
    case = "pairwise_matern_$(replace(string(ν), "." => "p"))"
    MANIFEST["cases"][case] = Dict("kind" => "matern", "order" => ν, "sill" => 0.9, "nugget" => 0.1, "ranges" => [9.0, 4.0, 2.0],
                                   "angle_deg" => 30.0, "X" => dump(case, "X", X),
                                   "C" => dump(case, "C", Float64.(ustrip.(GeoStatsProcesses._pairwise(f, pts)))))
  end
end

# ---- LUSIM: unconditional, conditional, bivariate conditional (lusim.jl:38-175)
function lusim_case(case, proc, grid, data, nv; seed)
  rng = MersenneTwister(seed)
  replay = copy(rng)
  real = rand(rng, proc, grid; method=LUSIM(), data=data)
  vars = [n for n in propertynames(real) if n != :geometry]
  nd = isnothing(data) ? 0 : length(unique(GeoStatsProcesses.initialize(proc, grid, data, NearestInit())[2][vars[1]] |> findall))
  ns = nelements(grid) - nd
  d = Dict{String,Any}("dims" => collect(size(grid)), "nvars" => nv, "vars" => string.(vars))
  for (j, v) in enumerate(vars)
    d["W$j"] = dump(case, "W$j", randn(replay, ns))
    d["Z$j"] = dump(case, "Z$j", field(real, v))
  end
  if !isnothing(data)
    X = reduce(hcat, [collect(ustrip.(to(centroid(domain(data), i)))) for i in 1:nelements(domain(data))])
    d["dcoords"] = dump(case, "dcoords", X)
    for v in vars
      d["dvals_$v"] = dump(case, "dvals_$v", Float64.(getproperty(data, v)))
    end
  end
  MANIFEST["cases"][case] = d
end
let grid = CartesianGrid(20, 15)
  proc = GaussianProcess(SphericalCovariance(range=8.0, sill=1.3, nugget=0.05), 0.7)
  MANIFEST["cases"]["lusim_uni_meta"] = Dict("kind" => "spherical", "range" => 8.0, "sill" => 1.3, "nugget" => 0.05, "mean" => 0.7)
  lusim_case("lusim_uni", proc, grid, nothing, 1; seed=11)
  pts = [(2.5, 2.5), (10.2, 7.9), (17.5, 12.5), (5.1, 13.3), (10.4, 7.6)]      # the last two snap to the same node: later wins
  data = georef((; Z=[0.3, -1.1, 0.8, 0.2, 1.9]), pts)
  lusim_case("lusim_cond", proc, grid, data, 1; seed=12)
  func = [1.0 0.7; 0.7 1.0] * SphericalCovariance(range=8.0)
  proc2 = GaussianProcess(func, [0.1, 0.2])
  data2 = georef((; Cu=[0.0, 0.1, 0.0], Zn=[0.1, 0.0, 0.1]), [(2.5, 2.5), (5.0, 7.5), (17.5, 5.0)])
  MANIFEST["cases"]["lusim_bi_meta"] = Dict("kind" => "spherical", "range" => 8.0, "C" => [1.0, 0.7, 0.7, 1.0], "mean" => [0.1, 0.2])
  lusim_case("lusim_bi", proc2, grid, data2, 2; seed=13)
end

# ---- FFTSIM: 2-D, 3-D, view; conditional with maxneighbors 3 and 26 (fftsim.jl:54-156)
function fftsim_case(case, proc, dom, data; seed, kw...)
  rng = MersenneTwister(seed)
  replay = copy(rng)
  real = rand(rng, proc, dom; method=FFTSIM(; kw...), data=data)
  grid = parent(dom)
  d = Dict{String,Any}("dims" => collect(size(grid)), "w" => dump(case, "w", rand(replay, Float64, size(grid))),
                       "Z" => dump(case, "Z", field(real, first(n for n in propertynames(real) if n != :geometry))))
  dom === grid || (d["inds"] = dump(case, "inds", Int64.(collect(parentindices(dom)))))
  if !isnothing(data)
    X = reduce(hcat, [collect(ustrip.(to(centroid(domain(data), i)))) for i in 1:nelements(domain(data))])
    d["dcoords"] = dump(case, "dcoords", X)
    d["dvals"] = dump(case, "dvals", Float64.(data.Z))
    for (k, v) in kw
      d[string(k)] = v
    end
  end
  MANIFEST["cases"][case] = d
end
let
  proc = GaussianProcess(ExponentialCovariance(range=6.0, sill=1.5), 0.2)
  MANIFEST["cases"]["fftsim_meta"] = Dict("kind" => "exponential", "range" => 6.0, "sill" => 1.5, "mean" => 0.2)
  g2, g3 = CartesianGrid(32, 20), CartesianGrid(16, 12, 10)
  fftsim_case("fftsim_2d", proc, g2, nothing; seed=21)
  fftsim_case("fftsim_3d", proc, g3, nothing; seed=22)
  fftsim_case("fftsim_view", proc, view(g2, 1:3:nelements(g2)), nothing; seed=23)
  rng = MersenneTwister(5)
  pts = [(rand(rng) * 32, rand(rng) * 20) for _ in 1:40]
  push!(pts, (8.5, 8.5), (9.5, 8.5), (8.5, 9.5), (9.5, 9.5))                   # equidistant from node (9, 9)'s corner: tie-breaking
  data = georef((; Z=randn(rng, length(pts)) .+ 0.2), pts)
  fftsim_case("fftsim_cond_k3", proc, g2, data; seed=24, maxneighbors=3)
  fftsim_case("fftsim_cond_k26", proc, g2, data; seed=25, maxneighbors=26)
end

open(joinpath(OUT, "manifest.toml"), "w") do io
  TOML.print(io, MANIFEST)
end
println("wrote ", length(MANIFEST["cases"]), " cases to ", OUT)
