# GeoStatsProcessesB200.jl - reference-side glue for libgspb200 (see INTEGRATION.md).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia, so boundary item (b) is UNTESTED on the Julia side
# (the same entry points with the same marshalling are what the Python ctypes tests drive).  The file is the binding a
# maintainer of GeoStatsProcesses.jl would add (as a package extension, exactly like
# ext/GeoStatsProcessesTuringPatternsExt.jl adds a method from outside): two method structs that plug
# into the unchanged `rand(process, domain, n; method=...)` (src/simulation/field.jl:47-124) through
# `preprocess` / `randsingle` (field.jl:64,67,87,90).  Every ccall below is declared in include/gsp_b200.h
# and exercised from Python (ctypes) by this repository's tests with identical argument marshalling.
module GeoStatsProcessesB200

using GeoStatsProcesses
using GeoStatsProcesses: FieldSimulationMethod, GaussianProcess, Ensemble, NearestInit, initialize
using GeoStatsFunctions
import GeoTables
using GeoTables: georef, nrow
using Meshes
using LinearAlgebra
using Random
using Unitful: ustrip, unit

import GeoStatsProcesses: preprocess, randsingle
# the statistics of src/ensembles.jl:42-52 are Distributions' generic functions (GeoStatsProcesses.jl:28-30 imports them);
# importing the bindings from GeoStatsProcesses extends exactly the functions `mean(e::Ensemble)` etc. dispatch on
import GeoStatsProcesses: mean, var, cdf, ccdf, quantile

const LIB = get(ENV, "GSP_B200_LIB", "libgspb200")

# ---------------------------------------------------------------- C structs (include/gsp_b200.h)
struct CStructure
  kind::Int32
  reserved::Int32
  sill::Float64
  A::NTuple{9,Float64}     # 3x3 row-major, u = |A * delta|
  param::Float64           # MaternCovariance / MaternVariogram: order nu; 0 otherwise
end

struct CCovModel
  nstruct::Int32
  reserved::Int32
  structs::Ptr{CStructure}
end

struct CDomain
  kind::Int32              # 0 points, 1 CartesianGrid
  dim::Int32
  nelems::Int64
  coords::Ptr{Float64}
  dims::NTuple{3,Int64}
  origin::NTuple{3,Float64}
  spacing::NTuple{3,Float64}
end

const KINDS = Dict(NuggetEffect => 0, SphericalCovariance => 1, ExponentialCovariance => 2, GaussianCovariance => 3,
                   CubicCovariance => 4, PentasphericalCovariance => 5, SineHoleCovariance => 6, CircularCovariance => 7,
                   MaternCovariance => 8, MaternVariogram => 8,
                   SphericalVariogram => 1, ExponentialVariogram => 2, GaussianVariogram => 3,
                   CubicVariogram => 4, PentasphericalVariogram => 5, SineHoleVariogram => 6, CircularVariogram => 7)

# ---------------------------------------------------------------- context (one per process, all visible GPUs)
mutable struct Context
  ptr::Ptr{Cvoid}
end

function Context(devices::Vector{Int32}=Int32[0])
  ref = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:gsp_ctx_create, LIB), Cint, (Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}), length(devices), devices, ref)
  rc == 0 || error("gsp_ctx_create failed with code $rc (no CPU fallback exists)")
  ctx = Context(ref[])
  finalizer(c -> ccall((:gsp_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
  ctx
end

const CTX = Ref{Union{Nothing,Context}}(nothing)
context() = (isnothing(CTX[]) && (CTX[] = Context()); CTX[])

lasterror(ctx) = unsafe_string(ccall((:gsp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr))

function check(ctx, rc)
  rc == 0 && return
  rc > 0 && throw(PosDefException(rc))                      # cholesky's failure mode (lusim.jl:92,98,103)
  -32 <= rc < 0 && throw(ArgumentError("libgspb200: argument $(-rc): $(lasterror(ctx))"))
  error("libgspb200 error $rc: $(lasterror(ctx))")
end

# ---------------------------------------------------------------- marshalling
# metric of one basic structure: isotropic range r -> I/r; MetricBall(radii, R) -> diag(1/radii) * R'
function metricmatrix(γ)
  b = GeoStatsFunctions.metricball(γ)
  r = ustrip.(Meshes.radii(b))
  R = Matrix(Meshes.rotation(b))
  d = length(r)
  A = zeros(3, 3)
  A[1:d, 1:d] = Diagonal(1 ./ r) * R'
  ntuple(i -> A'[i], 9)                                      # row-major
end

# flatten `structures(f)` (lusim.jl:133) of the marginal of variable j into C structures
function flatten(f, j)
  cₒ, cs, fs = GeoStatsFunctions.structures(f)
  out = CStructure[]
  nug = ustrip(cₒ[j, j])
  iszero(nug) || push!(out, CStructure(0, 0, nug, ntuple(i -> i in (1, 5, 9) ? 1.0 : 0.0, 9), 0.0))
  for (c, g) in zip(cs, fs)
    k = KINDS[typeof(g).name.wrapper]
    push!(out, CStructure(k, 0, ustrip(c[j, j]), metricmatrix(g), k == 8 ? Float64(g.order) : 0.0))
  end
  out
end

function cdomain(dom)
  g = parent(dom)
  if g isa CartesianGrid && dom === g
    d = embeddim(g)
    o = ustrip.(to(minimum(g)))
    s = ustrip.(spacing(g))
    pad(t, v) = ntuple(i -> i <= d ? t[i] : v, 3)
    return CDomain(1, d, nelements(g), C_NULL, pad(Int64.(size(g)), 1), pad(Float64.(o), 0.0), pad(Float64.(s), 1.0)), nothing
  end
  X = reduce(hcat, [collect(ustrip.(to(centroid(dom, i)))) for i in 1:nelements(dom)])  # dim x N (lusim.jl:81-82)
  CDomain(0, size(X, 1), size(X, 2), pointer(X), (1, 1, 1), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), X
end

# data nodes of one variable: `findall(mask[var])` and the values there (lusim.jl:71-75).  NearestInit on a CartesianGrid runs on
# the device (gsp_nearest_init: grid arithmetic instead of the KD-tree over all centroids that nearest.jl:16 builds); every other
# combination keeps the reference's `initialize` result.
function datanodes(domain, data, init, real, mask, var)
  if !isnothing(data) && init isa GeoStatsProcesses.NearestInit && domain isa CartesianGrid
    X = reduce(hcat, [collect(Float64.(ustrip.(to(centroid(GeoTables.domain(data), i))))) for i in 1:nrow(data)])   # dim x nd
    v = [ismissing(x) ? NaN : Float64(ustrip(x)) for x in getproperty(data, var)]
    nd = length(v)
    dinds, z₁, cnt = Vector{Int64}(undef, nd), Vector{Float64}(undef, nd), Ref{Int64}(0)
    cdom, _ = cdomain(domain)
    ctx = context()
    check(ctx, ccall((:gsp_nearest_init, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{CDomain}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}),
                     ctx.ptr, Ref(cdom), nd, X, v, dinds, z₁, cnt))
    return dinds[1:cnt[]], z₁[1:cnt[]]
  end
  dinds = Int64.(findall(mask[var]))                                     # ascending, 1-based (lusim.jl:71)
  dinds, Float64.(ustrip.(view(real[var], dinds)))
end

# ---------------------------------------------------------------- LUSIM on the GPU
"""
    LUSIM_B200(; batch=64, seed=nothing)

Drop-in replacement of `LUSIM()`.  `batch` realizations are produced by one device call (the L22*W
contraction is a tensor-core GEMM only when W is a matrix); `randsingle` hands them out one by one.
With `seed=nothing` the noise is drawn on the host with the `rng` given to `rand` (same draw order as
lusim.jl:160) and injected; with an integer seed the on-device counter RNG is used.
"""
Base.@kwdef struct LUSIM_B200 <: FieldSimulationMethod
  batch::Int = 64
  seed::Union{Nothing,UInt64} = nothing
end

mutable struct LUPre
  plans::Vector{Ptr{Cvoid}}
  vars::Vector{Symbol}
  units::Vector{Any}
  ρ::Float64
  N::Int
  Ns::Int
  buffer::Vector{Matrix{Float64}}    # per variable: N x batch
  cursor::Int
  served::Int
  lock::ReentrantLock
end

function preprocess(::AbstractRNG, process::GaussianProcess, method::LUSIM_B200, init, domain, data)
  f, μ = process.func, process.mean
  isvalid(f) = isstationary(f) && issymmetric(f) && isbanded(f)
  isvalid(f) || throw(ArgumentError("""
      LUSIM requires a geostatistical function that is stationary, symmetric and banded.
      Covariances or composite functions of covariances satisfy these properties.
    """))
  real, mask = initialize(process, domain, data, init)                     # lusim.jl:53
  vars = collect(keys(real))
  @assert length(vars) == nvariables(f) "incompatible number of variables for geostatistical function"
  @assert length(vars) ∈ (1, 2) "LUSIM only supports univariate and bivariate simulation"
  ctx = context()
  cdom, keep = cdomain(domain)
  plans = Ptr{Cvoid}[]
  keys = Tuple{Vector{CStructure},Vector{Int64}}[]
  Ns = 0
  GC.@preserve keep begin
    for (j, var) in enumerate(vars)
      dinds, z₁ = datanodes(domain, data, init, real, mask, var)
      structs = flatten(f, j)
      model = Ref(CCovModel(length(structs), 0, pointer(structs)))
      ref = Ref{Ptr{Cvoid}}(C_NULL)
      # the `map` at lusim.jl:66-107 assembles and factors once per variable; a variable whose marginal covariance and data nodes
      # equal an earlier one's (e.g. [1 ρ; ρ 1] * cov with shared data locations) shares that factor and only gets its own d₂
      twin = findfirst(k -> k == (structs, dinds), keys)
      GC.@preserve structs dinds z₁ begin
        rc = if isnothing(twin)
          ccall((:gsp_lu_plan_create, LIB), Cint,
                (Ptr{Cvoid}, Ptr{CCovModel}, Ptr{CDomain}, Int64, Ptr{Int64}, Ptr{Float64}, Float64, Ptr{Ptr{Cvoid}}),
                ctx.ptr, model, Ref(cdom), length(dinds), dinds, z₁, Float64(ustrip(μ[j])), ref)
        else
          ccall((:gsp_lu_plan_create_like, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Float64, Ptr{Ptr{Cvoid}}),
                plans[twin], length(dinds), dinds, z₁, Float64(ustrip(μ[j])), ref)
        end
      end
      check(ctx, rc)
      push!(plans, ref[])
      push!(keys, (structs, dinds))
      Ns = nelements(domain) - length(dinds)
    end
  end
  ρ = length(vars) == 2 ? Float64(GeoStatsProcesses._rho(f)) : NaN
  pre = LUPre(plans, vars, [unit(eltype(real[v])) for v in vars], ρ, nelements(domain), Ns, Matrix{Float64}[], 0, 0, ReentrantLock())
  finalizer(p -> foreach(h -> ccall((:gsp_lu_plan_destroy, LIB), Cint, (Ptr{Cvoid},), h), p.plans), pre)
  pre
end

function refill!(rng, method::LUSIM_B200, pre::LUPre)
  ctx = context()
  R, nv = method.batch, length(pre.plans)
  pre.buffer = [Matrix{Float64}(undef, pre.N, R) for _ in 1:nv]
  if isnothing(method.seed)
    W = [Matrix{Float64}(undef, pre.Ns, R) for _ in 1:nv]
    for r in 1:R, v in 1:nv                                                # per realization: w1 then w2 (lusim.jl:114-119,160)
      randn!(rng, view(W[v], :, r))
    end
    check(ctx, ccall((:gsp_lu_sample, LIB), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Float64}, UInt64, Int32, Int64, Float64, Ptr{Float64}, Ptr{Float64}),
                     pre.plans[1], R, W[1], 0, 0, pre.served, NaN, C_NULL, pre.buffer[1]))
    nv == 2 && check(ctx, ccall((:gsp_lu_sample, LIB), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Float64}, UInt64, Int32, Int64, Float64, Ptr{Float64}, Ptr{Float64}),
                     pre.plans[2], R, W[2], 0, 1, pre.served, pre.ρ, W[1], pre.buffer[2]))
  else
    for v in 1:nv
      check(ctx, ccall((:gsp_lu_sample, LIB), Cint,
                       (Ptr{Cvoid}, Int64, Ptr{Float64}, UInt64, Int32, Int64, Float64, Ptr{Float64}, Ptr{Float64}),
                       pre.plans[v], R, C_NULL, method.seed, v - 1, pre.served, v == 1 ? NaN : pre.ρ, C_NULL, pre.buffer[v]))
    end
  end
  pre.cursor = 0
end

function randsingle(rng::AbstractRNG, ::GaussianProcess, method::LUSIM_B200, domain, data, pre::LUPre)
  lock(pre.lock) do
    (isempty(pre.buffer) || pre.cursor == method.batch) && refill!(rng, method, pre)
    pre.cursor += 1
    pre.served += 1
    cols = (pre.vars[v] => pre.buffer[v][:, pre.cursor] .* pre.units[v] for v in eachindex(pre.vars))
    (; cols...)
  end
end

# ---------------------------------------------------------------- FFTSIM on the GPU (unconditional path)
Base.@kwdef struct FFTSIM_B200 <: FieldSimulationMethod
  batch::Int = 16
  seed::Union{Nothing,UInt64} = nothing
  minneighbors::Int = 1          # conditioning (fftsim.jl:47-52); `neighborhood` / `distance` other than the defaults stay with FFTSIM()
  maxneighbors::Int = 26
end

mutable struct FFTPre
  plan::Ptr{Cvoid}
  var::Symbol
  inds::Vector{Int64}
  dims::Dims
  buffer::Matrix{Float64}
  cursor::Int
  served::Int
  lock::ReentrantLock
end

function preprocess(::AbstractRNG, process::GaussianProcess, method::FFTSIM_B200, init, domain, data)
  f = process.func
  @assert isstationary(f) "geostatistical function must be stationary"
  real, mask = initialize(process, domain, data, init)
  @assert length(keys(real)) == 1 "FFTSIM does not support multivariate simulation"
  grid = parent(domain)
  ctx = context()
  cdom, _ = cdomain(grid)
  structs = flatten(f, 1)
  model = Ref(CCovModel(length(structs), 0, pointer(structs)))
  ref = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve structs begin
    check(ctx, ccall((:gsp_fft_plan_create, LIB), Cint, (Ptr{Cvoid}, Ptr{CCovModel}, Ptr{CDomain}, Ptr{Ptr{Cvoid}}),
                     ctx.ptr, model, Ref(cdom), ref))
  end
  inds = domain === grid ? Int64[] : Int64.(collect(parentindices(domain)))
  if !isnothing(data)
    # fftsim.jl:94-104: zbar = simple Kriging of the data where they are (k nearest, Euclidean); dinds = findall(mask[var]).
    # The weights of the per-realization Kriging (fftsim.jl:140-149) depend on geometry only and are built here, once.
    var = first(keys(real))
    knodes, _ = datanodes(domain, data, init, real, mask, var)
    vals = getproperty(data, var)
    keep = findall(!ismissing, vals)
    X = reduce(hcat, [collect(Float64.(ustrip.(to(centroid(GeoTables.domain(data), i))))) for i in keep])   # dim x nd, where the data are
    v = Float64.(ustrip.(vals[keep]))
    check(ctx, ccall((:gsp_fft_plan_condition, LIB), Cint,
                     (Ptr{Cvoid}, Float64, Int32, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}),
                     ref[], Float64(ustrip(process.mean)), method.minneighbors, method.maxneighbors, length(v), X, v,
                     length(knodes), knodes, length(inds), isempty(inds) ? C_NULL : inds))
  end
  pre = FFTPre(ref[], first(keys(real)), inds, size(grid), Matrix{Float64}(undef, 0, 0), 0, 0, ReentrantLock())
  finalizer(p -> ccall((:gsp_fft_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.plan), pre)
  pre
end

function randsingle(rng::AbstractRNG, process::GaussianProcess, method::FFTSIM_B200, domain, data, pre::FFTPre)
  lock(pre.lock) do
    if isempty(pre.buffer) || pre.cursor == method.batch
      ctx, R = context(), method.batch
      n = isempty(pre.inds) ? prod(pre.dims) : length(pre.inds)
      pre.buffer = Matrix{Float64}(undef, n, R)
      w = isnothing(method.seed) ? rand(rng, Float64, prod(pre.dims), R) : nothing     # rand(rng, Float64, dims) (fftsim.jl:124)
      check(ctx, ccall((:gsp_fft_sample, LIB), Cint,
                       (Ptr{Cvoid}, Int64, Ptr{Float64}, UInt64, Int64, Float64, Float64, Int64, Ptr{Int64}, Ptr{Float64}),
                       pre.plan, R, isnothing(w) ? C_NULL : w, something(method.seed, UInt64(0)), pre.served,
                       Float64(ustrip(sill(process.func))), Float64(ustrip(process.mean)), length(pre.inds),
                       isempty(pre.inds) ? C_NULL : pre.inds, pre.buffer))
      pre.cursor = 0
    end
    pre.cursor += 1
    pre.served += 1
    (; pre.var => pre.buffer[:, pre.cursor] .* unit(process.mean))
  end
end

# ------------------------------------------------------------------------------------------------
# Posterior mean (src/expectation/field/gaussian.jl:21-25: simple Kriging of the data onto the domain, fitpredict's default search:
# k nearest neighbours, maxneighbors = 10) with the device Kriging that conditions FFTSIM.  `mean(process, domain; data)` has no
# method argument to dispatch on, so the offload is a separate function.
function mean_b200(process::GaussianProcess, domain; data, init=NearestInit(), minneighbors=1, maxneighbors=10)
  pre = preprocess(Random.default_rng(), process, FFTSIM_B200(; minneighbors, maxneighbors), init, domain, data)
  n = isempty(pre.inds) ? prod(pre.dims) : length(pre.inds)
  z̄ = Vector{Float64}(undef, n)
  check(context(), ccall((:gsp_fft_plan_condmean, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), pre.plan, z̄))
  georef((; pre.var => z̄ .* unit(process.mean)), domain)
end

# ------------------------------------------------------------------------------------------------
# Device-resident ensembles: Ensemble(domain, reals; fetch) (src/ensembles.jl:10-16) whose realizations stay in
# HBM.  `reals` are lightweight handles, `fetch` downloads one realization (ensembles.jl:27-31), and the
# statistics of ensembles.jl:42-52 are overloaded to run on the device instead of the O(n R) `ereduce` loops.
struct DeviceReals
  ptr::Ptr{Cvoid}          # gsp_ensemble*
  var::Symbol
  n::Int
  R::Int
end
struct DeviceReal          # element of `reals`
  parent::DeviceReals
  r::Int                   # 0-based
end
Base.length(d::DeviceReals) = d.R
Base.getindex(d::DeviceReals, i::Int) = DeviceReal(d, i - 1)
Base.first(d::DeviceReals) = d[1]

function fetchreal(x::DeviceReal)
  z = Vector{Float64}(undef, x.parent.n)
  check(context(), ccall((:gsp_ensemble_fetch, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), x.parent.ptr, x.r, 1, z))
  (; x.parent.var => z)
end

"""
    rand_resident(process, domain, nreals; method=FFTSIM_B200(seed=0), data=nothing, init=NearestInit())

Like `rand(process, domain, nreals; method)` (src/simulation/field.jl:72-91) but the ensemble stays on the GPUs.
"""
function rand_resident(process::GaussianProcess, domain, nreals::Int; method=FFTSIM_B200(), data=nothing, init=NearestInit())
  rng = Random.default_rng()
  pre = preprocess(rng, process, method, init, domain, data)
  ctx = context()
  n = nelements(domain)
  ref = Ref{Ptr{Cvoid}}(C_NULL)
  check(ctx, ccall((:gsp_ensemble_create, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Ptr{Cvoid}}), ctx.ptr, n, nreals, ref))
  seed = something(method.seed, rand(rng, UInt64))
  if method isa FFTSIM_B200
    check(ctx, ccall((:gsp_fft_sample_ensemble, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, UInt64, Int64, Float64, Float64, Int64, Ptr{Int64}),
                     pre.plan, ref[], C_NULL, seed, 0, Float64(ustrip(sill(process.func))), Float64(ustrip(process.mean)),
                     length(pre.inds), isempty(pre.inds) ? C_NULL : pre.inds))
    var = pre.var
  else  # LUSIM_B200, first variable (a second variable gets its own ensemble with stream = 1 and rho)
    check(ctx, ccall((:gsp_lu_sample_ensemble, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, UInt64, Int32, Int64, Float64, Ptr{Float64}),
                     pre.plans[1], ref[], C_NULL, seed, 0, 0, NaN, C_NULL))
    var = pre.vars[1]
  end
  reals = DeviceReals(ref[], var, n, nreals)
  Ensemble(domain, reals; fetch=fetchreal)
end

# one explicit ccall per statistic: ccall needs a literal (symbol, library) pair and a literal argument-type tuple
function devvec(e::Ensemble{<:Any,DeviceReals}, call)
  z = Vector{Float64}(undef, e.reals.n)
  check(context(), call(e.reals.ptr, z))
  georef((; e.reals.var => z), e.domain)
end
mean(e::Ensemble{<:Any,DeviceReals}) =                                                      # ensembles.jl:42
  devvec(e, (h, z) -> ccall((:gsp_ensemble_mean, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, z))
var(e::Ensemble{<:Any,DeviceReals}) =                                                       # ensembles.jl:44
  devvec(e, (h, z) -> ccall((:gsp_ensemble_var, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, z))
cdf(e::Ensemble{<:Any,DeviceReals}, x::Number) =                                            # ensembles.jl:46
  devvec(e, (h, z) -> ccall((:gsp_ensemble_cdf, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), h, Float64(x), z))
ccdf(e::Ensemble{<:Any,DeviceReals}, x::Number) =                                           # ensembles.jl:48
  devvec(e, (h, z) -> ccall((:gsp_ensemble_ccdf, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), h, Float64(x), z))
function quantile(e::Ensemble{<:Any,DeviceReals}, ps::AbstractVector)                       # ensembles.jl:50-52
  n = e.reals.n
  q = Matrix{Float64}(undef, n, length(ps))
  pv = Float64.(ps)
  check(context(), ccall((:gsp_ensemble_quantile, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
                         e.reals.ptr, length(pv), pv, q))
  [georef((; e.reals.var => q[:, k]), e.domain) for k in eachindex(pv)]
end
quantile(e::Ensemble{<:Any,DeviceReals}, p::Number) = first(quantile(e, [p]))
release!(e::Ensemble{<:Any,DeviceReals}) = ccall((:gsp_ensemble_destroy, LIB), Cint, (Ptr{Cvoid},), e.reals.ptr)

export LUSIM_B200, FFTSIM_B200, rand_resident, mean_b200, release!

end # module
