# MLECuda.jl -- the reference-side binding of libmle_cuda: what a MultilevelEstimators.jl maintainer adds so that
# `Estimator(index_set, sample_method, sample_function, distributions; options...)` / `run(estimator, tol)` keep working
# unchanged while `parallel_sample!` runs on B200s.
#
# NOTE: Julia is not installed in the build container or on the GPU image, so this file has never been executed; it is
# deliberately thin and mechanical (one `ccall` per entry point of include/mle_cuda.h, checked statically against the header by
# tests/test_abi.py).  The same calls, in the same order, are exercised by the Python twin
# (multilevelestimators.jl_b200/host.py + engine.py) in tests/.
#
# ---- how it plugs in ---------------------------------------------------------------------------------------------------------
# The device returns sufficient statistics (n, mean, M2) instead of raw samples.  Everything the reference computes from its
# stored sample vectors goes through `length`, `isempty`, `Statistics.mean`, `Statistics.var` and `append!`
# (src/internals.jl:66-86, src/methods.jl:44-68, :121-122, :282, :336-351, src/history.jl:63-76), so the binding does not
# replace those functions one by one: it supplies a vector type, `MomentVector <: AbstractVector{Float64}`, that answers exactly
# those five queries from a moment triplet, and the estimator stores MomentVectors where it stored `Vector{Float64}`.
# `has_samples_at_index`, `all_keys`, the `interp` filter, `mean/var/mean0/var0/varest`, the U accumulator statistics and the
# History keys E/V/dE/dV then work UNMODIFIED.  The maintainer's patch is four places:
#
#  P1  src/estimator.jl:132 -- pass the store type to the internals:
#          store = sample_function isa MLECuda.DeviceModel ? MLECuda.MomentVector : Vector{Float64}
#          internals = EstimatorInternals(index_set, sample_method, options, store)
#  P2  src/internals.jl:17-22, :36-46, :205-214 -- thread `store` through EstimatorInternals / DefaultInternals /
#      IndexSetInternals(::U) and build the sample stores and the U accumulator with `store()` instead of `Vector{T}(undef, 0)`.
#  P3  src/sample.jl:39 and :79 (first line of both `parallel_sample!` methods):
#          estimator.sample_function isa MLECuda.DeviceModel && return MLECuda.device_parallel_sample!(estimator, index, Istart, Iend)
#      `sample!` (src/sample.jl:21-34) keeps timing the call and doing the nb_of_samples / total_work / total_time bookkeeping.
#  P4  (optional, QMC with a cost model) src/sample.jl:8, first line of `update_samples`:
#          estimator.sample_function isa MLECuda.DeviceModel && !(estimator[:cost_model] isa EmptyFunction) &&
#              return MLECuda.device_update_samples!(estimator, n_opt)
#      all due indices in ONE library call with one synchronisation.
#
# `sample_function` must be a `Function` (the field is declared `sample_function::Function`, src/estimator.jl:115), so the device
# model handle is a callable struct (`struct DeviceModel <: Function`, precedent: `struct EmptyFunction <: Function end`,
# src/parse.jl:178).  `nb_of_workers` (src/parse.jl:224-234) becomes the number of GPUs of the context: `Context(0:7)` is ONE
# multi-GPU context (mle_init_multi) and the library splits every request over its devices.
# `save_samples = true` (src/history.jl:92-95) makes the device return the raw samples as well; a MomentVector then also keeps
# them (`raw`), so `h[:samples]` / `h[:samples_diff]` hold real values.
module MLECuda

using MultilevelEstimators
using Statistics
import MultilevelEstimators: Estimator, AbstractIndexSet, MC, QMC, U, Index, parallel_sample!, nb_of_samples, distributions,
                             generator, Uniform, Normal, TruncatedNormal, Weibull

const libmle = "libmle_cuda"   # multilevelestimators.jl_b200/lib/libmle_cuda.so on the library path
const ABI_VERSION = 1           # MLE_ABI_VERSION of include/mle_cuda.h this file was written against

# ---- the sample store: (n, mean, M2) standing where a Vector{Float64} of samples stood -----------------------------------------
mutable struct MomentVector <: AbstractVector{Float64}
    n::Int
    mean::Float64
    M2::Float64                 # sum (x - mean)^2
    raw::Vector{Float64}        # the samples themselves, only with save_samples = true
end
MomentVector() = MomentVector(0, NaN, 0.0, Float64[])
MomentVector(t::AbstractVector{Float64}) = MomentVector(Int(t[1]), t[2], t[3], Float64[])   # one (n, mean, M2) triplet of the ABI
Base.size(v::MomentVector) = (v.n,)
Base.length(v::MomentVector) = v.n
Base.isempty(v::MomentVector) = v.n == 0
Base.getindex(v::MomentVector, i::Int) = isempty(v.raw) ? error("MomentVector keeps (n, mean, M2), not the samples (set save_samples = true)") : v.raw[i]
Base.copy(v::MomentVector) = MomentVector(v.n, v.mean, v.M2, copy(v.raw))
Statistics.mean(v::MomentVector) = v.n == 0 ? NaN : v.mean
Statistics.var(v::MomentVector; corrected::Bool = true) = v.n == 0 ? NaN : v.M2 / (corrected ? v.n - 1 : v.n)   # n = 1: 0/0 = NaN like var([x])
# append!: Chan / Golub / LeVeque merge of two triplets by the library's own mle_moments_merge (calls are merged in call order,
# with the same rounding as the Python twin)
function Base.append!(v::MomentVector, p::MomentVector)
    if p.n > 0
        a = Float64[v.n, v.n == 0 ? 0.0 : v.mean, v.M2]
        b = Float64[p.n, p.mean, p.M2]
        ccall((:mle_moments_merge, libmle), Cint, (Ptr{Float64}, Ptr{Float64}, Csize_t), a, b, 1)
        v.n, v.mean, v.M2 = Int(a[1]), a[2], a[3]
        append!(v.raw, p.raw)
    end
    v
end

struct MLEDist
    kind::Int32
    reserved::Int32
    p::NTuple{4, Float64}
end
MLEDist(d::Uniform) = MLEDist(0, 0, (Float64(d.a), Float64(d.b), 0.0, 0.0))
MLEDist(d::Normal) = MLEDist(1, 0, (Float64(d.μ), Float64(d.σ), 0.0, 0.0))
MLEDist(d::TruncatedNormal) = MLEDist(2, 0, (Float64(d.μ), Float64(d.σ), Float64(d.a), Float64(d.b)))
MLEDist(d::Weibull) = MLEDist(3, 0, (Float64(d.k), Float64(d.λ), 0.0, 0.0))

function check(ctx, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:mle_last_error, libmle), Cstring, (Ptr{Cvoid},), ctx))
    # MLE_ERR_ARG (-1) and MLE_ERR_RANGE (-3) are the reference's ArgumentError (src/check_input.jl)
    (rc == -1 || rc == -3) ? throw(ArgumentError(msg)) : error("libmle_cuda error $rc: $msg")
end

# One context = the reference's worker pool.  `Context(0)`: one GPU; `Context(0:7)` / `Context(nb_of_workers = 8)`: ONE multi-GPU
# context (mle_init_multi) -- handles are replicated on every device and each request is split inside the library (whole shifts
# per GPU; small requests stay on one GPU), so the estimator code never sees more than one handle.
mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(devices = 0; nb_of_workers::Integer = 0)
        devs = nb_of_workers > 0 ? collect(Cint, 0:nb_of_workers - 1) : collect(Cint, devices)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = length(devs) == 1 ?
            ccall((:mle_init, libmle), Cint, (Cint, Ref{Ptr{Cvoid}}), devs[1], ref) :
            ccall((:mle_init_multi, libmle), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), devs, length(devs), ref)
        rc == 0 || error(unsafe_string(ccall((:mle_last_error, libmle), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        ccall((:mle_abi_version, libmle), Cint, ()) == ABI_VERSION || error("libmle_cuda: ABI version mismatch")
        finalizer(c -> ccall((:mle_destroy, libmle), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        ctx
    end
end

"""
    DeviceModel(ctx, name, ndims, params; bases = Dict{Int, Matrix{Float64}}())

A built-in CUDA sample function ("expsum", "problem18", "lognormal1d", "lognormal2d").  `bases[level]` is the KL basis of the
level, `size = (nodes, terms)`; it is handed over row-major (`permutedims`).
"""
mutable struct DeviceModel <: Function
    ctx::Context
    ptr::Ptr{Cvoid}
    nqoi::Int
    lattice::Ptr{Cvoid}      # created lazily from the estimator's point generator
    dists::Ptr{Cvoid}
end

# `bases`: KL basis per level (key = level) or, for the multi-index 2-D model (name "lognormal2d", ndims = 2), per multi-index:
# the C ABI packs the zero-based index (lx, ly) into one Int as lx + 16*ly (see `basis_key`).
basis_key(level::Integer) = Int(level)
basis_key(index::NTuple{2, <:Integer}) = Int(index[1]) + 16 * Int(index[2])
basis_key(index::Index{2}) = basis_key(index.I)

function DeviceModel(ctx::Context, name::String, ndims::Integer, params::Vector{Float64} = Float64[];
                     bases::Dict = Dict{Int, Matrix{Float64}}())
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.ptr, ccall((:mle_model_create, libmle), Cint, (Ptr{Cvoid}, Cstring, Cint, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
                         ctx.ptr, name, ndims, params, length(params), ref))
    for (level, Φ) in bases
        rowmajor = permutedims(Φ)            # Julia is column-major; the ABI wants [nodes][terms] row-major
        check(ctx.ptr, ccall((:mle_model_set_basis, libmle), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}),
                             ctx.ptr, ref[], basis_key(level), size(Φ, 1), size(Φ, 2), rowmajor))
    end
    q = Ref{Cint}(0)
    ccall((:mle_model_nqoi, libmle), Cint, (Ptr{Cvoid}, Ref{Cint}), ref[], q)
    m = DeviceModel(ctx, ref[], q[], C_NULL, C_NULL)
    finalizer(m) do x
        x.lattice == C_NULL || ccall((:mle_lattice_destroy, libmle), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.ptr, x.lattice)
        x.dists == C_NULL || ccall((:mle_dists_destroy, libmle), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.ptr, x.dists)
        ccall((:mle_model_destroy, libmle), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.ptr, x.ptr)
    end
    m
end

# Calling the handle like the closure it replaces is an error: samples are taken in batches on the device.
(m::DeviceModel)(index, ξ) = error("DeviceModel is evaluated by libmle_cuda inside parallel_sample!, not per sample")

function ensure_handles!(m::DeviceModel, estimator)
    if m.dists == C_NULL
        d = MLEDist.(distributions(estimator))
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(m.ctx.ptr, ccall((:mle_dists_create, libmle), Cint, (Ptr{Cvoid}, Ptr{MLEDist}, Cint, Ref{Ptr{Cvoid}}),
                               m.ctx.ptr, d, length(d), ref))
        m.dists = ref[]
    end
end

function ensure_lattice!(m::DeviceModel, rule)
    if m.lattice == C_NULL
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(m.ctx.ptr, ccall((:mle_lattice_create, libmle), Cint, (Ptr{Cvoid}, Ptr{UInt32}, Cint, UInt64, Ref{Ptr{Cvoid}}),
                               m.ctx.ptr, rule.z, length(rule.z), UInt64(rule.n), ref))
        m.lattice = ref[]
    end
end

# The generating vectors the library embeds (the reference's src/generating_vectors/*.txt): "K_3600_32", "CKN_250_20".
function genvec(name::String)
    z, len = Ref{Ptr{UInt32}}(C_NULL), Ref{Cint}(0)
    ccall((:mle_genvec, libmle), Cint, (Cstring, Ref{Ptr{UInt32}}, Ref{Cint}), name, z, len) == 0 || throw(ArgumentError("unknown generating vector $name"))
    copy(unsafe_wrap(Array, z[], len[]))
end

"""
    device_points(estimator, index, Istart, Iend) -> Array{Float64, 3}  (m x n x R)

Parity hook for stages 1-2: `transform.(distributions[1:m], get_point(generator(estimator, index, r), k)[1:m])` for
`k = Istart-1 : Iend-1` and every shift `r` (src/sample.jl:84-88), computed on the device.
"""
function device_points(estimator::Estimator{I, <:QMC}, index::Index, Istart::Integer, Iend::Integer) where I<:AbstractIndexSet
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    R = estimator[:nb_of_shifts](index)
    gens = [generator(estimator, index, r) for r in 1:R]
    ensure_lattice!(model, gens[1].lattice_rule)
    shifts = reduce(hcat, [Vector{Float64}(g.shift) for g in gens])
    n = Iend - Istart + 1
    out = Array{Float64}(undef, m, n, R)
    GC.@preserve shifts out check(model.ctx.ptr,
        ccall((:mle_points, libmle), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64, Ptr{Float64}),
              model.ctx.ptr, model.lattice, model.dists, shifts, R, UInt64(Istart - 1), UInt64(n), m, UInt64(0), out))
    out
end

# ---- results into the estimator: the reference's own append functions, fed with MomentVectors -----------------------------------
# mom: (3, 2, q, R) = (n, mean, M2) x (dQ, Qf) x QoI x shift;  smp (optional): (n, 2, q, R)
function part(mom, X, q, r, smp)
    v = MomentVector(view(mom, :, X, q, r))
    smp === nothing || (v.raw = smp[:, X, q, r])
    v
end
function store_moments!(estimator::Estimator{<:AbstractIndexSet, <:QMC}, index::Index, mom, smp = nothing)
    for r in axes(mom, 4), q in axes(mom, 3)
        MultilevelEstimators.append_samples_diff!(estimator, q, r, index, part(mom, 1, q, r, smp))   # src/internals.jl:80
        MultilevelEstimators.append_samples!(estimator, q, r, index, part(mom, 2, q, r, smp))        # src/internals.jl:66
    end
end
function store_moments!(estimator::Estimator{<:AbstractIndexSet, <:MC}, index::Index, mom, smp = nothing)
    for q in axes(mom, 3)
        MultilevelEstimators.append_samples_diff!(estimator, q, index, part(mom, 1, q, 1, smp))
        MultilevelEstimators.append_samples!(estimator, q, index, part(mom, 2, q, 1, smp))
    end
end

shift_matrix(estimator, index, R) = reduce(hcat, [Vector{Float64}(generator(estimator, index, r).shift) for r in 1:R])   # [s x R]

# MC inputs: the device's uniforms are keyed by (seed, point number, coordinate) and point numbers restart at 0 on every
# index, so every index gets its own stream (the reference draws rand(m) from one global stream, src/sample.jl:42).
# Same mixing as host.py:index_seed.
const MC_SEED = Ref{UInt64}(1)
function index_seed(seed::UInt64, index::Index)
    x = seed
    for c in index.I
        x += 0x9e3779b97f4a7c15 * UInt64(c + 1)
        x = (x ⊻ (x >> 30)) * 0xbf58476d1ce4e5b9
        x = (x ⊻ (x >> 27)) * 0x94d049bb133111eb
        x ⊻= x >> 31
    end
    x
end

# ---- the hot path: one ccall per parallel_sample! (replaces the pmap of src/sample.jl:39-56, :79-102) -----------------------------
function device_parallel_sample!(estimator::Estimator{I, S}, index::Index, Istart::Integer, Iend::Integer) where {I<:AbstractIndexSet, S<:Union{MC, QMC}}
    I <: U && return device_sample_unbiased!(estimator, index, Istart, Iend)
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    qmc = S <: QMC
    R = qmc ? estimator[:nb_of_shifts](index) : 1
    qmc && ensure_lattice!(model, generator(estimator, index, 1).lattice_rule)
    shifts = qmc ? shift_matrix(estimator, index, R) : Float64[]
    n = Iend - Istart + 1
    mom = Array{Float64}(undef, 3, 2, model.nqoi, R)
    smp = estimator[:save_samples] ? Array{Float64}(undef, n, 2, model.nqoi, R) : nothing
    idx = Int32[index.I...]
    GC.@preserve shifts mom smp idx check(model.ctx.ptr,
        ccall((:mle_sample_batch, libmle), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64,
               Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              model.ctx.ptr, model.ptr, qmc ? model.lattice : C_NULL, model.dists, idx, length(idx),
              qmc ? pointer(shifts) : Ptr{Float64}(C_NULL), R, UInt64(Istart - 1), UInt64(n), m,
              qmc ? UInt64(0) : index_seed(MC_SEED[], index), mom, smp === nothing ? Ptr{Float64}(C_NULL) : pointer(smp),
              Ptr{Float64}(C_NULL)))
    store_moments!(estimator, index, mom, smp)
end

# ---- all due indices in one call: update_samples (src/sample.jl:8-13) with ONE synchronisation (patch P4) ------------------------
# Needs a deterministic `cost_model`: with the default cost model the reference charges every sample! call its own wall time
# (src/sample.jl:27); here the call's wall time is shared out over the indices in proportion to their modelled work.
function device_update_samples!(estimator::Estimator{I, <:QMC}, n_opt::Dict) where I<:AbstractIndexSet
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    due = [(τ, n_opt[τ] - nb_of_samples(estimator, τ)) for τ in keys(n_opt) if n_opt[τ] > nb_of_samples(estimator, τ)]
    isempty(due) && return
    d = length(first(due)[1].I)
    indices = Int32[c for (τ, _) in due for c in τ.I]                      # [count x d], entry e at indices + e*d
    Rs = Cint[estimator[:nb_of_shifts](τ) for (τ, _) in due]
    ms = Cint[estimator[:nb_of_uncertainties](τ) for (τ, _) in due]
    k0 = UInt64[nb_of_samples(estimator, τ) for (τ, _) in due]              # zero-based first point = samples taken so far
    ns = UInt64[n for (_, n) in due]
    shifts = [shift_matrix(estimator, τ, R) for ((τ, _), R) in zip(due, Rs)]
    ensure_lattice!(model, generator(estimator, first(due)[1], 1).lattice_rule)
    moms = [Array{Float64}(undef, 3, 2, model.nqoi, R) for R in Rs]
    t = @elapsed GC.@preserve shifts moms begin
        sp = Ptr{Float64}[pointer(x) for x in shifts]
        mp = Ptr{Float64}[pointer(x) for x in moms]
        check(model.ctx.ptr,
            ccall((:mle_update_samples, libmle), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Ptr{Ptr{Float64}}, Ptr{Cint}, Ptr{UInt64},
                   Ptr{UInt64}, Ptr{Cint}, UInt64, Ptr{Ptr{Float64}}),
                  model.ctx.ptr, model.ptr, model.lattice, model.dists, length(due), indices, d, sp, Rs, k0, ns, ms, UInt64(0), mp))
    end
    work = [n * estimator[:cost_model](τ) for (τ, n) in due]
    for (((τ, n), mom), w) in zip(zip(due, moms), work)
        store_moments!(estimator, τ, mom)
        # the bookkeeping of sample! (src/sample.jl:29-31, src/internals.jl:90,98,106)
        MultilevelEstimators.add_to_nb_of_samples(estimator, τ, n)
        MultilevelEstimators.add_to_total_work(estimator, τ, n)
        MultilevelEstimators.add_to_total_time(estimator, τ, t * w / sum(work))
    end
end

# ---- unbiased estimators (src/sample.jl:65-74, :113-124): ONE call evaluates every idx in zero(index):index on the same points --
# and reduces the accumulator  Z = sum_idx 1/Prob(idx) * dQ_idx  per sample on the device (mle_sample_unbiased).  The per-idx
# moments are appended exactly where the reference's append_samples! for U appends the samples; the accumulator (a MomentVector
# too, patch P2) takes the (n, mean, M2) of Z: `append_to_accumulator!` + `add_to_accumulator!` collapse into one append.
function device_sample_unbiased!(estimator::Estimator{I, S}, index::Index, Istart::Integer, Iend::Integer) where {I<:U, S<:Union{MC, QMC}}
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    qmc = S <: QMC
    R = qmc ? estimator[:nb_of_shifts](index) : 1
    qmc && ensure_lattice!(model, generator(estimator, index, 1).lattice_rule)
    shifts = qmc ? shift_matrix(estimator, index, R) : Float64[]
    n = Iend - Istart + 1
    box = CartesianIndices(index.I .+ 1)                                   # column-major: the first entry runs fastest
    inv_prob = Float64[1 / MultilevelEstimators.Prob(estimator, Index((Tuple(c) .- 1)...)) for c in box]
    moms = Array{Float64}(undef, 3, 2, model.nqoi, R, length(box))
    acc = Array{Float64}(undef, 3, model.nqoi, R)
    idx = Int32[index.I...]
    GC.@preserve shifts moms acc idx inv_prob check(model.ctx.ptr,
        ccall((:mle_sample_unbiased, libmle), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64,
               Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              model.ctx.ptr, model.ptr, qmc ? model.lattice : C_NULL, model.dists, idx, length(idx),
              qmc ? pointer(shifts) : Ptr{Float64}(C_NULL), R, UInt64(Istart - 1), UInt64(n), m,
              qmc ? UInt64(0) : index_seed(MC_SEED[], index), inv_prob, moms, acc, Ptr{Float64}(C_NULL)))
    for (i, c) in enumerate(box)
        store_moments!(estimator, Index((Tuple(c) .- 1)...), view(moms, :, :, :, :, i))
    end
    for r in 1:R, q in 1:model.nqoi
        append!(MultilevelEstimators.accumulator(estimator, q, r), MomentVector(view(acc, :, q, r)))
    end
end

end # module
