# MLECuda.jl -- the reference-side binding of libmle_cuda: what a MultilevelEstimators.jl maintainer adds so that
# `Estimator(index_set, sample_method, sample_function, distributions; options...)` / `run(estimator, tol)` keep working
# unchanged while `parallel_sample!` runs on a B200.
#
# NOTE: Julia is not installed in the build container or on the GPU image, so this file has never been executed; it is
# deliberately thin and mechanical (one `ccall` per entry point of include/mle_cuda.h).  The same calls, in the same order,
# are exercised by the Python twin (multilevelestimators.jl_b200/host.py + engine.py) in tests/.
#
# Hook: `sample_function` must be a `Function` (the field is declared `sample_function::Function`, src/estimator.jl:115),
# so the device model handle is a callable struct (`struct DeviceModel <: Function`, precedent:
# `struct EmptyFunction <: Function end`, src/parse.jl:178).  The field is NOT a type parameter of `Estimator`
# (src/estimator.jl:113-119), so `parallel_sample!` cannot dispatch on it; the maintainer adds ONE line at the top of each
# of the two methods in src/sample.jl (:39 MC, :79 QMC):
#
#     estimator.sample_function isa MLECuda.DeviceModel && return MLECuda.device_parallel_sample!(estimator, index, Istart, Iend)
#
# and the statistics accessors of src/methods.jl:44-57, :282 read the moment table below instead of the sample vectors
# when the sample function is a DeviceModel (see `_mean`, `_var`, `_varest_qmc` at the end of this file).
module MLECuda

using MultilevelEstimators
import MultilevelEstimators: Estimator, AbstractIndexSet, MC, QMC, Index, parallel_sample!, nb_of_samples, distributions,
                             generator, Uniform, Normal, TruncatedNormal, Weibull

const libmle = "libmle_cuda"   # multilevelestimators.jl_b200/lib/libmle_cuda.so on the library path
const ABI_VERSION = 1           # MLE_ABI_VERSION of include/mle_cuda.h this file was written against

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

mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:mle_init, libmle), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref)
        rc == 0 || error(unsafe_string(ccall((:mle_last_error, libmle), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        ccall((:mle_abi_version, libmle), Cint, ()) == ABI_VERSION || error("libmle_cuda: ABI version mismatch")
        finalizer(c -> ccall((:mle_destroy, libmle), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        ctx
    end
end

"""
    DeviceModel(ctx, name, ndims, params; bases = Dict{Int, Matrix{Float64}}())

A built-in CUDA sample function ("expsum", "problem18", "lognormal1d").  `bases[level]` is the KL basis of the level,
`size = (nodes, terms)`; it is handed over row-major (`permutedims`).
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

# Moment state replaces the stored sample vectors (src/internals.jl:44-45): one (n, mean, M2) triplet per
# (X ∈ {dQ, Qf}, qoi, shift), merged across calls by mle_moments_merge.  Kept in a side table keyed by the estimator.
const MOMENTS = IdDict{Any, Dict{Any, Array{Float64, 4}}}()
moments(estimator, index) = get!(() -> Dict{Any, Array{Float64, 4}}(), MOMENTS, estimator)[index]

function merge_moments!(estimator, index, part::Array{Float64, 4})
    tab = get!(() -> Dict{Any, Array{Float64, 4}}(), MOMENTS, estimator)
    if haskey(tab, index)
        ccall((:mle_moments_merge, libmle), Cint, (Ptr{Float64}, Ptr{Float64}, Csize_t), tab[index], part, length(part) ÷ 3)
    else
        tab[index] = part
    end
end

# Index sets: the reference's own code keeps driving everything above parallel_sample! -- SL, ML, FT, TD, HC, ZC and AD need
# nothing else.  Unbiased estimators (`U`, src/sample.jl:65-74, :113-124) expect the sample function to return (dQ, Qf) for every
# idx <= index from one input vector: for those, call mle_sample_batch once per idx in `zero(index):index` with the SAME shifts
# and point range and a `samples` buffer, append the raw samples to `samples_diff[idx]` / `samples[idx]` and add
# `dQ_idx / Prob(idx)` to the accumulator, exactly as host.py:Estimator._sample_unbiased does (tested there).
#
# ---- the hot path: one ccall per parallel_sample! (replaces the pmap of src/sample.jl:79-102) ----------------------------
function device_parallel_sample!(estimator::Estimator{I, <:QMC}, index::Index, Istart::Integer, Iend::Integer) where I<:AbstractIndexSet
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    R = estimator[:nb_of_shifts](index)
    gens = [generator(estimator, index, r) for r in 1:R]
    rule = gens[1].lattice_rule
    ensure_lattice!(model, rule)
    shifts = reduce(hcat, [Vector{Float64}(g.shift) for g in gens])          # [s x R], shift r in column r
    mom = Array{Float64}(undef, 3, 2, model.nqoi, R)
    idx = Int32[index.I...]
    t = Ref{Float64}(0.0)
    GC.@preserve shifts mom idx check(model.ctx.ptr,
        ccall((:mle_sample_batch, libmle), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64,
               Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
              model.ctx.ptr, model.ptr, model.lattice, model.dists, idx, length(idx), shifts, R,
              UInt64(Istart - 1), UInt64(Iend - Istart + 1), m, UInt64(0), mom, C_NULL, t))
    merge_moments!(estimator, index, mom)
end

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

function device_parallel_sample!(estimator::Estimator{I, <:MC}, index::Index, Istart::Integer, Iend::Integer) where I<:AbstractIndexSet
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    mom = Array{Float64}(undef, 3, 2, model.nqoi, 1)
    idx = Int32[index.I...]
    GC.@preserve mom idx check(model.ctx.ptr,
        ccall((:mle_sample_batch, libmle), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64,
               Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              model.ctx.ptr, model.ptr, C_NULL, model.dists, idx, length(idx), C_NULL, 1,
              UInt64(Istart - 1), UInt64(Iend - Istart + 1), m, index_seed(MC_SEED[], index), mom, C_NULL, C_NULL))
    merge_moments!(estimator, index, mom)
end

# ---- all due indices in one call: update_samples (src/sample.jl:8-13) with ONE synchronisation -----------------------------
# Needs a deterministic `cost_model` (with the default cost model the reference times every sample! call separately,
# src/sample.jl:27).  The maintainer's hook is one line at the top of `update_samples`:
#
#     estimator.sample_function isa MLECuda.DeviceModel && !(estimator[:cost_model] isa EmptyFunction) &&
#         return MLECuda.device_update_samples!(estimator, n_opt)
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
    shifts = [reduce(hcat, [Vector{Float64}(generator(estimator, τ, r).shift) for r in 1:R]) for ((τ, _), R) in zip(due, Rs)]
    ensure_lattice!(model, generator(estimator, first(due)[1], 1).lattice_rule)
    moms = [Array{Float64}(undef, 3, 2, model.nqoi, R) for R in Rs]
    GC.@preserve shifts moms begin
        sp = Ptr{Float64}[pointer(x) for x in shifts]
        mp = Ptr{Float64}[pointer(x) for x in moms]
        check(model.ctx.ptr,
            ccall((:mle_update_samples, libmle), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Ptr{Ptr{Float64}}, Ptr{Cint}, Ptr{UInt64},
                   Ptr{UInt64}, Ptr{Cint}, UInt64, Ptr{Ptr{Float64}}),
                  model.ctx.ptr, model.ptr, model.lattice, model.dists, length(due), indices, d, sp, Rs, k0, ns, ms, UInt64(0), mp))
    end
    for ((τ, n), mom) in zip(due, moms)
        merge_moments!(estimator, τ, mom)
        # the bookkeeping of sample! (src/sample.jl:29-31, src/internals.jl:90,98); total_time is not measured per index here
        MultilevelEstimators.add_to_nb_of_samples(estimator, τ, n)
        MultilevelEstimators.add_to_total_work(estimator, τ, n)
    end
end

# ---- unbiased estimators: every idx <= index on the same points, raw samples back (src/sample.jl:65-74, :113-124) ----------
# Returns samples[idx] = Array (n x 2 x q x R): [:, 1, :, :] = dQ_idx, [:, 2, :, :] = Qf_idx; the caller appends them and adds
# dQ_idx / Prob(idx) to the accumulator exactly as the reference's append_samples! for U does.
function device_sample_unbiased(estimator::Estimator{I, <:QMC}, index::Index, Istart::Integer, Iend::Integer) where I<:AbstractIndexSet
    model = estimator.sample_function
    ensure_handles!(model, estimator)
    m = estimator[:nb_of_uncertainties](index)
    R = estimator[:nb_of_shifts](index)
    gens = [generator(estimator, index, r) for r in 1:R]
    ensure_lattice!(model, gens[1].lattice_rule)
    shifts = reduce(hcat, [Vector{Float64}(g.shift) for g in gens])
    n = Iend - Istart + 1
    out = Dict{Any, Array{Float64, 4}}()
    for idx in CartesianIndices(index.I .+ 1)
        sub = Int32[Tuple(idx)...] .- Int32(1)
        mom = Array{Float64}(undef, 3, 2, model.nqoi, R)
        smp = Array{Float64}(undef, n, 2, model.nqoi, R)
        GC.@preserve shifts mom smp sub check(model.ctx.ptr,
            ccall((:mle_sample_batch, libmle), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Float64}, Cint, UInt64, UInt64, Cint, UInt64,
                   Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  model.ctx.ptr, model.ptr, model.lattice, model.dists, sub, length(sub), shifts, R,
                  UInt64(Istart - 1), UInt64(n), m, UInt64(0), mom, smp, C_NULL))
        out[Index(sub...)] = smp
    end
    out
end

# ---- statistics from the moment state (replace src/methods.jl:44-57, :282 for DeviceModel estimators) ---------------------
# mom[:, X, q, r] = (n, mean, M2);  X = 1: dQ (samples_diff), X = 2: Qf (samples)
_mean(mom, X, q) = sum(mom[2, X, q, r] for r in axes(mom, 4)) / size(mom, 4)
_var(mom, X, q) = sum(mom[3, X, q, r] / (mom[1, X, q, r] - 1) for r in axes(mom, 4)) / size(mom, 4)
_varest_qmc(mom, q) = (R = size(mom, 4); μ = [mom[2, 1, q, r] for r in 1:R]; sum(abs2, μ .- sum(μ) / R) / (R - 1) / R)

end # module
