# CUDABackend.jl -- `ccall` glue between BayesNets.jl and libbn_cuda.so (include/bn_cuda.h).
#
# Add to BayesNets.jl as src/Inference/CUDABackend/CUDABackend.jl and `include` it from src/BayesNets.jl
# after src/Inference/*.jl and src/DiscreteBayesNet/device_layout.jl.  There is no CUDA.jl dependency and
# no CPU fallback: every method below is one ccall; if the library cannot be loaded the call throws.
# Julia is not available in the build image, so this file is reviewed, not executed; its executable twin
# is bayesnets.jl_b200/{inference,sampling,factors}.py, which binds the same symbols through ctypes and is
# what tests/ drives (INTEGRATION.md).
#
# Dispatch points used (the reference's own extension mechanism):
#   infer(::InferenceMethod, ::InferenceState)      src/ProbabilisticGraphicalModels/inference.jl:48-49
#   Base.:*, Base.sum, LinearAlgebra.normalize!     src/Factors/factors_dims.jl:269,100,45
#   Random.rand(bn, ::BayesNetSampler, n)           src/sampling.jl:24-41

const libbn_cuda = get(ENV, "BAYESNETS_LIBBN_CUDA", "libbn_cuda.so")

const BN_OK = Int32(0)
const BN_BAD_ARG = Int32(-1)
const BN_DIM_MISMATCH = Int32(-2)
const BN_NOT_IN_FACTOR = Int32(-3)
const BN_ZERO_WEIGHT = Int32(-5)
const BN_OOM = Int32(-6)
const BN_UNSUPPORTED = Int32(-7)
const BN_BOUNDS = Int32(-8)
const BN_NCCL_ERR = Int32(-9)

_bn_msg() = unsafe_string(ccall((:bn_last_error, libbn_cuda), Cstring, ()))

"Map bn_status to the exception the reference throws at the corresponding site (include/bn_cuda.h)."
function _bn_check(rc::Int32)
    rc == BN_OK && return nothing
    msg = _bn_msg()
    rc == BN_DIM_MISMATCH && throw(DimensionMismatch(msg))            # factors_dims.jl:197-199
    (rc == BN_BAD_ARG || rc == BN_NOT_IN_FACTOR) && throw(ArgumentError(msg))  # inference.jl:10-11, errors.jl:12
    rc == BN_BOUNDS && throw(BoundsError(msg))                         # factors_access.jl:59-61
    rc == BN_OOM && throw(OutOfMemoryError())
    rc == BN_NCCL_ERR && error("NCCL: $msg")
    error("libbn_cuda status $rc: $msg")
end

"""
    set_lw_mode(mode::Symbol=:auto; prune::Bool=true)
    set_gibbs_mode(mode::Symbol=:auto)

Kernel selection for this thread (include/bn_cuda.h): `:auto` uses the network-specialised kernel as soon as its cubin
is at hand (in the process or in the on-disk cache); until then the generic interpreter kernels serve the calls while
NVRTC compiles on a background thread -- no call waits for a compile.  `:generic` / `:specialized` force one or the
other (`:specialized` blocks on the first compile). `prune` lets the specialised LW kernel skip barren nodes (never
changes a count bit).
"""
const _KERNEL_MODES = Dict(:auto => Int32(0), :generic => Int32(1), :specialized => Int32(2))
set_lw_mode(mode::Symbol=:auto; prune::Bool=true) =
    _bn_check(ccall((:bn_set_lw_mode, libbn_cuda), Int32, (Int32, Int32), _KERNEL_MODES[mode], prune))
set_gibbs_mode(mode::Symbol=:auto) =
    _bn_check(ccall((:bn_set_gibbs_mode, libbn_cuda), Int32, (Int32,), _KERNEL_MODES[mode]))

"""
    cuda_pool_trim() -> bytes

Return the device memory that freed factors / per-call temporaries left parked inside the library to the CUDA driver
(for callers that share the GPU with another allocator).
"""
function cuda_pool_trim()
    n = Ref{UInt64}(0)
    _bn_check(ccall((:bn_pool_trim, libbn_cuda), Int32, (Ref{UInt64},), n))
    return Int(n[])
end

# Seeds: the reference draws fresh random numbers on every call (Julia's global RNG).  `seed = nothing` (the default
# everywhere below) takes the next value of a process-wide counter; an explicit seed reproduces a call bit for bit.
const _seed_counter = Ref{UInt64}(0)
_next_seed(seed::Nothing) = (_seed_counter[] += 1; UInt64(0x5EED000000000000) + _seed_counter[])
_next_seed(seed::Integer) = UInt64(seed)

# ------------------------------------------------------------------------------------------------
# Multi-GPU: one `infer` call for the whole machine (src/ProbabilisticGraphicalModels/inference.jl:48-49).
# A Julia session is ONE process, so it uses bn_comm_init_all: libbn_cuda.so creates one NCCL communicator and one
# stream per listed device, bn_net_create_comm replicates the network on all of them, and bn_lw_infer / bn_gibbs_infer
# on such a net shard the global particle / chain range over the GPUs and issue the one ncclAllReduce themselves.
# (Process-per-GPU launches -- MPI.jl, Distributed -- call bn_comm_unique_id on rank 0, broadcast the 128 bytes and
# call bn_comm_init_rank on every rank; the rest is identical.)
mutable struct Communicator
    handle::UInt64
    devices::Vector{Int32}
    function Communicator(handle::UInt64, devices::Vector{Int32})
        c = new(handle, devices)
        finalizer(x -> ccall((:bn_comm_destroy, libbn_cuda), Int32, (UInt64,), x.handle), c)
        return c
    end
end

"One process drives all of `devices` (ncclCommInitAll): what a Julia session uses."
function Communicator(devices::AbstractVector{<:Integer})
    devs = Int32[d for d in devices]
    h = Ref{UInt64}(0)
    GC.@preserve devs begin
        _bn_check(ccall((:bn_comm_init_all, libbn_cuda), Int32, (Ptr{Int32}, Int32, Ref{UInt64}), devs, length(devs), h))
    end
    return Communicator(h[], devs)
end

"Communicator of one process-per-GPU rank (`id` = rank 0's `comm_unique_id()`, shipped by the launcher)."
function Communicator(id::Vector{UInt8}, rank::Integer, world::Integer, device::Integer)
    length(id) == 128 || throw(ArgumentError("NCCL id must be 128 bytes"))
    h = Ref{UInt64}(0)
    GC.@preserve id begin
        _bn_check(ccall((:bn_comm_init_rank, libbn_cuda), Int32, (Ptr{UInt8}, Int32, Int32, Int32, Ref{UInt64}),
                        id, rank, world, device, h))
    end
    return Communicator(h[], Int32[device])
end

function comm_unique_id()
    id = zeros(UInt8, 128)
    _bn_check(ccall((:bn_comm_unique_id, libbn_cuda), Int32, (Ptr{UInt8},), id))
    return id
end

"(world, first_rank, n_local) of a communicator."
function comm_info(c::Communicator)
    w = Ref{Int32}(0); f = Ref{Int32}(0); n = Ref{Int32}(0)
    _bn_check(ccall((:bn_comm_info, libbn_cuda), Int32, (UInt64, Ref{Int32}, Ref{Int32}, Ref{Int32}), c.handle, w, f, n))
    return Int(w[]), Int(f[]), Int(n[])
end
comm_barrier(c::Communicator) = _bn_check(ccall((:bn_comm_barrier, libbn_cuda), Int32, (UInt64,), c.handle))

# one communicator per device list, made on first use
const _communicators = Dict{Vector{Int32},Communicator}()
_communicator(devices) = get!(() -> Communicator(devices), _communicators, Int32[d for d in devices])

# Resident copies of a network: one per (network CONTENTS, device or device list).  Keyed weakly by the BayesNet object
# and, inside, by a hash of the flattened layout, so a CPD edited in place is re-uploaded instead of served stale.
const _device_nets = WeakKeyDict{Any,Dict{Any,DeviceNet}}()
function device_net(bn::DiscreteBayesNet; device::Integer=0, devices=nothing)
    lay = DeviceLayout(bn)
    fp = hash((lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, lay.names))
    where_ = devices === nothing ? Int32(device) : Int32[d for d in devices]
    cache = get!(() -> Dict{Any,DeviceNet}(), _device_nets, bn)
    for k in collect(keys(cache))                  # contents changed: drop the stale copies
        k[1] == fp || delete!(cache, k)
    end
    return get!(cache, (fp, where_)) do
        devices === nothing ? DeviceNet(lay; device=device) : DeviceNet(lay, _communicator(devices))
    end
end

# ------------------------------------------------------------------------------------------------
# Likelihood weighting                                     replaces src/Inference/likelihood.jl:16-58
# `devices = 0:7` runs the ONE call on all eight GPUs (particles sharded by global index + one all-reduce inside the
# library); `devices = nothing` uses the single GPU `device`.
@with_kw struct CUDALikelihoodWeighting <: InferenceMethod
    nsamples::Int = 500
    seed::Union{UInt64,Nothing} = nothing
    device::Int = 0
    devices::Union{Nothing,AbstractVector{<:Integer}} = nothing
end

function infer(im::CUDALikelihoodWeighting, inf::InferenceState{BN}) where {BN<:DiscreteBayesNet}
    bn = inf.pgm
    net = device_net(bn; device=im.device, devices=im.devices)
    ev = evidence_vector(net.layout, bn, inf.evidence)
    q = query_indices(bn, inf.query)
    ϕ = Factor(inf.query, map(n -> ncategories(bn, n), inf.query))     # likelihood.jl:22
    stats = zeros(Float64, 2)
    GC.@preserve ev q ϕ stats begin
        _bn_check(ccall((:bn_lw_infer, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, Ptr{Int32}, Int32, UInt64, UInt64, UInt64, Ptr{Float64}, Ptr{Float64}),
                        net.handle, ev, q, length(q), 0, im.nsamples, _next_seed(im.seed), ϕ.potential, stats))
    end
    normalize!(ϕ)                                                       # likelihood.jl:56 (0/0 -> NaN as before)
    return ϕ
end

# ------------------------------------------------------------------------------------------------
# Gibbs inference                                   replaces src/Inference/gibbs.jl:66-145 / :161-234
@with_kw mutable struct CUDAGibbsNodewise <: InferenceMethod
    nsamples::Int = 2000
    burn_in::Int = 500
    thin::Int = 3
    state::Assignment = Assignment()
    n_chains::Int = 1          # 1 = the reference's single chain; >1 = independent chains, counts summed
    seed::Union{UInt64,Nothing} = nothing
    device::Int = 0
    devices::Union{Nothing,AbstractVector{<:Integer}} = nothing   # e.g. 0:7: chains sharded over the GPUs
end
@with_kw mutable struct CUDAGibbsFull <: InferenceMethod
    nsamples::Int = 2000
    burn_in::Int = 500
    thin::Int = 3
    state::Assignment = Assignment()
    n_chains::Int = 1
    seed::Union{UInt64,Nothing} = nothing
    device::Int = 0
    devices::Union{Nothing,AbstractVector{<:Integer}} = nothing
end
_gibbs_mode(::CUDAGibbsNodewise) = Int32(0)
_gibbs_mode(::CUDAGibbsFull) = Int32(1)

function infer(im::Union{CUDAGibbsNodewise,CUDAGibbsFull}, inf::InferenceState{BN}) where {BN<:DiscreteBayesNet}
    im.burn_in < im.nsamples || throw(ArgumentError("Burn in ($(im.burn_in)) " *
                "must be less than number of samples ($(im.nsamples))"))    # gibbs.jl:70-71
    bn = inf.pgm
    net = device_net(bn; device=im.device, devices=im.devices)
    lay = net.layout
    ev = evidence_vector(lay, bn, inf.evidence)
    q = query_indices(bn, inf.query)
    ϕ = Factor(inf.query, map(n -> ncategories(bn, n), inf.query))
    n = length(lay.names)
    st = zeros(UInt8, n, im.n_chains)           # chain-major on the device = column per chain here
    init = Int32(0)
    if !isempty(im.state)                        # resume from im.state (gibbs.jl:83-88)
        for c in 1:im.n_chains, (i, nm) in enumerate(lay.names)
            x = get(inf.evidence, nm, im.state[nm])
            1 <= x <= lay.card[i] || throw(BoundsError(1:lay.card[i], x))   # the state indexes the CPTs on the device
            st[i, c] = UInt8(x - 1)
        end
        init = Int32(1)
    end
    GC.@preserve ev q ϕ st begin
        _bn_check(ccall((:bn_gibbs_infer, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, Ptr{Int32}, Int32, UInt64, UInt64, Int64, Int64, Int64, Int32, UInt64,
                         Ptr{UInt8}, Int32, Ptr{Float64}, Ptr{Float64}),
                        net.handle, ev, q, length(q), 0, im.n_chains, im.nsamples, im.burn_in, im.thin,
                        _gibbs_mode(im), _next_seed(im.seed), st, init, ϕ.potential, C_NULL))
    end
    for (i, nm) in enumerate(lay.names)          # im.state = x (gibbs.jl:141)
        im.state[nm] = Int(st[i, 1]) + 1
    end
    normalize!(ϕ)
    return ϕ
end

# ------------------------------------------------------------------------------------------------
# Exact inference                                           replaces src/Inference/exact.jl:6-33
@with_kw struct CUDAExact <: InferenceMethod
    fused::Bool = true        # one product-and-sum kernel per eliminated variable (no materialised product)
    device::Int = 0
end

function infer(im::CUDAExact, inf::InferenceState{BN}) where {BN<:DiscreteBayesNet}
    bn = inf.pgm
    net = device_net(bn; device=im.device)
    ev = evidence_vector(net.layout, bn, inf.evidence)
    q = query_indices(bn, inf.query)
    isempty(q) && throw(ArgumentError("ExactInference needs at least one query variable"))
    out = zeros(Float64, prod(ncategories(bn, n) for n in inf.query))
    od = zeros(Int32, length(q)); oc = zeros(Int32, length(q))
    GC.@preserve ev q out od oc begin
        _bn_check(ccall((:bn_exact_infer, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, Ptr{Int32}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                        net.handle, ev, q, length(q), im.fused, od, oc, out))
    end
    # the result's dimension order depends on the join order, exactly as in the reference (exact.jl:31)
    return Factor(net.layout.names[od .+ 1], reshape(out, Int.(oc)...))
end

# ------------------------------------------------------------------------------------------------
# Loopy belief propagation                           replaces src/Inference/loopy_belief.jl:24-218
@with_kw struct CUDALoopyBelief <: InferenceMethod
    nsamples::Int = 500
    tol::Float64 = 1e-8
    iters_for_convergence::Int = 6
    device::Int = 0
end

"""
    infer_batch(im::CUDALoopyBelief, bn, queries, evidences) -> Matrix{Float64}

Many (query, evidence) pairs of one network in ONE kernel launch (one warp per pair).  Column `b` holds
P(queries[b] | evidences[b]), zero-padded to the largest cardinality.  Same numbers as `infer(LoopyBelief(), ...)`.
"""
function infer_batch(im::CUDALoopyBelief, bn::DiscreteBayesNet, queries::NodeNames, evidences::Vector{<:Assignment})
    length(queries) == length(evidences) || throw(ArgumentError("one evidence Assignment per query"))
    net = device_net(bn; device=im.device)
    lay = net.layout
    nn, B = length(lay.names), length(queries)
    ev = Matrix{Int8}(undef, nn, B)                       # instance-major rows, as the C ABI expects
    q = Vector{Int32}(undef, B)
    for b in 1:B
        inf = InferenceState(bn, queries[b], evidences[b])   # ArgumentError if query ∉ bn or ∈ evidence
        ev[:, b] = evidence_vector(lay, bn, inf.evidence)
        q[b] = query_indices(bn, inf.query)[1]
    end
    stride = Int32(maximum(lay.card))
    out = zeros(Float64, stride, B)
    iters = zeros(Int32, B)
    GC.@preserve ev q out iters begin
        _bn_check(ccall((:bn_loopy_infer, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, Ptr{Int32}, UInt64, Int32, Float64, Int32, Int32, Ptr{Float64}, Ptr{Int32}),
                        net.handle, ev, q, B, im.nsamples, im.tol, im.iters_for_convergence, stride, out, iters))
    end
    return out
end

function infer(im::CUDALoopyBelief, inf::InferenceState{BN}) where {BN<:DiscreteBayesNet}
    length(inf.query) == 1 || throw(ArgumentError("There can only be one query variable"))   # loopy_belief.jl:26
    query = first(inf.query)
    out = infer_batch(im, inf.pgm, [query], [inf.evidence])
    return Factor([query], out[1:ncategories(inf.pgm, query), 1])
end

# ------------------------------------------------------------------------------------------------
# Table samplers                                      replaces src/sampling.jl:24-41,52-57,72-93,153-174
struct CUDADirectSampler <: BayesNetSampler
    seed::Union{UInt64,Nothing}
end
CUDADirectSampler() = CUDADirectSampler(nothing)
struct CUDARejectionSampler <: BayesNetSampler
    evidence::Assignment
    max_nsamples::Int
    seed::Union{UInt64,Nothing}
end
CUDARejectionSampler(evidence::Assignment; max_nsamples::Int=100) = CUDARejectionSampler(evidence, max_nsamples, nothing)
struct CUDALikelihoodWeightedSampler <: BayesNetSampler
    evidence::Assignment
    seed::Union{UInt64,Nothing}
end
CUDALikelihoodWeightedSampler(evidence::Assignment) = CUDALikelihoodWeightedSampler(evidence, nothing)

function _sample_table(bn::DiscreteBayesNet, ev::Vector{Int8}, kind::Integer, attempts::Integer, n::Integer, seed)
    net = device_net(bn)
    nn = length(net.layout.names)
    states = Matrix{UInt8}(undef, n, nn)         # node-major on the device = one column per node
    w = Vector{Float64}(undef, n); ok = Vector{UInt8}(undef, n)
    GC.@preserve ev states w ok begin
        _bn_check(ccall((:bn_sample, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, Int32, Int32, UInt64, UInt64, UInt64, Ptr{UInt8}, Ptr{Float64}, Ptr{UInt8}),
                        net.handle, ev, kind, attempts, 0, n, _next_seed(seed), states, w, ok))
    end
    return states, w, ok
end

function Random.rand(bn::DiscreteBayesNet, s::CUDADirectSampler, nsamples::Integer)
    states, _, _ = _sample_table(bn, fill(Int8(-1), length(bn)), 0, 1, nsamples, s.seed)
    return DataFrame([name(cpd) => Int.(states[:, i]) .+ 1 for (i, cpd) in enumerate(bn.cpds)])
end

function Random.rand(bn::DiscreteBayesNet, s::CUDARejectionSampler, nsamples::Integer)
    ev = evidence_vector(device_net(bn).layout, bn, s.evidence)
    states, _, ok = _sample_table(bn, ev, 1, s.max_nsamples, nsamples, s.seed)
    all(!iszero, ok) || throw(KeyError(name(first(bn.cpds))))   # the reference's empty!(a) then a[n] (sampling.jl:36,87)
    return DataFrame([name(cpd) => Int.(states[:, i]) .+ 1 for (i, cpd) in enumerate(bn.cpds)])
end

function Random.rand(bn::DiscreteBayesNet, s::CUDALikelihoodWeightedSampler, nsamples::Integer)
    ev = evidence_vector(device_net(bn).layout, bn, s.evidence)
    states, w, _ = _sample_table(bn, ev, 2, 1, nsamples, s.seed)
    df = DataFrame([name(cpd) => Int.(states[:, i]) .+ 1 for (i, cpd) in enumerate(bn.cpds)])
    df[!, :p] = w ./ sum(w)                                       # sampling.jl:171-172
    return df
end

"""
    cuda_get_weighted_sample(bn, evidence) -> (Assignment, weight)

get_weighted_sample (src/sampling.jl:102-121): one likelihood-weighted particle, evidence clamped.
"""
function cuda_get_weighted_sample(bn::DiscreteBayesNet, evidence::Assignment; seed=nothing)
    lay = device_net(bn).layout
    ev = evidence_vector(lay, bn, evidence)
    states, w, _ = _sample_table(bn, ev, 2, 1, 1, seed)
    a = Assignment()
    for (i, nm) in enumerate(lay.names)
        a[nm] = Int(states[1, i]) + 1
    end
    return (a, w[1])
end
cuda_get_weighted_sample(bn::DiscreteBayesNet, pair::Pair{NodeName}...; seed=nothing) =
    cuda_get_weighted_sample(bn, Assignment(pair); seed=seed)

# gibbs_sample(bn, nsamples, burn_in; ...) on the device    replaces src/gibbs.jl:289-374 (discrete nets)
# Same keyword surface, validation order and messages as the reference (:289-324).  `time_limit` is validated but never
# triggers (the device loop has no wall clock; the reference only uses it to cut a slow CPU loop short, :218-220) and
# `max_cache_size` is accepted and unused (the string-keyed conditional cache, :43-71, is a CPU optimisation).
# Extras: `seed`, `n_chains` (1 = the reference's single chain; more = independent chains, nsamples rows each).
function cuda_gibbs_sample(bn::DiscreteBayesNet, nsamples::Integer, burn_in::Integer;
        thinning::Integer=0,
        consistent_with::Assignment=Assignment(),
        variable_order::Union{Vector{Symbol},Nothing}=nothing,
        time_limit::Union{Int,Nothing}=nothing,
        error_if_time_out::Bool=true,
        initial_sample::Union{Assignment,Nothing}=nothing,
        max_cache_size::Union{Int,Nothing}=nothing,
        seed=nothing, n_chains::Integer=1)
    nsamples > 0 || throw(ArgumentError("nsamples parameter less than 1"))       # gibbs.jl:299-301
    burn_in >= 0 || throw(ArgumentError("Negative burn_in parameter"))
    thinning >= 0 || throw(ArgumentError("Negative thinning parameter"))
    bn_names = names(bn)
    if variable_order !== nothing                                                  # gibbs.jl:302-311
        for name in bn_names
            name in variable_order || throw(ArgumentError("Gibbs sample variable_order must contain all variables in the Bayes Net"))
        end
        for name in variable_order
            name in bn_names || throw(ArgumentError("Gibbs sample variable_order contains a variable not in the Bayes Net"))
        end
    end
    if time_limit !== nothing                                                      # gibbs.jl:312-314
        time_limit > 0 || throw(ArgumentError(join(["Invalid time_limit specified (", time_limit, ")"])))
    end
    if initial_sample !== nothing                                                  # gibbs.jl:315-324
        for name in bn_names
            haskey(initial_sample, name) || throw(ArgumentError("Gibbs sample initial_sample must be an assignment with all variables in the Bayes Net"))
        end
        for name in keys(consistent_with)
            initial_sample[name] == consistent_with[name] || throw(ArgumentError("Gibbs sample initial_sample was inconsistent with consistent_with"))
        end
        pdf(bn, initial_sample) > 0 || throw(ArgumentError("Gibbs sample initial_sample has a pdf value of zero"))
    end
    net = device_net(bn)
    lay = net.layout
    ev = evidence_vector(lay, bn, consistent_with)
    free = [nm for (i, nm) in enumerate(lay.names) if ev[i] < 0]
    order = C_NULL
    if variable_order !== nothing       # evidence variables are never resampled (gibbs.jl:210): keep the free ones, in order
        order = Int32[findfirst(==(v), free) - 1 for v in variable_order if v in free]
    end
    init = C_NULL
    if initial_sample !== nothing
        init = repeat(UInt8[initial_sample[nm] - 1 for nm in lay.names], 1, n_chains)
    end
    out = Matrix{UInt8}(undef, n_chains * nsamples, length(lay.names))
    GC.@preserve ev order init out begin
        _bn_check(ccall((:bn_gibbs_sample, libbn_cuda), Int32,
                        (UInt64, Ptr{Int8}, UInt64, UInt64, Int64, Int64, Int64, Ptr{Int32}, UInt64, Ptr{UInt8}, Ptr{UInt8}),
                        net.handle, ev, 0, n_chains, nsamples, burn_in, thinning, order, _next_seed(seed), init, out))
    end
    samples = DataFrame([nm => Int.(out[:, i]) .+ 1 for (i, nm) in enumerate(lay.names) if ev[i] < 0])
    # columns for the variables that were conditioned on, Float64 like the reference (gibbs.jl:368-373)
    evidence = Dict{Symbol,Array{Float64,1}}()
    for varname in keys(consistent_with)
        evidence[varname] = ones(size(samples)[1]) * consistent_with[varname]
    end
    return hcat(samples, DataFrame(evidence))
end

"""
GibbsSampler on the device (src/gibbs.jl:410-451): same fields and defaults as the reference's `GibbsSampler`.
"""
mutable struct CUDAGibbsSampler <: BayesNetSampler
    evidence::Assignment
    burn_in::Int
    thinning::Int
    variable_order::Union{Vector{Symbol},Nothing}
    time_limit::Union{Int,Nothing}
    error_if_time_out::Bool
    initial_sample::Union{Assignment,Nothing}
    max_cache_size::Union{Int,Nothing}

    function CUDAGibbsSampler(evidence::Assignment=Assignment();
        burn_in::Int=100,
        thinning::Int=0,
        variable_order::Union{Vector{Symbol},Nothing}=nothing,
        time_limit::Union{Int,Nothing}=nothing,
        error_if_time_out::Bool=true,
        initial_sample::Union{Assignment,Nothing}=nothing,
        max_cache_size::Union{Int,Nothing}=nothing)
        new(evidence, burn_in, thinning, variable_order, time_limit, error_if_time_out, initial_sample, max_cache_size)
    end
end

Random.rand(bn::DiscreteBayesNet, sampler::CUDAGibbsSampler, nsamples::Integer) =
    cuda_gibbs_sample(bn, nsamples, sampler.burn_in; thinning=sampler.thinning,
        consistent_with=sampler.evidence, variable_order=sampler.variable_order,
        time_limit=sampler.time_limit, error_if_time_out=sampler.error_if_time_out,
        initial_sample=sampler.initial_sample, max_cache_size=sampler.max_cache_size)

function Random.rand!(a::Assignment, bn::DiscreteBayesNet, sampler::CUDAGibbsSampler)   # gibbs.jl:454-460
    df = rand(bn, sampler, 1)
    for name in names(bn)
        a[name] = df[1, name]
    end
    return a
end

# ------------------------------------------------------------------------------------------------
# Device-resident factors: * / sum / normalize!      replaces src/Factors/factors_dims.jl:45-57,60-100,185-269
mutable struct DeviceFactor
    handle::UInt64
    dimensions::NodeNames
    function DeviceFactor(h::UInt64, dims::NodeNames)
        f = new(h, dims)
        finalizer(x -> ccall((:bn_factor_free, libbn_cuda), Int32, (UInt64,), x.handle), f)
        return f
    end
end

# Symbols <-> int32 ids: a process-wide interning table (ids only need to agree between operands)
const _var_ids = Dict{Symbol,Int32}()
var_id(s::Symbol) = get!(() -> Int32(length(_var_ids)), _var_ids, s)
const _var_names = Dict{Int32,Symbol}()
function _ids(dims::NodeNames)
    ids = Int32[var_id(d) for d in dims]
    for (d, i) in zip(dims, ids); _var_names[i] = d; end
    return ids
end

function DeviceFactor(ϕ::Factor; device::Integer=0)
    h = Ref{UInt64}(0)
    ids = _ids(ϕ.dimensions); card = Int32[size(ϕ.potential)...]
    GC.@preserve ids card ϕ begin
        _bn_check(ccall((:bn_factor_upload, libbn_cuda), Int32,
                        (Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32, Ref{UInt64}),
                        length(ids), ids, card, ϕ.potential, device, h))
    end
    return DeviceFactor(h[], copy(ϕ.dimensions))
end

function _wrap(h::UInt64)
    nd = Ref{Int32}(0)
    _bn_check(ccall((:bn_factor_ndims, libbn_cuda), Int32, (UInt64, Ref{Int32}), h, nd))
    ids = zeros(Int32, nd[]); card = zeros(Int32, nd[])
    _bn_check(ccall((:bn_factor_shape, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Ptr{Int32}), h, ids, card))
    return DeviceFactor(h, Symbol[_var_names[i] for i in ids])
end

function Base.convert(::Type{Factor}, d::DeviceFactor)
    nd = length(d.dimensions)
    ids = zeros(Int32, nd); card = zeros(Int32, nd)
    _bn_check(ccall((:bn_factor_shape, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Ptr{Int32}), d.handle, ids, card))
    pot = Array{Float64}(undef, Int.(card)...)
    _bn_check(ccall((:bn_factor_download, libbn_cuda), Int32, (UInt64, Ptr{Float64}), d.handle, pot))
    return Factor(d.dimensions, pot)
end

function Base.:*(a::DeviceFactor, b::DeviceFactor)
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_product, libbn_cuda), Int32, (UInt64, UInt64, Ref{UInt64}), a.handle, b.handle, h))
    return _wrap(h[])
end

function Base.sum(a::DeviceFactor, dims::NodeNameUnion)
    ids = _ids(nodeconvert(NodeNames, dims))
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_sum, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Int32, Ref{UInt64}),
                    a.handle, ids, length(ids), h))
    return _wrap(h[])
end

function LinearAlgebra.normalize!(a::DeviceFactor; p::Int=1)
    _bn_check(ccall((:bn_factor_normalize, libbn_cuda), Int32, (UInt64, Int32), a.handle, p))
    return a
end

# The rest of src/Factors/factors_dims.jl on the device (not used by variable elimination; csrc/factor_ops.cu).
const _BN_OP = Dict{Any,Int32}((*) => 0, (/) => 1, (+) => 2, (-) => 3, max => 4, min => 5)

"join(op, a, b, :outer) for op in * / + - (factors_dims.jl:185-234,269-272), larger-first swap included."
function Base.join(op, a::DeviceFactor, b::DeviceFactor, kind::Symbol=:outer)
    kind == :outer || error("Inner joins are (still) umimplemented")
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_join, libbn_cuda), Int32, (UInt64, UInt64, Int32, Ref{UInt64}), a.handle, b.handle, _BN_OP[op], h))
    return _wrap(h[])
end
Base.:/(a::DeviceFactor, b::DeviceFactor) = join(/, a, b)
Base.:+(a::DeviceFactor, b::DeviceFactor) = join(+, a, b)
Base.:-(a::DeviceFactor, b::DeviceFactor) = join(-, a, b)

"reducedim(op, a, dims) for op in + * max min (factors_dims.jl:60-85,100-107)."
function reducedim(op, a::DeviceFactor, dims::NodeNameUnion)
    ids = _ids(nodeconvert(NodeNames, dims))
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_reduce, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Int32, Int32, Ref{UInt64}),
                    a.handle, ids, length(ids), _BN_OP[op], h))
    return _wrap(h[])
end
Base.prod(a::DeviceFactor, dims::NodeNameUnion) = reducedim(*, a, dims)
Base.maximum(a::DeviceFactor, dims::NodeNameUnion) = reducedim(max, a, dims)
Base.minimum(a::DeviceFactor, dims::NodeNameUnion) = reducedim(min, a, dims)

"normalize!(a, dims; p) (factors_dims.jl:24-43)."
function LinearAlgebra.normalize!(a::DeviceFactor, dims::NodeNameUnion; p::Int=1)
    ids = _ids(unique(nodeconvert(NodeNames, dims)))
    _bn_check(ccall((:bn_factor_normalize_dims, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Int32, Int32), a.handle, ids, length(ids), p))
    return a
end

"broadcast!(op, a, dims, values) in place (factors_dims.jl:131-165); values are Float64 scalars or vectors."
function Base.broadcast!(op, a::DeviceFactor, dims::NodeNameUnion, values)
    if isa(dims, NodeName)
        dims = [dims]; values = [values]
    end
    length(dims) == length(values) ||
        throw(ArgumentError("Number of dimensions does not match number of values to broadcast"))
    ids = _ids(dims)
    lens = Int32[length(v) for v in values]
    flat = Float64[x for v in values for x in v]
    GC.@preserve ids lens flat begin
        _bn_check(ccall((:bn_factor_broadcast, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Int32}, Int32),
                        a.handle, ids, length(ids), flat, lens, _BN_OP[op]))
    end
    return a
end

# getindex / setindex! with an Assignment (factors_access.jl:19-46): dims the factor lacks are ignored, unassigned dims
# are Colons; states are 1-based here and 0-based across the ABI.
function _assignment_args(a::DeviceFactor, asg::Assignment)
    ids = Int32[]; states = Int32[]
    for d in a.dimensions
        haskey(asg, d) || continue
        isa(asg[d], Integer) || throw(TypeError(:getindex, "Invalid state for dimension $d", Int, asg[d]))
        push!(ids, var_id(d)); push!(states, Int32(asg[d] - 1))
    end
    return ids, states
end

function Base.getindex(a::DeviceFactor, asg::Assignment)
    ids, states = _assignment_args(a, asg)
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_slice, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Ptr{Int32}, Int32, Ref{UInt64}),
                    a.handle, ids, states, length(ids), h))
    return _wrap(h[])
end

"phi[assignment] = v: one value for every cell of the slab, or an array of the slab's shape (`.= v`)."
function Base.setindex!(a::DeviceFactor, v, asg::Assignment)
    ids, states = _assignment_args(a, asg)
    vals = Float64[x for x in v]                               # column-major, like the slab
    GC.@preserve ids states vals begin
        _bn_check(ccall((:bn_factor_setindex, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Int64),
                        a.handle, ids, states, length(ids), vals, length(vals)))
    end
    return v
end

"permutedims(phi, perm) / permutedims!(phi, perm) (factors_main.jl:160-166)."
function Base.permutedims(a::DeviceFactor, perm)
    p0 = Int32[p - 1 for p in perm]
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_permutedims, libbn_cuda), Int32, (UInt64, Ptr{Int32}, Int32, Ref{UInt64}),
                    a.handle, p0, length(p0), h))
    return _wrap(h[])
end
function _rebind!(a::DeviceFactor, b::DeviceFactor)           # in-place ops rebind the table, as the reference rebinds phi.potential
    ccall((:bn_factor_free, libbn_cuda), Int32, (UInt64,), a.handle)
    a.handle, a.dimensions = b.handle, b.dimensions
    b.handle = UInt64(0)                                       # handle 0 is never issued: b's finalizer gets BN_BAD_ARG back and ignores it
    return a
end
Base.permutedims!(a::DeviceFactor, perm) = _rebind!(a, permutedims(a, perm))

"push!(phi, dim, length) (factors_main.jl:148-158): a new last dimension along which the table repeats."
function Base.push!(a::DeviceFactor, dim::NodeName, length::Int)
    dim in a.dimensions && error("Dimension $(dim) already exists")
    h = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_push, libbn_cuda), Int32, (UInt64, Int32, Int32, Ref{UInt64}), a.handle, _ids([dim])[1], length, h))
    return _rebind!(a, _wrap(h[]))
end
Base.sum!(a::DeviceFactor, dims::NodeNameUnion) = _rebind!(a, sum(a, dims))
Base.prod!(a::DeviceFactor, dims::NodeNameUnion) = _rebind!(a, prod(a, dims))
Base.maximum!(a::DeviceFactor, dims::NodeNameUnion) = _rebind!(a, maximum(a, dims))
Base.minimum!(a::DeviceFactor, dims::NodeNameUnion) = _rebind!(a, minimum(a, dims))

"sum(reduce(*, factors), h) in one kernel: the elimination step of src/Inference/exact.jl:27."
function eliminate(factors::Vector{DeviceFactor}, h::NodeName)
    hs = UInt64[f.handle for f in factors]
    ids = _ids([h]); out = Ref{UInt64}(0)
    _bn_check(ccall((:bn_factor_eliminate, libbn_cuda), Int32, (Ptr{UInt64}, Int32, Ptr{Int32}, Int32, Ref{UInt64}),
                    hs, length(hs), ids, 1, out))
    return _wrap(out[])
end

# ---------------------------------------------------------------------------------------------------------------
# statistics(bn, data) on the device          replaces src/DiscreteBayesNet/discrete_bayes_net.jl:151-172,221-231
# One pass over the table counts every node.  `data`: n_nodes x m Matrix{Int} with 1-based states (the reference's
# `Matrix{Int}(data)'`); the structure (parents in cpd.parents order, cardinalities) is the network's.  Returns the
# reference's Vector{Matrix{Int}} (r_i x q_i, first parent fastest).  For the reference's DataFrame front ends, which
# take parents = inneighbors(dag, i) and ncategories from the data, build the structure-only DeviceNet accordingly
# (bayesnets.jl_b200/learning.py does exactly that).
function cuda_statistics(bn::DiscreteBayesNet, data::AbstractMatrix{<:Integer})
    net = device_net(bn)
    lay = net.layout
    n, m = size(data)
    n == length(lay.names) || throw(DimensionMismatch("statistics' dag and data must be of the same dimension, $(length(lay.names)) ≠ $n"))
    states = Matrix{UInt8}(undef, m, n)                    # node-major on the device = one column per node
    for i in 1:n, r in 1:m
        1 <= data[i, r] <= lay.card[i] || throw(BoundsError(1:lay.card[i], data[i, r]))
        states[r, i] = UInt8(data[i, r] - 1)
    end
    counts = Vector{Int64}(undef, lay.cpt_off[end])
    GC.@preserve states counts begin
        _bn_check(ccall((:bn_statistics, libbn_cuda), Int32, (UInt64, Ptr{UInt8}, UInt64, Ptr{Int64}),
                        net.handle, states, m, counts))
    end
    return [reshape(counts[lay.cpt_off[i]+1:lay.cpt_off[i+1]], Int(lay.card[i]), :) for i in 1:n]
end

# ---------------------------------------------------------------------------------------------------------------
# logpdf(bn, df) / pdf(bn, df) on the device
#                      replaces src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-104 for a DiscreteBayesNet
# The reference adds log P(x_i | pa_i) row by row; the device counts the table once (bn_statistics' kernel) and
# returns sum_b N[b] * log(cpt[b]).  Columns are taken by name like `a[name] = df[i, name]` (:91).
function cuda_logpdf(bn::DiscreteBayesNet, df::DataFrame)
    net = device_net(bn)
    lay = net.layout
    n, m = length(lay.names), nrow(df)
    states = Matrix{UInt8}(undef, m, n)                    # node-major on the device = one column per node
    isparent = falses(n); isparent[lay.par_idx .+ 1] .= true
    impossible = false
    for (i, name) in enumerate(lay.names), r in 1:m
        x = df[r, name]
        if !(1 <= x <= lay.card[i])                        # pdf(Categorical, x) = 0 outside the support ...
            isparent[i] && throw(BoundsError(1:lay.card[i], x))   # ... but a parent's value indexes the CPT (categorical_cpd.jl:36-46)
            impossible = true; x = 1
        end
        states[r, i] = UInt8(x - 1)
    end
    impossible && return -Inf
    out = Ref{Float64}(0.0)
    GC.@preserve states begin
        _bn_check(ccall((:bn_logpdf, libbn_cuda), Int32, (UInt64, Ptr{UInt8}, UInt64, Ref{Float64}), net.handle, states, m, out))
    end
    return out[]
end
cuda_pdf(bn::DiscreteBayesNet, df::DataFrame) = exp(cuda_logpdf(bn, df))
