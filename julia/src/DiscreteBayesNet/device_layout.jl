# device_layout.jl -- flatten a DiscreteBayesNet into the compact layout libbn_cuda.so consumes.
#
# Add to BayesNets.jl as src/DiscreteBayesNet/device_layout.jl (include after discrete_bayes_net.jl).
# This is the Julia twin of bayesnets.jl_b200/device_layout.py; the two must stay field-for-field equal.
# Julia is not available in the build image, so this file is reviewed, not executed (INTEGRATION.md).
#
# Layout (include/bn_cuda.h):
#   nodes in bn.cpds order = topological (src/bayes_nets.jl:26-48), states 0-based on the device
#   card[n]            ncategories(cpd)
#   par_ptr / par_idx  CSR of parents in cpd.parents order (first parent fastest in the CPT,
#                      src/CPDs/categorical_cpd.jl:36-46), 0-based node indices
#   cpt_off[n+1], cpt  vcat(d.p for d in cpd.distributions) per node (src/Factors/factors_main.jl:62-67)

struct DeviceLayout
    names::NodeNames
    card::Vector{Int32}
    par_ptr::Vector{Int32}
    par_idx::Vector{Int32}
    cpt_off::Vector{Int64}
    cpt::Vector{Float64}
end

function DeviceLayout(bn::DiscreteBayesNet)
    n = length(bn.cpds)
    card = Vector{Int32}(undef, n)
    par_ptr = Int32[0]
    par_idx = Int32[]
    cpt_off = Int64[0]
    cpt = Float64[]
    for (i, cpd) in enumerate(bn.cpds)
        card[i] = ncategories(cpd)
        for (p, pc) in zip(cpd.parents, cpd.parental_ncategories)
            j = bn.name_to_index[p]
            j < i || throw(ArgumentError("BayesNet is not topologically ordered"))
            ncategories(bn.cpds[j]) == pc ||
                throw(DimensionMismatch("CPD $(name(cpd)): parent $p has $(ncategories(bn.cpds[j])) categories, CPD says $pc"))
            push!(par_idx, Int32(j - 1))
        end
        push!(par_ptr, Int32(length(par_idx)))
        for d in cpd.distributions
            append!(cpt, probs(d))
        end
        push!(cpt_off, Int64(length(cpt)))
    end
    return DeviceLayout(names(bn), card, par_ptr, par_idx, cpt_off, cpt)
end

"""
    evidence_vector(lay, bn, evidence) -> Vector{Int8}

`-1` = free, otherwise the clamped 0-based state.  Names that are not nodes are ignored, exactly as
`haskey(evidence, s) for s in nodes` does (src/Inference/likelihood.jl:25).
"""
function evidence_vector(lay::DeviceLayout, bn::DiscreteBayesNet, evidence::Assignment)
    ev = fill(Int8(-1), length(lay.names))
    for (k, v) in evidence
        haskey(bn.name_to_index, k) || continue
        i = bn.name_to_index[k]
        v isa Integer || throw(TypeError(:evidence_vector, "state of $k", Integer, v))
        1 <= v <= lay.card[i] || throw(BoundsError(1:lay.card[i], v))
        v - 1 <= 127 || throw(BoundsError(1:128, v))   # int8 on the ABI (BN_MAX_EVIDENCE_STATE): never wrap to "free"
        ev[i] = Int8(v - 1)
    end
    return ev
end

query_indices(bn::DiscreteBayesNet, query::NodeNames) = Int32[bn.name_to_index[q] - 1 for q in query]

"""
A DiscreteBayesNet resident in HBM (opaque `bn_net_t`).  One upload per (network contents, device or communicator)
-- see `device_net` in CUDABackend.jl; `bn_net_destroy` runs from the finalizer.  With a `Communicator` the network is
replicated on every GPU of the communicator (`bn_net_create_comm`) and `device` is its first GPU.
"""
mutable struct DeviceNet
    handle::UInt64
    layout::DeviceLayout
    device::Int32
    comm::Any                      # Communicator or nothing (kept alive as long as the net)
    function DeviceNet(lay::DeviceLayout; device::Integer=0)
        maximum(lay.card) <= 255 || throw(ArgumentError("a node has more than 255 states (uint8 states on the device)"))
        h = Ref{UInt64}(0)
        par_idx = isempty(lay.par_idx) ? Int32[0] : lay.par_idx
        GC.@preserve lay par_idx begin
            _bn_check(ccall((:bn_net_create, libbn_cuda), Int32,
                            (Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{UInt64}),
                            length(lay.names), lay.card, lay.par_ptr, par_idx, lay.cpt_off, lay.cpt, device, h))
        end
        net = new(h[], lay, Int32(device), nothing)
        finalizer(x -> ccall((:bn_net_destroy, libbn_cuda), Int32, (UInt64,), x.handle), net)
        return net
    end
    function DeviceNet(lay::DeviceLayout, comm)
        maximum(lay.card) <= 255 || throw(ArgumentError("a node has more than 255 states (uint8 states on the device)"))
        h = Ref{UInt64}(0)
        par_idx = isempty(lay.par_idx) ? Int32[0] : lay.par_idx
        GC.@preserve lay par_idx begin
            _bn_check(ccall((:bn_net_create_comm, libbn_cuda), Int32,
                            (Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, UInt64, Ref{UInt64}),
                            length(lay.names), lay.card, lay.par_ptr, par_idx, lay.cpt_off, lay.cpt, comm.handle, h))
        end
        net = new(h[], lay, comm.devices[1], comm)
        finalizer(x -> ccall((:bn_net_destroy, libbn_cuda), Int32, (UInt64,), x.handle), net)
        return net
    end
end
DeviceNet(bn::DiscreteBayesNet; device::Integer=0) = DeviceNet(DeviceLayout(bn); device=device)
