/*
 * bn_cuda.h -- C ABI of libbn_cuda.so: the B200 (sm_100a) implementation of BayesNets.jl's
 * data-parallel discrete-inference hot path.
 *
 * The reference (sisl/BayesNets.jl v3.5.2, pure Julia) has no FFI of its own: its extension points are
 * multiple dispatch on InferenceMethod (src/ProbabilisticGraphicalModels/inference.jl:4,48-49) and on
 * Base.:* / Base.sum / normalize! for Factor (src/Factors/factors_dims.jl:100,269,45).  Each entry point
 * below names the reference method whose inner loop it replaces; INTEGRATION.md shows the `ccall`
 * a maintainer adds on the Julia side and the ctypes stub the Python mirror uses.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; every function returns bn_status
 *     (0 = ok, <0 = error; bn_last_error() gives a thread-local message)
 *   - networks are passed flat, nodes in topological order (src/bayes_nets.jl:26-48), states 0-based:
 *       card[n], par_ptr[n+1]/par_idx[] (CSR, parents in cpd.parents order), cpt_off[n+1], cpt[]
 *       CPT entry P(x_i = s | pa) = cpt[cpt_off[i] + s + card[i] * sum_k state(parent_k) * stride_k],
 *       stride_0 = 1, stride_k = prod_{j<k} card(parent_j)      (src/CPDs/categorical_cpd.jl:36-46,
 *       src/Factors/factors_main.jl:62-67: vcat(d.p for d in distributions))
 *   - evidence[n] is int8: negative = free, otherwise the clamped 0-based state; sampled tables are uint8 with a
 *     leading 0xFF = "no row".  Hence at most BN_MAX_CARD = 255 states per node (bn_net_create rejects more), and only
 *     states 0..BN_MAX_EVIDENCE_STATE = 127 of a node can be clamped as evidence (the host mirrors raise a BoundsError
 *     for a larger evidence state instead of letting it wrap to "free")
 *   - factors are column-major Float64 exactly as Factor.potential (src/Factors/factors_main.jl:20);
 *     variables are int32 ids (the Julia glue maps Symbol <-> id)
 *   - host pointers unless the name ends in _dev; the caller owns every host buffer, the library owns
 *     device memory behind opaque handles; no host pointer is retained after return
 *   - calls are synchronous unless they end in _dev (those enqueue on the stream set by
 *     bn_set_stream and return immediately)
 *   - threading: any number of host threads / streams may use the same net or factor handles concurrently for
 *     READ-style calls (infer, sample, statistics, products): per-call staging comes from the stream-ordered pool
 *     and the per-net kernel cache is locked.  Calls that write a handle in place (bn_factor_normalize, _setindex,
 *     _broadcast, _relabel) and the destroy / free calls need external ordering against other users of THAT handle.
 *     A comm-bearing net (bn_net_create_comm) is driven by one host thread at a time (NCCL's rule for a communicator)
 *   - random streams: Philox4x32-10, key = seed, counter = (index_lo, index_hi, draw/4, stream id);
 *     particle / chain indices are GLOBAL (first_* + i), so sharding a job over ranks reproduces the
 *     single-GPU stream bit for bit
 */
#ifndef BN_CUDA_H
#define BN_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_MAX_CARD 255
#define BN_MAX_EVIDENCE_STATE 127
typedef uint64_t bn_net_t;    /* device-resident flattened DiscreteBayesNet */
typedef uint64_t bn_factor_t; /* device-resident Factor (dims + column-major fp64 potential) */

typedef enum {
  BN_OK = 0,
  BN_BAD_ARG = -1,        /* ArgumentError */
  BN_DIM_MISMATCH = -2,   /* DimensionMismatch (src/Factors/factors_dims.jl:197-199) */
  BN_NOT_IN_FACTOR = -3,  /* ArgumentError "... is not a valid dimension" (src/Factors/errors.jl:12) */
  BN_CUDA_ERR = -4,
  BN_ZERO_WEIGHT = -5,    /* every particle had weight 0 / no initial Gibbs sample (src/gibbs.jl:332-334) */
  BN_OOM = -6,
  BN_UNSUPPORTED = -7,
  BN_BOUNDS = -8,         /* BoundsError (src/Factors/factors_access.jl:59-61) */
  BN_NCCL_ERR = -9        /* libnccl.so.2 missing or a NCCL call failed (multi-GPU entry points only) */
} bn_status;

enum { BN_SAMPLE_ANCESTRAL = 0, BN_SAMPLE_REJECTION = 1, BN_SAMPLE_WEIGHTED = 2 };
enum { BN_GIBBS_NODEWISE = 0, BN_GIBBS_FULL = 1 };

/* ---- library ------------------------------------------------------------------------------ */
const char* bn_last_error(void);
int32_t bn_version(void);
int32_t bn_device_count(int32_t* out_n);
/* Stream (a cudaStream_t cast to void*) used by subsequent calls from this host thread; NULL = default. */
int32_t bn_set_stream(void* cuda_stream);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t bn_launch_count(void);

/* Wait for everything this host thread has enqueued on its current stream (the _dev entry points are asynchronous). */
int32_t bn_synchronize(void);
/* Device memory of freed factors / per-call temporaries is parked inside the library, keyed by (device, stream, size),
 * and handed to the next allocation of that size on that stream without a driver call (at most BN_POOL_CACHE_MB MiB,
 * default 16384; given back automatically when the driver reports out-of-memory).  bn_pool_trim returns all parked
 * blocks to the driver now (synchronises the device); out_bytes (may be NULL) receives how many bytes were parked. */
int32_t bn_pool_trim(uint64_t* out_bytes);

/* ---- plain device buffers for callers without a CUDA array library (the Julia glue, the Python mirror): the *_dev
 *      entry points take their pointers.  Stream-ordered on the calling thread's current stream; bn_dev_download
 *      synchronises that stream before returning, bn_dev_upload returns once `src` may be reused. ---------------- */
int32_t bn_dev_alloc(int32_t device, uint64_t bytes, void** out_ptr);
int32_t bn_dev_free(int32_t device, void* ptr);
int32_t bn_dev_memset(void* ptr, int32_t value, uint64_t bytes);
int32_t bn_dev_upload(void* dst_dev, const void* src, uint64_t bytes);
int32_t bn_dev_download(void* dst, const void* src_dev, uint64_t bytes);

/* ---- multi-GPU: infer(im, bn, query; evidence) is ONE call for the whole machine
 *      (src/ProbabilisticGraphicalModels/inference.jl:48-49), so the fan-out over GPUs and the one collective the
 *      sampling paths need live behind the C ABI.  NCCL (libnccl.so.2, dlopen'ed at first use; BN_NCCL_PATH overrides)
 *      is called directly from C: no torch.distributed / CUDA.jl / MPI in the data path. ---------------------------
 * A communicator spans `world` GPUs.  Two ways to make one:
 *   bn_comm_init_all  : ONE process drives n_devices GPUs (ncclCommInitAll, one stream per device) -- what a Julia
 *                       session needs; world = n_devices, all of them local.
 *   bn_comm_init_rank : one process PER GPU (torchrun / MPI): rank 0 calls bn_comm_unique_id, ships the 128 bytes to
 *                       the other ranks out of band (file, TCP store, MPI_Bcast), then every rank calls
 *                       bn_comm_init_rank(id, rank, world, device) -- collective, returns when all ranks joined.
 * bn_net_create_comm uploads the network to every LOCAL device of the communicator.  On such a net
 *   bn_lw_infer[_dev], bn_gibbs_infer[_dev]: first_* / n_* describe the GLOBAL job; GPU g of `world` takes the
 *       contiguous global index range bn_comm_shard(first, n, g, world) (Philox counters are global indices, so any
 *       world size consumes the same streams), every local GPU is launched, and ONE ncclAllReduce(sum, fp64) of the
 *       (cells + 2) doubles leaves the whole job's counts on every GPU before the call returns them.
 *       _dev variants: out_dev lives on the FIRST local device and is enqueued on the caller's current stream.
 *       bn_gibbs_infer: state_inout / out_chain_counts rows belong to chains; each process fills the rows of the chains
 *       its local GPUs ran and leaves the others untouched.
 *   every other entry point (factors, exact inference, statistics, tables ...) runs on the first local device.  */
typedef uint64_t bn_comm_t;
#define BN_COMM_ID_BYTES 128
enum { BN_DTYPE_F64 = 0, BN_DTYPE_I64 = 1 };
enum { BN_REDUCE_SUM = 0, BN_REDUCE_MAX = 1 };
int32_t bn_comm_init_all(const int32_t* devices, int32_t n_devices, bn_comm_t* out_comm);
int32_t bn_comm_unique_id(uint8_t* out_id /* BN_COMM_ID_BYTES */);
int32_t bn_comm_init_rank(const uint8_t* id, int32_t rank, int32_t world, int32_t device, bn_comm_t* out_comm);
/* world = GPUs in the communicator, first_rank = global rank of this process's first local GPU, n_local = local GPUs */
int32_t bn_comm_info(bn_comm_t comm, int32_t* out_world, int32_t* out_first_rank, int32_t* out_n_local);
int32_t bn_comm_destroy(bn_comm_t comm);
/* The sharding rule, host-only (no GPU, no NCCL): rank g of `world` gets [first + n*g/world, first + n*(g+1)/world). */
int32_t bn_comm_shard(uint64_t first, uint64_t n, int32_t rank, int32_t world, uint64_t* out_first, uint64_t* out_n);
/* Element-wise reduction over the PROCESSES of the communicator of one host buffer per process (in place; a
 * single-process communicator returns the buffer unchanged): timing maxima, user-side bookkeeping.  Synchronous. */
int32_t bn_comm_allreduce(bn_comm_t comm, void* host_inout, int64_t count, int32_t dtype, int32_t op);
/* The same on a device buffer of the first local GPU, enqueued on the caller's current stream (one local GPU per
 * process only): e.g. the int64 sufficient statistics of a row-sharded table (bn_statistics_dev). */
int32_t bn_comm_allreduce_dev(bn_comm_t comm, void* dev_inout, int64_t count, int32_t dtype, int32_t op);
/* Every GPU of the communicator has finished what was enqueued before the call (an all-reduce of one double + sync). */
int32_t bn_comm_barrier(bn_comm_t comm);

/* ---- network: replaces nothing; it is the device_layout.jl upload (DiscreteBayesNet -> HBM) ---- */
int32_t bn_net_create(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                      const int64_t* cpt_off, const double* cpt, int32_t device, bn_net_t* out_net);
/* The same network replicated on every local device of `comm` (see the multi-GPU section above).  The communicator
 * must outlive the net. */
int32_t bn_net_create_comm(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                           const int64_t* cpt_off, const double* cpt, bn_comm_t comm, bn_net_t* out_net);
int32_t bn_net_destroy(bn_net_t net);

/* ---- likelihood weighting: infer(::LikelihoodWeightingInference, ::InferenceState)
 *      src/Inference/likelihood.jl:16-58 (the :34-54 particle loop) ------------------------------
 * out_counts: prod(card[query]) doubles, column-major over the query cards (first query fastest),
 *             UNNORMALISED weighted counts  phi.potential[q_ind...] += w  (:53)
 * out_stats : [sum w, sum w^2]  (N_eff = stats[0]^2 / stats[1])
 * Particles first_particle .. first_particle + n_particles - 1.  Normalisation (:56) is the caller's
 * (bn_factor_normalize or one division) so that shards can be summed first. */
int32_t bn_lw_infer(bn_net_t net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                    uint64_t first_particle, uint64_t n_particles, uint64_t seed,
                    double* out_counts, double* out_stats);
/* Same, asynchronous, result left in device memory: out_dev[0..ncell) counts, [ncell] sum w,
 * [ncell+1] sum w^2.  out_dev is ZEROED by the call.  This is what the multi-GPU path all-reduces. */
int32_t bn_lw_infer_dev(bn_net_t net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                        uint64_t first_particle, uint64_t n_particles, uint64_t seed, double* out_dev);

/* LW kernel selection for subsequent calls from this host thread.
 * mode : 0 = auto: the network-specialised kernel is used as soon as its cubin is at hand -- in the process, or in the
 *            on-disk cache a previous process left; otherwise the generic interpreter kernel serves the call and, once
 *            the cumulative work requested for the same (structure, evidence mask, query) reaches 5e7 particles/rows
 *            (2e8 node updates for Gibbs), NVRTC compiles the kernel on a BACKGROUND host thread (~10 ms per node);
 *            later calls switch over when it is ready.  No call ever waits for a compile;
 *        1 = generic interpreter only, 2 = specialised required (compiled on first use, blocking; error if unavailable).
 * prune: != 0 lets the specialised kernel skip "barren" nodes (not an ancestor of a query or evidence node);
 *        draw indices are static, so the counts are bit-identical either way.  Default: mode 0, prune 1. */
int32_t bn_set_lw_mode(int32_t mode, int32_t prune);
/* Same selection for bn_gibbs_infer and bn_gibbs_sample: 0 = auto, 1 = generic interpreter kernel, 2 = specialised
 * (bn_gibbs_infer: required; bn_gibbs_sample: used whenever the network qualifies -- <= 8 states per node -- else the
 * interpreter serves the call).  Chains, counts and tables are identical either way.
 * Both kinds of kernel draw the nodes with small Markov blankets from integer threshold tables built once per net
 * (csrc/gibbs_tables.cu): the same decision as the fp64 cumulative scan for every 32-bit draw, so nothing observable
 * changes; BN_GIBBS_TABLE_ROWS=0 turns the tables off. */
int32_t bn_set_gibbs_mode(int32_t mode);
/* Number of NVRTC compiles since load (kernels are cached process-wide by generated source). */
uint64_t bn_lw_spec_compile_count(void);
/* Kernel-cache counters since load: out4 = [NVRTC compiles, of which started on a background thread ("auto" modes),
 * cubins found in the on-disk cache (BN_CUBIN_CACHE, default ~/.cache/bn_cuda; "off" disables), kernels loaded]. */
int32_t bn_spec_cache_stats(uint64_t* out4);
/* Host-only (no GPU): emit the CUDA source the specialised LW kernel would be built from for this flat network
 * and, when cubin_size != NULL, compile it with NVRTC for sm_100a.  out_src may be NULL to query *out_len. */
int32_t bn_lw_spec_source(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                          const int64_t* cpt_off, const double* cpt, const int8_t* evidence, const int32_t* query,
                          int32_t n_query, int32_t prune, char* out_src, int64_t cap, int64_t* out_len,
                          int64_t* cubin_size, int32_t* n_live);

int32_t bn_gibbs_spec_source(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                             const int64_t* cpt_off, const double* cpt, const int8_t* evidence, const int32_t* query,
                             int32_t n_query, int32_t full, int32_t threads, char* out_src, int64_t cap, int64_t* out_len,
                             int64_t* cubin_size);

/* ---- table samplers: rand(bn, n) / rand(bn, n, evidence) / get_weighted_dataframe
 *      src/sampling.jl:24-41,52-57 (ancestral), :72-93 (rejection, <= max_attempts tries),
 *      :102-115,153-174 (likelihood-weighted rows) -------------------------------------------------
 * out_states: n_nodes x n uint8, node-major (out_states[i*n + row]) = one DataFrame column per node
 * out_w     : n doubles (weighted kind; may be NULL otherwise)
 * out_ok    : n bytes, 0 where rejection gave up (the reference returns an empty assignment, :87) */
int32_t bn_sample(bn_net_t net, const int8_t* evidence, int32_t kind, int32_t max_attempts,
                  uint64_t first_row, uint64_t n_rows, uint64_t seed,
                  uint8_t* out_states, double* out_w, uint8_t* out_ok);
int32_t bn_sample_dev(bn_net_t net, const int8_t* evidence, int32_t kind, int32_t max_attempts,
                      uint64_t first_row, uint64_t n_rows, uint64_t seed,
                      uint8_t* out_states_dev, double* out_w_dev, uint8_t* out_ok_dev);

/* ---- Gibbs inference: infer(::GibbsSamplingNodewise / ::GibbsSamplingFull, ::InferenceState)
 *      src/Inference/gibbs.jl:66-145 / :161-234, n_chains independent chains ---------------------
 * state_inout: optional n_chains x n_nodes uint8 (chain-major).  init_from_state != 0 resumes the
 *              chains from it (im.state, :83-88); final states are written back when non-NULL.
 * out_counts : prod(card[query]) doubles summed over chains (unnormalised)
 * out_chain_counts: optional n_chains x ncell (chain-major) for between-chain variance */
int32_t bn_gibbs_infer(bn_net_t net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                       uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in,
                       int64_t thin, int32_t mode, uint64_t seed, uint8_t* state_inout,
                       int32_t init_from_state, double* out_counts, double* out_chain_counts);
int32_t bn_gibbs_infer_dev(bn_net_t net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                           uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in,
                           int64_t thin, int32_t mode, uint64_t seed, double* out_dev);

/* ---- gibbs_sample(bn, nsamples, burn_in; thinning, consistent_with, variable_order, initial_sample)
 *      src/gibbs.jl:289-374 (main loop :183-245), n_chains chains each returning nsamples rows ----
 * variable_order: NULL = fresh random permutation every sweep (:222-234), else n_free positions into
 *                 the list of non-evidence nodes (topological order)
 * init_state    : NULL = 50 likelihood-weighted particles, one resampled (:330-335), else
 *                 n_chains x n_nodes uint8
 * out_states    : n_nodes x (n_chains*nsamples) uint8, node-major; row = chain*nsamples + s */
int32_t bn_gibbs_sample(bn_net_t net, const int8_t* evidence, uint64_t first_chain, uint64_t n_chains,
                        int64_t nsamples, int64_t burn_in, int64_t thinning,
                        const int32_t* variable_order, uint64_t seed, const uint8_t* init_state,
                        uint8_t* out_states);

/* ---- factors: Factor(dims, potential) on the device (src/Factors/factors_main.jl:18-36) -------- */
int32_t bn_factor_upload(int32_t ndims, const int32_t* dims, const int32_t* card, const double* data,
                         int32_t device, bn_factor_t* out);
int32_t bn_factor_alloc(int32_t ndims, const int32_t* dims, const int32_t* card, int32_t device,
                        bn_factor_t* out);               /* uninitialised potential */
int32_t bn_factor_from_cpd(bn_net_t net, int32_t node, bn_factor_t* out); /* convert(Factor, cpd) :62-68,
                                                         dims = [node, parents...], ids = node indices */
int32_t bn_factor_relabel(bn_factor_t f, const int32_t* new_dims); /* rename the dims in place (ndims ids, unique) */
int32_t bn_factor_ndims(bn_factor_t f, int32_t* out_ndims);
int32_t bn_factor_shape(bn_factor_t f, int32_t* out_dims, int32_t* out_card); /* ndims entries each */
int32_t bn_factor_length(bn_factor_t f, int64_t* out_len);
int32_t bn_factor_download(bn_factor_t f, double* out);
int32_t bn_factor_dev_ptr(bn_factor_t f, void** out_ptr);
int32_t bn_factor_free(bn_factor_t f);

/* phi1 * phi2 = join(*, phi1, phi2, :outer): src/Factors/factors_dims.jl:185-234,269.
 * Larger operand first (:189-191), dims = union (:203), BN_DIM_MISMATCH on unequal common dims. */
int32_t bn_factor_product(bn_factor_t f1, bn_factor_t f2, bn_factor_t* out);
/* sum(phi, dims) = reducedim(+, phi, dims): src/Factors/factors_dims.jl:60-85,100. */
int32_t bn_factor_sum(bn_factor_t f, const int32_t* sum_dims, int32_t n_sum, bn_factor_t* out);
/* phi[assignment] (getindex): src/Factors/factors_access.jl:19-35,48-72; dims not in phi are ignored. */
int32_t bn_factor_slice(bn_factor_t f, const int32_t* dims, const int32_t* states, int32_t n_assign,
                        bn_factor_t* out);
/* phi[assignment] = v (setindex!): src/Factors/factors_access.jl:39-46 over _translate_index (:48-72); dims not in phi
 * are ignored, unassigned dims are Colons.  n_values == 1: every cell of the slab gets values[0]; n_values == the slab's
 * size: an array over the free dims, column-major (`.= v`); otherwise BN_DIM_MISMATCH.  In place.  values: host. */
int32_t bn_factor_setindex(bn_factor_t f, const int32_t* dims, const int32_t* states, int32_t n_assign,
                           const double* values, int64_t n_values);
/* permutedims(phi, perm): src/Factors/factors_main.jl:160-166; out.dims[i] = phi.dims[perm[i]], perm 0-based. */
int32_t bn_factor_permutedims(bn_factor_t f, const int32_t* perm, int32_t n, bn_factor_t* out);
/* push!(phi, dim, length): src/Factors/factors_main.jl:148-158 (duplicate, auxiliary.jl:50-65): a new last dimension
 * along which the table is repeated; BN_BAD_ARG if phi already has `dim`. */
int32_t bn_factor_push(bn_factor_t f, int32_t dim, int32_t length, bn_factor_t* out);
/* normalize!(phi; p): src/Factors/factors_dims.jl:45-57 (p in {1,2}; p=2 divides by sum(abs2)). */
int32_t bn_factor_normalize(bn_factor_t f, int32_t p);
/* The two halves of normalize!, for a factor that is split over several GPUs (bayesnets.jl_b200/sharded_factors.py):
 * bn_factor_norm returns sum(abs, phi) (p = 1) or sum(abs2, phi) (p = 2) of the local part, the caller adds the parts
 * up (one scalar all-reduce) and bn_factor_div_scalar applies phi ./= total, the division normalize! performs. */
int32_t bn_factor_norm(bn_factor_t f, int32_t p, double* out_total);
int32_t bn_factor_div_scalar(bn_factor_t f, double total);
/* sum(reduce(*, factors), var) in ONE pass (the elimination step of src/Inference/exact.jl:27):
 * multiplies left to right like the reference's fold, never materialising the product.
 * n_sum = 0 gives the plain n-ary product. */
int32_t bn_factor_eliminate(const bn_factor_t* factors, int32_t n_factors, const int32_t* sum_dims,
                            int32_t n_sum, bn_factor_t* out);

/* ---- the rest of the Factor algebra (src/Factors/factors_dims.jl), not used by variable elimination -------------
 * bn_factor_join: join(op, phi1, phi2, :outer) for op in * / + - (:185-234,269-272).  As in the reference the LARGER
 *   factor becomes phi1 before op is applied (:189-191), so / and - are operand-order dependent exactly like there.
 * bn_factor_reduce: reducedim(op, phi, dims) for op in + * max min (:60-85,100-107), dims dropped; sequential
 *   accumulation over the reduced dims in column-major order; max / min propagate NaN like Base.max / Base.min.
 * bn_factor_normalize_dims: normalize!(phi, dims; p) (:24-43): every slice over `dims` is divided by its sum(abs)
 *   (p = 1) or sum(abs2) (p = 2 -- no square root, as in the reference); empty dims = no-op.
 * bn_factor_broadcast: broadcast!(op, phi, dims, values) (:131-165) in place.  values = the vectors back to back,
 *   value_len[j] = 1 (a scalar) or the length of dims[j] (else BN_DIM_MISMATCH); several dims apply left to right. */
enum { BN_OP_MUL = 0, BN_OP_DIV = 1, BN_OP_ADD = 2, BN_OP_SUB = 3, BN_OP_MAX = 4, BN_OP_MIN = 5 };
int32_t bn_factor_join(bn_factor_t f1, bn_factor_t f2, int32_t op, bn_factor_t* out);
int32_t bn_factor_reduce(bn_factor_t f, const int32_t* dims, int32_t n_dims, int32_t op, bn_factor_t* out);
int32_t bn_factor_normalize_dims(bn_factor_t f, const int32_t* dims, int32_t n_dims, int32_t p);
int32_t bn_factor_broadcast(bn_factor_t f, const int32_t* dims, int32_t n_dims, const double* values,
                            const int32_t* value_len, int32_t op);

/* ---- infer(::ExactInference, ::InferenceState): src/Inference/exact.jl:6-33 ---------------------
 * Variable elimination entirely on the device.  out_dims/out_card: n_query entries (the result's dim
 * order is join-order dependent, as in the reference -- compare by name); out: prod(card) doubles.
 * fused != 0 uses bn_factor_eliminate per hidden variable, 0 the reference's pairwise product+sum. */
int32_t bn_exact_infer(bn_net_t net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                       int32_t fused, int32_t* out_dims, int32_t* out_card, double* out);

/* ---- statistics(parent_list, ncategories, data): src/DiscreteBayesNet/discrete_bayes_net.jl:151-172,221-231 ----
 * Sufficient statistics of a data table under the structure of `net` (its CPT values are not read), all nodes in
 * one pass.  states: uint8, node-major (states[i * pitch + row]), 0-based, i.e. exactly what bn_sample[_dev] writes;
 * rows whose first byte is 0xFF (rejection sampler gave up) are skipped.  out_counts: cpt_off[n] int64 in the flat
 * CPT layout, counts[cpt_off[i] + k + card_i * j] = N_i[k, j] = the reference's r x q matrix, column-major
 * (first parent fastest, :131-136).  bn_statistics validates states < card (BN_BOUNDS); the _dev variant trusts
 * the table (out-of-range states are dropped or miscounted, never written out of bounds), zeroes out_counts_dev
 * and enqueues on the current stream.  pitch >= n_rows; TMA staging needs pitch % 16 == 0 and a 16-byte aligned
 * table (otherwise plain loads). */
int32_t bn_statistics(bn_net_t net, const uint8_t* states, uint64_t n_rows, int64_t* out_counts);
int32_t bn_statistics_dev(bn_net_t net, const uint8_t* states_dev, uint64_t n_rows, uint64_t pitch,
                          int64_t* out_counts_dev);

/* ---- logpdf(bn, df::DataFrame) / pdf(bn, df): src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-104
 *      over logpdf(bn, assignment), src/bayes_nets.jl:322-328 ---------------------------------------------------
 * Log-likelihood of a data table under `net`: sum over rows and nodes of log P(x_i | pa_i).  Same table layout and
 * 0xFF-row rule as bn_statistics.  Computed as sum_b N[b] * log(cpt[b]) over the bins with N[b] > 0, N = the table's
 * sufficient statistics (one pass of the counting kernel, then one small deterministic reduction); a row the net
 * gives probability 0 makes the result -Inf, as in the reference.  Zero rows -> 0.0.  bn_logpdf validates
 * states < card (BN_BOUNDS); the _dev variant trusts the table, writes one double to out_logl_dev and enqueues on
 * the current stream. */
int32_t bn_logpdf(bn_net_t net, const uint8_t* states, uint64_t n_rows, double* out_logl);
int32_t bn_logpdf_dev(bn_net_t net, const uint8_t* states_dev, uint64_t n_rows, uint64_t pitch, double* out_logl_dev);
/* The second half alone: log-likelihood from sufficient statistics already on the device (bn_statistics_dev's
 * layout), e.g. counts all-reduced over the ranks of a row-sharded table -- every rank then gets the same bits as one
 * GPU counting the whole table.  Enqueues on the current stream. */
int32_t bn_logpdf_counts_dev(bn_net_t net, const int64_t* counts_dev, double* out_logl_dev);

/* ---- infer(::LoopyBelief, ::InferenceState): src/Inference/loopy_belief.jl:24-218 -------------------------
 * Loopy belief propagation, batched: instance b has evidence row evidence[b * n_nodes .. ] (NULL = no evidence
 * anywhere) and ONE query node query[b] (:26); all instances run concurrently, one warp each, messages
 * L1 / L2 resident.  nsamples = iteration cap, tol / iters_for_convergence = the reference's early stop (:196; tol < 0 never
 * stops).  out[b * out_stride + s] = P(query[b] = s | evidence_b), s < card(query[b]) (rest zero-filled);
 * out_iters[b] (may be NULL) = iterations run.  Reproduces the reference's message aliasing (csrc/loopy.cu header),
 * i.e. the same numbers as the reference, which on nets whose roots have several children are not textbook LBP.
 * bn_loopy_infer validates like InferenceState (query in the net and not in the evidence -> BN_BAD_ARG; state out
 * of range -> BN_BOUNDS); the _dev variant takes device pointers, trusts them (out-of-range evidence = free,
 * bad query -> NaN row) and enqueues on the current stream. */
int32_t bn_loopy_infer(bn_net_t net, const int8_t* evidence, const int32_t* query, uint64_t n_inst, int32_t nsamples,
                       double tol, int32_t iters_for_convergence, int32_t out_stride, double* out, int32_t* out_iters);
int32_t bn_loopy_infer_dev(bn_net_t net, const int8_t* evidence_dev, const int32_t* query_dev, uint64_t n_inst,
                           int32_t nsamples, double tol, int32_t iters_for_convergence, int32_t out_stride,
                           double* out_dev, int32_t* out_iters_dev);

#ifdef __cplusplus
}
#endif
#endif /* BN_CUDA_H */
