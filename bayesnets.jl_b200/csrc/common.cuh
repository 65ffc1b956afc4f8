// common.cuh -- shared host/device helpers for libbn_cuda.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bn_cuda.h"

namespace bn {

// ---------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern thread_local cudaStream_t tl_stream;
extern std::atomic<uint64_t> g_launches;

#define BN_CUDA_TRY(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      bn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (_e == cudaErrorMemoryAllocation) ? BN_OOM : BN_CUDA_ERR;                      \
    }                                                                                       \
  } while (0)

#define BN_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      bn::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

#define BN_TRY(expr)              \
  do {                            \
    int32_t _rc = (expr);         \
    if (_rc != BN_OK) return _rc; \
  } while (0)

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------- device properties
struct DevProps {
  int sm_count = 148;
  int max_smem_optin = 232448;  // 227 KB
  int smem_per_sm = 233472;     // 228 KB
};
const DevProps& dev_props(int device);

// ---------------------------------------------------------------------------- device buffers
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
  }
  cudaError_t upload(const T* host, size_t count, cudaStream_t s) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess) return e;
    if (count == 0) return cudaSuccess;
    return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

// Stream-ordered pool allocation (cudaMallocAsync) for factor potentials: variable elimination allocates and
// frees 2^20..2^27-entry intermediates back to back; the default mempool keeps them cached (release
// threshold = max), so after warm-up an allocation is a pointer bump and frees never synchronise the device.
void ensure_pool(int device);
// Block cache in front of cudaMallocAsync / cudaFreeAsync (core.cu).  A block released on stream s is parked under
// (device, s, bytes) and handed to the next allocation of exactly that size on that stream without any driver call:
// stream order makes the reuse safe, and back-to-back factor ops lose no time to the stream-ordered free
// (measured: 1.7 us per op on 35 us kernels, tools/probe/alloc_gap.cu).  Bounded by BN_POOL_CACHE_MB (default 16384);
// beyond it, and when the driver reports out-of-memory, blocks go back to the driver pool.
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t s, int* device_out);
void pool_free(void* p, size_t bytes, cudaStream_t s, int device);
size_t pool_trim();  // returns the bytes that were parked
template <typename T>
struct PoolBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owns = true;  // false: a read-only view of memory owned elsewhere (borrow)
  cudaStream_t s = nullptr;  // the stream the allocation was made on; the free is ordered on the same stream, so a
                             // finalizer running after bn_set_stream() switched streams cannot free under a kernel
  int device = 0;
  size_t bytes = 0;
  PoolBuf() = default;
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
  ~PoolBuf() { release(); }
  void release() {
    if (p && owns) pool_free(p, bytes, s, device);
    p = nullptr; n = 0; owns = true;
  }
  void borrow(T* ptr, size_t count) {
    release();
    p = ptr; n = count; owns = false;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    s = tl_stream;
    bytes = (count ? count : 1) * sizeof(T);
    const cudaError_t e = pool_alloc((void**)&p, bytes, s, &device);
    if (e != cudaSuccess) { p = nullptr; n = 0; }  // nothing to release later
    return e;
  }
  cudaError_t upload(const T* host, size_t count, cudaStream_t st) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess) return e;
    if (count == 0) return cudaSuccess;
    return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
  }
};
// Per-call temporaries of the host-buffer entry points (bn_sample, bn_gibbs_*, bn_statistics): pool allocations, so a
// call costs no cudaMalloc / cudaFree (each ~0.1-1 ms and a device synchronisation) after the first one.
template <typename T>
using TmpBuf = PoolBuf<T>;

// ---------------------------------------------------------------------------- objects behind handles
struct SpecKernel;  // specialize.cu: one NVRTC-compiled LW kernel per (evidence mask, query)
struct Net;
void spec_cache_clear(Net* net);
struct LoopyPlan;  // loopy.cu: message layout + task list of the LoopyBelief kernel (depends on the structure only)
void loopy_plan_free(LoopyPlan* p);

struct Comm;  // comm.cu: NCCL communicator(s) + one stream per local device

struct StatsPlan;  // stats.cu: node groups, warp counts and node records of the row-lane statistics kernel
void stats_plan_free(StatsPlan* p);

// gibbs_tables.cu: Markov-blanket conditionals as integer threshold tables (nodes with small blankets)
struct GibbsTables {
  uint32_t row_limit = 0;
  std::vector<int64_t> tab_off;                 // per node: word offset of its table, -1 = keeps the fp64 conditional
  std::vector<std::vector<int>> members;        // blanket members (ascending node index) ...
  std::vector<std::vector<uint32_t>> strides;   // ... and their row strides
  uint64_t words = 0;
  DevBuf<uint32_t> tab;
  bool has_never = false;  // some tabled row needs the unrepresentable threshold 2^32: the caller keeps the fp64 kernel
};
void gibbs_tables_free(GibbsTables* t);
void gibbs_tables_layout(const Net& net, uint32_t row_limit, GibbsTables* g);
int32_t gibbs_tables_get(Net* net, uint32_t row_limit, const GibbsTables** out);
uint32_t gibbs_table_row_limit(const Net& net);  // specialize.cu: rows a blanket may have to be tabulated (0 = no tables)

// Threading contract (include/bn_cuda.h): any number of host threads / streams may use one net concurrently.  Per-call
// staging (programs, tables, evidence bytes, results) is allocated from the stream-ordered pool per call, never shared;
// the only mutable per-net state is the kernel cache below, guarded by `mu`.
struct Net {
  ~Net();
  LoopyPlan* loopy = nullptr;
  GibbsTables* gtab = nullptr;
  StatsPlan* stats_plan[2] = {nullptr, nullptr};  // stats.cu: row-lane kernel plan without / with TMA tiles
  mutable std::atomic<int> gibbs_row_limit{-1};  // specialize.cu: blanket-table row limit, decided at first use
  std::mutex mu;  // guards spec_cache, spec_demand, loopy, gtab
  std::unordered_map<std::string, SpecKernel*> spec_cache;
  std::unordered_map<std::string, double> spec_demand;  // work requested per key before its kernel was compiled
  // multi-GPU (comm.cu): a net created with bn_net_create_comm is the replica on the communicator's first local
  // device and owns the replicas on the other local devices; sharded entry points fan out over all of them
  Comm* comm = nullptr;       // not owned
  std::vector<Net*> peers;    // owned replicas on local devices 1..n_local-1
  int device = 0;
  int n = 0;
  // host copies (planning is host-side)
  std::vector<int32_t> card, par_ptr, par_idx;
  std::vector<int64_t> cpt_off;
  std::vector<double> cpt;
  std::vector<int32_t> child_ptr, child_idx;  // children in ascending topological index
  // device copies
  DevBuf<int32_t> d_card;
  DevBuf<double> d_cpt;
};

// RAII: run a scope on another (device, stream) pair, e.g. a peer replica of a multi-GPU net
struct StreamScope {
  cudaStream_t saved;
  explicit StreamScope(cudaStream_t s) : saved(tl_stream) { tl_stream = s; }
  ~StreamScope() { tl_stream = saved; }
};

// comm.cu: fan a sharded job [first, first + n) out over the local GPUs of net->comm (launch(replica, local index, lo,
// count, out) runs with the replica's device and stream current and zeroes + fills n_out doubles), then ONE
// ncclAllReduce(sum, fp64); out_dev0 lives on the first local device and is ordered on the caller's stream.
int32_t comm_fanout(Net* net, uint64_t first, uint64_t n, size_t n_out, double* out_dev0,
                    const std::function<int32_t(Net*, int, uint64_t, uint64_t, double*)>& launch);
int32_t comm_sync_peers(Net* net);
void comm_describe(const Net* net, int* world, int* first_rank, int* n_local);
cudaStream_t comm_stream(const Net* net, int local_index);
void comm_shard_of(const Net* net, int local_index, uint64_t first, uint64_t n, uint64_t* lo, uint64_t* cnt);

struct Factor {
  int device = 0;
  std::vector<int32_t> dims, card;
  int64_t len = 1;
  PoolBuf<double> data;
};

Net* get_net(bn_net_t h);
bn_net_t register_net(Net* net);
int32_t make_net(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                 const int64_t* cpt_off, const double* cpt, int32_t device, Net** out);
void build_children(Net* net);
Factor* get_factor(bn_factor_t h);
bn_factor_t register_factor(Factor* f);
Factor* take_factor(bn_factor_t h);  // removes the handle, caller deletes
int32_t new_factor(const std::vector<int32_t>& dims, const std::vector<int32_t>& card, int device,
                   Factor** out);
// convert(Factor, cpd) WITHOUT the copy: a read-only view of the node's CPT inside the net (dims = [node, parents...]).
// Internal to variable elimination, whose kernels only read their operands; never handed to callers, never normalised
// in place.  The view must not outlive the net.
int32_t factor_view_of_cpd(Net* net, int32_t node, bn_factor_t* out);
// factor_stream.cu: one-shot streaming kernels; *handled = false -> nothing launched, use the contraction kernel
int32_t sum_segmented(const double* in, const std::vector<int32_t>& card, const std::vector<char>& summed, double* out,
                      bool* handled);
int32_t push_duplicate(const double* in, long long n_in, int32_t length, double* out, bool* handled);
// permute.cu: tiled transpose behind bn_factor_permutedims; *handled = false -> nothing launched, use the contraction
int32_t permute_tiled(int device, const double* in, const std::vector<int32_t>& card, const int32_t* perm, double* out,
                      bool* handled);

// ---------------------------------------------------------------------------- sampling program
// Host-compiled form of (net, evidence) shared by the generic particle kernel (sampling.cu) and the
// network-specialised one (specialize.cu).
struct NodePlan {
  bool is_evidence = false;
  int card = 0;
  uint32_t toff = 0;       // word offset of the node's table inside the table area
  uint32_t row_words = 0;  // words per table row (thresholds padded; 2 for an fp64 evidence row)
  std::vector<std::pair<int, uint32_t>> parents;  // FREE parents: (node, stride in rows)
};
struct SampleProgram {
  std::vector<uint32_t> words;  // [prog | tables]
  uint32_t prog_words = 0;      // multiple of 4
  uint32_t table_words = 0;     // multiple of 4
  uint32_t n_free = 0;
  std::vector<NodePlan> plan;   // one per node, topological order
};
int32_t check_evidence(const Net* net, const int8_t* evidence);  // sampling.cu: states < card, else BN_BOUNDS
int32_t compile_program(const Net& net, const int8_t* evidence, bool treat_all_free, uint32_t state_stride,
                        SampleProgram* out);
// specialize.cu: returns BN_OK and sets *used when the NVRTC-specialised kernel handled the launch
int32_t launch_lw_specialized(Net* net, const int8_t* evidence, const int32_t* query, uint32_t n_query,
                              const std::vector<int64_t>& qstride, uint32_t ncell, uint64_t first, uint64_t n,
                              uint64_t seed, double* out_dev, bool* used);
int32_t launch_table_specialized(Net* net, const int8_t* evidence, int32_t kind, uint32_t max_attempts, uint64_t first,
                                 uint64_t n, uint64_t seed, uint8_t* out_states, double* out_w, uint8_t* out_ok,
                                 bool* used);

// specialize.cu: network-specialised Gibbs inference (same counts as gibbs_infer_kernel)
int32_t launch_gibbs_specialized(Net* net, const int8_t* evidence, const int32_t* query, uint32_t n_query,
                                 const std::vector<int64_t>& qstride, uint32_t ncell, bool full, uint64_t first_chain,
                                 uint64_t n_chains, int64_t burn_in, int64_t thin, int64_t total, uint64_t seed,
                                 double* out_dev, double* chain_counts_dev, uint8_t* state_dev, int32_t init_from_state,
                                 bool* used);

int32_t launch_gibbs_sample_specialized(Net* net, const int8_t* evidence, const int32_t* orders_dev, uint32_t order_stride,
                                        const uint8_t* init_dev, uint64_t first_chain, uint64_t n_chains, int64_t burn_in,
                                        int64_t nsamples, int64_t thinning, uint64_t seed, uint8_t* out_states_dev,
                                        uint64_t n_rows, bool* used);

// ---------------------------------------------------------------------------- Philox4x32-10
// Salmon et al. SC'11.  key bumps are warp-uniform so only 2 IMAD.WIDE + 2 LOP3 per round are
// per-thread work.  Counter layout is the library-wide spec (include/bn_cuda.h).
struct Philox {
  uint32_t k0, k1;
};
enum : uint32_t { STREAM_SAMPLE = 0, STREAM_GIBBS_INIT = 1, STREAM_GIBBS_UPDATE = 2, STREAM_PERM = 3 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 31-bit quantised CDF threshold, identical to oracle/bn_oracle.c cdf_threshold().
inline uint32_t cdf_threshold(double c) {
  double t = __builtin_floor(c * 2147483648.0 + 0.5);
  if (t < 0.0) t = 0.0;
  if (t > 2147483648.0) t = 2147483648.0;
  return (uint32_t)t;
}

}  // namespace bn
