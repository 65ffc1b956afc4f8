// gibbs.cu -- batched Gibbs: thousands of independent chains, one chain per thread.
//
// Replaces the scalar loops of src/Inference/gibbs.jl:66-145 (Nodewise), :161-234 (Full) and
// src/gibbs.jl:183-245 (gibbs_sample main loop).
//
// Markov-blanket conditional (src/Inference/gibbs.jl:36-50):
//     p[v] = ( prod_{c in children(n)} P(x_c | pa_c, x_n = v) ) * P(v | pa_n)     -- fp64, same multiply
// order as the reference's foldl over [children..., n]; draw by StatsBase.sample's cumulative scan.
// Chain states are u8 [node][thread] in shared memory; CPTs stay fp64 in global memory behind L1
// (the whole 200-node table is ~50 KB), walked through a host-compiled "update program" so the
// kernel never touches CSR metadata.  All control flow (sweep order, burn-in, thinning) is
// chain-independent, hence warp-uniform.
#include "common.cuh"

namespace bn {

// node record  : [card, cpt_off, npar, (parent*T, stride*card)...]
// update record: [node, node_rec, nchild, (child_rec, child*T, stride_of_node_in_child*card_child)...]
struct GibbsProgram {
  std::vector<uint32_t> words;
  std::vector<uint32_t> upd_off;  // per free node (topological order) -> offset of its update record
  uint32_t n_free = 0;
};

// gt != nullptr: nodes with a tabulated conditional (gibbs_tables.cu) get a TABLE update record instead:
//   [node | 0x80000000, table offset (words), K | nmem << 16, (member * T, stride * row words) x nmem]
static int32_t compile_gibbs(const Net& net, const int8_t* evidence, uint32_t T, GibbsProgram* out,
                             const GibbsTables* gt = nullptr) {
  const int n = net.n;
  BN_REQUIRE(net.cpt_off[n] < (int64_t)0x7fffffff, BN_UNSUPPORTED, "CPT too large for the Gibbs kernel");
  std::vector<uint32_t>& w = out->words;
  std::vector<uint32_t> rec(n);
  for (int i = 0; i < n; ++i) {
    rec[i] = (uint32_t)w.size();
    w.push_back((uint32_t)net.card[i]);
    w.push_back((uint32_t)net.cpt_off[i]);
    w.push_back((uint32_t)(net.par_ptr[i + 1] - net.par_ptr[i]));
    int64_t stride = 1;
    for (int k = net.par_ptr[i]; k < net.par_ptr[i + 1]; ++k) {
      int p = net.par_idx[k];
      w.push_back((uint32_t)p * T);
      w.push_back((uint32_t)(stride * net.card[i]));
      stride *= net.card[p];
    }
  }
  for (int i = 0; i < n; ++i) {
    if (evidence && evidence[i] >= 0) {
      BN_REQUIRE(evidence[i] < net.card[i], BN_BOUNDS, "evidence state %d out of range for node %d", (int)evidence[i], i);
      continue;
    }
    out->upd_off.push_back((uint32_t)w.size());
    if (gt && gt->tab_off[i] >= 0) {
      const int K = net.card[i];
      const uint32_t W = K <= 2 ? 1u : (K == 3 ? 2u : (K == 4 ? 4u : 8u));
      w.push_back((uint32_t)i | 0x80000000u);
      w.push_back((uint32_t)gt->tab_off[i]);
      w.push_back((uint32_t)K | ((uint32_t)gt->members[i].size() << 16));
      for (size_t m = 0; m < gt->members[i].size(); ++m) {
        w.push_back((uint32_t)gt->members[i][m] * T);
        w.push_back(gt->strides[i][m] * W);
      }
      continue;
    }
    w.push_back((uint32_t)i);
    w.push_back(rec[i]);
    w.push_back((uint32_t)(net.child_ptr[i + 1] - net.child_ptr[i]));
    for (int k = net.child_ptr[i]; k < net.child_ptr[i + 1]; ++k) {
      int c = net.child_idx[k];
      int64_t stride = 1, mine = 0;
      for (int k2 = net.par_ptr[c]; k2 < net.par_ptr[c + 1]; ++k2) {
        if (net.par_idx[k2] == i) mine = stride;
        stride *= net.card[net.par_idx[k2]];
      }
      w.push_back(rec[c]);
      w.push_back((uint32_t)c * T);
      w.push_back((uint32_t)(mine * net.card[c]));
    }
  }
  out->n_free = (uint32_t)out->upd_off.size();
  return BN_OK;
}

// ------------------------------------------------------------------------------------------ device
struct GibbsArgs {
  const uint32_t* prog;     // program words
  const uint32_t* upd_off;  // [n_free]
  const double* cpt;
  const uint32_t* tab;      // gibbs_tables.cu thresholds (nullptr: no table records in the program)
  const int8_t* evidence;   // [n_nodes]
  const int32_t* card;      // [n_nodes]
  uint32_t n_nodes, n_free;
  uint64_t first_chain, n_chains;
  uint32_t seed_lo, seed_hi;
  // infer
  int64_t burn_in, thin, total;
  int32_t full;
  const uint32_t* query;  // pairs (node*T, stride)
  uint32_t n_query, ncell;
  double* out;                // [ncell] summed over chains
  double* out_chain_counts;   // optional [n_chains x ncell]
  uint8_t* state_inout;       // optional chain-major
  int32_t init_from_state;
  // sample
  const int32_t* orders;      // [n_sweeps (or 1) x n_free] positions into the free list
  uint32_t order_stride;      // n_free, or 0 for a fixed order
  int64_t nsamples, thinning;
  uint8_t* out_states;        // node-major [n_nodes x n_rows]
  uint64_t n_rows;
};

struct ChainRng {
  uint32_t c0, c1, stream, k0, k1;
  uint32_t buf[4];
  uint32_t d;
  __device__ __forceinline__ void init(uint64_t idx, uint32_t strm, uint32_t s0, uint32_t s1) {
    c0 = (uint32_t)idx; c1 = (uint32_t)(idx >> 32); stream = strm; k0 = s0; k1 = s1; d = 0;
  }
  __device__ __forceinline__ uint32_t next() {
    if ((d & 3u) == 0u) philox4x32_10(c0, c1, d >> 2, stream, k0, k1, buf);
    const uint32_t r = (d & 2u) ? ((d & 1u) ? buf[3] : buf[2]) : ((d & 1u) ? buf[1] : buf[0]);
    ++d;
    return r;
  }
};

// p[j] = conditional weight of state v0 + j for j < 8 (1.0-initialised product, children first, own CPD last)
__device__ __forceinline__ void mb_chunk(const GibbsArgs& a, const uint32_t* __restrict__ u, const uint8_t* st,
                                         uint32_t T, uint32_t v0, uint32_t K, double (&p)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = 1.0;
  const uint32_t node = __ldg(u + 0);
  const uint32_t nrec = __ldg(u + 1);
  const uint32_t nchild = __ldg(u + 2);
  const uint32_t xn = st[node * T];
  for (uint32_t k = 0; k < nchild; ++k) {
    const uint32_t crec = __ldg(u + 3 + 3 * k);
    const uint32_t cT = __ldg(u + 4 + 3 * k);
    const uint32_t scn = __ldg(u + 5 + 3 * k);
    const uint32_t* cr = a.prog + crec;
    const uint32_t npar = __ldg(cr + 2);
    uint32_t base = __ldg(cr + 1) + st[cT];
    for (uint32_t q = 0; q < npar; ++q) base += __ldg(cr + 4 + 2 * q) * (uint32_t)st[__ldg(cr + 3 + 2 * q)];
    base -= scn * xn;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (v0 + j < K) p[j] *= __ldg(a.cpt + base + (v0 + j) * scn);
  }
  const uint32_t* nr = a.prog + nrec;
  const uint32_t npar = __ldg(nr + 2);
  uint32_t base = __ldg(nr + 1);
  for (uint32_t q = 0; q < npar; ++q) base += __ldg(nr + 4 + 2 * q) * (uint32_t)st[__ldg(nr + 3 + 2 * q)];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (v0 + j < K) p[j] *= __ldg(a.cpt + base + v0 + j);
}

// x_n ~ P(x_n | mb(x_n)): eval_mb_cpd + sample(Weights) (src/Inference/gibbs.jl:110-112)
__device__ __forceinline__ void gibbs_update(const GibbsArgs& a, uint32_t upd, uint8_t* st, uint32_t T, uint32_t r) {
  const uint32_t* u = a.prog + upd;
  const uint32_t w0 = __ldg(u + 0);
  if (w0 & 0x80000000u) {  // tabulated conditional: pick = #{v : r >= R_v}, the same decision as the scan below
    const uint32_t info = __ldg(u + 2), K = info & 0xffffu, nmem = info >> 16;
    uint32_t off = __ldg(u + 1);
    for (uint32_t m = 0; m < nmem; ++m) off += __ldg(u + 4 + 2 * m) * (uint32_t)st[__ldg(u + 3 + 2 * m)];
    const uint32_t* tp = a.tab + off;
    uint32_t pick = 0;
    for (uint32_t v = 0; v + 1 < K; ++v) pick += (uint32_t)(r >= __ldg(tp + v));
    st[(w0 & 0x7fffffffu) * T] = (uint8_t)pick;
    return;
  }
  const uint32_t node = w0;
  const uint32_t K = __ldg(a.prog + __ldg(u + 1));
  const double uu = ((double)r + 0.5) * (1.0 / 4294967296.0);
  double p[8];
  uint32_t pick = 0;
  if (K <= 8) {
    mb_chunk(a, u, st, T, 0, K, p);
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if ((uint32_t)j < K) sum += p[j];
    const double t = uu * sum;
    double cw = p[0];
    bool go = true;
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      go = go && (cw < t) && ((uint32_t)j < K);
      if (go) { cw += p[j]; pick = j; }
    }
  } else {
    double sum = 0.0;
    for (uint32_t v0 = 0; v0 < K; v0 += 8) {
      mb_chunk(a, u, st, T, v0, K, p);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v0 + j < K) sum += p[j];
    }
    const double t = uu * sum;
    double cw = 0.0;
    bool go = true;
    for (uint32_t v0 = 0; v0 < K && go; v0 += 8) {
      mb_chunk(a, u, st, T, v0, K, p);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t v = v0 + j;
        if (v == 0) { cw = p[0]; continue; }
        go = go && (cw < t) && (v < K);
        if (go) { cw += p[j]; pick = v; }
      }
    }
  }
  st[node * T] = (uint8_t)pick;
}

__device__ __forceinline__ void init_uniform(const GibbsArgs& a, uint64_t chain, uint8_t* st, uint32_t T) {
  // _init_gibbs_sample: src/Inference/gibbs.jl:6-20
  ChainRng rng;
  rng.init(chain, STREAM_GIBBS_INIT, a.seed_lo, a.seed_hi);
  for (uint32_t i = 0; i < a.n_nodes; ++i) {
    const int e = a.evidence[i];
    st[i * T] = e >= 0 ? (uint8_t)e : (uint8_t)__umulhi(rng.next(), (uint32_t)a.card[i]);
  }
}

__global__ void __launch_bounds__(256) gibbs_infer_kernel(const GibbsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(smem_raw);
  uint8_t* st = reinterpret_cast<uint8_t*>(bins + a.ncell) + tid;
  for (uint32_t c = tid; c < a.ncell; c += T) bins[c] = 0ull;
  __syncthreads();
  const uint64_t ci = (uint64_t)blockIdx.x * T + tid;
  const bool live = ci < a.n_chains;
  if (live) {
    const uint64_t chain = a.first_chain + ci;
    if (a.state_inout && a.init_from_state)
      for (uint32_t i = 0; i < a.n_nodes; ++i) st[i * T] = a.state_inout[ci * a.n_nodes + i];
    else
      init_uniform(a, chain, st, T);
    ChainRng rng;
    rng.init(chain, STREAM_GIBBS_UPDATE, a.seed_lo, a.seed_hi);
    // a sample is kept at iterations burn_in + k * thin, k >= 1 (src/Inference/gibbs.jl:118-121): a countdown instead
    // of a 64-bit modulo per update
    int64_t until = a.burn_in + a.thin, num_smpls = 0;
    bool finished = (a.n_free == 0 && !a.full);
    while (!finished) {
      for (uint32_t f = 0; f < a.n_free; ++f) {
        gibbs_update(a, __ldg(a.upd_off + f), st, T, rng.next());
        if (!a.full) {
          if (--until == 0) {  // one node update = one iteration (src/Inference/gibbs.jl:120)
            until = a.thin;
            uint32_t cell = 0;
            for (uint32_t j = 0; j < a.n_query; ++j) cell += a.query[2 * j + 1] * (uint32_t)st[a.query[2 * j]];
            atomicAdd(&bins[cell], 1ull);
            if (a.out_chain_counts) a.out_chain_counts[ci * a.ncell + cell] += 1.0;
            ++num_smpls;
          }
          if (num_smpls >= a.total) { finished = true; break; }
        }
      }
      if (a.full) {
        if (--until == 0) {  // one sweep = one iteration (:214)
          until = a.thin;
          uint32_t cell = 0;
          for (uint32_t j = 0; j < a.n_query; ++j) cell += a.query[2 * j + 1] * (uint32_t)st[a.query[2 * j]];
          atomicAdd(&bins[cell], 1ull);
          if (a.out_chain_counts) a.out_chain_counts[ci * a.ncell + cell] += 1.0;
          ++num_smpls;
        }
        if (num_smpls >= a.total) finished = true;
      }
    }
    if (a.state_inout)
      for (uint32_t i = 0; i < a.n_nodes; ++i) a.state_inout[ci * a.n_nodes + i] = st[i * T];
  }
  __syncthreads();
  for (uint32_t c = tid; c < a.ncell; c += T)
    if (bins[c]) atomicAdd(&a.out[c], (double)bins[c]);  // integers < 2^53: exact, order-independent
}

// Importance-resample one of 50 weighted particles per chain (src/gibbs.jl:330-335,
// src/sampling.jl:182-197).  states: node-major [n_nodes x 50*n_chains]; out: chain-major.
__global__ void gibbs_resample_kernel(const uint8_t* states, const double* w, uint32_t n_nodes, uint64_t n_chains,
                                      uint64_t first_chain, uint32_t seed_lo, uint32_t seed_hi, uint8_t* init,
                                      int* zero_flag) {
  const uint64_t ci = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_chains) return;
  const double* wc = w + ci * 50;
  double sw = 0.0;
  for (int j = 0; j < 50; ++j) sw += wc[j];
  if (!(sw > 0.0)) { atomicExch(zero_flag, 1); sw = 1.0; }
  uint32_t r[4];
  const uint64_t chain = first_chain + ci;
  philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), 0, STREAM_GIBBS_INIT, seed_lo, seed_hi, r);
  const double t = (((double)r[0] + 0.5) * (1.0 / 4294967296.0)) * sw;
  double c = wc[0];
  int pick = 0;
  while (c < t && pick < 49) c += wc[++pick];
  const uint64_t n_rows = n_chains * 50;
  for (uint32_t i = 0; i < n_nodes; ++i) init[ci * n_nodes + i] = states[(uint64_t)i * n_rows + ci * 50 + pick];
}

__global__ void __launch_bounds__(256) gibbs_sample_kernel(const GibbsArgs a, const uint8_t* init) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  uint8_t* st = smem_raw + tid;
  const uint64_t ci = (uint64_t)blockIdx.x * T + tid;
  if (ci >= a.n_chains) return;
  const uint64_t chain = a.first_chain + ci;
  for (uint32_t i = 0; i < a.n_nodes; ++i) st[i * T] = init[ci * a.n_nodes + i];
  ChainRng rng;
  rng.init(chain, STREAM_GIBBS_UPDATE, a.seed_lo, a.seed_hi);
  int64_t sweep = 0;
  auto do_sweep = [&]() {
    const int32_t* ord = a.orders + (uint64_t)sweep * a.order_stride;
    for (uint32_t f = 0; f < a.n_free; ++f) gibbs_update(a, __ldg(a.upd_off + __ldg(ord + f)), st, T, rng.next());
    ++sweep;
  };
  for (int64_t b = 0; b < a.burn_in; ++b) do_sweep();  // burn-in sweeps, all discarded (src/gibbs.jl:339)
  for (int64_t s = 0; s < a.nsamples; ++s) {
    for (int64_t k = 0; k <= a.thinning; ++k) do_sweep();  // `thinning` skipped + 1 kept (:226-240)
    for (uint32_t i = 0; i < a.n_nodes; ++i)
      a.out_states[(uint64_t)i * a.n_rows + ci * (uint64_t)a.nsamples + (uint64_t)s] = st[i * T];
  }
}

// ------------------------------------------------------------------------------------------ host
static void sweep_permutation(uint64_t seed, uint64_t sweep, int32_t m, int32_t* perm) {
  // identical to oracle/bn_oracle.c orc_sweep_permutation
  uint32_t buf[4] = {0, 0, 0, 0};
  int64_t blk = -1;
  for (int j = 0; j < m; ++j) perm[j] = j;
  for (int j = 0; j + 1 < m; ++j) {
    if ((j >> 2) != blk) {
      blk = j >> 2;
      philox4x32_10((uint32_t)sweep, (uint32_t)(sweep >> 32), (uint32_t)blk, STREAM_PERM, (uint32_t)seed,
                    (uint32_t)(seed >> 32), buf);
    }
    uint32_t r = buf[j & 3];
    int k = j + (int)(((uint64_t)r * (uint64_t)(m - j)) >> 32);
    int32_t t = perm[j]; perm[j] = perm[k]; perm[k] = t;
  }
}

struct GibbsDevice {
  TmpBuf<uint32_t> prog, upd_off, query;
  TmpBuf<int8_t> evidence;
  TmpBuf<int32_t> orders;
};

static int32_t stage_gibbs(Net* net, const int8_t* evidence, uint32_t T, const int32_t* query, int32_t n_query,
                           const std::vector<int64_t>& qstride, GibbsDevice* dev, GibbsArgs* a) {
  // tabulated conditionals for small-blanket nodes, shared with the specialised kernels (built once per net)
  const GibbsTables* gt = nullptr;
  const uint32_t row_limit = gibbs_table_row_limit(*net);
  if (row_limit > 0) {
    BN_TRY(gibbs_tables_get(net, row_limit, &gt));
    if (gt && (gt->words == 0 || gt->has_never)) gt = nullptr;
  }
  GibbsProgram gp;
  BN_TRY(compile_gibbs(*net, evidence, T, &gp, gt));
  a->tab = gt ? gt->tab.p : nullptr;
  cudaStream_t s = tl_stream;
  BN_CUDA_TRY(dev->prog.upload(gp.words.data(), gp.words.size(), s));
  BN_CUDA_TRY(dev->upd_off.upload(gp.upd_off.data(), gp.upd_off.size(), s));
  std::vector<int8_t> ev(net->n, -1);
  if (evidence) ev.assign(evidence, evidence + net->n);
  BN_CUDA_TRY(dev->evidence.upload(ev.data(), ev.size(), s));
  std::vector<uint32_t> q(2 * (size_t)n_query);
  for (int j = 0; j < n_query; ++j) { q[2 * j] = (uint32_t)query[j] * T; q[2 * j + 1] = (uint32_t)qstride[j]; }
  BN_CUDA_TRY(dev->query.upload(q.data(), q.size(), s));
  a->prog = dev->prog.p;
  a->upd_off = dev->upd_off.p;
  a->cpt = net->d_cpt.p;
  a->evidence = dev->evidence.p;
  a->card = net->d_card.p;
  a->n_nodes = (uint32_t)net->n;
  a->n_free = gp.n_free;
  a->query = dev->query.p;
  a->n_query = (uint32_t)n_query;
  return BN_OK;
}

static int gibbs_threads(const DevProps& dp, int n_nodes, size_t extra, uint64_t n_chains) {
  // one chain per thread; the largest CTA whose state fits and that still gives every SM >= 4 CTAs to balance
  // (65536 chains in 256-thread CTAs are 1.7 CTAs per SM: a third of the SMs would run twice as long)
  int fit = 0;
  for (int t : {256, 128, 64, 32})
    if ((size_t)n_nodes * t + extra + 64 <= (size_t)dp.max_smem_optin) {
      if (!fit) fit = t;
      if (n_chains >= (uint64_t)t * dp.sm_count * 4) return t;
    }
  return fit ? (fit > 64 ? 64 : fit) : 0;
}

}  // namespace bn

using namespace bn;

extern "C" {

// One GPU's share of a Gibbs inference job: chains [first_chain, first_chain + n_chains) on `net`'s device, enqueued
// on the current stream (asynchronous: per-call staging is stream-ordered pool memory).
static int32_t gibbs_infer_impl(Net* net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                                uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in,
                                int64_t thin, int32_t mode, uint64_t seed, uint8_t* d_state, int32_t init_from_state,
                                double* out_dev, double* d_chain_counts) {
  // src/Inference/gibbs.jl:70-71
  BN_REQUIRE(burn_in < nsamples, BN_BAD_ARG, "Burn in (%lld) must be less than number of samples (%lld)",
             (long long)burn_in, (long long)nsamples);
  BN_REQUIRE(thin >= 1 && burn_in >= 0, BN_BAD_ARG, "thin must be >= 1 and burn_in >= 0");
  BN_REQUIRE(mode == BN_GIBBS_NODEWISE || mode == BN_GIBBS_FULL, BN_BAD_ARG, "bad Gibbs mode %d", mode);
  BN_CUDA_TRY(cudaSetDevice(net->device));
  int64_t ncell = 1;
  std::vector<int64_t> qstride;
  for (int j = 0; j < n_query; ++j) {
    BN_REQUIRE(query && query[j] >= 0 && query[j] < net->n, BN_BAD_ARG,
               "Query %d is not in the probabilistic graphical model", query ? query[j] : -1);
    BN_REQUIRE(!(evidence && evidence[query[j]] >= 0), BN_BAD_ARG, "Query %d is part of the evidence", query[j]);
    qstride.push_back(ncell);
    ncell *= net->card[query[j]];
    BN_REQUIRE(ncell <= 4096, BN_UNSUPPORTED, "query table has more than 4096 cells");
  }
  BN_TRY(check_evidence(net, evidence));
  BN_CUDA_TRY(cudaMemsetAsync(out_dev, 0, (size_t)ncell * 8, tl_stream));
  if (n_chains == 0) return BN_OK;
  {  // network-specialised kernel first (specialize.cu); identical counts
    bool used = false;
    BN_TRY(launch_gibbs_specialized(net, evidence, query, (uint32_t)n_query, qstride, (uint32_t)ncell,
                                    mode == BN_GIBBS_FULL, first_chain, n_chains, burn_in, thin,
                                    (nsamples - burn_in + thin - 1) / thin, seed, out_dev, d_chain_counts, d_state,
                                    init_from_state, &used));
    if (used) return BN_OK;
  }
  const DevProps& dp = dev_props(net->device);
  const int T = gibbs_threads(dp, net->n, (size_t)ncell * 8, n_chains);
  BN_REQUIRE(T > 0, BN_UNSUPPORTED, "network too large for the Gibbs kernel (%d nodes)", net->n);
  GibbsDevice dev;  // stream-ordered pool staging: freed after the kernel, in stream order
  GibbsArgs a{};
  BN_TRY(stage_gibbs(net, evidence, (uint32_t)T, query, n_query, qstride, &dev, &a));
  a.first_chain = first_chain;
  a.n_chains = n_chains;
  a.seed_lo = (uint32_t)seed;
  a.seed_hi = (uint32_t)(seed >> 32);
  a.burn_in = burn_in;
  a.thin = thin;
  a.total = (nsamples - burn_in + thin - 1) / thin;  // :74
  a.full = mode == BN_GIBBS_FULL;
  a.ncell = (uint32_t)ncell;
  a.out = out_dev;
  a.out_chain_counts = d_chain_counts;
  a.state_inout = d_state;
  a.init_from_state = init_from_state;
  const size_t smem = (size_t)ncell * 8 + (size_t)net->n * T;
  BN_CUDA_TRY(cudaFuncSetAttribute(gibbs_infer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((n_chains + T - 1) / T);
  gibbs_infer_kernel<<<grid, T, smem, tl_stream>>>(a);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

static int64_t query_cells(const Net* net, const int32_t* query, int32_t n_query) {
  int64_t ncell = 1;
  for (int j = 0; j < n_query; ++j)
    if (query && query[j] >= 0 && query[j] < net->n) ncell *= net->card[query[j]];
  return ncell;
}

int32_t bn_gibbs_infer_dev(bn_net_t h, const int8_t* evidence, const int32_t* query, int32_t n_query,
                           uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in,
                           int64_t thin, int32_t mode, uint64_t seed, double* out_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_gibbs_infer: unknown net handle");
  BN_REQUIRE(out_dev, BN_BAD_ARG, "bn_gibbs_infer_dev: output is NULL");
  if (!net->comm)
    return gibbs_infer_impl(net, evidence, query, n_query, first_chain, n_chains, nsamples, burn_in, thin, mode, seed,
                            nullptr, 0, out_dev, nullptr);
  // multi-GPU net: chains are sharded by global chain index, integer counts all-reduced (exactly the 1-GPU counts)
  BN_CUDA_TRY(cudaSetDevice(net->device));
  const int64_t ncell = query_cells(net, query, n_query);
  BN_REQUIRE(ncell <= 4096, BN_UNSUPPORTED, "query table has more than 4096 cells");
  return comm_fanout(net, first_chain, n_chains, (size_t)ncell, out_dev,
                     [&](Net* r, int, uint64_t lo, uint64_t cnt, double* out) {
                       return gibbs_infer_impl(r, evidence, query, n_query, lo, cnt, nsamples, burn_in, thin, mode, seed,
                                               nullptr, 0, out, nullptr);
                     });
}

int32_t bn_gibbs_infer(bn_net_t h, const int8_t* evidence, const int32_t* query, int32_t n_query,
                       uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in, int64_t thin,
                       int32_t mode, uint64_t seed, uint8_t* state_inout, int32_t init_from_state,
                       double* out_counts, double* out_chain_counts) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_gibbs_infer: unknown net handle");
  BN_REQUIRE(out_counts, BN_BAD_ARG, "bn_gibbs_infer: out_counts is NULL");
  BN_REQUIRE(!(init_from_state && !state_inout), BN_BAD_ARG, "init_from_state needs state_inout");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  const int64_t ncell = query_cells(net, query, n_query);
  BN_REQUIRE(ncell <= 4096, BN_UNSUPPORTED, "query table has more than 4096 cells");
  const size_t nn = (size_t)net->n;
  if (init_from_state)  // a resumed state indexes the CPTs on the device: it must be inside the cardinalities
    for (uint64_t c = 0; c < n_chains; ++c)
      for (size_t i = 0; i < nn; ++i)
        BN_REQUIRE(state_inout[c * nn + i] < net->card[i], BN_BOUNDS, "chain %llu: state %d of node %d outside 0..%d",
                   (unsigned long long)c, (int)state_inout[c * nn + i], (int)i, net->card[i] - 1);
  int world = 1, first_rank = 0, nl = 1;
  comm_describe(net, &world, &first_rank, &nl);
  // per local GPU: the rows (chains) of state_inout / out_chain_counts that GPU runs
  struct Shard { TmpBuf<uint8_t> st; TmpBuf<double> cc; uint64_t lo = 0, cnt = 0; cudaStream_t s = nullptr; int device = 0; };
  std::vector<Shard> sh((size_t)nl);
  auto launch = [&](Net* r, int d, uint64_t lo, uint64_t cnt, double* out) -> int32_t {
    Shard& S = sh[(size_t)d];
    S.lo = lo; S.cnt = cnt; S.s = tl_stream; S.device = r->device;
    const size_t row0 = (size_t)(lo - first_chain);
    if (state_inout) {
      BN_CUDA_TRY(S.st.alloc((size_t)cnt * nn));
      if (init_from_state && cnt)
        BN_CUDA_TRY(cudaMemcpyAsync(S.st.p, state_inout + row0 * nn, (size_t)cnt * nn, cudaMemcpyHostToDevice, tl_stream));
    }
    if (out_chain_counts) {
      BN_CUDA_TRY(S.cc.alloc((size_t)cnt * ncell));
      BN_CUDA_TRY(cudaMemsetAsync(S.cc.p, 0, (size_t)cnt * ncell * 8, tl_stream));
    }
    return gibbs_infer_impl(r, evidence, query, n_query, lo, cnt, nsamples, burn_in, thin, mode, seed,
                            state_inout ? S.st.p : nullptr, init_from_state, out, out_chain_counts ? S.cc.p : nullptr);
  };
  TmpBuf<double> d_out;
  BN_CUDA_TRY(d_out.alloc((size_t)ncell));
  int32_t rc = net->comm ? comm_fanout(net, first_chain, n_chains, (size_t)ncell, d_out.p, launch)
                         : launch(net, 0, first_chain, n_chains, d_out.p);
  if (rc == BN_OK) {
    BN_CUDA_TRY(cudaMemcpyAsync(out_counts, d_out.p, (size_t)ncell * 8, cudaMemcpyDeviceToHost, tl_stream));
    for (Shard& S : sh) {
      BN_CUDA_TRY(cudaSetDevice(S.device));
      const size_t row0 = (size_t)(S.lo - first_chain);
      if (state_inout && S.cnt)
        BN_CUDA_TRY(cudaMemcpyAsync(state_inout + row0 * nn, S.st.p, (size_t)S.cnt * nn, cudaMemcpyDeviceToHost, S.s));
      if (out_chain_counts && S.cnt)
        BN_CUDA_TRY(cudaMemcpyAsync(out_chain_counts + row0 * ncell, S.cc.p, (size_t)S.cnt * ncell * 8,
                                    cudaMemcpyDeviceToHost, S.s));
    }
  }
  for (Shard& S : sh) {  // wait for every local GPU, release its staging on its own device
    if (cudaSetDevice(S.device) != cudaSuccess) continue;
    cudaError_t e = cudaStreamSynchronize(S.s);
    S.st.release();
    S.cc.release();
    if (e != cudaSuccess && rc == BN_OK) {
      set_error("bn_gibbs_infer: %s", cudaGetErrorString(e));
      rc = BN_CUDA_ERR;
    }
  }
  cudaSetDevice(net->device);
  if (rc == BN_OK) BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return rc;
}

int32_t bn_gibbs_sample(bn_net_t h, const int8_t* evidence, uint64_t first_chain, uint64_t n_chains,
                        int64_t nsamples, int64_t burn_in, int64_t thinning, const int32_t* variable_order,
                        uint64_t seed, const uint8_t* init_state, uint8_t* out_states) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_gibbs_sample: unknown net handle");
  BN_REQUIRE(out_states, BN_BAD_ARG, "bn_gibbs_sample: out_states is NULL");
  // src/gibbs.jl:299-301
  BN_REQUIRE(nsamples > 0, BN_BAD_ARG, "nsamples parameter less than 1");
  BN_REQUIRE(burn_in >= 0, BN_BAD_ARG, "Negative burn_in parameter");
  BN_REQUIRE(thinning >= 0, BN_BAD_ARG, "Negative thinning parameter");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  if (n_chains == 0) return BN_OK;
  const DevProps& dp = dev_props(net->device);
  const int T = gibbs_threads(dp, net->n, 0, n_chains);
  BN_REQUIRE(T > 0, BN_UNSUPPORTED, "network too large for the Gibbs kernel (%d nodes)", net->n);
  cudaStream_t s = tl_stream;
  GibbsDevice dev;
  GibbsArgs a{};
  std::vector<int64_t> none;
  BN_TRY(stage_gibbs(net, evidence, (uint32_t)T, nullptr, 0, none, &dev, &a));
  const int32_t n_free = (int32_t)a.n_free;
  // sweep orders
  const int64_t n_sweeps = burn_in + nsamples * (thinning + 1);
  std::vector<int32_t> orders;
  if (variable_order) {
    std::vector<char> seen(n_free > 0 ? n_free : 1, 0);
    for (int j = 0; j < n_free; ++j) {
      BN_REQUIRE(variable_order[j] >= 0 && variable_order[j] < n_free && !seen[variable_order[j]], BN_BAD_ARG,
                 "variable_order must be a permutation of the non-evidence nodes");
      seen[variable_order[j]] = 1;
    }
    orders.assign(variable_order, variable_order + n_free);
    a.order_stride = 0;
  } else {
    BN_REQUIRE((double)n_sweeps * (double)(n_free > 0 ? n_free : 1) < 5e8, BN_UNSUPPORTED,
               "too many sweeps x variables for precomputed random orders");
    orders.resize((size_t)n_sweeps * (size_t)(n_free > 0 ? n_free : 1));
    for (int64_t sw = 0; sw < n_sweeps; ++sw) sweep_permutation(seed, (uint64_t)sw, n_free, orders.data() + sw * n_free);
    a.order_stride = (uint32_t)n_free;
  }
  if (orders.empty()) orders.push_back(0);
  BN_CUDA_TRY(dev.orders.upload(orders.data(), orders.size(), s));
  a.orders = dev.orders.p;

  // initial states
  TmpBuf<uint8_t> d_init, d_cand, d_out;
  TmpBuf<double> d_w;
  TmpBuf<int> d_flag;
  const size_t nst = (size_t)n_chains * net->n;
  BN_CUDA_TRY(d_init.alloc(nst));
  if (init_state) {
    BN_CUDA_TRY(cudaMemcpyAsync(d_init.p, init_state, nst, cudaMemcpyHostToDevice, s));
  } else {
    const uint64_t n_cand = n_chains * 50;
    BN_CUDA_TRY(d_cand.alloc((size_t)net->n * n_cand));
    BN_CUDA_TRY(d_w.alloc(n_cand));
    BN_CUDA_TRY(d_flag.alloc(1));
    BN_CUDA_TRY(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));
    BN_TRY(bn_sample_dev(h, evidence, BN_SAMPLE_WEIGHTED, 1, first_chain * 50, n_cand, seed, d_cand.p, d_w.p, nullptr));
    gibbs_resample_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, s>>>(
        d_cand.p, d_w.p, (uint32_t)net->n, n_chains, first_chain, (uint32_t)seed, (uint32_t)(seed >> 32), d_init.p,
        d_flag.p);
    count_launch();
    BN_CUDA_TRY(cudaGetLastError());
    int flag = 0;
    BN_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    BN_CUDA_TRY(cudaStreamSynchronize(s));
    BN_REQUIRE(!flag, BN_ZERO_WEIGHT,
               "Gibbs Sampler was unable to find an inital sample with non-zero probability, please supply an inital sample");
  }
  a.first_chain = first_chain;
  a.n_chains = n_chains;
  a.seed_lo = (uint32_t)seed;
  a.seed_hi = (uint32_t)(seed >> 32);
  a.burn_in = burn_in;
  a.nsamples = nsamples;
  a.thinning = thinning;
  a.n_rows = n_chains * (uint64_t)nsamples;
  BN_CUDA_TRY(d_out.alloc((size_t)net->n * a.n_rows));
  a.out_states = d_out.p;
  bool used = false;  // network-specialised kernel first (specialize.cu); same tables either way
  BN_TRY(launch_gibbs_sample_specialized(net, evidence, dev.orders.p, a.order_stride, d_init.p, first_chain, n_chains, burn_in,
                                         nsamples, thinning, seed, d_out.p, a.n_rows, &used));
  if (!used) {
    const size_t smem = (size_t)net->n * T;
    BN_CUDA_TRY(cudaFuncSetAttribute(gibbs_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_sample_kernel<<<(unsigned)((n_chains + T - 1) / T), T, smem, s>>>(a, d_init.p);
    count_launch();
    BN_CUDA_TRY(cudaGetLastError());
  }
  BN_CUDA_TRY(cudaMemcpyAsync(out_states, d_out.p, (size_t)net->n * a.n_rows, cudaMemcpyDeviceToHost, s));
  BN_CUDA_TRY(cudaStreamSynchronize(s));
  return BN_OK;
}

}  // extern "C"
