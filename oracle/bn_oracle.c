/*
 * bn_oracle.c -- CPU restatement of the BayesNets.jl discrete-inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The shipped path is bayesnets.jl_b200/csrc (CUDA, sm_100a) behind include/bn_cuda.h.
 *
 * Parity status: the reference (pure Julia, un-vendored deps) cannot run in this image, so this
 * restatement is pinned against the reference's own golden vectors / closed-form posteriors
 * (tests/test_oracle_golden.py).  RNG bit-streams of the reference
 * (Julia MersenneTwister/Xoshiro via Distributions.rand / StatsBase.sample) are NOT pinned by any
 * reference test -> sampled paths are "parity unpinned" at the bit-stream level and are checked
 * statistically (4 sigma) instead; the deterministic fp64 paths (Factor * / sum / normalize!,
 * ExactInference) are pinned by the 324-entry golden product and friends.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Layout conventions (shared with the device layout, bayesnets.jl_b200/device_layout.py):
 *   - nodes are in topological order (src/bayes_nets.jl:26-48); states are 0-based here
 *   - CPT of node i: cpt[cpt_off[i] + s + card[i] * q], q = sum_k state(parent_k) * stride_k,
 *     stride_0 = 1, stride_k = prod_{j<k} card(parent_j)  -- first parent fastest,
 *     exactly vcat(d.p for d in distributions) (src/CPDs/categorical_cpd.jl:36-46,
 *     src/Factors/factors_main.jl:62-67)
 *   - factors are column-major Float64 (src/Factors/factors_main.jl:18-36)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#include <float.h>
#endif

#define ORC_MAX_CARD 64
#define ORC_MAXD 64

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 * The reference draws from Julia's global RNG (src/Inference/likelihood.jl:42 -> Distributions.rand;
 * src/Inference/gibbs.jl:112 -> StatsBase.sample); the stream itself is unpinned, so the device
 * spec substitutes a counter-based generator.  Known-answer vectors: tests/test_oracle_golden.py.
 * ---------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* Draw number `d` of logical stream (idx, stream) under `seed`:
 *   counter = (idx_lo, idx_hi, d >> 2, stream), key = (seed_lo, seed_hi), word d & 3.
 * A tiny one-block cache makes sequential d cheap. */
typedef struct {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t buf[4];
  int64_t  blk;
} orc_stream;

static inline void stream_init(orc_stream* s, uint64_t seed, uint64_t idx, uint32_t stream) {
  s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
  s->ctr[0] = (uint32_t)idx;  s->ctr[1] = (uint32_t)(idx >> 32);
  s->ctr[2] = 0;              s->ctr[3] = stream;
  s->blk = -1;
}
static inline uint32_t stream_draw(orc_stream* s, uint64_t d) {
  int64_t b = (int64_t)(d >> 2);
  if (b != s->blk) {
    s->ctr[2] = (uint32_t)b;
    orc_philox4x32_10(s->ctr, s->key, s->buf);
    s->blk = b;
  }
  return s->buf[d & 3];
}

enum { ORC_STREAM_SAMPLE = 0, ORC_STREAM_GIBBS_INIT = 1, ORC_STREAM_GIBBS_UPDATE = 2, ORC_STREAM_PERM = 3 };

/* ------------------------------------------------------------------------------------------
 * Flat network view.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n;
  const int32_t* card;
  const int32_t* par_ptr;
  const int32_t* par_idx;
  const int64_t* cpt_off;
  const double*  cpt;
} orc_net;

/* CPT row lookup: src/CPDs/categorical_cpd.jl:36-46 -- LinearIndices(parental_ncategories)[sub...],
 * column-major over parents, first parent fastest. */
static inline const double* cpt_row(const orc_net* bn, int i, const uint8_t* x) {
  int64_t q = 0, stride = 1;
  for (int k = bn->par_ptr[i]; k < bn->par_ptr[i + 1]; ++k) {
    int p = bn->par_idx[k];
    q += stride * x[p];
    stride *= bn->card[p];
  }
  return bn->cpt + bn->cpt_off[i] + (int64_t)bn->card[i] * q;
}

/* Device-spec quantised CDF threshold: T_k = round(c_k * 2^31), c_k = fl(p_0 + ... + p_k). */
static inline uint32_t cdf_threshold(double c) {
  double t = floor(c * 2147483648.0 + 0.5);
  if (t < 0.0) t = 0.0;
  if (t > 2147483648.0) t = 2147483648.0;
  return (uint32_t)t;
}

/* Categorical draw.
 *  mode 0 (REFERENCE_SCAN): Distributions.jl rand(::Categorical) -- one uniform, inverse-CDF linear
 *      scan in fp64 (call sites src/Inference/likelihood.jl:42, src/CPDs/cpds.jl:97), u = r31 / 2^31.
 *  mode 1 (DEVICE_SPEC): same scan over 31-bit quantised thresholds; this is what the CUDA kernels
 *      compute (state = #{k : r31 >= T_k}), so GPU-vs-oracle is bit-exact in this mode. */
static inline int draw_categorical(const double* p, int card, uint32_t r, int mode) {
  uint32_t r31 = r >> 1;
  if (mode == 0) {
    double u = (double)r31 * (1.0 / 2147483648.0);
    double cp = p[0];
    int i = 0;
    while (cp <= u && i < card - 1) cp += p[++i];
    return i;
  } else {
    double c = 0.0;
    int s = 0;
    for (int k = 0; k < card - 1; ++k) {
      c += p[k];
      if (r31 >= cdf_threshold(c)) s = k + 1; /* thresholds are non-decreasing */
    }
    return s;
  }
}

/* ------------------------------------------------------------------------------------------
 * Likelihood weighting: src/Inference/likelihood.jl:16-58.
 *   per particle: topological sweep (:37); free node -> sample[nn] = rand(d) (:42);
 *   evidence node -> w *= pdf(d, sample[nn]) (:44); phi[q_ind...] += w (:53).
 * Returns UNNORMALISED weighted counts (column-major over the query cards, first query fastest)
 * plus stats[0] = sum w, stats[1] = sum w^2; normalisation (:56) is orc_factor_normalize.
 * Particle ids [first, first+n) index the Philox counter so shards reproduce the 1-process stream.
 * ---------------------------------------------------------------------------------------- */
int orc_lw_infer(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                 const int64_t* cpt_off, const double* cpt, const int8_t* evidence,
                 const int32_t* query, int32_t n_query, uint64_t first, uint64_t n, uint64_t seed,
                 int32_t mode, double* out_counts, double* out_stats) {
  orc_net bn = {n_nodes, card, par_ptr, par_idx, cpt_off, cpt};
  int64_t ncell = 1;
  for (int j = 0; j < n_query; ++j) ncell *= card[query[j]];
  for (int64_t c = 0; c < ncell; ++c) out_counts[c] = 0.0;
  double sw = 0.0, sw2 = 0.0;
#pragma omp parallel
  {
    double* bins = (double*)calloc((size_t)ncell, sizeof(double));
    uint8_t* x = (uint8_t*)malloc((size_t)n_nodes);
    double lsw = 0.0, lsw2 = 0.0;
#pragma omp for schedule(static)
    for (int64_t pi = 0; pi < (int64_t)n; ++pi) {
      orc_stream rs;
      stream_init(&rs, seed, first + (uint64_t)pi, ORC_STREAM_SAMPLE);
      uint64_t d = 0;
      double w = 1.0;
      for (int i = 0; i < n_nodes; ++i) {
        const double* row = cpt_row(&bn, i, x);
        if (evidence && evidence[i] >= 0) {
          x[i] = (uint8_t)evidence[i];
          w *= row[x[i]];
        } else {
          x[i] = (uint8_t)draw_categorical(row, card[i], stream_draw(&rs, d++), mode);
        }
      }
      int64_t cell = 0, stride = 1;
      for (int j = 0; j < n_query; ++j) { cell += stride * x[query[j]]; stride *= card[query[j]]; }
      bins[cell] += w;
      lsw += w; lsw2 += w * w;
    }
#pragma omp critical
    {
      for (int64_t c = 0; c < ncell; ++c) out_counts[c] += bins[c];
      sw += lsw; sw2 += lsw2;
    }
    free(bins); free(x);
  }
  if (out_stats) { out_stats[0] = sw; out_stats[1] = sw2; }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Table-producing samplers: src/sampling.jl.
 *   kind 0 ANCESTRAL  rand!(a, bn, ::DirectSampler) :52-57
 *   kind 1 REJECTION  rand!(a, bn, ::RejectionSampler) :79-87 (<= max_attempts tries, else ok=0)
 *   kind 2 WEIGHTED   get_weighted_sample! :102-115 (evidence clamped, w = prod pdf)
 * out_states is node-major: out_states[i * n + row] (one DataFrame column per node, :30-41).
 * Rejection attempt `a` of row r uses stream id (a << 8) | ORC_STREAM_SAMPLE.
 * ---------------------------------------------------------------------------------------- */
int orc_sample(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
               const int64_t* cpt_off, const double* cpt, const int8_t* evidence, int32_t kind,
               int32_t max_attempts, uint64_t first, uint64_t n, uint64_t seed, int32_t mode,
               uint8_t* out_states, double* out_w, uint8_t* out_ok) {
  orc_net bn = {n_nodes, card, par_ptr, par_idx, cpt_off, cpt};
#pragma omp parallel
  {
    uint8_t* x = (uint8_t*)malloc((size_t)n_nodes);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
      double w = 1.0;
      int ok = 1;
      if (kind == 1) {
        ok = 0;
        for (int a = 0; a < max_attempts && !ok; ++a) {
          orc_stream rs;
          stream_init(&rs, seed, first + (uint64_t)r, ((uint32_t)a << 8) | ORC_STREAM_SAMPLE);
          for (int i = 0; i < n_nodes; ++i)
            x[i] = (uint8_t)draw_categorical(cpt_row(&bn, i, x), card[i], stream_draw(&rs, (uint64_t)i), mode);
          ok = 1; /* consistent(a, evidence): src/ProbabilisticGraphicalModels/assignments.jl:13-22 */
          for (int i = 0; i < n_nodes; ++i)
            if (evidence && evidence[i] >= 0 && x[i] != (uint8_t)evidence[i]) { ok = 0; break; }
        }
      } else {
        orc_stream rs;
        stream_init(&rs, seed, first + (uint64_t)r, ORC_STREAM_SAMPLE);
        uint64_t d = 0;
        for (int i = 0; i < n_nodes; ++i) {
          const double* row = cpt_row(&bn, i, x);
          if (kind == 2 && evidence && evidence[i] >= 0) {
            x[i] = (uint8_t)evidence[i];
            w *= row[x[i]];
          } else {
            x[i] = (uint8_t)draw_categorical(row, card[i], stream_draw(&rs, d++), mode);
          }
        }
      }
      for (int i = 0; i < n_nodes; ++i) out_states[(int64_t)i * (int64_t)n + r] = ok ? x[i] : 0xFF;
      if (out_w) out_w[r] = w;
      if (out_ok) out_ok[r] = (uint8_t)ok;
    }
    free(x);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Gibbs.  Markov-blanket conditional: src/Inference/gibbs.jl:36-50 (eval_mb_cpd):
 *     p[v] = foldl(*, pdf(cpd, x | x[n]=v) for cpd in [children(n)..., n])     (:23, :42-46)
 * children in ascending topological index (Graphs.outneighbors is sorted, src/bayes_nets.jl:96-99).
 * Draw: StatsBase.sample(Weights(p)) (:112): t = rand()*sum(p); first i with cumsum >= t.
 * Device spec: u = (r + 0.5) * 2^-32 so t > 0 whenever sum > 0 (a zero-weight leading state is never
 * picked); an all-zero conditional picks state 0, as the reference's scan does.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t* ptr; int32_t* idx; } orc_children;

static orc_children build_children(const orc_net* bn) {
  orc_children ch;
  ch.ptr = (int32_t*)calloc((size_t)bn->n + 1, sizeof(int32_t));
  for (int i = 0; i < bn->n; ++i)
    for (int k = bn->par_ptr[i]; k < bn->par_ptr[i + 1]; ++k) ch.ptr[bn->par_idx[k] + 1]++;
  for (int i = 0; i < bn->n; ++i) ch.ptr[i + 1] += ch.ptr[i];
  ch.idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ch.ptr[bn->n] > 0 ? ch.ptr[bn->n] : 1));
  int32_t* fill = (int32_t*)calloc((size_t)bn->n, sizeof(int32_t));
  for (int i = 0; i < bn->n; ++i) /* ascending child index */
    for (int k = bn->par_ptr[i]; k < bn->par_ptr[i + 1]; ++k) {
      int p = bn->par_idx[k];
      /* a node may list the same parent only once (names are unique) */
      ch.idx[ch.ptr[p] + fill[p]++] = i;
    }
  free(fill);
  return ch;
}

static inline int gibbs_update(const orc_net* bn, const orc_children* ch, int node, uint8_t* x,
                               uint32_t r, int log_domain) {
  double p[ORC_MAX_CARD];
  int card = bn->card[node];
  uint8_t old = x[node];
  for (int v = 0; v < card; ++v) {
    x[node] = (uint8_t)v;
    if (!log_domain) {
      double prod = 1.0;
      for (int k = ch->ptr[node]; k < ch->ptr[node + 1]; ++k) {
        int c = ch->idx[k];
        prod *= cpt_row(bn, c, x)[x[c]];
      }
      prod *= cpt_row(bn, node, x)[v];
      p[v] = prod;
    } else { /* src/gibbs.jl:60-66: exp(sum logpdf) -- constant CPDs of the blanket cancel */
      double s = 0.0;
      for (int k = ch->ptr[node]; k < ch->ptr[node + 1]; ++k) {
        int c = ch->idx[k];
        s += log(cpt_row(bn, c, x)[x[c]]);
      }
      s += log(cpt_row(bn, node, x)[v]);
      p[v] = exp(s);
    }
  }
  x[node] = old;
  double sum = 0.0;
  for (int v = 0; v < card; ++v) sum += p[v];
  double u = ((double)r + 0.5) * (1.0 / 4294967296.0);
  double t = u * sum;
  int i = 0;
  double cw = p[0];
  while (cw < t && i < card - 1) cw += p[++i];
  return i;
}

/* Init: src/Inference/gibbs.jl:6-20 -- uniform over 1:ncat for non-evidence nodes.
 * Device spec: state = (r * card) >> 32 with draw index = ordinal of the free node. */
static void gibbs_init(const orc_net* bn, const int8_t* evidence, uint64_t seed, uint64_t chain, uint8_t* x) {
  orc_stream rs;
  stream_init(&rs, seed, chain, ORC_STREAM_GIBBS_INIT);
  uint64_t d = 0;
  for (int i = 0; i < bn->n; ++i) {
    if (evidence && evidence[i] >= 0) x[i] = (uint8_t)evidence[i];
    else x[i] = (uint8_t)(((uint64_t)stream_draw(&rs, d++) * (uint64_t)bn->card[i]) >> 32);
  }
}

/* infer(::GibbsSamplingNodewise) src/Inference/gibbs.jl:66-145 (full=0) and
 * infer(::GibbsSamplingFull) :161-234 (full=1), run for `n_chains` independent chains
 * [first_chain, first_chain + n_chains); counts are summed over chains (unnormalised).
 *   total = ceil((nsamples - burn_in)/thin) kept samples per chain (:74)
 *   nodewise: num_iters counts node updates (:120); full: counts sweeps (:214)
 *   sweep order = non-evidence nodes in topological order (:80,:108)
 * state_inout (optional, n_chains x n_nodes, chain-major): if init_from_state != 0 the chains
 * resume from it (im.state, :83-88); final states are written back.
 * per_chain_counts (optional): n_chains x ncell, for between-chain variance. */
int orc_gibbs_infer(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                    const int64_t* cpt_off, const double* cpt, const int8_t* evidence,
                    const int32_t* query, int32_t n_query, uint64_t first_chain, uint64_t n_chains,
                    int64_t nsamples, int64_t burn_in, int64_t thin, int32_t full, uint64_t seed,
                    uint8_t* state_inout, int32_t init_from_state, double* out_counts,
                    double* per_chain_counts) {
  if (!(burn_in < nsamples) || thin < 1) return -1; /* ArgumentError, :70-71 */
  orc_net bn = {n_nodes, card, par_ptr, par_idx, cpt_off, cpt};
  orc_children ch = build_children(&bn);
  int64_t ncell = 1;
  for (int j = 0; j < n_query; ++j) ncell *= card[query[j]];
  for (int64_t c = 0; c < ncell; ++c) out_counts[c] = 0.0;
  int64_t total = (nsamples - burn_in + thin - 1) / thin;
  int n_free = 0;
  int32_t* free_nodes = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_nodes > 0 ? n_nodes : 1));
  for (int i = 0; i < n_nodes; ++i)
    if (!(evidence && evidence[i] >= 0)) free_nodes[n_free++] = i;
#pragma omp parallel
  {
    double* bins = (double*)calloc((size_t)ncell, sizeof(double));
    uint8_t* x = (uint8_t*)malloc((size_t)n_nodes);
#pragma omp for schedule(static)
    for (int64_t ci = 0; ci < (int64_t)n_chains; ++ci) {
      uint64_t chain = first_chain + (uint64_t)ci;
      if (state_inout && init_from_state) memcpy(x, state_inout + ci * n_nodes, (size_t)n_nodes);
      else gibbs_init(&bn, evidence, seed, chain, x);
      orc_stream rs;
      stream_init(&rs, seed, chain, ORC_STREAM_GIBBS_UPDATE);
      uint64_t d = 0;
      int64_t num_iters = 0, num_smpls = 0;
      int finished = (n_free == 0 && !full) ? 1 : 0; /* reference would spin forever; we stop */
      while (!finished) {
        for (int fi = 0; fi < n_free; ++fi) {
          int node = free_nodes[fi];
          x[node] = (uint8_t)gibbs_update(&bn, &ch, node, x, stream_draw(&rs, d++), 0);
          if (!full) {
            num_iters++;
            if (num_iters > burn_in && ((num_iters - burn_in) % thin) == 0) {
              int64_t cell = 0, stride = 1;
              for (int j = 0; j < n_query; ++j) { cell += stride * x[query[j]]; stride *= card[query[j]]; }
              bins[cell] += 1.0;
              if (per_chain_counts) per_chain_counts[ci * ncell + cell] += 1.0;
              num_smpls++;
            }
            if (num_smpls >= total) { finished = 1; break; }
          }
        }
        if (full) {
          num_iters++;
          if (num_iters > burn_in && ((num_iters - burn_in) % thin) == 0) {
            int64_t cell = 0, stride = 1;
            for (int j = 0; j < n_query; ++j) { cell += stride * x[query[j]]; stride *= card[query[j]]; }
            bins[cell] += 1.0;
            if (per_chain_counts) per_chain_counts[ci * ncell + cell] += 1.0;
            num_smpls++;
          }
          if (num_smpls >= total) finished = 1;
        }
      }
      if (state_inout) memcpy(state_inout + ci * n_nodes, x, (size_t)n_nodes);
    }
#pragma omp critical
    for (int64_t c = 0; c < ncell; ++c) out_counts[c] += bins[c];
    free(bins); free(x);
  }
  free(free_nodes); free(ch.ptr); free(ch.idx);
  return 0;
}

/* Fisher-Yates permutation of 0..m-1 for sweep `sweep` (shared by every chain; the reference
 * reshuffles its single chain's order each sweep, src/gibbs.jl:222-234).
 * Draw j of stream (idx = sweep, ORC_STREAM_PERM): k = j + ((r * (m - j)) >> 32). */
void orc_sweep_permutation(uint64_t seed, uint64_t sweep, int32_t m, int32_t* perm) {
  orc_stream rs;
  stream_init(&rs, seed, sweep, ORC_STREAM_PERM);
  for (int j = 0; j < m; ++j) perm[j] = j;
  for (int j = 0; j + 1 < m; ++j) {
    uint32_t r = stream_draw(&rs, (uint64_t)j);
    int k = j + (int)(((uint64_t)r * (uint64_t)(m - j)) >> 32);
    int32_t t = perm[j]; perm[j] = perm[k]; perm[k] = t;
  }
}

/* gibbs_sample: src/gibbs.jl:289-374 with gibbs_sample_main_loop :183-245.
 *   burn_in sweeps (main loop with thinning 0 keeps every sweep, all discarded) then for each of
 *   `nsamples` rows: `thinning` discarded sweeps + 1 kept sweep (:226-240).
 *   Order: variable_order (positions into the free-node list) if given, else a fresh permutation
 *   before every sweep (:222-234; the reference shuffles once more than it sweeps, which only
 *   re-randomises an already uniform permutation).
 *   Initial state (:330-335): 50 likelihood-weighted particles, one importance-resampled
 *   (src/sampling.jl:153-174,182-197), unless init_from_state.
 * out_states is node-major over rows: out_states[i * (n_chains*nsamples) + chain*nsamples + s].
 * Evidence columns hold the clamped value (:368-373). */
int orc_gibbs_sample(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                     const int64_t* cpt_off, const double* cpt, const int8_t* evidence,
                     uint64_t first_chain, uint64_t n_chains, int64_t nsamples, int64_t burn_in,
                     int64_t thinning, const int32_t* variable_order, uint64_t seed,
                     const uint8_t* init_state, int32_t log_domain, uint8_t* out_states) {
  if (nsamples < 1 || burn_in < 0 || thinning < 0) return -1; /* :299-301 */
  orc_net bn = {n_nodes, card, par_ptr, par_idx, cpt_off, cpt};
  orc_children ch = build_children(&bn);
  int n_free = 0;
  int32_t* free_nodes = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_nodes > 0 ? n_nodes : 1));
  for (int i = 0; i < n_nodes; ++i)
    if (!(evidence && evidence[i] >= 0)) free_nodes[n_free++] = i;
  int64_t n_rows = (int64_t)n_chains * nsamples;
  int64_t sweeps_total = burn_in + nsamples * (thinning + 1);
  /* per-sweep orders are chain-independent: precompute */
  int32_t* orders = (int32_t*)malloc(sizeof(int32_t) * (size_t)((sweeps_total > 0 ? sweeps_total : 1) * (n_free > 0 ? n_free : 1)));
  for (int64_t s = 0; s < sweeps_total; ++s) {
    if (variable_order) for (int j = 0; j < n_free; ++j) orders[s * n_free + j] = variable_order[j];
    else orc_sweep_permutation(seed, (uint64_t)s, n_free, orders + s * n_free);
  }
  int status = 0;
#pragma omp parallel
  {
    uint8_t* x = (uint8_t*)malloc((size_t)n_nodes);
    uint8_t* cand = (uint8_t*)malloc((size_t)n_nodes * 50);
    double wts[50];
#pragma omp for schedule(static)
    for (int64_t ci = 0; ci < (int64_t)n_chains; ++ci) {
      uint64_t chain = first_chain + (uint64_t)ci;
      if (init_state) memcpy(x, init_state + ci * n_nodes, (size_t)n_nodes);
      else {
        /* 50 weighted particles, ids chain*50 + j in the sampling stream; resample one */
        double sw = 0.0;
        for (int j = 0; j < 50; ++j) {
          orc_stream rs;
          stream_init(&rs, seed, chain * 50 + (uint64_t)j, ORC_STREAM_SAMPLE);
          uint64_t d = 0;
          double w = 1.0;
          uint8_t* y = cand + (size_t)j * n_nodes;
          for (int i = 0; i < n_nodes; ++i) {
            const double* row = cpt_row(&bn, i, y);
            if (evidence && evidence[i] >= 0) { y[i] = (uint8_t)evidence[i]; w *= row[y[i]]; }
            else y[i] = (uint8_t)draw_categorical(row, card[i], stream_draw(&rs, d++), 1);
          }
          wts[j] = w; sw += w;
        }
        if (!(sw > 0.0)) {
#pragma omp atomic write
          status = -5; /* "unable to find an initial sample with non-zero probability" :332-334 */
          sw = 1.0;
        }
        orc_stream ri;
        stream_init(&ri, seed, chain, ORC_STREAM_GIBBS_INIT);
        double u = ((double)stream_draw(&ri, 0) + 0.5) * (1.0 / 4294967296.0);
        double t = u * sw, c = wts[0];
        int pick = 0;
        while (c < t && pick < 49) c += wts[++pick];
        memcpy(x, cand + (size_t)pick * n_nodes, (size_t)n_nodes);
      }
      orc_stream rs;
      stream_init(&rs, seed, chain, ORC_STREAM_GIBBS_UPDATE);
      uint64_t d = 0;
      int64_t sweep = 0;
      for (int64_t b = 0; b < burn_in; ++b, ++sweep)
        for (int j = 0; j < n_free; ++j) {
          int node = free_nodes[orders[sweep * n_free + j]];
          x[node] = (uint8_t)gibbs_update(&bn, &ch, node, x, stream_draw(&rs, d++), log_domain);
        }
      for (int64_t s = 0; s < nsamples; ++s) {
        for (int64_t k = 0; k <= thinning; ++k, ++sweep)
          for (int j = 0; j < n_free; ++j) {
            int node = free_nodes[orders[sweep * n_free + j]];
            x[node] = (uint8_t)gibbs_update(&bn, &ch, node, x, stream_draw(&rs, d++), log_domain);
          }
        for (int i = 0; i < n_nodes; ++i) out_states[(int64_t)i * n_rows + ci * nsamples + s] = x[i];
      }
    }
    free(x); free(cand);
  }
  free(orders); free(free_nodes); free(ch.ptr); free(ch.idx);
  return status;
}

/* ------------------------------------------------------------------------------------------
 * Factors (src/Factors).  Variables are int ids; data column-major.
 * ---------------------------------------------------------------------------------------- */

/* join(*, phi1, phi2, :outer) = `*`: src/Factors/factors_dims.jl:185-234,269.
 *   - larger factor first (:189-191); common dims must agree (:197-199, returns -2 = DimensionMismatch)
 *   - new_dims = union(phi1.dims, phi2.dims) (:203)
 *   - phi2 is permuted to (common in phi1 order, unique2) and materialised (:213), then broadcast (:233)
 * Call with out_data == NULL to get only the shape (out_nd, out_dims, out_card). */
int orc_factor_product(int32_t nd1, const int32_t* dims1, const int32_t* card1, const double* a,
                       int32_t nd2, const int32_t* dims2, const int32_t* card2, const double* b,
                       int32_t* out_nd, int32_t* out_dims, int32_t* out_card, double* out_data) {
  int64_t n1 = 1, n2 = 1;
  for (int i = 0; i < nd1; ++i) n1 *= card1[i];
  for (int i = 0; i < nd2; ++i) n2 *= card2[i];
  if (n1 < n2) { /* swap */
    const int32_t* t; const double* td; int32_t ti; int64_t tn;
    t = dims1; dims1 = dims2; dims2 = t;
    t = card1; card1 = card2; card2 = t;
    td = a; a = b; b = td;
    ti = nd1; nd1 = nd2; nd2 = ti;
    tn = n1; n1 = n2; n2 = tn;
  }
  if (nd1 > ORC_MAXD || nd2 > ORC_MAXD) return -1;
  /* position of each phi2 dim in phi1 (or -1) */
  int pos2in1[ORC_MAXD];
  int ncommon = 0, nuniq2 = 0;
  int common2[ORC_MAXD], uniq2[ORC_MAXD];
  for (int j = 0; j < nd2; ++j) {
    pos2in1[j] = -1;
    for (int i = 0; i < nd1; ++i) if (dims1[i] == dims2[j]) pos2in1[j] = i;
  }
  /* common, in phi1 order */
  for (int i = 0; i < nd1; ++i)
    for (int j = 0; j < nd2; ++j)
      if (pos2in1[j] == i) {
        if (card1[i] != card2[j]) return -2;
        common2[ncommon++] = j;
      }
  for (int j = 0; j < nd2; ++j) if (pos2in1[j] < 0) uniq2[nuniq2++] = j;
  int nd = nd1 + nuniq2;
  *out_nd = nd;
  for (int i = 0; i < nd1; ++i) { out_dims[i] = dims1[i]; out_card[i] = card1[i]; }
  for (int k = 0; k < nuniq2; ++k) { out_dims[nd1 + k] = dims2[uniq2[k]]; out_card[nd1 + k] = card2[uniq2[k]]; }
  if (!out_data) return 0;

  /* temp = permutedims(phi2, [common2..., uniq2...]) */
  int perm[ORC_MAXD];
  for (int k = 0; k < ncommon; ++k) perm[k] = common2[k];
  for (int k = 0; k < nuniq2; ++k) perm[ncommon + k] = uniq2[k];
  int64_t stride2[ORC_MAXD];
  { int64_t s = 1; for (int j = 0; j < nd2; ++j) { stride2[j] = s; s *= card2[j]; } }
  double* temp = (double*)malloc(sizeof(double) * (size_t)(n2 > 0 ? n2 : 1));
  {
    int digit[ORC_MAXD] = {0};
    int64_t src = 0;
    for (int64_t t = 0; t < n2; ++t) {
      temp[t] = b[src];
      for (int k = 0; k < nd2; ++k) { /* odometer in permuted order */
        int j = perm[k];
        if (++digit[k] < card2[j]) { src += stride2[j]; break; }
        src -= stride2[j] * (card2[j] - 1);
        digit[k] = 0;
      }
    }
  }
  int64_t ncomm = 1;
  for (int k = 0; k < ncommon; ++k) ncomm *= card2[common2[k]];
  int64_t nu2 = n2 / (ncomm > 0 ? ncomm : 1);
  /* stride of each phi1 dim inside temp's common block (0 if not common) */
  int64_t cstride1[ORC_MAXD];
  for (int i = 0; i < nd1; ++i) cstride1[i] = 0;
  { int64_t s = 1; for (int k = 0; k < ncommon; ++k) { cstride1[pos2in1[common2[k]]] = s; s *= card2[common2[k]]; } }
  /* broadcast!(*, new_v, phi1.potential, temp) */
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < nu2; ++u) {
    int digit[ORC_MAXD] = {0};
    int64_t ic = 0;
    const double* trow = temp + u * ncomm;
    double* orow = out_data + u * n1;
    for (int64_t i = 0; i < n1; ++i) {
      orow[i] = a[i] * trow[ic];
      for (int k = 0; k < nd1; ++k) {
        if (++digit[k] < card1[k]) { ic += cstride1[k]; break; }
        ic -= cstride1[k] * (card1[k] - 1);
        digit[k] = 0;
      }
    }
  }
  free(temp);
  return 0;
}

/* sum(phi, dims) = reducedim(+, phi, dims): src/Factors/factors_dims.jl:60-85,100.
 * Accumulates in increasing column-major input order (Base mapreducedim!), drops the dims.
 * Returns -3 if a dim is not in the factor (ArgumentError via _check_dims_valid, auxiliary.jl:7-14). */
int orc_factor_sum(int32_t nd, const int32_t* dims, const int32_t* card, const double* a,
                   int32_t nsum, const int32_t* sum_dims, int32_t* out_nd, int32_t* out_dims,
                   int32_t* out_card, double* out_data) {
  if (nd > ORC_MAXD) return -1;
  int is_sum[ORC_MAXD];
  for (int i = 0; i < nd; ++i) is_sum[i] = 0;
  for (int s = 0; s < nsum; ++s) {
    int found = 0;
    for (int i = 0; i < nd; ++i) if (dims[i] == sum_dims[s]) { is_sum[i] = 1; found = 1; }
    if (!found) return -3;
  }
  int ond = 0;
  int64_t n = 1, no = 1;
  for (int i = 0; i < nd; ++i) {
    n *= card[i];
    if (!is_sum[i]) { out_dims[ond] = dims[i]; out_card[ond] = card[i]; no *= card[i]; ond++; }
  }
  *out_nd = ond;
  if (!out_data) return 0;
  /* per output element, walk the summed digits in column-major order (same add order as a
   * linear sweep over the input accumulating into R[...]) */
  int kept[ORC_MAXD], summed[ORC_MAXD], nk = 0, ns = 0;
  int64_t istride[ORC_MAXD];
  { int64_t s = 1; for (int i = 0; i < nd; ++i) { istride[i] = s; s *= card[i]; if (is_sum[i]) summed[ns++] = i; else kept[nk++] = i; } }
  int64_t nsumel = n / (no > 0 ? no : 1);
#pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < no; ++o) {
    int64_t base = 0, r = o;
    for (int k = 0; k < nk; ++k) { int i = kept[k]; base += (r % card[i]) * istride[i]; r /= card[i]; }
    int digit[ORC_MAXD] = {0};
    int64_t off = 0;
    double acc = 0.0;
    for (int64_t t = 0; t < nsumel; ++t) {
      acc = (t == 0) ? a[base + off] : acc + a[base + off];
      for (int k = 0; k < ns; ++k) {
        int i = summed[k];
        if (++digit[k] < card[i]) { off += istride[i]; break; }
        off -= istride[i] * (card[i] - 1);
        digit[k] = 0;
      }
    }
    out_data[o] = acc;
  }
  return 0;
}

/* Base.sum pairwise (block 1024), as Julia's mapreduce_impl. */
static double pairwise_sum(const double* x, int64_t lo, int64_t hi, int p) {
  if (hi - lo <= 1024) {
    double s = 0.0;
    for (int64_t i = lo; i < hi; ++i) s += (p == 1) ? fabs(x[i]) : x[i] * x[i];
    return s;
  }
  int64_t mid = lo + (hi - lo) / 2;
  return pairwise_sum(x, lo, mid, p) + pairwise_sum(x, mid, hi, p);
}

/* normalize!(phi; p): src/Factors/factors_dims.jl:45-57.  p=1: sum(abs); p=2: sum(abs2) (no sqrt --
 * test/test_factors.jl:79-81 pins [1/30 2/30; 1/10 4/30]).  Other p -> -1 (ArgumentError :51). */
int orc_factor_normalize(int64_t n, double* x, int32_t p) {
  if (p != 1 && p != 2) return -1;
  double total = pairwise_sum(x, 0, n, p);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] /= total;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * statistics(parent_list, ncategories, data): src/DiscreteBayesNet/discrete_bayes_net.jl:151-172 (one node:
 * js = (data[parents,:] .- 1)' * stridevec .+ 1, stridevec[k] = prod of the earlier parents' ncategories;
 * N = sparse(data[target,:], js, 1, r, q) -- duplicates add) and :221-231 (all nodes).
 * data: uint8, node-major data[i * pitch + row], 0-based; rows whose first byte is 0xFF are skipped (the device
 * sampler's "rejection gave up" marker).  out: cpt_off[n] int64 in the flat CPT layout = each N_i column-major.
 * OpenMP over nodes (each node's table is private to one thread: no atomics, deterministic).
 * ---------------------------------------------------------------------------------------- */
int orc_statistics(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                   const int64_t* cpt_off, const uint8_t* data, uint64_t n_rows, uint64_t pitch, int64_t* out) {
  for (int64_t b = 0; b < cpt_off[n_nodes]; ++b) out[b] = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n_nodes; ++i) {
    int64_t* N = out + cpt_off[i];
    const uint8_t* xi = data + (uint64_t)i * pitch;
    for (uint64_t r = 0; r < n_rows; ++r) {
      if (data[r] == 0xFF) continue;
      int64_t j = 0, stride = 1;
      for (int k = par_ptr[i]; k < par_ptr[i + 1]; ++k) {
        const int p = par_idx[k];
        j += stride * data[(uint64_t)p * pitch + r];
        stride *= card[p];
      }
      N[xi[r] + (int64_t)card[i] * j] += 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * logpdf(bn, df::DataFrame): src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-99 -- `logl = 0.0`,
 * then for each row in order `logl += logpdf(bn, a)`, where logpdf(bn, a) (src/bayes_nets.jl:322-328) starts at 0.0
 * and adds logpdf(cpd, a) = log(cpd(a).p[a[name]]) for the cpds in topological order (src/CPDs/cpds.jl:110,
 * CPT row = src/CPDs/categorical_cpd.jl:36-46).  Same order of additions here, single thread; rows whose first byte
 * is 0xFF are skipped as in orc_statistics.  pdf(bn, df) = exp(logpdf(bn, df)) (:104).
 * ---------------------------------------------------------------------------------------- */
int orc_logpdf_table(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                     const int64_t* cpt_off, const double* cpt, const uint8_t* data, uint64_t n_rows, uint64_t pitch,
                     double* out) {
  double logl = 0.0;
  for (uint64_t r = 0; r < n_rows; ++r) {
    if (data[r] == 0xFF) continue;
    double row = 0.0;
    for (int i = 0; i < n_nodes; ++i) {
      int64_t j = 0, stride = 1;
      for (int k = par_ptr[i]; k < par_ptr[i + 1]; ++k) {
        const int p = par_idx[k];
        j += stride * data[(uint64_t)p * pitch + r];
        stride *= card[p];
      }
      row += log(cpt[cpt_off[i] + data[(uint64_t)i * pitch + r] + (int64_t)card[i] * j]);
    }
    logl += row;
  }
  *out = logl;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Loopy belief propagation: infer(::LoopyBelief), src/Inference/loopy_belief.jl:24-218, one (query, evidence)
 * instance after the other (OpenMP over instances).  Second restatement beside oracle/loopy_np.py: the numpy one keeps
 * the reference's array aliasing literally; this one states its CONSEQUENCES explicitly --
 *   - a parentless node's pi is its factor's potential, rewritten in place per child, every iteration (:125,:170-175)
 *     -> R[] persists across iterations;
 *   - all children of a node receive the array left after the last child's update (:177) -> one pi message per node;
 *   - from iteration 2 on both message dictionaries hold a root's array: its children read the message of the current
 *     iteration (nodes run in topological order, roots first) and its own change is norm(px - px, Inf)
 * -- and tests/test_oracle_golden.py requires the two to agree bit for bit.  Arithmetic order as the reference's calls:
 * n-ary broadcast `*` is a left fold (factors_dims.jl:131-165), sum! accumulates column-major (:87-98), normalize!(phi)
 * divides by sum(abs) (:45-57), normalize!(v, 1) multiplies by inv(norm1) (LinearAlgebra.__normalize!), norm(v) and
 * norm(v, Inf) are generic_norm2 / generic_normInf.
 * ---------------------------------------------------------------------------------------- */
static double lbp_norm2_diff(const double* a, const double* b, int n) {
  double maxabs = 0.0;
  int nan = 0;
  for (int i = 0; i < n; ++i) {
    double d = fabs(a[i] - b[i]);
    if (d != d) nan = 1; else if (d > maxabs) maxabs = d;
  }
  if (nan) return NAN;
  if (maxabs == 0.0 || isinf(maxabs)) return maxabs;
  double mm = maxabs * maxabs, s = 0.0;
  if (isfinite((double)n * mm) && mm != 0.0) {
    for (int i = 0; i < n; ++i) { double d = a[i] - b[i]; s = s + d * d; }
    return sqrt(s);
  }
  for (int i = 0; i < n; ++i) { double r = fabs(a[i] - b[i]) / maxabs; s = s + r * r; }
  return maxabs * sqrt(s);
}

int orc_loopy_infer(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                    const int64_t* cpt_off, const double* cpt, const int8_t* evidence, const int32_t* query, int64_t n_inst,
                    int32_t nsamples, double tol, int32_t iters_for_convergence, int32_t out_stride, double* out,
                    int32_t* out_iters) {
  if (iters_for_convergence < 1 || nsamples < 0) return -1;
  orc_net bn = {n_nodes, card, par_ptr, par_idx, cpt_off, cpt};
  orc_children ch = build_children(&bn);
  const int n = n_nodes, n_edges = par_ptr[n];
  int64_t S_c = 0, S_l = 0;
  int64_t* poff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t* loff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_edges + 1));
  int max_card = 1;
  for (int i = 0; i < n; ++i) { poff[i] = S_c; S_c += card[i]; if (card[i] > max_card) max_card = card[i]; }
  for (int e = 0; e < n_edges; ++e) { loff[e] = S_l; S_l += card[par_idx[e]]; }
  /* edge (child -> p) of the k-th child of p */
  int32_t* cedge = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_edges + 1));
  for (int p = 0; p < n; ++p)
    for (int kk = ch.ptr[p]; kk < ch.ptr[p + 1]; ++kk) {
      int c = ch.idx[kk];
      for (int e = par_ptr[c]; e < par_ptr[c + 1]; ++e) if (par_idx[e] == p) cedge[kk] = e;
    }
  const double DELTA = 1.0 / DBL_MAX, EPS_DELTA = DBL_EPSILON / DELTA;
#pragma omp parallel
  {
    double* PI = (double*)malloc(sizeof(double) * (size_t)(2 * S_c + 1));
    double* R = (double*)malloc(sizeof(double) * (size_t)(S_c + 1));
    double* LAM = (double*)malloc(sizeof(double) * (size_t)(2 * S_l + 1));
    double* lam = (double*)malloc(sizeof(double) * (size_t)max_card);
    double* lx = (double*)malloc(sizeof(double) * (size_t)max_card * (size_t)max_card);
    int* dig = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < n_inst; ++b) {
      const int8_t* ev = evidence ? evidence + b * n : NULL;
#define EV(i) (ev ? (int)ev[i] : -1)
      for (int i = 0; i < n; ++i)
        for (int x = 0; x < card[i]; ++x) {
          double v = EV(i) >= 0 ? (x == EV(i) ? 1.0 : 0.0) : 1.0 / (double)card[i];   /* :69-85 */
          PI[poff[i] + x] = PI[S_c + poff[i] + x] = v;
          R[poff[i] + x] = par_ptr[i + 1] == par_ptr[i] ? cpt[cpt_off[i] + x] : 0.0;
        }
      for (int e = 0; e < n_edges; ++e)
        for (int x = 0; x < card[par_idx[e]]; ++x) LAM[loff[e] + x] = LAM[S_l + loff[e] + x] = 1.0 / (double)card[par_idx[e]];
      int consec = (INFINITY <= tol) ? iters_for_convergence : 0, iters = 0;
      for (int t = 1; t <= nsamples; ++t) {
        iters = t;
        double* PIr = PI + ((t - 1) & 1) * S_c; double* PIw = PI + (t & 1) * S_c;
        const double* LAMr = LAM + ((t - 1) & 1) * S_l; double* LAMw = LAM + (t & 1) * S_l;
        double mc = -INFINITY;
        for (int i = 0; i < n; ++i) {                                                   /* :100 */
          const int c = card[i], k = par_ptr[i + 1] - par_ptr[i], m = ch.ptr[i + 1] - ch.ptr[i];
          const double* cp = cpt + cpt_off[i];
          int64_t Q = (cpt_off[i + 1] - cpt_off[i]) / c;
          /* lambda: evidence one-hot, ones, or the fold of the children's messages (:114-121) */
          for (int x = 0; x < c; ++x) {
            if (EV(i) >= 0) lam[x] = x == EV(i) ? 1.0 : 0.0;
            else if (m == 0) lam[x] = 1.0;
            else {
              double l = LAMr[loff[cedge[ch.ptr[i]]] + x];
              for (int k2 = 1; k2 < m; ++k2) l = l * LAMr[loff[cedge[ch.ptr[i] + k2]] + x];
              lam[x] = l;
            }
          }
          /* lambda messages to the parents (:136-160) */
          for (int j = 0; j < k; ++j) {
            const int e = par_ptr[i] + j, pc = card[par_idx[e]];
            for (int z = 0; z < c * pc; ++z) lx[z] = 0.0;
            for (int jj = 0; jj < k; ++jj) dig[jj] = 0;
            for (int64_t q = 0; q < Q; ++q) {                 /* ascending q = column-major over the parents */
              for (int x = 0; x < c; ++x) {
                double tv = cp[x + c * q];
                for (int jj = 0; jj < k; ++jj)
                  if (jj != j) tv = tv * PIr[poff[par_idx[par_ptr[i] + jj]] + dig[jj]];
                double* slot = lx + x + c * dig[j];
                *slot = *slot + tv;                            /* 0 + tv == tv: same as starting the sum at the first term */
              }
              for (int jj = 0; jj < k; ++jj) { if (++dig[jj] < card[par_idx[par_ptr[i] + jj]]) break; dig[jj] = 0; }
            }
            double total = 0.0;
            for (int u = 0; u < pc; ++u) {
              double msg = lx[c * u] * lam[0];
              for (int x = 1; x < c; ++x) msg = msg + lx[x + c * u] * lam[x];
              LAMw[loff[e] + u] = msg;
              total = u ? total + fabs(msg) : fabs(msg);
            }
            for (int u = 0; u < pc; ++u) LAMw[loff[e] + u] = LAMw[loff[e] + u] / total;
            double delta = lbp_norm2_diff(LAMw + loff[e], LAMr + loff[e], pc);
            if (delta > mc) mc = delta;
          }
          /* pi messages to the children (:163-185) */
          if (EV(i) >= 0 || m == 0) continue;
          double* px;
          if (k == 0) px = R + poff[i];
          else {
            px = PIw + poff[i];
            for (int x = 0; x < c; ++x) {
              double acc = 0.0;
              for (int jj = 0; jj < k; ++jj) dig[jj] = 0;
              for (int64_t q = 0; q < Q; ++q) {
                double tv = cp[x + c * q];
                for (int jj = 0; jj < k; ++jj) tv = tv * PIr[poff[par_idx[par_ptr[i] + jj]] + dig[jj]];
                acc = q ? acc + tv : tv;
                for (int jj = 0; jj < k; ++jj) { if (++dig[jj] < card[par_idx[par_ptr[i] + jj]]) break; dig[jj] = 0; }
              }
              px[x] = acc;
            }
          }
          for (int kk = 0; kk < m; ++kk) {
            double nrm = 0.0;
            for (int x = 0; x < c; ++x) {
              double v = px[x];
              if (m > 1) {
                double p = 0.0; int first = 1;
                for (int k2 = 0; k2 < m; ++k2) {
                  if (k2 == kk) continue;
                  double l = LAMr[loff[cedge[ch.ptr[i] + k2]] + x];
                  p = first ? l : p * l; first = 0;
                }
                v = v * p; px[x] = v;
              }
              nrm = x ? nrm + fabs(v) : fabs(v);
            }
            int nan = 0; double dm = 0.0;
            for (int x = 0; x < c; ++x) {
              double v = px[x];
              if (nrm >= DELTA) v = v * (1.0 / nrm);
              else { v = v * EPS_DELTA; v = v * (1.0 / (nrm * EPS_DELTA)); }
              px[x] = v;
              double old = k == 0 ? (t == 1 ? 1.0 / (double)c : v) : PIr[poff[i] + x];
              double d = fabs(v - old);
              if (d != d) nan = 1; else if (d > dm) dm = d;
            }
            if (!nan && dm > mc) mc = dm;
          }
          if (k == 0)
            for (int x = 0; x < c; ++x) { PIw[poff[i] + x] = px[x]; if (t >= 2) PIr[poff[i] + x] = px[x]; }
        }
        consec = (mc <= tol) ? (consec + 1 < iters_for_convergence ? consec + 1 : iters_for_convergence) : 0;
        if (consec >= iters_for_convergence) break;
      }
      /* belief of the query (:201-217) */
      {
        const int qn = query[b], c = card[qn], k = par_ptr[qn + 1] - par_ptr[qn], m = ch.ptr[qn + 1] - ch.ptr[qn];
        const double* PIr = PI + (iters & 1) * S_c; const double* LAMr = LAM + (iters & 1) * S_l;
        const double* cp = cpt + cpt_off[qn];
        int64_t Q = (cpt_off[qn + 1] - cpt_off[qn]) / c;
        double* o = out + b * out_stride; double total = 0.0;
        for (int x = 0; x < c; ++x) {
          double l = 1.0;
          if (m) {
            l = LAMr[loff[cedge[ch.ptr[qn]]] + x];
            for (int k2 = 1; k2 < m; ++k2) l = l * LAMr[loff[cedge[ch.ptr[qn] + k2]] + x];
          }
          double pi;
          if (k == 0) pi = R[poff[qn] + x];
          else {
            pi = 0.0;
            for (int jj = 0; jj < k; ++jj) dig[jj] = 0;
            for (int64_t q = 0; q < Q; ++q) {
              double tv = cp[x + c * q];
              for (int jj = 0; jj < k; ++jj) tv = tv * PIr[poff[par_idx[par_ptr[qn] + jj]] + dig[jj]];
              pi = q ? pi + tv : tv;
              for (int jj = 0; jj < k; ++jj) { if (++dig[jj] < card[par_idx[par_ptr[qn] + jj]]) break; dig[jj] = 0; }
            }
          }
          o[x] = l * pi;
          total = x ? total + fabs(o[x]) : fabs(o[x]);
        }
        for (int x = 0; x < c; ++x) o[x] = o[x] / total;
        for (int x = c; x < out_stride; ++x) o[x] = 0.0;
        if (out_iters) out_iters[b] = iters;
      }
#undef EV
    }
    free(PI); free(R); free(LAM); free(lam); free(lx); free(dig);
  }
  free(poff); free(loff); free(cedge); free(ch.ptr); free(ch.idx);
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}
