// sampling.cu -- particle kernels: likelihood weighting (infer), ancestral / weighted / rejection tables.
//
// Replaces the scalar loops of src/Inference/likelihood.jl:34-54 and src/sampling.jl:24-115.
//
// Design (B200): one particle per thread, persistent grid (SMs x resident CTAs).  Per call the host
// "compiles" (net, evidence) into a small program + integer CDF tables:
//   * evidence parents are folded into each table's base offset, so only FREE parents are looked up
//   * free node  -> rows of 31-bit CDF thresholds T_k = round(cdf_k * 2^31); state = #{k : r31 >= T_k}
//   * evidence node -> rows of fp64 P(e | free parents); w *= row   (fp64: bit-exact vs the oracle)
// Program and tables are staged into shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier);
// particle states live in shared memory as u8 [node][thread] (conflict-free: 32 lanes = 8 words).
// Weighted counts: fp64 shared-memory bins per CTA, then one global fp64 atomic per (CTA, cell).
// Not HBM-bound: tables are smem-resident; the bound is issue slots (Philox = 4 ops / round).
#include "common.cuh"

namespace bn {

// program record header: kind | card << 8 | n_free_parents << 16
enum : uint32_t { K_FREE = 0, K_EVID = 1 };
constexpr uint32_t THRESH_PAD = 0x80000000u;  // r31 < 2^31 always

static inline uint32_t row_words_free(int card) {
  if (card <= 1) return 1;
  if (card == 2) return 1;
  if (card == 3) return 2;
  return (uint32_t)((card - 1 + 3) / 4 * 4);
}

// Build program + tables.  `treat_all_free`: ancestral / rejection sampling ignore the evidence while
// sampling (src/sampling.jl:52-57,79-87).  state_stride = threads per CTA (row pitch of the u8 states).
int32_t compile_program(const Net& net, const int8_t* evidence, bool treat_all_free,
                        uint32_t state_stride, SampleProgram* out) {
  const int n = net.n;
  std::vector<uint32_t> prog, tab;
  prog.reserve((size_t)n * 8);
  out->plan.clear();
  uint32_t n_free = 0;
  for (int i = 0; i < n; ++i) {
    const bool is_ev = !treat_all_free && evidence && evidence[i] >= 0;
    const int card = net.card[i];
    if (evidence && evidence[i] >= 0)
      BN_REQUIRE(evidence[i] < card, BN_BOUNDS, "evidence state %d out of range for node %d (card %d)",
                 (int)evidence[i], i, card);
    // split parents into evidence (folded) and free (looked up)
    int64_t base_q = 0, stride = 1;
    std::vector<std::pair<int, int64_t>> fp;  // (parent, stride in CPT rows)
    for (int k = net.par_ptr[i]; k < net.par_ptr[i + 1]; ++k) {
      int p = net.par_idx[k];
      bool p_ev = !treat_all_free && evidence && evidence[p] >= 0;
      if (p_ev) base_q += stride * evidence[p];
      else fp.emplace_back(p, stride);
      stride *= net.card[p];
    }
    // enumerate free-parent configurations (first free parent fastest)
    int64_t qf = 1;
    for (auto& pr : fp) qf *= net.card[pr.first];
    const uint32_t rw = is_ev ? 2u : row_words_free(card);
    BN_REQUIRE((int64_t)tab.size() + qf * rw + 8 < (int64_t)0x3fffffff, BN_UNSUPPORTED,
               "node %d: sampling table too large", i);
    // align the node's table to its row width (16 B for uint4 rows, 8 B for uint2 / double rows)
    uint32_t align = rw >= 4 ? 4u : rw;
    while (tab.size() % align) tab.push_back(THRESH_PAD);
    const uint32_t toff = (uint32_t)tab.size();
    std::vector<int> digit(fp.size(), 0);
    const double* cpt = net.cpt.data() + net.cpt_off[i];
    for (int64_t q = 0; q < qf; ++q) {
      int64_t row = base_q;
      for (size_t k = 0; k < fp.size(); ++k) row += fp[k].second * digit[k];
      const double* pr = cpt + row * card;
      if (is_ev) {
        double v = pr[evidence[i]];
        uint64_t bits;
        memcpy(&bits, &v, 8);
        tab.push_back((uint32_t)bits);
        tab.push_back((uint32_t)(bits >> 32));
      } else {
        double c = 0.0;
        uint32_t wrote = 0;
        for (int k = 0; k < card - 1; ++k) {
          c += pr[k];
          tab.push_back(cdf_threshold(c));
          wrote++;
        }
        while (wrote < rw) { tab.push_back(THRESH_PAD); wrote++; }
      }
      for (size_t k = 0; k < fp.size(); ++k) {
        if (++digit[k] < net.card[fp[k].first]) break;
        digit[k] = 0;
      }
    }
    BN_REQUIRE(fp.size() < 65536, BN_UNSUPPORTED, "node %d has too many parents", i);
    {
      NodePlan np;
      np.is_evidence = is_ev; np.card = card; np.toff = toff; np.row_words = rw;
      int64_t fs = 1;
      for (auto& pr : fp) { np.parents.emplace_back(pr.first, (uint32_t)fs); fs *= net.card[pr.first]; }
      out->plan.push_back(std::move(np));
    }
    prog.push_back((is_ev ? K_EVID : K_FREE) | ((uint32_t)card << 8) | ((uint32_t)fp.size() << 16));
    prog.push_back(toff);
    int64_t fstride = 1;
    for (auto& pr : fp) {
      prog.push_back((uint32_t)pr.first * state_stride);
      prog.push_back((uint32_t)(fstride * rw));
      fstride *= net.card[pr.first];
    }
    if (!is_ev) n_free++;
  }
  while (prog.size() % 4) prog.push_back(0);
  while (tab.size() % 4) tab.push_back(THRESH_PAD);
  out->prog_words = (uint32_t)prog.size();
  out->table_words = (uint32_t)tab.size();
  out->n_free = n_free;
  out->words = std::move(prog);
  out->words.insert(out->words.end(), tab.begin(), tab.end());
  return BN_OK;
}

// ------------------------------------------------------------------------------------------ device
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One-shot TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    // chunks of <= 64 KB keep each bulk request modest
    uint32_t done = 0;
    while (done < bytes) {
      uint32_t chunk = bytes - done < 65536u ? bytes - done : 65536u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32((const char*)dst_smem + done)),
          "l"((const char*)src_gmem + done), "r"(chunk), "r"(b)
          : "memory");
      done += chunk;
    }
  }
  // everyone waits for phase 0
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(b)
        : "memory");
  }
}

struct SampleArgs {
  const uint32_t* words;  // device copy of SampleProgram::words
  uint32_t prog_words, table_words;
  uint32_t n_nodes;
  uint32_t tables_in_smem;
  uint64_t first, n;  // particle / row range
  uint32_t seed_lo, seed_hi;
  // MODE_INFER
  const uint32_t* query;  // [n_query] pairs (node * state_stride, cell stride)
  uint32_t n_query, ncell;
  double* out;  // [ncell + 2]
  // table modes
  uint8_t* out_states;  // [n_nodes x n]
  double* out_w;
  uint8_t* out_ok;
  const int8_t* evidence;  // device copy, for rejection checks / clamped columns
  uint32_t max_attempts;
};

enum { MODE_INFER = 0, MODE_TABLE = 1, MODE_REJECT = 2 };

template <int MODE>
__global__ void __launch_bounds__(1024, 1) particle_kernel(const SampleArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t T = blockDim.x;
  const uint32_t tid = threadIdx.x;
  // layout: [prog | tables?] [bins fp64 (ncell+2)] [states u8 n_nodes x T]
  uint32_t* prog = reinterpret_cast<uint32_t*>(smem_raw);
  const uint32_t staged_words = a.prog_words + (a.tables_in_smem ? a.table_words : 0u);
  const uint32_t* tab = a.tables_in_smem ? prog + a.prog_words : a.words + a.prog_words;
  double* bins = reinterpret_cast<double*>(prog + staged_words);
  const uint32_t nb = (MODE == MODE_INFER) ? a.ncell + 2 : 0;
  uint8_t* st = reinterpret_cast<uint8_t*>(bins + nb) + tid;

  tma_bulk_load(prog, a.words, staged_words * 4u, &bar);
  if (MODE == MODE_INFER)
    for (uint32_t c = tid; c < nb; c += T) bins[c] = 0.0;
  __syncthreads();

  const uint64_t stride = (uint64_t)gridDim.x * T;
  double sw = 0.0, sw2 = 0.0;
  for (uint64_t row = (uint64_t)blockIdx.x * T + tid; row < a.n; row += stride) {
    const uint64_t pid = a.first + row;
    double w = 1.0;
    uint32_t ok = 1;
    const uint32_t attempts = (MODE == MODE_REJECT) ? a.max_attempts : 1u;
    for (uint32_t att = 0; att < attempts; ++att) {
      const uint32_t* pc = prog;
      uint32_t d = 0;
      uint32_t r[4];
      w = 1.0;
      for (uint32_t i = 0; i < a.n_nodes; ++i) {
        const uint32_t hdr = pc[0];
        uint32_t off = pc[1];
        const uint32_t nfp = hdr >> 16;
        const uint32_t card = (hdr >> 8) & 0xffu;
        for (uint32_t k = 0; k < nfp; ++k) off += pc[3 + 2 * k] * (uint32_t)st[pc[2 + 2 * k]];
        pc += 2 + 2 * nfp;
        if ((hdr & 3u) == K_FREE) {
          if ((d & 3u) == 0u)
            philox4x32_10((uint32_t)pid, (uint32_t)(pid >> 32), d >> 2, (att << 8) | STREAM_SAMPLE, a.seed_lo,
                          a.seed_hi, r);
          const uint32_t rr = (d & 2u) ? ((d & 1u) ? r[3] : r[2]) : ((d & 1u) ? r[1] : r[0]);
          const uint32_t r31 = rr >> 1;
          ++d;
          uint32_t s;
          if (card == 2) {
            s = r31 >= tab[off];
          } else if (card == 3) {
            const uint2 t = *reinterpret_cast<const uint2*>(tab + off);
            s = (r31 >= t.x) + (r31 >= t.y);
          } else {
            s = 0;
            for (uint32_t k = 0; k + 1 < card; k += 4) {
              const uint4 t = *reinterpret_cast<const uint4*>(tab + off + k);
              s += (r31 >= t.x) + (r31 >= t.y) + (r31 >= t.z) + (r31 >= t.w);
            }
          }
          st[i * T] = (uint8_t)s;
        } else {
          w *= *reinterpret_cast<const double*>(tab + off);
        }
      }
      if (MODE == MODE_REJECT) {
        // consistent(a, evidence): src/ProbabilisticGraphicalModels/assignments.jl:13-22
        ok = 1;
        for (uint32_t i = 0; i < a.n_nodes; ++i) {
          const int e = a.evidence[i];
          if (e >= 0 && (uint32_t)e != st[i * T]) { ok = 0; break; }
        }
        if (ok) break;
      }
    }
    if (MODE == MODE_INFER) {
      uint32_t cell = 0;
      for (uint32_t j = 0; j < a.n_query; ++j) cell += a.query[2 * j + 1] * (uint32_t)st[a.query[2 * j]];
      atomicAdd(&bins[cell], w);
      sw += w;
      sw2 += w * w;
    } else {
      for (uint32_t i = 0; i < a.n_nodes; ++i) {
        uint8_t s = st[i * T];
        if (MODE == MODE_TABLE && a.evidence && a.evidence[i] >= 0) s = (uint8_t)a.evidence[i];
        if (MODE == MODE_REJECT && !ok) s = 0xFF;
        a.out_states[(uint64_t)i * a.n + row] = s;
      }
      if (a.out_w) a.out_w[row] = w;
      if (a.out_ok) a.out_ok[row] = (uint8_t)ok;
    }
  }
  if (MODE == MODE_INFER) {
    // warp-reduce the two scalar sums, then one smem atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      sw2 += __shfl_xor_sync(0xffffffffu, sw2, o);
    }
    if ((tid & 31u) == 0) {
      atomicAdd(&bins[a.ncell], sw);
      atomicAdd(&bins[a.ncell + 1], sw2);
    }
    __syncthreads();
    for (uint32_t c = tid; c < nb; c += T)
      if (bins[c] != 0.0) atomicAdd(&a.out[c], bins[c]);
  }
}

// ------------------------------------------------------------------------------------------ host
struct LaunchShape {
  int threads = 0, ctas_per_sm = 0, grid = 0;
  size_t smem = 0;
  bool tables_in_smem = true;
};

// Pick threads/CTA and CTAs/SM maximising resident threads under the 227 KB/CTA and 228 KB/SM limits.
static bool pick_shape(const DevProps& dp, uint32_t n_nodes, size_t fixed_bytes, size_t table_bytes,
                       size_t bins_bytes, uint64_t n_rows, LaunchShape* out) {
  static const int cand[] = {1024, 768, 512, 384, 256, 192, 128, 64, 32};
  LaunchShape best;
  long best_resident = 0;
  for (int pass = 0; pass < 2 && best_resident == 0; ++pass) {
    const bool in_smem = pass == 0;
    for (int t : cand) {
      size_t need = fixed_bytes + (in_smem ? table_bytes : 0) + bins_bytes + (size_t)n_nodes * t + 16;
      if (need > (size_t)dp.max_smem_optin) continue;
      int per_sm = (int)((size_t)dp.smem_per_sm / (need + 1024));  // 1 KB reserved per CTA
      if (per_sm < 1) per_sm = 1;
      if (per_sm * t > 2048) per_sm = 2048 / t;
      if (per_sm > 32) per_sm = 32;
      long resident = (long)per_sm * t;
      if (resident > best_resident) {
        best_resident = resident;
        best.threads = t; best.ctas_per_sm = per_sm; best.smem = need; best.tables_in_smem = in_smem;
      }
    }
  }
  if (best_resident == 0) return false;
  long want_ctas = (long)((n_rows + best.threads - 1) / best.threads);
  long full = (long)dp.sm_count * best.ctas_per_sm;
  best.grid = (int)(want_ctas < full ? (want_ctas > 0 ? want_ctas : 1) : full);
  *out = best;
  return true;
}

template <int MODE>
static int32_t launch_particles(Net* net, const int8_t* evidence, bool all_free, SampleArgs a, uint32_t n_query,
                                const int32_t* query, const std::vector<int64_t>& qstride) {
  const DevProps& dp = dev_props(net->device);
  cudaStream_t s = tl_stream;
  // first pass: table size does not depend on the stride, so compile once with stride 1 to size things
  SampleProgram probe;
  BN_TRY(compile_program(*net, evidence, all_free, 1, &probe));
  LaunchShape shp;
  const size_t bins_bytes = (MODE == MODE_INFER) ? (size_t)(a.ncell + 2) * 8 : 0;
  BN_REQUIRE(pick_shape(dp, (uint32_t)net->n, (size_t)probe.prog_words * 4, (size_t)probe.table_words * 4,
                        bins_bytes, a.n, &shp),
             BN_UNSUPPORTED, "network too large for the shared-memory particle kernel (%d nodes)", net->n);
  SampleProgram prg;
  BN_TRY(compile_program(*net, evidence, all_free, (uint32_t)shp.threads, &prg));

  // per-call device staging from the stream-ordered pool: [program words | query pairs | evidence]; freed (in stream
  // order, after the kernel) when `stage` goes out of scope, so concurrent calls on one net never share it
  const size_t words_bytes = prg.words.size() * 4;
  const size_t q_off = (words_bytes + 15) / 16 * 16;
  const size_t ev_off = q_off + ((size_t)n_query * 8 + 15) / 16 * 16;
  const size_t total = ev_off + (size_t)net->n + 16;
  TmpBuf<unsigned char> stage;
  BN_CUDA_TRY(stage.alloc(total));
  std::vector<unsigned char> host(total, 0);
  memcpy(host.data(), prg.words.data(), words_bytes);
  uint32_t* hq = reinterpret_cast<uint32_t*>(host.data() + q_off);
  for (uint32_t j = 0; j < n_query; ++j) {
    hq[2 * j] = (uint32_t)query[j] * (uint32_t)shp.threads;
    hq[2 * j + 1] = (uint32_t)qstride[j];
  }
  int8_t* hev = reinterpret_cast<int8_t*>(host.data() + ev_off);
  for (int i = 0; i < net->n; ++i) hev[i] = evidence ? evidence[i] : (int8_t)-1;
  BN_CUDA_TRY(cudaMemcpyAsync(stage.p, host.data(), total, cudaMemcpyHostToDevice, s));

  a.words = reinterpret_cast<const uint32_t*>(stage.p);
  a.prog_words = prg.prog_words;
  a.table_words = prg.table_words;
  a.n_nodes = (uint32_t)net->n;
  a.tables_in_smem = shp.tables_in_smem ? 1u : 0u;
  a.query = reinterpret_cast<const uint32_t*>(stage.p + q_off);
  a.n_query = n_query;
  a.evidence = reinterpret_cast<const int8_t*>(stage.p + ev_off);

  BN_CUDA_TRY(cudaFuncSetAttribute(particle_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)shp.smem));
  particle_kernel<MODE><<<shp.grid, shp.threads, shp.smem, s>>>(a);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

// evidence states must index the node's CPT: a state >= card would read out of bounds on the device
int32_t check_evidence(const Net* net, const int8_t* evidence) {
  if (!evidence) return BN_OK;
  for (int i = 0; i < net->n; ++i)
    BN_REQUIRE(evidence[i] < net->card[i], BN_BOUNDS, "evidence state %d of node %d outside 0..%d", (int)evidence[i], i,
               net->card[i] - 1);
  return BN_OK;
}

static int32_t check_query(const Net* net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                           std::vector<int64_t>* qstride, int64_t* ncell) {
  BN_REQUIRE(n_query >= 0 && (n_query == 0 || query), BN_BAD_ARG, "query is NULL");
  int64_t cells = 1;
  qstride->clear();
  for (int j = 0; j < n_query; ++j) {
    // src/ProbabilisticGraphicalModels/inference.jl:10-11
    BN_REQUIRE(query[j] >= 0 && query[j] < net->n, BN_BAD_ARG, "Query %d is not in the probabilistic graphical model",
               query[j]);
    BN_REQUIRE(!(evidence && evidence[query[j]] >= 0), BN_BAD_ARG, "Query %d is part of the evidence", query[j]);
    for (int k = 0; k < j; ++k) BN_REQUIRE(query[k] != query[j], BN_BAD_ARG, "query has duplicate node %d", query[j]);
    qstride->push_back(cells);
    cells *= net->card[query[j]];
    BN_REQUIRE(cells <= 16384, BN_UNSUPPORTED, "query table has more than 16384 cells");
  }
  *ncell = cells;
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

// One GPU's share of a likelihood-weighting job: particles [first, first + n) on `net`'s device, current stream.
static int32_t lw_infer_local(Net* net, const int8_t* evidence, const int32_t* query, int32_t n_query,
                              const std::vector<int64_t>& qstride, int64_t ncell, uint64_t first_particle,
                              uint64_t n_particles, uint64_t seed, double* out_dev) {
  BN_CUDA_TRY(cudaMemsetAsync(out_dev, 0, (size_t)(ncell + 2) * 8, tl_stream));
  if (n_particles == 0) return BN_OK;
  // network-specialised kernel first (specialize.cu); the generic interpreter covers everything else
  bool used = false;
  BN_TRY(launch_lw_specialized(net, evidence, query, (uint32_t)n_query, qstride, (uint32_t)ncell, first_particle,
                               n_particles, seed, out_dev, &used));
  if (used) return BN_OK;
  SampleArgs a{};
  a.first = first_particle;
  a.n = n_particles;
  a.seed_lo = (uint32_t)seed;
  a.seed_hi = (uint32_t)(seed >> 32);
  a.ncell = (uint32_t)ncell;
  a.out = out_dev;
  return launch_particles<MODE_INFER>(net, evidence, false, a, (uint32_t)n_query, query, qstride);
}

int32_t bn_lw_infer_dev(bn_net_t h, const int8_t* evidence, const int32_t* query, int32_t n_query,
                        uint64_t first_particle, uint64_t n_particles, uint64_t seed, double* out_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_lw_infer: unknown net handle");
  BN_REQUIRE(out_dev, BN_BAD_ARG, "bn_lw_infer: output is NULL");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  BN_TRY(check_evidence(net, evidence));
  std::vector<int64_t> qstride;
  int64_t ncell = 1;
  BN_TRY(check_query(net, evidence, query, n_query, &qstride, &ncell));
  if (!net->comm)
    return lw_infer_local(net, evidence, query, n_query, qstride, ncell, first_particle, n_particles, seed, out_dev);
  // multi-GPU net: every GPU of the communicator takes its share of the global particle range, then one all-reduce
  return comm_fanout(net, first_particle, n_particles, (size_t)ncell + 2, out_dev,
                     [&](Net* r, int, uint64_t lo, uint64_t cnt, double* out) {
                       return lw_infer_local(r, evidence, query, n_query, qstride, ncell, lo, cnt, seed, out);
                     });
}

int32_t bn_lw_infer(bn_net_t h, const int8_t* evidence, const int32_t* query, int32_t n_query,
                    uint64_t first_particle, uint64_t n_particles, uint64_t seed, double* out_counts,
                    double* out_stats) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_lw_infer: unknown net handle");
  BN_REQUIRE(out_counts, BN_BAD_ARG, "bn_lw_infer: out_counts is NULL");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  int64_t ncell = 1;
  for (int j = 0; j < n_query; ++j)
    if (query && query[j] >= 0 && query[j] < net->n) ncell *= net->card[query[j]];
  BN_REQUIRE(ncell <= 16384, BN_UNSUPPORTED, "query table has more than 16384 cells");
  TmpBuf<double> result;  // per call: concurrent calls on one net never share a result buffer
  BN_CUDA_TRY(result.alloc((size_t)ncell + 2));
  BN_TRY(bn_lw_infer_dev(h, evidence, query, n_query, first_particle, n_particles, seed, result.p));
  std::vector<double> host((size_t)ncell + 2);
  BN_CUDA_TRY(cudaMemcpyAsync(host.data(), result.p, host.size() * 8, cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  if (net->comm) BN_TRY(comm_sync_peers(net));
  memcpy(out_counts, host.data(), (size_t)ncell * 8);
  if (out_stats) { out_stats[0] = host[ncell]; out_stats[1] = host[ncell + 1]; }
  return BN_OK;
}

int32_t bn_sample_dev(bn_net_t h, const int8_t* evidence, int32_t kind, int32_t max_attempts,
                      uint64_t first_row, uint64_t n_rows, uint64_t seed, uint8_t* out_states_dev,
                      double* out_w_dev, uint8_t* out_ok_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_sample: unknown net handle");
  BN_REQUIRE(out_states_dev, BN_BAD_ARG, "bn_sample: out_states is NULL");
  BN_REQUIRE(kind >= BN_SAMPLE_ANCESTRAL && kind <= BN_SAMPLE_WEIGHTED, BN_BAD_ARG, "bn_sample: bad kind %d", kind);
  BN_REQUIRE(kind != BN_SAMPLE_REJECTION || max_attempts >= 1, BN_BAD_ARG, "bn_sample: max_attempts must be >= 1");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  if (n_rows == 0) return BN_OK;
  {  // network-specialised kernel first (specialize.cu); same rows bit for bit
    bool used = false;
    BN_TRY(launch_table_specialized(net, evidence, kind, (uint32_t)(max_attempts > 0 ? max_attempts : 1), first_row, n_rows,
                                    seed, out_states_dev, out_w_dev, out_ok_dev, &used));
    if (used) return BN_OK;
  }
  SampleArgs a{};
  a.first = first_row;
  a.n = n_rows;
  a.seed_lo = (uint32_t)seed;
  a.seed_hi = (uint32_t)(seed >> 32);
  a.out_states = out_states_dev;
  a.out_w = out_w_dev;
  a.out_ok = out_ok_dev;
  a.max_attempts = (uint32_t)(max_attempts > 0 ? max_attempts : 1);
  std::vector<int64_t> none;
  if (kind == BN_SAMPLE_REJECTION)
    return launch_particles<MODE_REJECT>(net, evidence, true, a, 0, nullptr, none);
  // ancestral ignores evidence entirely; weighted clamps it
  return launch_particles<MODE_TABLE>(net, kind == BN_SAMPLE_WEIGHTED ? evidence : nullptr,
                                      kind != BN_SAMPLE_WEIGHTED, a, 0, nullptr, none);
}

int32_t bn_sample(bn_net_t h, const int8_t* evidence, int32_t kind, int32_t max_attempts, uint64_t first_row,
                  uint64_t n_rows, uint64_t seed, uint8_t* out_states, double* out_w, uint8_t* out_ok) {
  Net* net = get_net(h);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_sample: unknown net handle");
  BN_REQUIRE(out_states, BN_BAD_ARG, "bn_sample: out_states is NULL");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  const size_t nst = (size_t)net->n * n_rows;
  TmpBuf<uint8_t> d_states, d_ok;
  TmpBuf<double> d_w;
  BN_CUDA_TRY(d_states.alloc(nst));
  BN_CUDA_TRY(d_w.alloc(n_rows));
  BN_CUDA_TRY(d_ok.alloc(n_rows));
  BN_TRY(bn_sample_dev(h, evidence, kind, max_attempts, first_row, n_rows, seed, d_states.p, d_w.p, d_ok.p));
  if (n_rows) {
    BN_CUDA_TRY(cudaMemcpyAsync(out_states, d_states.p, nst, cudaMemcpyDeviceToHost, tl_stream));
    if (out_w) BN_CUDA_TRY(cudaMemcpyAsync(out_w, d_w.p, n_rows * 8, cudaMemcpyDeviceToHost, tl_stream));
    if (out_ok) BN_CUDA_TRY(cudaMemcpyAsync(out_ok, d_ok.p, n_rows, cudaMemcpyDeviceToHost, tl_stream));
  }
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

}  // extern "C"
