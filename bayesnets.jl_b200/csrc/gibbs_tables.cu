// gibbs_tables.cu -- Markov-blanket conditionals of the Gibbs samplers as integer threshold tables.
//
// A Gibbs update (src/Inference/gibbs.jl:36-50, :107-113) draws x_n from p[v] = prod_children P(x_c | pa_c, x_n = v) *
// P(v | pa_n) with StatsBase's cumulative scan: pick = #{v >= 1 : cw_{v-1} < t}, cw = cumsum(p), t = u * sum(p),
// u = (r + 0.5) / 2^32 for the 32-bit draw r.  For a fixed configuration of the blanket every cw_v and sum is a fixed
// fp64 number and t is monotone in r, so each comparison is equivalent to r >= R_v for ONE integer R_v in [0, 2^32]:
//     R_v = min { r : cw_v < fl(fl((r + 0.5) * 2^-32) * sum) }                       (2^32 if there is none)
// Tabulating R_v per (node, blanket configuration) turns the update into: index from the blanket's states, one vector
// load, K-1 integer compares -- no fp64 chain, ~1/3 of the instructions and of the code bytes of the straight-line
// conditional (specialize.cu), and the decisions are IDENTICAL to the fp64 scan for every r, because R_v is found by
// bisection over r with exactly the arithmetic of the kernels and the oracle (same multiply order: children first, own
// CPD last; same sums; this file is compiled -fmad=false like everything else).
//
// Only nodes whose blanket has <= row_limit configurations get a table (the hubs keep the fp64 code).  R_v = 2^32
// ("never": the states after v have no mass at all) does not fit 32 bits; a net in which that occurs in any tabled row
// (deterministic CPDs) reports has_never and the caller keeps the fp64 kernel for it.
#include "common.cuh"

namespace bn {

constexpr int GT_MAX_MEMBERS = 16;

// program per tabled node: [node, K, W, tab_off, nrows, nmem, nterms, (member card) x nmem,
//                           per term: (cpt_off, step_self, nidx, (member index, stride) x nidx)]
__global__ void __launch_bounds__(128) gibbs_table_kernel(const uint32_t* __restrict__ prog, const uint32_t* __restrict__ row_node,
                                                          const uint32_t* __restrict__ node_prog, const uint32_t* __restrict__ row_base,
                                                          uint32_t n_blocks_rows, const double* __restrict__ cpt,
                                                          uint32_t* __restrict__ tab, uint32_t* __restrict__ never_flag) {
  // row_node[b] = tabled-node slot of row block b (128 rows per block), row_base[b] = first row of that block
  const uint32_t b = blockIdx.x;
  if (b >= n_blocks_rows) return;
  const uint32_t* u = prog + node_prog[row_node[b]];
  const uint32_t K = u[1], W = u[2], tab_off = u[3], nrows = u[4], nmem = u[5], nterms = u[6];
  const uint32_t row = row_base[b] + threadIdx.x;
  if (row >= nrows) return;
  uint32_t st[GT_MAX_MEMBERS];
  {
    uint32_t r = row;
    for (uint32_t m = 0; m < nmem; ++m) {
      const uint32_t c = u[7 + m];
      st[m] = r % c;
      r /= c;
    }
  }
  double p[8];
  const uint32_t* t = u + 7 + nmem;
  for (uint32_t k = 0; k < nterms; ++k) {
    uint32_t base = t[0];
    const uint32_t step = t[1], nidx = t[2];
    for (uint32_t q = 0; q < nidx; ++q) base += st[t[3 + 2 * q]] * t[4 + 2 * q];
    for (uint32_t v = 0; v < K; ++v) {
      const double x = cpt[base + v * step];
      p[v] = k == 0 ? x : p[v] * x;
    }
    t += 3 + 2 * nidx;
  }
  double sum = p[0];
  for (uint32_t v = 1; v < K; ++v) sum += p[v];
  double cw = p[0];
  uint32_t* out = tab + tab_off + (size_t)row * W;
  for (uint32_t v = 0; v + 1 < K; ++v) {
    if (v) cw += p[v];
    unsigned long long lo = 0ull, hi = 1ull << 32;  // smallest r with cw < t(r); 2^32 = none
    while (lo < hi) {
      const unsigned long long mid = (lo + hi) >> 1;
      const double uu = ((double)(uint32_t)mid + 0.5) * (1.0 / 4294967296.0);
      const double tt = uu * sum;
      if (cw < tt) hi = mid; else lo = mid + 1;
    }
    if (lo >> 32) { atomicOr(never_flag, 1u); lo = 0xFFFFFFFFull; }
    out[v] = (uint32_t)lo;
  }
  for (uint32_t v = K - 1; v < W; ++v) out[v] = 0xFFFFFFFFu;  // padding, never read as a threshold
}

void gibbs_tables_free(GibbsTables* t) { delete t; }

static uint32_t row_words(int K) { return K <= 2 ? 1u : (K == 3 ? 2u : (K == 4 ? 4u : 8u)); }

// Layout (structure + cardinalities only): which nodes get a table, their blanket members and strides.
void gibbs_tables_layout(const Net& net, uint32_t row_limit, GibbsTables* g) {
  const int n = net.n;
  g->row_limit = row_limit;
  g->tab_off.assign(n, -1);
  g->members.assign(n, {});
  g->strides.assign(n, {});
  g->words = 0;
  for (int i = 0; i < n; ++i) {
    const int K = net.card[i];
    if (K < 2 || K > 8) continue;
    std::vector<int> mb;
    auto add = [&](int m) { if (m != i && std::find(mb.begin(), mb.end(), m) == mb.end()) mb.push_back(m); };
    for (int k = net.par_ptr[i]; k < net.par_ptr[i + 1]; ++k) add(net.par_idx[k]);
    for (int k = net.child_ptr[i]; k < net.child_ptr[i + 1]; ++k) {
      const int c = net.child_idx[k];
      add(c);
      for (int k2 = net.par_ptr[c]; k2 < net.par_ptr[c + 1]; ++k2) add(net.par_idx[k2]);
    }
    std::sort(mb.begin(), mb.end());
    if ((int)mb.size() > GT_MAX_MEMBERS) continue;
    uint64_t rows = 1;
    bool fits = true;
    for (int m : mb) {
      rows *= (uint64_t)net.card[m];
      if (rows > (uint64_t)row_limit) { fits = false; break; }
    }
    if (!fits) continue;
    std::vector<uint32_t> str;
    uint64_t s = 1;
    for (int m : mb) { str.push_back((uint32_t)s); s *= (uint64_t)net.card[m]; }
    g->tab_off[i] = (int64_t)g->words;
    g->members[i] = mb;
    g->strides[i] = str;
    g->words += rows * row_words(K);
    g->words = (g->words + 3) / 4 * 4;  // 16-byte aligned rows for the vector loads
  }
}

// Builds (once per net) and returns the tables; nullptr in *out when no node qualifies.
int32_t gibbs_tables_get(Net* net, uint32_t row_limit, const GibbsTables** out) {
  *out = nullptr;
  std::lock_guard<std::mutex> lk(net->mu);
  if (net->gtab && net->gtab->row_limit == row_limit) { *out = net->gtab; return BN_OK; }
  if (net->gtab) { gibbs_tables_free(net->gtab); net->gtab = nullptr; }
  struct Holder {  // the net gets the tables only when they are complete: an error below leaves it without any
    GibbsTables* g = new GibbsTables();
    ~Holder() { if (g) gibbs_tables_free(g); }
  } hold;
  GibbsTables* g = hold.g;
  gibbs_tables_layout(*net, row_limit, g);
  if (g->words == 0) { net->gtab = g; hold.g = nullptr; *out = g; return BN_OK; }
  BN_REQUIRE(g->words < (1ull << 31), BN_UNSUPPORTED, "Gibbs conditional tables too large (%llu words)", (unsigned long long)g->words);
  // programs
  const int n = net->n;
  std::vector<uint32_t> prog, node_prog, row_node, row_base;
  for (int i = 0; i < n; ++i) {
    if (g->tab_off[i] < 0) continue;
    const std::vector<int>& mb = g->members[i];
    auto member_index = [&](int m) { return (uint32_t)(std::find(mb.begin(), mb.end(), m) - mb.begin()); };
    const int K = net->card[i];
    uint64_t rows = 1;
    for (int m : mb) rows *= (uint64_t)net->card[m];
    const uint32_t slot = (uint32_t)node_prog.size();
    node_prog.push_back((uint32_t)prog.size());
    prog.insert(prog.end(), {(uint32_t)i, (uint32_t)K, row_words(K), (uint32_t)g->tab_off[i], (uint32_t)rows, (uint32_t)mb.size(),
                             (uint32_t)(net->child_ptr[i + 1] - net->child_ptr[i] + 1)});
    for (int m : mb) prog.push_back((uint32_t)net->card[m]);
    auto term = [&](int c, bool own) {
      const int Kc = net->card[c];
      std::vector<uint32_t> idx;
      uint32_t step = own ? 1u : 0u;
      int64_t stride = 1;
      for (int k = net->par_ptr[c]; k < net->par_ptr[c + 1]; ++k) {
        const int pa = net->par_idx[k];
        if (!own && pa == i) step = (uint32_t)(stride * Kc);
        else { idx.push_back(member_index(pa)); idx.push_back((uint32_t)(stride * Kc)); }
        stride *= net->card[pa];
      }
      if (!own) { idx.push_back(member_index(c)); idx.push_back(1u); }
      prog.push_back((uint32_t)net->cpt_off[c]);
      prog.push_back(step);
      prog.push_back((uint32_t)(idx.size() / 2));
      prog.insert(prog.end(), idx.begin(), idx.end());
    };
    for (int k = net->child_ptr[i]; k < net->child_ptr[i + 1]; ++k) term(net->child_idx[k], false);
    term(i, true);
    for (uint64_t r0 = 0; r0 < rows; r0 += 128) { row_node.push_back(slot); row_base.push_back((uint32_t)r0); }
  }
  BN_CUDA_TRY(cudaSetDevice(net->device));
  cudaStream_t s = tl_stream;
  BN_CUDA_TRY(g->tab.alloc((size_t)g->words));
  DevBuf<uint32_t> d_prog, d_np, d_rn, d_rb, d_flag;
  BN_CUDA_TRY(d_prog.upload(prog.data(), prog.size(), s));
  BN_CUDA_TRY(d_np.upload(node_prog.data(), node_prog.size(), s));
  BN_CUDA_TRY(d_rn.upload(row_node.data(), row_node.size(), s));
  BN_CUDA_TRY(d_rb.upload(row_base.data(), row_base.size(), s));
  BN_CUDA_TRY(d_flag.alloc(1));
  BN_CUDA_TRY(cudaMemsetAsync(d_flag.p, 0, 4, s));
  BN_CUDA_TRY(cudaMemsetAsync(g->tab.p, 0xFF, (size_t)g->words * 4, s));
  gibbs_table_kernel<<<(unsigned)row_node.size(), 128, 0, s>>>(d_prog.p, d_rn.p, d_np.p, d_rb.p, (uint32_t)row_node.size(),
                                                              net->d_cpt.p, g->tab.p, d_flag.p);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  uint32_t flag = 0;
  BN_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag.p, 4, cudaMemcpyDeviceToHost, s));
  BN_CUDA_TRY(cudaStreamSynchronize(s));
  g->has_never = flag != 0;
  net->gtab = g;
  hold.g = nullptr;
  *out = g;
  return BN_OK;
}

}  // namespace bn
