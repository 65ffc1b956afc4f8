// exact.cu -- infer(::ExactInference, ::InferenceState): variable elimination with every intermediate
// factor resident in HBM (src/Inference/exact.jl:6-33).
//
// Host side is only bookkeeping (which factors mention the variable being eliminated); all arithmetic
// is bn_factor_* kernels.  fused=1 performs `sum(reduce(*, contain_h), h)` (:27) as one
// bn_factor_eliminate pass, never materialising the product; fused=0 follows the reference literally
// (pairwise products, then the sum) and is what the CPU baseline is compared with.
#include "common.cuh"

#include <algorithm>
#include <set>

namespace bn {

// Greedy minimum-degree ordering of the moralised graph over non-evidence nodes; query nodes are never
// eliminated.  The reference asks CliqueTrees.jl for an MMD ordering with the query clique last
// (src/Inference/exact.jl:37-84); no reference test pins the order (posteriors only) and it affects
// only fp64 rounding and intermediate sizes.
static std::vector<int> elimination_order(const Net& net, const int8_t* evidence, const std::vector<int>& query) {
  const int n = net.n;
  std::vector<std::set<int>> adj(n);
  std::vector<char> alive(n, 0), is_q(n, 0);
  for (int i = 0; i < n; ++i) alive[i] = !(evidence && evidence[i] >= 0);
  for (int q : query) is_q[q] = 1;
  for (int j = 0; j < n; ++j) {
    std::vector<int> fam;
    for (int k = net.par_ptr[j]; k < net.par_ptr[j + 1]; ++k)
      if (alive[net.par_idx[k]]) fam.push_back(net.par_idx[k]);
    if (alive[j]) fam.push_back(j);
    for (size_t a = 0; a < fam.size(); ++a)
      for (size_t b = a + 1; b < fam.size(); ++b) { adj[fam[a]].insert(fam[b]); adj[fam[b]].insert(fam[a]); }
  }
  std::vector<int> todo, order;
  for (int i = 0; i < n; ++i) if (alive[i] && !is_q[i]) todo.push_back(i);
  while (!todo.empty()) {
    size_t best = 0;
    for (size_t t = 1; t < todo.size(); ++t)
      if (adj[todo[t]].size() < adj[todo[best]].size()) best = t;  // ties: lowest index (todo is ascending)
    const int v = todo[best];
    std::vector<int> nb(adj[v].begin(), adj[v].end());
    for (size_t a = 0; a < nb.size(); ++a)
      for (size_t b = a + 1; b < nb.size(); ++b) { adj[nb[a]].insert(nb[b]); adj[nb[b]].insert(nb[a]); }
    for (int a : nb) adj[a].erase(v);
    adj[v].clear();
    todo.erase(todo.begin() + (long)best);
    order.push_back(v);
  }
  return order;
}

struct HandleBag {  // frees every live handle on scope exit
  std::vector<bn_factor_t> hs;
  ~HandleBag() { for (bn_factor_t h : hs) bn_factor_free(h); }
  void drop(bn_factor_t h) {
    hs.erase(std::remove(hs.begin(), hs.end(), h), hs.end());
    bn_factor_free(h);
  }
};

}  // namespace bn

using namespace bn;

extern "C" int32_t bn_exact_infer(bn_net_t hnet, const int8_t* evidence, const int32_t* query, int32_t n_query,
                                  int32_t fused, int32_t* out_dims, int32_t* out_card, double* out) {
  Net* net = get_net(hnet);
  BN_REQUIRE(net, BN_BAD_ARG, "bn_exact_infer: unknown net handle");
  BN_REQUIRE(out && n_query >= 1 && query, BN_BAD_ARG, "bn_exact_infer: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  std::vector<int> q;
  for (int j = 0; j < n_query; ++j) {
    // src/ProbabilisticGraphicalModels/inference.jl:10-11
    BN_REQUIRE(query[j] >= 0 && query[j] < net->n, BN_BAD_ARG, "Query %d is not in the probabilistic graphical model", query[j]);
    BN_REQUIRE(!(evidence && evidence[query[j]] >= 0), BN_BAD_ARG, "Query %d is part of the evidence", query[j]);
    BN_REQUIRE(std::find(q.begin(), q.end(), query[j]) == q.end(), BN_BAD_ARG, "query has duplicate node %d", query[j]);
    q.push_back(query[j]);
  }
  std::vector<int32_t> ev_dims, ev_states;
  for (int i = 0; i < net->n; ++i)
    if (evidence && evidence[i] >= 0) {
      BN_REQUIRE(evidence[i] < net->card[i], BN_BOUNDS, "evidence state %d out of range for node %d", (int)evidence[i], i);
      ev_dims.push_back(i);
      ev_states.push_back(evidence[i]);
    }
  const std::vector<int> hidden = elimination_order(*net, evidence, q);

  HandleBag bag;
  // factors = map(n -> Factor(bn, n, evidence), nodes)   (:16)
  // The CPTs are used where they lie: a factor is a read-only VIEW of its node's table inside the net (no copy; the
  // C5 net's 2^24-entry CPT alone would be 128 MB read + written per call), and only factors that actually mention an
  // evidence variable are sliced into a new table.
  for (int i = 0; i < net->n; ++i) {
    bn_factor_t f = 0, g = 0;
    BN_TRY(factor_view_of_cpd(net, i, &f));
    bag.hs.push_back(f);
    bool touched = false;
    for (int32_t d : get_factor(f)->dims) touched = touched || (evidence && evidence[d] >= 0);
    if (touched) {
      BN_TRY(bn_factor_slice(f, ev_dims.data(), ev_states.data(), (int32_t)ev_dims.size(), &g));
      bag.drop(f);
      bag.hs.push_back(g);
    }
  }
  auto contains = [](bn_factor_t h, int v) {
    Factor* f = get_factor(h);
    return std::find(f->dims.begin(), f->dims.end(), v) != f->dims.end();
  };
  for (int hvar : hidden) {  // :20-29
    std::vector<bn_factor_t> with;
    for (bn_factor_t h : bag.hs) if (contains(h, hvar)) with.push_back(h);
    if (with.empty()) continue;
    const int32_t sd = hvar;
    bn_factor_t res = 0;
    if (fused && with.size() <= 6) {
      BN_TRY(bn_factor_eliminate(with.data(), (int32_t)with.size(), &sd, 1, &res));
    } else {
      bn_factor_t acc = with[0];
      HandleBag tmp;  // the running product is owned here, so an error return cannot leak it
      for (size_t k = 1; k < with.size(); ++k) {
        bn_factor_t nxt = 0;
        BN_TRY(bn_factor_product(acc, with[k], &nxt));
        if (!tmp.hs.empty()) tmp.drop(acc);
        tmp.hs.push_back(nxt);
        acc = nxt;
      }
      BN_TRY(bn_factor_sum(acc, &sd, 1, &res));
    }
    for (bn_factor_t h : with) bag.drop(h);
    bag.hs.push_back(res);  // push!(factors, ...) -- appended last, like the reference
  }
  // phi = normalize!(reduce(*, factors))   (:31)
  bn_factor_t acc = bag.hs[0];
  HandleBag tmp;  // owns the running product / the copy below on every exit path
  for (size_t k = 1; k < bag.hs.size(); ++k) {
    bn_factor_t nxt = 0;
    BN_TRY(bn_factor_product(acc, bag.hs[k], &nxt));
    if (!tmp.hs.empty()) tmp.drop(acc);
    tmp.hs.push_back(nxt);
    acc = nxt;
  }
  if (!get_factor(acc)->data.owns) {  // a lone CPT view (single-factor network): normalise a copy, never the net
    bn_factor_t cp = 0;
    BN_TRY(bn_factor_slice(acc, nullptr, nullptr, 0, &cp));
    tmp.hs.push_back(cp);
    acc = cp;
  }
  BN_TRY(bn_factor_normalize(acc, 1));
  Factor* f = get_factor(acc);
  for (size_t i = 0; i < f->dims.size(); ++i) {
    if (out_dims) out_dims[i] = f->dims[i];
    if (out_card) out_card[i] = f->card[i];
  }
  return bn_factor_download(acc, out);
}
