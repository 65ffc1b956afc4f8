// loopy.cu -- infer(::LoopyBelief) batched over evidence sets (src/Inference/loopy_belief.jl:24-218).
//
// The reference answers one (query, evidence) pair per call with Dict-of-Vector messages on one core.  Here one WARP
// owns one (query, evidence) instance: all of its lambda / pi messages are double-buffered fp64 planes private to the
// warp, the 32 lanes split one iteration's work into TASKS -- "pi messages of node i" and "lambda message of node i to
// its j-th parent" -- packed onto the lanes by the host (longest task first onto the least loaded lane); a lane walks
// its list one CPT cell per trip of ONE flat loop, so lanes with differently shaped tasks do not wait for each other's
// loop nests.  The CPTs, a byte table of parent-state digits and the task records form a read-only plan shared by all
// warps.  Plan, CPTs and message planes can each live in shared memory or in (L1 / L2 resident) global memory -- a
// template choice; measured on B200 the all-global placement wins because it lets an SM hold 32 instances.
//
// Bug-compatible on purpose.  The reference keeps messages by reference, and three aliases change its numbers
// (oracle/loopy_np.py spells them out; the kernel reproduces them):
//   (1) a parentless node's pi IS its factor's potential, updated in place once per child on every iteration
//       (:125,:170-175)  -> plane R below holds that persistent vector;
//   (2) every child of a node is handed the same array, i.e. the message left after the LAST child's update (:177)
//       -> pi messages are stored per node, not per edge;
//   (3) from the second iteration on both dictionaries hold a root's array, so its children read the message of the
//       CURRENT iteration and its own change is norm(px - px, Inf) -> roots run first (phase A) and also write the
//       read plane.
// Arithmetic order (left folds, column-major sequential sums, x ./= sum(abs) for factors, x .*= inv(norm1) for
// vectors, LinearAlgebra's generic_norm2 / generic_normInf) follows the oracle term by term; compiled -fmad=false, so
// beliefs and iteration counts are bit-identical to it.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace bn {

struct LoopyPlan {
  std::vector<uint32_t> words;  // [nodes | edges | child edges | tasks A | tasks P | task records B by lane | lane ptr | digit rows]
  uint32_t off_node = 0, off_edge = 0, off_child = 0, off_taskA = 0, off_taskB = 0, off_taskP = 0, off_taskPP = 0, off_lane = 0, off_dig = 0;
  uint32_t nA = 0, nB = 0, nP = 0, nPP = 0, n_edges = 0, S_c = 0, S_l = 0, S_pp = 0, S_r = 0, max_card = 0;
  DevBuf<uint32_t> dev;
};
void loopy_plan_free(LoopyPlan* p) { delete p; }

enum : uint32_t { NODE_WORDS = 12, EDGE_WORDS = 8, TASK_WORDS = 12, NO_PP = 0xffffffffu, PP_MIN_CHILDREN = 5 };
// phase-B task record (TASK_WORDS words): everything the flat loop needs, so a lane switches task with three
// 128-bit loads.  w0 cpt offset | w1 c + k<<8 + pc<<16 + flags<<24 | w2 stride | w3 n_hi (lambda) / node (pi) |
// w4 digit-row offset | w5 destination offset | w6 offset of the node's lambda_i / pi | w7 digit mask |
// w8..w11 pi-plane offsets of parents 0..3 (the ONE slot for absent / skipped parents); k > 4: w8 = first edge, w9 = j
enum : uint32_t { TF_LAMBDA = 1u << 24, TF_SLOW = 2u << 24 };

struct LoopyArgs {
  const uint32_t* plan;
  const double* cpt;
  const int8_t* evidence;
  const int32_t* query;
  double* out;
  int32_t* out_iters;
  unsigned char* scratch;  // message planes of every warp when they do not fit shared memory
  unsigned long long n_inst;
  double tol, delta, eps_delta;
  uint32_t plan_words, cpt_len;
  uint32_t n_nodes, n_edges, S_c, S_l, S_pp, S_r;
  uint32_t off_node, off_edge, off_child, off_taskA, off_taskB, off_taskP, off_taskPP, off_lane, off_dig, nA, nB, nP, nPP;
  uint32_t shared_bytes, per_warp_bytes, out_stride;
  int32_t nsamples, K;
};

// One parent-state row of the digit table: (k + 3) / 4 words, one byte per parent.
__device__ __forceinline__ double times_parent_messages(double t, const uint32_t* __restrict__ ed0,
                                                        const uint32_t* __restrict__ dgrow, const double* __restrict__ PIr,
                                                        uint32_t k, uint32_t skip) {
  uint32_t w = 0;
  for (uint32_t j = 0; j < k; ++j) {
    if ((j & 3u) == 0u) w = dgrow[j >> 2];
    const double f = PIr[ed0[j * EDGE_WORDS + 4] + (w & 0xffu)];
    w >>= 8;
    t = t * (j == skip ? 1.0 : f);  // x * 1.0 is exact: the skipped parent leaves no trace
  }
  return t;
}

// pi[x] = sum over parent configurations q (ascending) of ((cpt[x,q] * pi_1[u_1]) * pi_2[u_2]) ...
// broadcast(*, phi, parents, pis) then sum!(.., parents): loopy_belief.jl:130-132 (final belief only; the iteration
// runs the same sum inside the flat task loop)
template <bool CPT_SMEM>
__device__ __forceinline__ double pi_entry(const double* __restrict__ cpt, const uint32_t* __restrict__ ed0,
                                           const uint32_t* __restrict__ dg, const double* __restrict__ PIr, uint32_t x,
                                           uint32_t c, uint32_t Q, uint32_t k) {
  const uint32_t kw = (k + 3u) >> 2;
  double acc = 0.0;
  for (uint32_t q = 0; q < Q; ++q) {
    double t = CPT_SMEM ? cpt[x + c * q] : __ldg(cpt + x + c * q);
    t = times_parent_messages(t, ed0, dg + q * kw, PIr, k, 0xffffffffu);
    acc = q ? acc + t : t;
  }
  return acc;
}

struct LoopyCtx {
  const uint32_t* nodes;
  const uint32_t* edges;
  const uint32_t* childe;
  const double* PIr;
  double* PIw;
  double* Rp;
  const double* LAMr;
  double* LAMw;
  const double* PP;
  double delta, eps_delta;
  int32_t t;
};

// pi messages of node i to its children, px already holding pi (or the root's persistent prior):
// loopy_belief.jl:163-185.  Returns the largest change (-inf if none counts).
__device__ __forceinline__ double pi_messages(const LoopyCtx g, uint32_t i, double mc) {
  const uint32_t* nd = g.nodes + i * NODE_WORDS;
  const uint32_t c = nd[0], k = nd[2], m = nd[4], poff = nd[6];
  const bool root = (k == 0);
  double* px = root ? g.Rp + nd[9] : g.PIw + poff;
  const uint32_t* ce = g.childe + nd[5];
  for (uint32_t kk = 0; kk < m; ++kk) {
    double nrm = 0.0;
    for (uint32_t x = 0; x < c; ++x) {
      double v = px[x];
      if (m > 1) {  // px .*= reduce(.*, lambdas of the other children)
        double p;
        if (m >= PP_MIN_CHILDREN) p = g.PP[g.edges[ce[kk] * EDGE_WORDS + 7] + x];  // folded in phase A
        else {
          const uint32_t first = kk == 0 ? 1u : 0u;
          p = g.LAMr[g.edges[ce[first] * EDGE_WORDS + 1] + x];
          for (uint32_t k2 = first + 1; k2 < m; ++k2)
            if (k2 != kk) p = p * g.LAMr[g.edges[ce[k2] * EDGE_WORDS + 1] + x];
        }
        v = v * p;
        px[x] = v;
      }
      nrm = x ? nrm + fabs(v) : fabs(v);
    }
    // normalize!(px, 1): LinearAlgebra.__normalize!
    bool nan = false;
    double dm = 0.0;
    const bool safe = nrm >= g.delta;
    const double inv = safe ? 1.0 / nrm : 1.0 / (nrm * g.eps_delta);
    for (uint32_t x = 0; x < c; ++x) {
      double v = px[x];
      if (safe) v = v * inv;
      else { v = v * g.eps_delta; v = v * inv; }
      px[x] = v;
      // change against the message of the previous iteration (root: see (3) in the header)
      const double old = root ? (g.t == 1 ? 1.0 / (double)c : v) : g.PIr[poff + x];
      const double d = fabs(v - old);
      if (d != d) nan = true;
      else dm = fmax(dm, d);
    }
    if (!nan && dm > mc) mc = dm;
  }
  if (root)
    for (uint32_t x = 0; x < c; ++x) {
      const double v = px[x];
      g.PIw[poff + x] = v;
      if (g.t >= 2) const_cast<double*>(g.PIr)[poff + x] = v;
    }
  return mc;
}

// normalize!(lx), then norm(lx - old) as LinearAlgebra.generic_norm2: loopy_belief.jl:150-159
__device__ __forceinline__ double lambda_finish(const LoopyCtx g, uint32_t loff, uint32_t pc, double mc) {
  double total = 0.0;
  for (uint32_t u = 0; u < pc; ++u) total = u ? total + fabs(g.LAMw[loff + u]) : fabs(g.LAMw[loff + u]);
  bool nan = false;
  double maxabs = 0.0;
  for (uint32_t u = 0; u < pc; ++u) {
    const double v = g.LAMw[loff + u] / total;
    g.LAMw[loff + u] = v;
    const double d = fabs(v - g.LAMr[loff + u]);
    if (d != d) nan = true;
    else maxabs = fmax(maxabs, d);
  }
  double delta;
  if (nan) delta = NAN;
  else if (maxabs == 0.0 || isinf(maxabs)) delta = maxabs;
  else {
    const double mm = maxabs * maxabs;
    const bool plain = isfinite((double)pc * mm) && mm != 0.0;
    double s = 0.0;
    for (uint32_t u = 0; u < pc; ++u) {
      const double d = g.LAMw[loff + u] - g.LAMr[loff + u];
      if (plain) s = s + d * d;
      else { const double r = fabs(d) / maxabs; s = s + r * r; }
    }
    delta = plain ? sqrt(s) : maxabs * sqrt(s);
  }
  return delta > mc ? delta : mc;
}

// Where things live.  PLAN_SMEM / CPT_SMEM: the plan (task records, node / edge tables, digit rows) and the CPTs are
// staged into shared memory once per CTA, else read through L1 / L2.  MSG_SMEM: every warp's message planes are in
// shared memory, else a slice of a global scratch buffer (L1 / L2 resident: 18 KB per warp on the C2 net).  Same code
// and results in every combination; what changes is how many warps an SM can hold.
template <bool PLAN_SMEM, bool CPT_SMEM, bool MSG_SMEM, int MINB = 0>
__global__ void __launch_bounds__(MINB > 0 ? 256 : 320, MINB) loopy_kernel(const LoopyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t* plan = a.plan;
  const double* cpt = a.cpt;
  {
    size_t at = 0;
    if (PLAN_SMEM) {
      uint32_t* plan_s = reinterpret_cast<uint32_t*>(smem_raw);
      for (uint32_t i = threadIdx.x; i < a.plan_words; i += blockDim.x) plan_s[i] = a.plan[i];
      plan = plan_s;
      at = ((size_t)a.plan_words * 4 + 15) & ~(size_t)15;
    }
    if (CPT_SMEM) {
      double* cpt_s = reinterpret_cast<double*>(smem_raw + at);
      for (uint32_t i = threadIdx.x; i < a.cpt_len; i += blockDim.x) cpt_s[i] = a.cpt[i];
      cpt = cpt_s;
    }
    if (PLAN_SMEM || CPT_SMEM) __syncthreads();
  }
  const uint32_t* nodes = plan + a.off_node;
  const uint32_t* edges = plan + a.off_edge;
  const uint32_t* taskA = plan + a.off_taskA;
  const uint4* taskB = reinterpret_cast<const uint4*>(plan + a.off_taskB);
  const uint32_t* taskP = plan + a.off_taskP;
  const uint32_t* taskPP = plan + a.off_taskPP;
  const uint32_t* dig = plan + a.off_dig;

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  unsigned char* mine = MSG_SMEM ? smem_raw + a.shared_bytes + (size_t)warp * a.per_warp_bytes
                                 : a.scratch + ((size_t)blockIdx.x * wpc + warp) * a.per_warp_bytes;
  // planes: PI[0], PI[1] (S_c + 1 each, the last entry is the constant 1.0), R (roots only), LAMN, PP, LAM[0], LAM[1]
  const uint32_t P1 = a.S_c + 1u;
  double* PI0 = reinterpret_cast<double*>(mine);
  double* LAMN = PI0 + 2 * P1 + a.S_r;  // lambda_i[x] of the current iteration
  double* PPw = LAMN + a.S_c;           // products of the sibling lambdas (nodes with more than two children)
  double* LAM0 = PPw + a.S_pp;
  int8_t* ev = reinterpret_cast<int8_t*>(LAM0 + 2 * a.S_l);
  LoopyCtx g;
  g.nodes = nodes;
  g.edges = edges;
  g.childe = plan + a.off_child;
  g.Rp = PI0 + 2 * P1;
  g.PP = PPw;
  g.delta = a.delta;
  g.eps_delta = a.eps_delta;
  // this lane's share of the phase-B tasks (packed by the host so that every lane carries the same load)
  const uint32_t my_begin = plan[a.off_lane + lane], my_end = plan[a.off_lane + lane + 1];

  for (unsigned long long inst = (unsigned long long)blockIdx.x * wpc + warp; inst < a.n_inst;
       inst += (unsigned long long)gridDim.x * wpc) {
    // ---- initial messages: loopy_belief.jl:69-90
    const int8_t* ev_g = a.evidence + inst * a.n_nodes;
    for (uint32_t i = lane; i < a.n_nodes; i += 32) {
      const uint32_t* nd = nodes + i * NODE_WORDS;
      const uint32_t c = nd[0], poff = nd[6];
      int e = ev_g[i];
      if (e >= (int)c) e = -1;  // host entry validates; the _dev entry treats junk as free
      ev[i] = (int8_t)e;
      const double u = 1.0 / (double)c;
      for (uint32_t x = 0; x < c; ++x) {
        const double v = e >= 0 ? ((int)x == e ? 1.0 : 0.0) : u;
        PI0[poff + x] = v;
        PI0[P1 + poff + x] = v;
        if (nd[2] == 0) g.Rp[nd[9] + x] = CPT_SMEM ? cpt[nd[1] + x] : __ldg(cpt + nd[1] + x);
      }
    }
    if (lane == 0) PI0[a.S_c] = PI0[P1 + a.S_c] = 1.0;
    for (uint32_t e = lane; e < a.n_edges; e += 32) {
      const uint32_t* ed = edges + e * EDGE_WORDS;
      const double u = 1.0 / (double)ed[3];
      for (uint32_t x = 0; x < ed[3]; ++x) {
        LAM0[ed[1] + x] = u;
        LAM0[a.S_l + ed[1] + x] = u;
      }
    }
    __syncwarp();

    int32_t consec = (INFINITY <= a.tol) ? a.K : 0;
    int32_t iters = 0;
    for (int32_t t = 1; t <= a.nsamples; ++t) {
      iters = t;
      g.t = t;
      g.PIr = PI0 + ((t - 1) & 1) * P1;
      g.PIw = PI0 + (t & 1) * P1;
      g.LAMr = LAM0 + ((t - 1) & 1) * a.S_l;
      g.LAMw = LAM0 + (t & 1) * a.S_l;
      double mc = -INFINITY;

      // ---- phase A, everything that only reads last iteration's lambdas:
      //  * every node with parents folds its children's lambdas (or its evidence) into lambda_i[x] (:114-121);
      //  * for nodes with at least PP_MIN_CHILDREN children, the product of the OTHER children's lambdas that each child's pi
      //    message needs (:172-173), one lane per child instead of one lane per node in phase C;
      //  * then unobserved roots update their persistent priors in place (aliases (1) and (3))
      for (uint32_t i = lane; i < a.n_nodes; i += 32) {
        const uint32_t* nd = nodes + i * NODE_WORDS;
        if (nd[2] == 0) continue;
        const uint32_t c = nd[0], m = nd[4], poff = nd[6];
        const uint32_t* ce = g.childe + nd[5];
        const int evi = ev[i];
        for (uint32_t x = 0; x < c; ++x) {
          double lam;
          if (evi >= 0) lam = ((int)x == evi) ? 1.0 : 0.0;
          else if (m == 0) lam = 1.0;
          else {
            lam = g.LAMr[edges[ce[0] * EDGE_WORDS + 1] + x];
            for (uint32_t k2 = 1; k2 < m; ++k2) lam = lam * g.LAMr[edges[ce[k2] * EDGE_WORDS + 1] + x];
          }
          LAMN[poff + x] = lam;
        }
      }
      for (uint32_t it = lane; it < a.nPP; it += 32) {  // one (child edge, state) pair per lane, most siblings first
        const uint32_t item = taskPP[it], e = item & 0xffffffu, x = item >> 24;
        const uint32_t* ed = edges + e * EDGE_WORDS;
        if (ev[ed[0]] >= 0) continue;
        const uint32_t* nd = nodes + ed[0] * NODE_WORDS;  // the parent whose children are being multiplied
        const uint32_t m = nd[4];
        const uint32_t* ce = g.childe + nd[5];
        double p = 0.0;
        bool first = true;
        for (uint32_t k2 = 0; k2 < m; ++k2) {
          const uint32_t e2 = ce[k2];
          if (e2 == e) continue;
          const double l = g.LAMr[edges[e2 * EDGE_WORDS + 1] + x];
          p = first ? l : p * l;
          first = false;
        }
        PPw[ed[7] + x] = p;
      }
      __syncwarp();
      for (uint32_t ti = lane; ti < a.nA; ti += 32) {
        const uint32_t i = taskA[ti];
        if (ev[i] < 0) mc = pi_messages(g, i, mc);
      }
      __syncwarp();

      // ---- phase B: the sums over parent configurations.  Every lane walks its own task list one CPT cell per trip
      // of one flat loop, so the warp's time is the largest per-lane sum of cells, not a product of the largest loop
      // bounds of 32 differently shaped tasks.  Up to four parents the cell is straight-line code: one digit word,
      // four independent message loads (absent or skipped parents read the constant 1.0), four multiplies.
      uint32_t tp = my_begin;
      bool active = false;
      uint32_t flags = 0, c = 0, pc = 1, stride = 1, n_hi = 1, jump = 0, wmask = 0, po0 = 0, po1 = 0, po2 = 0, po3 = 0;
      uint32_t lo = 0, hi = 0, x = 0, u = 0, q = 0;
      const double* cp = cpt;
      const uint32_t* dg = dig;
      const double* lamn = LAMN;
      double* dst = g.PIw;
      double acc = 0.0, msg = 0.0;
      while (true) {
        while (!active && tp < my_end) {
          const uint4 r0 = taskB[tp * 3], r1 = taskB[tp * 3 + 1], r2 = taskB[tp * 3 + 2];
          ++tp;
          flags = r0.y;
          if (!(flags & TF_LAMBDA) && ev[r0.w] >= 0) continue;  // an evidence node keeps its one-hot message
          c = flags & 0xffu;
          pc = (flags >> 16) & 0xffu;
          cp = cpt + r0.x;
          stride = r0.z;
          n_hi = (flags & TF_LAMBDA) ? r0.w : 1u;
          jump = stride * (pc - 1u);
          dg = dig + r1.x;
          dst = ((flags & TF_LAMBDA) ? g.LAMw : g.PIw) + r1.y;
          lamn = LAMN + r1.z;
          wmask = r1.w;
          po0 = r2.x; po1 = r2.y; po2 = r2.z; po3 = r2.w;
          lo = hi = x = u = q = 0;
          active = true;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
          double tv = CPT_SMEM ? cp[x + c * q] : __ldg(cp + x + c * q);
          if (!(flags & TF_SLOW)) {
            const uint32_t w = dg[q] & wmask;
            const double f0 = g.PIr[po0 + (w & 0xffu)];
            const double f1 = g.PIr[po1 + ((w >> 8) & 0xffu)];
            const double f2 = g.PIr[po2 + ((w >> 16) & 0xffu)];
            const double f3 = g.PIr[po3 + (w >> 24)];
            tv = tv * f0;  // x * 1.0 is exact: absent parents leave no trace
            tv = tv * f1;
            tv = tv * f2;
            tv = tv * f3;
          } else {
            const uint32_t k = (flags >> 8) & 0xffu;
            tv = times_parent_messages(tv, edges + po0 * EDGE_WORDS, dg + q * ((k + 3u) >> 2), g.PIr, k, po1);
          }
          acc = (lo | hi) ? acc + tv : tv;
          ++q;
          if (++lo == stride) {
            lo = 0;
            q += jump;
            if (++hi == n_hi) {
              hi = 0;  // cell (x, u) is complete
              if (flags & TF_LAMBDA) {
                const double v = acc * lamn[x];  // broadcast!(*, lx, nn, lambda) then sum!(lx, nn)
                msg = x ? msg + v : v;
                if (++x == c) {
                  x = 0;
                  dst[u] = msg;
                  active = (++u != pc);
                }
              } else {
                dst[x] = acc;
                active = (++x != c);
              }
              q = stride * u;
            }
          }
        }
      }
      __syncwarp();

      // ---- phase C: finish the messages, one per lane: normalise + change of every lambda message, then the pi
      // messages of every unobserved non-root node to its children
      for (uint32_t e = lane; e < a.n_edges; e += 32) {
        const uint32_t* ed = edges + e * EDGE_WORDS;
        mc = lambda_finish(g, ed[1], ed[3], mc);
      }
      for (uint32_t ti = lane; ti < a.nP; ti += 32) {
        const uint32_t i = taskP[ti];
        if (ev[i] < 0) mc = pi_messages(g, i, mc);
      }
      __syncwarp();
#pragma unroll
      for (int o = 16; o; o >>= 1) mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o));
      consec = (mc <= a.tol) ? min(consec + 1, a.K) : 0;  // maximum(change_per_iter) <= tol (:196)
      if (consec >= a.K) break;
    }

    // ---- belief of the query: loopy_belief.jl:201-217 (messages of the last completed iteration)
    if (lane == 0) {
      const int32_t qn = a.query[inst];
      double* o = a.out + inst * a.out_stride;
      if (qn < 0 || (uint32_t)qn >= a.n_nodes || nodes[qn * NODE_WORDS] > a.out_stride) {
        for (uint32_t x = 0; x < a.out_stride; ++x) o[x] = NAN;
      } else {
        const double* PIr = PI0 + (iters & 1) * P1;
        const double* LAMr = LAM0 + (iters & 1) * a.S_l;
        const uint32_t* nd = nodes + (uint32_t)qn * NODE_WORDS;
        const uint32_t c = nd[0], k = nd[2], m = nd[4], poff = nd[6];
        const uint32_t* ce = g.childe + nd[5];
        double total = 0.0;
        for (uint32_t x = 0; x < c; ++x) {
          double lam = 1.0;
          if (m) {
            lam = LAMr[edges[ce[0] * EDGE_WORDS + 1] + x];
            for (uint32_t k2 = 1; k2 < m; ++k2) lam = lam * LAMr[edges[ce[k2] * EDGE_WORDS + 1] + x];
          }
          const double pi = k == 0 ? g.Rp[nd[9] + x]
                                   : pi_entry<CPT_SMEM>(cpt + nd[1], edges + nd[3] * EDGE_WORDS, dig + nd[7], PIr, x, c,
                                                        nd[8], k);
          const double v = lam * pi;
          o[x] = v;
          total = x ? total + fabs(v) : fabs(v);
        }
        for (uint32_t x = 0; x < c; ++x) o[x] = o[x] / total;
        for (uint32_t x = c; x < a.out_stride; ++x) o[x] = 0.0;
      }
      if (a.out_iters) a.out_iters[inst] = iters;
    }
    __syncwarp();
  }
}

static int32_t loopy_plan(Net* net) {
  std::lock_guard<std::mutex> lk(net->mu);  // built once per net; read-only afterwards
  if (net->loopy) return BN_OK;
  const int n = net->n;
  auto* P = new LoopyPlan();
  std::vector<uint32_t> pi_off((size_t)n), lam_off, dig_off((size_t)n);
  uint64_t S_c = 0, S_l = 0, dig_words = 0;
  for (int i = 0; i < n; ++i) {
    pi_off[(size_t)i] = (uint32_t)S_c;
    S_c += (uint64_t)net->card[i];
    P->max_card = std::max<uint32_t>(P->max_card, (uint32_t)net->card[i]);
  }
  const int32_t n_edges = net->par_ptr[n];
  lam_off.resize((size_t)n_edges);
  for (int i = 0; i < n; ++i)
    for (int e = net->par_ptr[i]; e < net->par_ptr[i + 1]; ++e) {
      lam_off[(size_t)e] = (uint32_t)S_l;
      S_l += (uint64_t)net->card[net->par_idx[e]];
    }
  for (int i = 0; i < n; ++i) {
    const int64_t Q = (net->cpt_off[i + 1] - net->cpt_off[i]) / net->card[i];
    dig_off[(size_t)i] = (uint32_t)dig_words;
    dig_words += (uint64_t)Q * (uint64_t)((net->par_ptr[i + 1] - net->par_ptr[i] + 3) / 4);
    if (dig_words > (16ull << 20) || net->cpt_off[n] > (int64_t)0x7fffffff || net->par_ptr[n] >= (1 << 24)) {
      delete P;
      set_error("LoopyBelief: CPTs too wide for the digit table (node %d has %lld parent configurations)", i, (long long)Q);
      return BN_UNSUPPORTED;
    }
  }
  auto& w = P->words;
  P->off_node = 0;
  w.resize((size_t)n * NODE_WORDS, 0u);
  for (int i = 0; i < n; ++i) {
    uint32_t* nd = &w[(size_t)i * NODE_WORDS];
    nd[0] = (uint32_t)net->card[i];
    nd[1] = (uint32_t)net->cpt_off[i];
    nd[2] = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
    nd[3] = (uint32_t)net->par_ptr[i];
    nd[4] = (uint32_t)(net->child_ptr[i + 1] - net->child_ptr[i]);
    nd[5] = (uint32_t)net->child_ptr[i];
    nd[6] = pi_off[(size_t)i];
    nd[7] = dig_off[(size_t)i];
    nd[8] = (uint32_t)((net->cpt_off[i + 1] - net->cpt_off[i]) / net->card[i]);
    if (nd[2] == 0) { nd[9] = P->S_r; P->S_r += nd[0]; }  // persistent prior of a root (plane R)
  }
  P->off_edge = (uint32_t)w.size();
  w.resize(w.size() + (size_t)std::max(n_edges, 1) * EDGE_WORDS, 0u);
  for (int i = 0; i < n; ++i) {
    uint32_t stride = 1;
    for (int e = net->par_ptr[i]; e < net->par_ptr[i + 1]; ++e) {
      uint32_t* ed = &w[P->off_edge + (size_t)e * EDGE_WORDS];
      const int p = net->par_idx[e];
      ed[0] = (uint32_t)p;
      ed[1] = lam_off[(size_t)e];
      ed[2] = stride;
      ed[3] = (uint32_t)net->card[p];
      ed[4] = pi_off[(size_t)p];
      ed[5] = (uint32_t)i;
      ed[6] = (uint32_t)(e - net->par_ptr[i]);
      ed[7] = NO_PP;
      stride *= (uint32_t)net->card[p];
    }
  }
  // children of p in ascending order -> the edge (child -> p) that carries lambdas[(child, p)]
  P->off_child = (uint32_t)w.size();
  w.resize(w.size() + (size_t)std::max(n_edges, 1), 0u);
  for (int p = 0; p < n; ++p)
    for (int kk = net->child_ptr[p]; kk < net->child_ptr[p + 1]; ++kk) {
      const int ch = net->child_idx[kk];
      for (int e = net->par_ptr[ch]; e < net->par_ptr[ch + 1]; ++e)
        if (net->par_idx[e] == p) w[P->off_child + (size_t)kk] = (uint32_t)e;
    }
  // products of sibling lambdas are tabulated only where a node has PP_MIN_CHILDREN or more children
  uint64_t S_pp = 0;
  for (int p = 0; p < n; ++p) {
    if (net->child_ptr[p + 1] - net->child_ptr[p] < (int)PP_MIN_CHILDREN) continue;
    for (int kk = net->child_ptr[p]; kk < net->child_ptr[p + 1]; ++kk) {
      w[P->off_edge + (size_t)w[P->off_child + (size_t)kk] * EDGE_WORDS + 7] = (uint32_t)S_pp;
      S_pp += (uint64_t)net->card[p];
    }
  }
  P->S_pp = (uint32_t)S_pp;
  // tasks.  Phase A (unobserved roots) and phase C (pi messages) are dealt round-robin, most expensive first; the
  // phase-B records are packed onto the 32 lanes by longest-processing-time-first, so every lane walks about the same
  // number of CPT cells per iteration.
  struct Rec { uint64_t cost; uint32_t w[TASK_WORDS]; };
  std::vector<std::pair<uint64_t, uint32_t>> A, Pm;
  std::vector<Rec> B;
  for (int i = 0; i < n; ++i) {
    const uint32_t c = (uint32_t)net->card[i], k = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
    const uint64_t m = (uint64_t)(net->child_ptr[i + 1] - net->child_ptr[i]);
    const uint64_t cells = (uint64_t)(net->cpt_off[i + 1] - net->cpt_off[i]);
    const uint32_t Q = (uint32_t)(cells / c);
    if (m && !k) A.push_back({c * m * (m + 2), (uint32_t)i});
    if (m && k) Pm.push_back({c * m * (m + 2), (uint32_t)i});
    if (!k) continue;
    for (int jj = -1; jj < (int)k; ++jj) {  // -1: the node's own pi (only if it has children), else the lambda to parent jj
      if (jj < 0 && !m) continue;
      Rec r{};
      r.cost = cells * (k > 4 ? 3 * k : 8) + 24;
      const uint32_t* ed = jj < 0 ? nullptr : &w[P->off_edge + (size_t)(net->par_ptr[i] + jj) * EDGE_WORDS];
      const uint32_t pc = ed ? ed[3] : 1u;
      r.w[0] = (uint32_t)net->cpt_off[i];
      r.w[1] = c | (k << 8) | (pc << 16) | (ed ? (uint32_t)TF_LAMBDA : 0u) | (k > 4 ? (uint32_t)TF_SLOW : 0u);
      r.w[2] = ed ? ed[2] : Q;                             // stride of the parent the message goes to
      r.w[3] = ed ? Q / (ed[2] * pc) : (uint32_t)i;        // n_hi | node (evidence check)
      r.w[4] = dig_off[(size_t)i];
      r.w[5] = ed ? ed[1] : pi_off[(size_t)i];
      r.w[6] = pi_off[(size_t)i];
      if (k > 4) {
        r.w[8] = (uint32_t)net->par_ptr[i];
        r.w[9] = jj < 0 ? 0xffffffffu : (uint32_t)jj;
      } else {
        for (uint32_t t = 0; t < 4; ++t) {
          const bool use = t < k && (int)t != jj;
          r.w[8 + t] = use ? pi_off[(size_t)net->par_idx[net->par_ptr[i] + (int)t]] : (uint32_t)S_c;  // S_c: the ONE slot
          if (use) r.w[7] |= 0xffu << (8 * t);
        }
      }
      B.push_back(r);
    }
  }
  auto by_cost = [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) {
    return x.first != y.first ? x.first > y.first : x.second < y.second;
  };
  std::sort(A.begin(), A.end(), by_cost);
  std::sort(Pm.begin(), Pm.end(), by_cost);
  std::stable_sort(B.begin(), B.end(), [](const Rec& x, const Rec& y) { return x.cost > y.cost; });
  P->off_taskA = (uint32_t)w.size();
  for (auto& t : A) w.push_back(t.second);
  P->off_taskP = (uint32_t)w.size();
  for (auto& t : Pm) w.push_back(t.second);
  P->nA = (uint32_t)A.size();
  P->nP = (uint32_t)Pm.size();
  // (child edge, state) pairs whose sibling products are tabulated, nodes with the most children first
  {
    std::vector<std::pair<uint64_t, uint32_t>> items;
    for (int p = 0; p < n; ++p) {
      const int m = net->child_ptr[p + 1] - net->child_ptr[p];
      if (m < (int)PP_MIN_CHILDREN) continue;
      for (int kk = net->child_ptr[p]; kk < net->child_ptr[p + 1]; ++kk)
        for (int x = 0; x < net->card[p]; ++x) items.push_back({(uint64_t)m, w[P->off_child + (size_t)kk] | ((uint32_t)x << 24)});
    }
    std::stable_sort(items.begin(), items.end(), [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) { return x.first > y.first; });
    P->off_taskPP = (uint32_t)w.size();
    for (auto& t : items) w.push_back(t.second);
    P->nPP = (uint32_t)items.size();
  }
  P->nB = (uint32_t)B.size();
  std::vector<std::vector<const Rec*>> lanes(32);
  std::vector<uint64_t> load(32, 0);
  for (const Rec& r : B) {
    const size_t l = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
    lanes[l].push_back(&r);
    load[l] += r.cost;
  }
  while (w.size() % 4) w.push_back(0u);  // records are read as uint4
  P->off_taskB = (uint32_t)w.size();
  std::vector<uint32_t> lane_ptr(33, 0);
  for (int l = 0; l < 32; ++l) {
    for (const Rec* r : lanes[(size_t)l]) w.insert(w.end(), r->w, r->w + TASK_WORDS);
    lane_ptr[(size_t)l + 1] = lane_ptr[(size_t)l] + (uint32_t)lanes[(size_t)l].size();
  }
  P->off_lane = (uint32_t)w.size();
  for (uint32_t v : lane_ptr) w.push_back(v);
  // parent-state digits: row q of node i is (k + 3) / 4 words at dig_off[i] + q * kw, byte j = state of parent j
  P->off_dig = (uint32_t)w.size();
  w.resize(w.size() + (size_t)dig_words + 1, 0u);
  for (int i = 0; i < n; ++i) {
    const int k = net->par_ptr[i + 1] - net->par_ptr[i];
    const int kw = (k + 3) / 4;
    const int64_t Q = (net->cpt_off[i + 1] - net->cpt_off[i]) / net->card[i];
    for (int64_t q = 0; q < Q; ++q) {
      int64_t r = q;
      for (int j = 0; j < k; ++j) {
        const int pc = net->card[net->par_idx[net->par_ptr[i] + j]];
        w[P->off_dig + dig_off[(size_t)i] + (size_t)q * (size_t)kw + (size_t)(j / 4)] |= (uint32_t)(r % pc) << (8 * (j % 4));
        r /= pc;
      }
    }
  }
  while (w.size() % 4) w.push_back(0u);
  P->n_edges = (uint32_t)n_edges;
  P->S_c = (uint32_t)S_c;
  P->S_l = (uint32_t)S_l;
  cudaError_t e = P->dev.upload(w.data(), w.size(), tl_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(tl_stream);
  if (e != cudaSuccess) {
    delete P;
    set_error("LoopyBelief: plan upload failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BN_OOM : BN_CUDA_ERR;
  }
  net->loopy = P;
  return BN_OK;
}

static int32_t loopy_launch(Net* net, const int8_t* ev_dev, const int32_t* query_dev, uint64_t n_inst, int32_t nsamples,
                            double tol, int32_t K, int32_t out_stride, double* out_dev, int32_t* iters_dev) {
  BN_REQUIRE(nsamples >= 0, BN_BAD_ARG, "LoopyBelief: nsamples must be >= 0");
  BN_REQUIRE(K >= 1, BN_BAD_ARG, "LoopyBelief: iters_for_convergence must be >= 1");
  BN_REQUIRE(out_stride >= 1, BN_BAD_ARG, "LoopyBelief: out_stride must be >= 1");
  if (n_inst == 0) return BN_OK;
  BN_TRY(loopy_plan(net));
  const LoopyPlan& P = *net->loopy;
  const DevProps& dp = dev_props(net->device);
  LoopyArgs a{};
  a.plan = P.dev.p;
  a.plan_words = (uint32_t)P.words.size();
  a.cpt = net->d_cpt.p;
  a.cpt_len = (uint32_t)net->cpt.size();
  a.evidence = ev_dev;
  a.query = query_dev;
  a.out = out_dev;
  a.out_iters = iters_dev;
  a.n_inst = n_inst;
  a.tol = tol;
  a.delta = 1.0 / DBL_MAX;  // inv(prevfloat(typemax(Float64)))
  a.eps_delta = DBL_EPSILON / a.delta;
  a.n_nodes = (uint32_t)net->n;
  a.n_edges = P.n_edges;
  a.S_c = P.S_c;
  a.S_l = P.S_l;
  a.S_pp = P.S_pp;
  a.S_r = P.S_r;
  a.off_node = P.off_node; a.off_edge = P.off_edge; a.off_child = P.off_child;
  a.off_taskA = P.off_taskA; a.off_taskB = P.off_taskB; a.off_lane = P.off_lane; a.off_dig = P.off_dig;
  a.off_taskP = P.off_taskP; a.off_taskPP = P.off_taskPP;
  a.nA = P.nA; a.nB = P.nB; a.nP = P.nP; a.nPP = P.nPP;
  a.out_stride = (uint32_t)out_stride;
  a.nsamples = nsamples;
  a.K = K;
  // Placement (measured on the C2 net): the kernel is latency-bound, so residency beats shared-memory
  // latency -- by default plan, CPTs and message planes all stay in global memory (L1 / L2 resident) and an SM holds
  // 4 CTAs x 8 warps = 32 instances instead of the 9 whose messages fit its shared memory.  BN_LOOPY_MODE (bit 0 plan,
  // bit 1 CPTs, bit 2 messages in shared memory), BN_LOOPY_WARPS and BN_LOOPY_CTAS reproduce the other placements.
  const size_t budget = (size_t)dp.max_smem_optin - 1024;
  const size_t plan_b = ((size_t)a.plan_words * 4 + 15) & ~(size_t)15;
  const size_t cpt_b = ((size_t)a.cpt_len * 8 + 15) & ~(size_t)15;
  const size_t per_warp = (((size_t)3 * P.S_c + 2 + (size_t)P.S_r + (size_t)P.S_pp + (size_t)2 * P.S_l) * 8 + (size_t)net->n + 15) & ~(size_t)15;
  const int warps_env = getenv("BN_LOOPY_WARPS") ? atoi(getenv("BN_LOOPY_WARPS")) : 0;
  const int ctas_env = getenv("BN_LOOPY_CTAS") ? atoi(getenv("BN_LOOPY_CTAS")) : 0;
  const int mode_env = getenv("BN_LOOPY_MODE") ? atoi(getenv("BN_LOOPY_MODE")) : -1;
  a.per_warp_bytes = (uint32_t)per_warp;
  int warps = 8, per_sm = 4;
  bool plan_smem, cpt_smem, msg_smem;
  if (mode_env >= 0) {
    plan_smem = mode_env & 1; cpt_smem = mode_env & 2; msg_smem = mode_env & 4;
  } else {
    plan_smem = cpt_smem = msg_smem = false;
  }
  size_t shared_b = (plan_smem ? plan_b : 0) + (cpt_smem ? cpt_b : 0);
  if (msg_smem && shared_b + per_warp > budget) { msg_smem = false; }
  if (shared_b > budget) { plan_smem = cpt_smem = false; shared_b = 0; }
  if (msg_smem) {
    // up to 10 warps; the warp count that keeps the most warps resident per SM (plan + CPTs are paid once per CTA)
    const int max_warps = (int)std::min<size_t>(10, (budget - shared_b) / per_warp);
    int best = 0;
    for (int w = 1; w <= max_warps; ++w) {
      const size_t sm = shared_b + (size_t)w * per_warp + 1024;
      const int resident = (int)std::min<size_t>((size_t)dp.smem_per_sm / sm, (size_t)(2048 / (w * 32))) * w;
      if (resident >= best) { best = resident; warps = w; }
    }
    if (warps_env > 0) warps = std::min(max_warps, warps_env);
  } else if (warps_env > 0) {
    warps = std::min(10, warps_env);
  }
  warps = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)warps, n_inst));
  a.shared_bytes = (uint32_t)shared_b;
  const size_t smem = shared_b + (msg_smem ? (size_t)warps * per_warp : 0);
  void (*kern)(const LoopyArgs);
  switch ((plan_smem ? 1 : 0) | (cpt_smem ? 2 : 0) | (msg_smem ? 4 : 0)) {
    case 0: kern = loopy_kernel<false, false, false, 4>; break;  // 64 registers: 4 CTAs x 8 warps per SM
    case 1: kern = loopy_kernel<true, false, false>; break;
    case 2: kern = loopy_kernel<false, true, false>; break;
    case 3: kern = loopy_kernel<true, true, false>; break;
    case 4: kern = loopy_kernel<false, false, true>; break;
    case 5: kern = loopy_kernel<true, false, true>; break;
    case 6: kern = loopy_kernel<false, true, true>; break;
    default: kern = loopy_kernel<true, true, true>; break;
  }
  BN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin));
  per_sm = (int)std::min<size_t>((size_t)per_sm, (size_t)dp.smem_per_sm / (smem + 1024));
  if (msg_smem) per_sm = (int)((size_t)dp.smem_per_sm / (smem + 1024));
  if (ctas_env > 0) per_sm = ctas_env;
  per_sm = std::max(1, std::min(per_sm, 2048 / (warps * 32)));
  TmpBuf<unsigned char> scratch;
  const uint64_t ctas = (n_inst + (uint64_t)warps - 1) / (uint64_t)warps;
  const int grid = (int)std::min<uint64_t>(ctas, (uint64_t)dp.sm_count * (uint64_t)per_sm);
  if (!msg_smem) {
    BN_CUDA_TRY(scratch.alloc((size_t)grid * (size_t)warps * per_warp));
    a.scratch = scratch.p;
  }
  kern<<<grid, warps * 32, smem, tl_stream>>>(a);  // (a stream-ordered free of `scratch` follows the kernel)
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

int32_t bn_loopy_infer_dev(bn_net_t h, const int8_t* evidence_dev, const int32_t* query_dev, uint64_t n_inst,
                           int32_t nsamples, double tol, int32_t iters_for_convergence, int32_t out_stride,
                           double* out_dev, int32_t* out_iters_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net && (n_inst == 0 || (evidence_dev && query_dev && out_dev)), BN_BAD_ARG, "bn_loopy_infer_dev: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  return loopy_launch(net, evidence_dev, query_dev, n_inst, nsamples, tol, iters_for_convergence, out_stride, out_dev,
                      out_iters_dev);
}

int32_t bn_loopy_infer(bn_net_t h, const int8_t* evidence, const int32_t* query, uint64_t n_inst, int32_t nsamples,
                       double tol, int32_t iters_for_convergence, int32_t out_stride, double* out, int32_t* out_iters) {
  Net* net = get_net(h);
  BN_REQUIRE(net && (n_inst == 0 || (query && out)), BN_BAD_ARG, "bn_loopy_infer: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  const size_t n = (size_t)net->n;
  // InferenceState checks (src/ProbabilisticGraphicalModels/inference.jl:6-14) and state bounds, per instance
  for (uint64_t b = 0; b < n_inst; ++b) {
    const int32_t q = query[b];
    BN_REQUIRE(q >= 0 && q < net->n, BN_BAD_ARG, "instance %llu: query %d is not in the network", (unsigned long long)b, q);
    BN_REQUIRE(net->card[q] <= out_stride, BN_BAD_ARG, "instance %llu: out_stride %d < %d states of the query",
               (unsigned long long)b, out_stride, net->card[q]);
    if (!evidence) continue;
    const int8_t* e = evidence + b * n;
    BN_REQUIRE(e[q] < 0, BN_BAD_ARG, "instance %llu: query %d is part of the evidence", (unsigned long long)b, q);
    for (size_t i = 0; i < n; ++i)
      BN_REQUIRE(e[i] < net->card[i], BN_BOUNDS, "instance %llu: evidence state %d of node %zu outside 0:%d",
                 (unsigned long long)b, (int)e[i], i, net->card[i] - 1);
  }
  if (n_inst == 0) return BN_OK;
  TmpBuf<int8_t> d_ev;
  TmpBuf<int32_t> d_q, d_it;
  TmpBuf<double> d_out;
  BN_CUDA_TRY(d_ev.alloc(n_inst * n));
  BN_CUDA_TRY(d_q.alloc(n_inst));
  BN_CUDA_TRY(d_it.alloc(n_inst));
  BN_CUDA_TRY(d_out.alloc(n_inst * (size_t)out_stride));
  if (evidence) BN_CUDA_TRY(cudaMemcpyAsync(d_ev.p, evidence, n_inst * n, cudaMemcpyHostToDevice, tl_stream));
  else BN_CUDA_TRY(cudaMemsetAsync(d_ev.p, 0xFF, n_inst * n, tl_stream));
  BN_CUDA_TRY(cudaMemcpyAsync(d_q.p, query, n_inst * sizeof(int32_t), cudaMemcpyHostToDevice, tl_stream));
  BN_TRY(loopy_launch(net, d_ev.p, d_q.p, n_inst, nsamples, tol, iters_for_convergence, out_stride, d_out.p, d_it.p));
  BN_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, n_inst * (size_t)out_stride * sizeof(double), cudaMemcpyDeviceToHost, tl_stream));
  if (out_iters)
    BN_CUDA_TRY(cudaMemcpyAsync(out_iters, d_it.p, n_inst * sizeof(int32_t), cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

}  // extern "C"
