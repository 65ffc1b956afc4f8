// factor_ops.cu -- the rest of the Factor algebra: join with / + - (and *), reducedim with * max min (and +),
// normalize!(phi, dims), broadcast!(op, phi, dims, values).
//
// These are the operations of src/Factors/factors_dims.jl that variable elimination does NOT use (it needs `*`, `sum`
// and normalize!, which have the tuned contraction kernel of factor.cu); LoopyBelief is their only caller in the
// reference and has its own kernel (loopy.cu).  They exist so that a Factor never has to leave the device: one
// gather kernel for the elementwise ops, one for the reductions, both memory-bound, mixed-radix addressing with the
// dims merged on the host wherever operands and output are contiguous together.  Arithmetic order is the oracle's:
// op(phi1, phi2) AFTER the reference's "larger factor first" swap (factors_dims.jl:189-191 -- so `/` and `-` divide /
// subtract in the swapped order when the left factor is the smaller one, as the reference does), reductions
// accumulate sequentially over the reduced dims in column-major order, n-ary broadcast is a left fold.
#include <algorithm>

#include "common.cuh"

namespace bn {

namespace {

constexpr int GD = 48;   // merged loop dims
constexpr int GTHR = 256;

struct GatherParams {
  int nd;                 // output loop dims (merged)
  int nr;                 // reduced loop dims (merged)
  int32_t card[GD];       // output dims then reduced dims
  long long sa[GD], sb[GD];  // element strides of operand a / b along each loop dim (0 = absent)
  long long n_out, n_red;
  const double* a;
  const double* b;
  double* out;
  int op, pre;
};

__device__ __forceinline__ double apply_op(int op, double x, double y) {
  switch (op) {
    case BN_OP_MUL: return x * y;
    case BN_OP_DIV: return x / y;
    case BN_OP_ADD: return x + y;
    case BN_OP_SUB: return x - y;
    case BN_OP_MAX: return (x != x) ? x : ((y != y) ? y : fmax(x, y));  // Base.max propagates NaN
    default: return (x != x) ? x : ((y != y) ? y : fmin(x, y));
  }
}

__device__ __forceinline__ double apply_pre(int pre, double x) { return pre == 1 ? fabs(x) : (pre == 2 ? x * x : x); }

// out[o] = op(a[off_a(o)], b[off_b(o)])          (out may be a itself: every element is read once, then written)
__global__ void __launch_bounds__(GTHR) join_kernel(const GatherParams p) {
  for (long long o = (long long)blockIdx.x * GTHR + threadIdx.x; o < p.n_out; o += (long long)gridDim.x * GTHR) {
    long long r = o, oa = 0, ob = 0;
    for (int d = 0; d < p.nd; ++d) {
      const long long q = r / p.card[d];
      const long long dg = r - q * p.card[d];
      oa += dg * p.sa[d];
      ob += dg * p.sb[d];
      r = q;
    }
    p.out[o] = apply_op(p.op, p.a[oa], p.b[ob]);
  }
}

// out[o] = fold over the reduced dims, column-major, of pre(a[...])
__global__ void __launch_bounds__(GTHR) reduce_kernel(const GatherParams p) {
  for (long long o = (long long)blockIdx.x * GTHR + threadIdx.x; o < p.n_out; o += (long long)gridDim.x * GTHR) {
    long long r = o, base = 0;
    for (int d = 0; d < p.nd; ++d) {
      const long long q = r / p.card[d];
      base += (r - q * p.card[d]) * p.sa[d];
      r = q;
    }
    double acc = 0.0;
    for (long long s = 0; s < p.n_red; ++s) {
      long long rr = s, off = base;
      for (int d = 0; d < p.nr; ++d) {
        const long long q = rr / p.card[p.nd + d];
        off += (rr - q * p.card[p.nd + d]) * p.sa[p.nd + d];
        rr = q;
      }
      const double v = apply_pre(p.pre, p.a[off]);
      acc = s ? apply_op(p.op, acc, v) : v;
    }
    p.out[o] = acc;
  }
}

// setindex!(phi, v, assignment): phi[slab] = v (one value) or .= v (an array over the slab, column-major).
// o runs over the slab's cells; sa = strides of phi along the free dims, base = offset of the pinned states.
__global__ void __launch_bounds__(GTHR) slab_store_kernel(const GatherParams p, long long base, double scalar) {
  for (long long o = (long long)blockIdx.x * GTHR + threadIdx.x; o < p.n_out; o += (long long)gridDim.x * GTHR) {
    long long r = o, off = base;
    for (int d = 0; d < p.nd; ++d) {
      const long long q = r / p.card[d];
      off += (r - q * p.card[d]) * p.sa[d];
      r = q;
    }
    p.out[off] = p.b ? p.b[o] : scalar;
  }
}

int find(const std::vector<int32_t>& v, int32_t x) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == x) return (int)i;
  return -1;
}

std::vector<long long> strides_of(const std::vector<int32_t>& card) {
  std::vector<long long> st(card.size(), 1);
  for (size_t i = 1; i < card.size(); ++i) st[i] = st[i - 1] * card[i - 1];
  return st;
}

// append loop dims [card, sa, sb] to p starting at slot `at`, merging neighbours that are contiguous in both operands
int push_dims(GatherParams& p, int at, const std::vector<int32_t>& card, const std::vector<long long>& sa,
              const std::vector<long long>& sb) {
  int n = 0;
  for (size_t i = 0; i < card.size(); ++i) {
    if (card[i] == 1) continue;
    if (n > 0) {
      const int l = at + n - 1;
      if (sa[i] == p.sa[l] * p.card[l] && sb[i] == p.sb[l] * p.card[l] && (long long)p.card[l] * card[i] < (1ll << 30)) {
        p.card[l] *= card[i];
        continue;
      }
    }
    p.card[at + n] = card[i];
    p.sa[at + n] = sa[i];
    p.sb[at + n] = sb[i];
    ++n;
  }
  return n;
}

int grid_for(int device, long long n) {
  const long long want = (n + GTHR - 1) / GTHR;
  return (int)std::max<long long>(1, std::min<long long>(want, (long long)dev_props(device).sm_count * 8));
}

// out (dims od/oc, may alias a) = op(a laid out over (ad, ac), b over (bd, bc)); every dim of a and b is in od
int32_t launch_join(int device, const double* a, const std::vector<int32_t>& ad, const std::vector<int32_t>& ac,
                    const double* b, const std::vector<int32_t>& bd, const std::vector<int32_t>& bc,
                    const std::vector<int32_t>& od, const std::vector<int32_t>& oc, double* out, int op) {
  BN_REQUIRE((int)od.size() <= GD, BN_UNSUPPORTED, "factor op over %zu variables (max %d)", od.size(), GD);
  GatherParams p{};
  const std::vector<long long> ast = strides_of(ac), bst = strides_of(bc);
  std::vector<long long> sa(od.size(), 0), sb(od.size(), 0);
  long long n = 1;
  for (size_t i = 0; i < od.size(); ++i) {
    const int ia = find(ad, od[i]), ib = find(bd, od[i]);
    if (ia >= 0) sa[i] = ast[(size_t)ia];
    if (ib >= 0) sb[i] = bst[(size_t)ib];
    n *= oc[i];
  }
  p.nd = push_dims(p, 0, oc, sa, sb);
  p.nr = 0;
  p.n_out = n;
  p.n_red = 1;
  p.a = a; p.b = b; p.out = out; p.op = op; p.pre = 0;
  join_kernel<<<grid_for(device, n), GTHR, 0, tl_stream>>>(p);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

int32_t launch_reduce(int device, const Factor* f, const std::vector<int32_t>& rdims, int op, int pre, double* out,
                      std::vector<int32_t>* od, std::vector<int32_t>* oc) {
  BN_REQUIRE((int)f->dims.size() <= GD, BN_UNSUPPORTED, "factor op over %zu variables (max %d)", f->dims.size(), GD);
  const std::vector<long long> st = strides_of(f->card);
  std::vector<int32_t> kc, rc;
  std::vector<long long> ks, rs, zero;
  od->clear(); oc->clear();
  long long n_out = 1, n_red = 1;
  for (size_t i = 0; i < f->dims.size(); ++i) {
    if (find(rdims, f->dims[i]) >= 0) { rc.push_back(f->card[i]); rs.push_back(st[i]); n_red *= f->card[i]; }
    else { kc.push_back(f->card[i]); ks.push_back(st[i]); od->push_back(f->dims[i]); oc->push_back(f->card[i]); n_out *= f->card[i]; }
  }
  GatherParams p{};
  p.nd = push_dims(p, 0, kc, ks, ks);
  p.nr = push_dims(p, p.nd, rc, rs, rs);
  p.n_out = n_out;
  p.n_red = n_red;
  p.a = f->data.p; p.b = nullptr; p.out = out; p.op = op; p.pre = pre;
  reduce_kernel<<<grid_for(device, n_out), GTHR, 0, tl_stream>>>(p);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

int32_t check_dims(const Factor* f, const int32_t* dims, int32_t n, std::vector<int32_t>* out) {
  BN_REQUIRE(n >= 0 && (n == 0 || dims), BN_BAD_ARG, "bad dims argument");
  for (int j = 0; j < n; ++j) {
    BN_REQUIRE(find(f->dims, dims[j]) >= 0, BN_NOT_IN_FACTOR, "%d is not a valid dimension", dims[j]);
    out->push_back(dims[j]);
  }
  return BN_OK;
}

}  // namespace

}  // namespace bn

using namespace bn;

extern "C" {

int32_t bn_factor_join(bn_factor_t h1, bn_factor_t h2, int32_t op, bn_factor_t* out) {
  Factor *f1 = get_factor(h1), *f2 = get_factor(h2);
  BN_REQUIRE(f1 && f2 && out, BN_BAD_ARG, "bn_factor_join: bad arguments");
  BN_REQUIRE(op >= BN_OP_MUL && op <= BN_OP_SUB, BN_BAD_ARG, "bn_factor_join: op %d is not one of * / + -", op);
  BN_REQUIRE(f1->device == f2->device, BN_BAD_ARG, "factors live on different devices");
  BN_CUDA_TRY(cudaSetDevice(f1->device));
  if (f1->len < f2->len) std::swap(f1, f2);  // factors_dims.jl:189-191, before op is applied
  std::vector<int32_t> od = f1->dims, oc = f1->card;
  for (size_t j = 0; j < f2->dims.size(); ++j) {
    const int pos = find(f1->dims, f2->dims[j]);
    if (pos >= 0)
      BN_REQUIRE(f1->card[(size_t)pos] == f2->card[j], BN_DIM_MISMATCH,
                 "Common dimensions must have same size (dim %d: %d vs %d)", f2->dims[j], f1->card[(size_t)pos], f2->card[j]);
  }
  for (size_t j = 0; j < f2->dims.size(); ++j)
    if (find(f1->dims, f2->dims[j]) < 0) { od.push_back(f2->dims[j]); oc.push_back(f2->card[j]); }
  Factor* o = nullptr;
  BN_TRY(new_factor(od, oc, f1->device, &o));
  const int32_t rc = launch_join(f1->device, f1->data.p, f1->dims, f1->card, f2->data.p, f2->dims, f2->card, od, oc, o->data.p, op);
  if (rc != BN_OK) { delete o; return rc; }
  *out = register_factor(o);
  return BN_OK;
}

int32_t bn_factor_reduce(bn_factor_t h, const int32_t* dims, int32_t n, int32_t op, bn_factor_t* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out, BN_BAD_ARG, "bn_factor_reduce: bad arguments");
  BN_REQUIRE(op == BN_OP_ADD || op == BN_OP_MUL || op == BN_OP_MAX || op == BN_OP_MIN, BN_BAD_ARG,
             "bn_factor_reduce: op %d is not one of + * max min", op);
  std::vector<int32_t> rd;
  BN_TRY(check_dims(f, dims, n, &rd));
  BN_CUDA_TRY(cudaSetDevice(f->device));
  std::vector<int32_t> od, oc;
  for (size_t i = 0; i < f->dims.size(); ++i)
    if (find(rd, f->dims[i]) < 0) { od.push_back(f->dims[i]); oc.push_back(f->card[i]); }
  Factor* o = nullptr;
  BN_TRY(new_factor(od, oc, f->device, &o));
  std::vector<int32_t> d2, c2;
  const int32_t rc = launch_reduce(f->device, f, rd, op, 0, o->data.p, &d2, &c2);
  if (rc != BN_OK) { delete o; return rc; }
  *out = register_factor(o);
  return BN_OK;
}

int32_t bn_factor_normalize_dims(bn_factor_t h, const int32_t* dims, int32_t n, int32_t pnorm) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_normalize_dims: unknown factor");
  BN_REQUIRE(pnorm == 1 || pnorm == 2, BN_BAD_ARG, "p = %d is not supported", pnorm);
  std::vector<int32_t> rd;
  BN_TRY(check_dims(f, dims, n, &rd));
  if (rd.empty()) return BN_OK;  // factors_dims.jl:32 `if !isempty(inds)`
  BN_CUDA_TRY(cudaSetDevice(f->device));
  // total = sum(abs | abs2, potential, dims): a factor over the remaining dims; then potential ./= total
  std::vector<int32_t> td, tc;
  for (size_t i = 0; i < f->dims.size(); ++i)
    if (find(rd, f->dims[i]) < 0) { td.push_back(f->dims[i]); tc.push_back(f->card[i]); }
  Factor* t = nullptr;
  BN_TRY(new_factor(td, tc, f->device, &t));
  std::vector<int32_t> d2, c2;
  int32_t rc = launch_reduce(f->device, f, rd, BN_OP_ADD, pnorm, t->data.p, &d2, &c2);
  if (rc == BN_OK)
    rc = launch_join(f->device, f->data.p, f->dims, f->card, t->data.p, td, tc, f->dims, f->card, f->data.p, BN_OP_DIV);
  delete t;  // stream-ordered free: runs after the kernels above
  return rc;
}

int32_t bn_factor_broadcast(bn_factor_t h, const int32_t* dims, int32_t n, const double* values, const int32_t* value_len,
                            int32_t op) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && n >= 0 && (n == 0 || (dims && values && value_len)), BN_BAD_ARG, "bn_factor_broadcast: bad arguments");
  BN_REQUIRE(op >= BN_OP_MUL && op <= BN_OP_MIN, BN_BAD_ARG, "bn_factor_broadcast: unknown op %d", op);
  std::vector<int32_t> bd;
  for (int j = 0; j < n; ++j)
    for (int j2 = 0; j2 < j; ++j2) BN_REQUIRE(dims[j] != dims[j2], BN_BAD_ARG, "Dimensions must be unique");
  BN_TRY(check_dims(f, dims, n, &bd));
  size_t total = 0;
  for (int j = 0; j < n; ++j) {
    const int32_t c = f->card[(size_t)find(f->dims, dims[j])];
    BN_REQUIRE(value_len[j] == 1 || value_len[j] == c, BN_DIM_MISMATCH,
               "broadcast value for dimension %d has length %d, dimension has %d", dims[j], value_len[j], c);
    total += (size_t)value_len[j];
  }
  if (n == 0) return BN_OK;
  BN_CUDA_TRY(cudaSetDevice(f->device));
  TmpBuf<double> vals;
  BN_CUDA_TRY(vals.upload(values, total, tl_stream));
  size_t at = 0;
  for (int j = 0; j < n; ++j) {  // f(phi, v1, v2, ...) is a left fold: one in-place pass per value
    std::vector<int32_t> vd, vc;
    if (value_len[j] > 1) { vd.push_back(dims[j]); vc.push_back(value_len[j]); }
    BN_TRY(launch_join(f->device, f->data.p, f->dims, f->card, vals.p + at, vd, vc, f->dims, f->card, f->data.p, op));
    at += (size_t)value_len[j];
  }
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));  // `values` is a host buffer the caller may reuse
  return BN_OK;
}

// setindex!(phi, v, a::Assignment): src/Factors/factors_access.jl:39-46 with _translate_index (:48-72): dims of the
// assignment that phi lacks are ignored, unassigned dims are Colons.  n_values == 1 stores one value into every cell of
// the slab (`phi.potential[idx...] = v` / `.= v`), n_values == the slab's size stores an array over the free dims in
// column-major order (`.= v` with an array of the slab's shape); anything else is the DimensionMismatch `.=` throws.
int32_t bn_factor_setindex(bn_factor_t h, const int32_t* dims, const int32_t* states, int32_t n_assign,
                           const double* values, int64_t n_values) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && values && n_values >= 1, BN_BAD_ARG, "bn_factor_setindex: bad arguments");
  BN_REQUIRE(n_assign == 0 || (dims && states), BN_BAD_ARG, "bn_factor_setindex: NULL assignment");
  const std::vector<long long> fst = strides_of(f->card);
  long long base = 0, slab = 1;
  std::vector<int32_t> fc;
  std::vector<long long> fs, zero;
  for (size_t i = 0; i < f->dims.size(); ++i) {
    int hit = -1;
    for (int j = 0; j < n_assign; ++j) if (dims[j] == f->dims[i]) hit = j;
    if (hit >= 0) {
      BN_REQUIRE(states[hit] >= 0 && states[hit] < f->card[i], BN_BOUNDS, "BoundsError: state %d for dim %d (size %d)",
                 states[hit], f->dims[i], f->card[i]);
      base += fst[i] * states[hit];
    } else {
      fc.push_back(f->card[i]);
      fs.push_back(fst[i]);
      zero.push_back(0);
      slab *= f->card[i];
    }
  }
  BN_REQUIRE(n_values == 1 || n_values == slab, BN_DIM_MISMATCH,
             "DimensionMismatch: cannot broadcast %lld values into a slab of %lld cells", (long long)n_values, slab);
  BN_REQUIRE((int)fc.size() <= GD, BN_UNSUPPORTED, "factor op over %zu variables (max %d)", fc.size(), GD);
  BN_CUDA_TRY(cudaSetDevice(f->device));
  GatherParams p{};
  p.nd = push_dims(p, 0, fc, fs, zero);
  p.n_out = slab;
  p.out = f->data.p;
  TmpBuf<double> vals;
  if (n_values > 1) {
    BN_CUDA_TRY(vals.upload(values, (size_t)n_values, tl_stream));
    p.b = vals.p;
  }
  slab_store_kernel<<<grid_for(f->device, slab), GTHR, 0, tl_stream>>>(p, base, values[0]);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  if (n_values > 1) BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));  // `values` is a host buffer the caller may reuse
  return BN_OK;
}

}  // extern "C"
