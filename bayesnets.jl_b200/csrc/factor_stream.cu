// factor_stream.cu -- one-shot streaming kernels for the two Factor ops whose address pattern needs no offset tables:
//   sum(phi, dims) of ONE factor  (src/Factors/factors_dims.jl:60-85,100: reducedim(+, phi, dims), dims dropped)
//   push!(phi, dim, length)       (src/Factors/factors_main.jl:148-158: the table repeated along a new last dim)
//
// Why they exist beside contract_kernel (factor.cu): that kernel is a persistent grid whose CTAs first tabulate
// mixed-radix offsets (a few microseconds) and then walk tiles; on a 30-40 us streaming op the prologue, the static
// tile schedule's tail and ~80 instructions per 16 bytes of output cost 10-25% of the HBM roofline (S4 0.68, D1 0.75,
// S1-S3 0.83 against a same-size device copy at 0.925; profiles/r2a_bench.json).  A sum of one factor over a few dims
// needs none of that: the input index space is [K0][C1][K1][C2]...[Cm][Km] (Ki = product of the kept dims between two
// summed dims, Ci = a summed dim), so an output index splits into <= m + 1 segment digits by m divisions (shifts on
// binary nets) and every summed state is a constant offset.  Each thread owns two 16-byte outputs and issues all of
// their loads (<= 16 states each) before the first add: one-shot CTAs, no shared memory, no tables.
//
// Arithmetic is the reference's: states are added in column-major order of the summed dims (first summed dim fastest),
// exactly the order contract_kernel uses, so results are bit-identical to it and to the oracle (tests use array_equal).
#include "common.cuh"

#include <algorithm>

namespace bn {

constexpr int SEG_MAX = 6;      // kept segments (= summed dims + 1)
constexpr int SEG_MAX_NS = 16;  // joint states of the summed dims held in flight
constexpr int SEG_THR = 256;

struct SegSumParams {
  const double* in;
  double* out;
  uint32_t n_units;             // output units: double2 (pairs along the first kept segment) or double
  int32_t nseg;                 // kept segments
  uint32_t K[SEG_MAX];          // segment sizes in units for segment 0, in elements for the others
  int32_t shift[SEG_MAX];       // log2(K) or -1
  uint32_t instride[SEG_MAX];   // input stride (elements) of one step in segment i (segment 0: per unit)
  uint32_t ns;                  // joint summed states (pairs mode: ns / 2 loads)
  int32_t soff[SEG_MAX_NS];     // input offset of every joint state, first summed dim fastest
  uint32_t pair_step;           // pairs mode: input distance (elements) between the two outputs of a unit = C1
};

__device__ __forceinline__ uint32_t seg_offset(const SegSumParams& p, uint32_t u) {
  uint32_t off = 0, r = u;
#pragma unroll
  for (int i = 0; i < SEG_MAX; ++i) {
    if (i < p.nseg) {
      uint32_t q, d;
      if (i == p.nseg - 1) { q = 0; d = r; }
      else if (p.shift[i] >= 0) { q = r >> p.shift[i]; d = r & (p.K[i] - 1u); }
      else { q = r / p.K[i]; d = r - q * p.K[i]; }
      off += d * p.instride[i];
      r = q;
    }
  }
  return off;
}

// PAIRS = false: a unit is a double2 of outputs along the first kept segment (K0 even); every state is one 128-bit load.
// PAIRS = true : the innermost input dim is summed (K0 = 1, C1 even) and the next kept segment K1 is even: a unit is the
//                output pair (2v, 2v + 1) of K1, whose inputs are 2 * C1 consecutive doubles; a 128-bit load carries
//                states (2j, 2j + 1) of one output, the second output's loads sit `pair_step` = C1 elements further.
// JU units per thread, chosen so that >= 8 independent 128-bit loads are in flight per thread.
template <bool PAIRS, int JU>
__global__ void __launch_bounds__(SEG_THR) seg_sum_kernel(const SegSumParams p) {
  const uint32_t u0 = (blockIdx.x * JU) * SEG_THR + threadIdx.x;
  uint32_t off[JU];
  bool live[JU];
#pragma unroll
  for (int j = 0; j < JU; ++j) {
    const uint32_t u = u0 + j * SEG_THR;
    live[j] = u < p.n_units;
    off[j] = seg_offset(p, live[j] ? u : 0u);
  }
  const uint32_t nl = PAIRS ? p.ns / 2u : p.ns;  // 128-bit loads per output column
  constexpr int G = (PAIRS ? 4 : 8) / JU;        // load groups: JU x G x (PAIRS ? 2 : 1) = 8 loads before the first add
  double2 acc[JU];
  for (uint32_t s = 0; s < nl; s += G) {
    double2 v[JU][PAIRS ? 2 * G : G];
#pragma unroll
    for (int j = 0; j < JU; ++j)
#pragma unroll
      for (uint32_t t = 0; t < (uint32_t)G; ++t)
        if (s + t < nl) {
          const double* a = p.in + off[j] + p.soff[PAIRS ? 2u * (s + t) : s + t];
          if (PAIRS) {
            v[j][2 * t] = __ldg(reinterpret_cast<const double2*>(a));
            v[j][2 * t + 1] = __ldg(reinterpret_cast<const double2*>(a + p.pair_step));
          } else {
            v[j][t] = __ldg(reinterpret_cast<const double2*>(a));
          }
        }
#pragma unroll
    for (int j = 0; j < JU; ++j)
#pragma unroll
      for (uint32_t t = 0; t < (uint32_t)G; ++t)
        if (s + t < nl) {
          if (PAIRS) {
            acc[j].x = (s + t == 0u) ? v[j][2 * t].x : acc[j].x + v[j][2 * t].x;
            acc[j].x += v[j][2 * t].y;
            acc[j].y = (s + t == 0u) ? v[j][2 * t + 1].x : acc[j].y + v[j][2 * t + 1].x;
            acc[j].y += v[j][2 * t + 1].y;
          } else {
            if (s + t == 0u) acc[j] = v[j][t];
            else { acc[j].x += v[j][t].x; acc[j].y += v[j][t].y; }
          }
        }
  }
#pragma unroll
  for (int j = 0; j < JU; ++j) {
    const uint32_t u = u0 + j * SEG_THR;
    if (live[j]) reinterpret_cast<double2*>(p.out)[u] = acc[j];
  }
}

// sum of ONE factor over the dims flagged in `summed` (card / summed in the operand's dim order).  *handled = false: the
// shape does not fit this kernel (odd leading segment, > 16 joint states, ...) and nothing was launched.
int32_t sum_segmented(const double* in, const std::vector<int32_t>& card, const std::vector<char>& summed, double* out,
                      bool* handled) {
  *handled = false;
  if (getenv("BN_FACTOR_NO_STREAM")) return BN_OK;
  const size_t nd = card.size();
  SegSumParams p{};
  // segments: [K0][C1][K1]...[Cm][Km]
  std::vector<uint64_t> K(1, 1), Cs;
  for (size_t d = 0; d < nd; ++d) {
    if (summed[d]) { Cs.push_back((uint64_t)card[d]); K.push_back(1); }
    else K.back() *= (uint64_t)card[d];
  }
  const size_t m = Cs.size();
  if (m == 0 || m + 1 > (size_t)SEG_MAX) return BN_OK;
  uint64_t ns = 1, n_in = 1, n_out = 1;
  for (uint64_t c : Cs) ns *= c;
  for (uint64_t k : K) n_out *= k;
  n_in = n_out * ns;
  if (ns > (uint64_t)SEG_MAX_NS || n_in >= (1ull << 31)) return BN_OK;
  const bool pairs = K[0] == 1;
  if (pairs ? (Cs[0] % 2 != 0 || K[1] % 2 != 0) : (K[0] % 2 != 0)) return BN_OK;
  // input strides of the summed dims and of the kept segments
  uint64_t stride = 1;
  std::vector<uint64_t> seg_stride(m + 1), sum_stride(m);
  for (size_t i = 0; i <= m; ++i) {
    seg_stride[i] = stride;
    stride *= K[i];
    if (i < m) { sum_stride[i] = stride; stride *= Cs[i]; }
  }
  for (uint64_t s = 0; s < ns; ++s) {  // first summed dim fastest
    uint64_t r = s, o = 0;
    for (size_t i = 0; i < m; ++i) { o += (r % Cs[i]) * sum_stride[i]; r /= Cs[i]; }
    p.soff[s] = (int32_t)o;
  }
  p.ns = (uint32_t)ns;
  // kept segments as digit groups of the unit index; drop size-1 segments (their digit is always 0)
  int nseg = 0;
  for (size_t i = 0; i <= m; ++i) {
    uint64_t k = K[i], st = seg_stride[i];
    if (i == (pairs ? 1u : 0u)) { k /= 2; st *= 2; }  // a unit is a double2 of outputs along this segment
    if (k == 1 && !(nseg == 0 && i == m)) continue;
    p.K[nseg] = (uint32_t)k;
    p.instride[nseg] = (uint32_t)st;
    p.shift[nseg] = (k & (k - 1)) == 0 ? (int32_t)__builtin_ctzll(k) : -1;
    ++nseg;
  }
  if (nseg == 0) { p.K[0] = 1; p.instride[0] = 0; p.shift[0] = 0; nseg = 1; }
  p.nseg = nseg;
  const uint64_t n_units = n_out / 2;
  if (n_units >= (1ull << 32)) return BN_OK;
  p.n_units = (uint32_t)n_units;
  p.pair_step = pairs ? (uint32_t)Cs[0] : 0u;
  p.in = in;
  p.out = out;
  {  // the largest input element any thread touches must lie inside the table (a wrong plan falls back, never reads out)
    uint64_t hi = 0;
    for (int i = 0; i < p.nseg; ++i) hi += (uint64_t)(p.K[i] - 1u) * p.instride[i];
    int32_t smax = 0;
    for (uint32_t st = 0; st < p.ns; st += (pairs ? 2u : 1u)) smax = std::max(smax, p.soff[st]);  // states a load starts at
    hi += (uint64_t)smax + (pairs ? (uint64_t)p.pair_step : 0ull) + 1ull;
    if (hi >= n_in) return BN_OK;
  }
  const uint32_t nl = pairs ? p.ns / 2u : p.ns;
  const int ju = (pairs ? nl < 2u : nl < 3u) ? 4 : 2;
  const unsigned grid = (unsigned)((n_units + (uint64_t)SEG_THR * ju - 1) / ((uint64_t)SEG_THR * ju));
  if (pairs) {
    if (ju == 4) seg_sum_kernel<true, 4><<<grid, SEG_THR, 0, tl_stream>>>(p);
    else seg_sum_kernel<true, 2><<<grid, SEG_THR, 0, tl_stream>>>(p);
  } else {
    if (ju == 4) seg_sum_kernel<false, 4><<<grid, SEG_THR, 0, tl_stream>>>(p);
    else seg_sum_kernel<false, 2><<<grid, SEG_THR, 0, tl_stream>>>(p);
  }
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return BN_OK;
}

// push!: out[o + n_in * r] = in[o], r < length.  One 128-bit load, `length` 128-bit stores per thread.
__global__ void __launch_bounds__(SEG_THR) dup_kernel(const double2* __restrict__ in, double2* __restrict__ out, uint32_t n2,
                                                      uint32_t length) {
  const uint32_t i0 = (blockIdx.x * 4u) * SEG_THR + threadIdx.x;
  double2 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * SEG_THR < n2) v[j] = __ldg(in + i0 + j * SEG_THR);
  for (uint32_t r = 0; r < length; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j * SEG_THR < n2) out[(size_t)r * n2 + i0 + j * SEG_THR] = v[j];
}

int32_t push_duplicate(const double* in, long long n_in, int32_t length, double* out, bool* handled) {
  *handled = false;
  if (getenv("BN_FACTOR_NO_STREAM")) return BN_OK;
  if (n_in % 2 || n_in >= (1ll << 32) || (reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16) return BN_OK;
  const uint32_t n2 = (uint32_t)(n_in / 2);
  const unsigned grid = (n2 + 4u * SEG_THR - 1) / (4u * SEG_THR);
  dup_kernel<<<grid, SEG_THR, 0, tl_stream>>>(reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), n2,
                                               (uint32_t)length);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return BN_OK;
}

}  // namespace bn
