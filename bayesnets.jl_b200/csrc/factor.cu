// factor.cu -- dense fp64 factor algebra on the device: product (outer join), marginalise, slice,
// normalise and the fused n-ary "multiply-then-sum-out" elimination step.
//
// Replaces join/`*` (src/Factors/factors_dims.jl:185-234,269), reducedim/sum (:60-85,100),
// normalize! (:45-57), getindex(::Assignment) (src/Factors/factors_access.jl:19-35) and the
// `sum(reduce(*, contain_h), h)` step of variable elimination (src/Inference/exact.jl:27).
//
// ONE contraction kernel covers product / sum / slice / fused eliminate:
//     out[o] = sum_{s in S} prod_k F_k[ off_k(o) + off_k(s) ]
// The loop space is split   o = lo + Rlo * hi :
//   lo  = the first output dims (Rlo ~ 256..1024 contiguous outputs; also contiguous in the larger
//         operand, whose dims lead the union -- src/Factors/factors_dims.jl:203), one lo per thread,
//         its per-factor offsets held in registers;
//   hi  = the remaining output dims, H values per CTA tile; per-factor offsets of the tile's hi values
//         are computed once into shared memory and read back as warp-uniform broadcasts;
//   s   = the summed dims, offsets likewise tabulated in shared memory; accumulation runs in
//         column-major order like Base's mapreducedim!, products fold left to right like reduce(*).
// Every operand element is touched exactly once per use, output/leading-operand traffic is fully
// coalesced and 128-bit vectorised (double2 along the first output dim, or along the first summed
// dim when that is the contiguous one).  All of it is HBM-bound: 8*(sum N_k + N_out) bytes.
#include "common.cuh"

#include <algorithm>

namespace bn {

constexpr int MAXD = 48;  // loop dims (output + summed)
constexpr int MAXF = 6;   // operands per contraction
constexpr int MAX_NS = 2048;
constexpr int NTHR = 256;
constexpr int TLEV = 3;     // tile-offset table levels
constexpr int TLEV_MAX = 256;

// Loop dims are ordered [lo | j | t0 | t1 | t2 | s]:
//   lo: leading output dims, Rlo entries = V vector columns of W outputs; a thread owns ONE column for the
//       whole kernel, so its per-operand lo offsets live in registers
//   j : H rows per tile; G = NTHR / V row groups work on rows g, g+G, g+2G, ...  Output dims the LEADING operand
//       does not have are preferred here: rows that differ only in such a dim re-read the same leading-operand
//       bytes back to back (L1 hits) instead of once per value a whole factor later (HBM re-read)
//   t : remaining output dims = the tile index, split into <= 3 digit groups of <= 256 values each; per-operand
//       offsets of every digit value are tabulated once per CTA (no per-tile divisions by dim, no barriers)
//   s : summed dims, Ns joint states (offset table in smem), accumulated first-dim-fastest
struct ContractParams {
  int32_t nf, nlo, nj, ns;
  int32_t nt[TLEV], L[TLEV];       // dims / values per tile-digit group (L = 1 when unused)
  int32_t card[MAXD];
  uint32_t magic[MAXD];            // ceil(2^20 / card): digit extraction without division (numerators < 1024)
  uint32_t pmagic[MAXD];           // ceil(2^20 / product of the earlier cards of the dim's group): all digits of a
                                   // group value e < 1024 come out independently (no serial divide chain)
  int32_t fstride[MAXF + 1][MAXD]; // element stride of each operand along each loop dim (0 = absent);
                                   // row nf = the OUTPUT's strides (its rows need not be consecutive)
  const double* f[MAXF];
  double* out;
  long long n_tiles;
  int32_t H, Ns, V, G;
  int32_t lead, pf_bytes, pf_ns;  // TMA L2 prefetch of the leading operand's next-tile rows (pf_bytes = 0: off)
  int32_t vec128[MAXF];  // operand is contiguous + 16-byte aligned along the vector dim: one 128-bit load
  int32_t vstep[MAXF];   // otherwise two 64-bit loads `vstep` elements apart (0 = broadcast)
  int32_t waves;  // grid = waves x resident CTAs
  int32_t rowvec[MAXF];  // the tile's 4 rows are 4 CONSECUTIVE elements of this operand (its two fastest dims were
                         // chosen as the row dims): one thread fetches them as 2 x 128-bit instead of 4 x 64-bit
};

enum { VEC_NONE = 0, VEC_LO = 1, VEC_S = 2 };

// All offsets are 32-bit element indices (every operand and the output hold < 2^31 doubles = 16 GiB).
__device__ __forceinline__ double2 ld2(const double* __restrict__ f, int32_t off, int vec128, int32_t vstep) {
  double2 r;
  if (vec128) {
    r = __ldg(reinterpret_cast<const double2*>(f + off));
  } else {
    r.x = __ldg(f + off);
    r.y = __ldg(f + off + vstep);
  }
  return r;
}

// offsets of every joint value of loop dims [d0, d0 + nd) for each of NSLOT operands -> tab[k * count + e];
// count <= 1024, so digit = e - (e * magic >> 20) * card is exact (error bound 2^-10 < 1/255)
template <int NSLOT>
__device__ __forceinline__ void fill_table(int32_t* tab, int count, int d0, int nd, const ContractParams& p) {
  for (int e = threadIdx.x; e < count; e += NTHR) {
    uint32_t r = (uint32_t)e;
    int32_t off[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) off[k] = 0;
    for (int d = 0; d < nd; ++d) {
      const uint32_t q = (r * p.magic[d0 + d]) >> 20;
      const int32_t dg = (int32_t)(r - q * (uint32_t)p.card[d0 + d]);
      r = q;
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) off[k] += dg * p.fstride[k][d0 + d];
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) tab[k * count + e] = off[k];
  }
}

// offsets of ONE joint value e < 1024 of loop dims [d0, d0 + nd), every digit extracted independently
// (floor(e / P_d) by multiply-shift with pmagic, exact for e < 1024 and P_d <= 1024): no serial divide chain, so a
// one-shot CTA reaches its first load after ~100 cycles
template <int NSLOT>
__device__ __forceinline__ void add_offsets(int32_t (&off)[NSLOT], uint32_t e, int d0, int nd, const ContractParams& p) {
  for (int d = d0; d < d0 + nd; ++d) {
    const uint32_t q = (e * p.pmagic[d]) >> 20;
    const int32_t dg = (int32_t)(q - ((q * p.magic[d]) >> 20) * (uint32_t)p.card[d]);
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) off[k] += dg * p.fstride[k][d];
  }
}

// rows in flight per thread: keep NF * (loads per row) * JU independent loads within the register budget
// (<= 12 double2 or <= 16 double in flight)
template <int NF, int VEC, int NS>
struct RowsInFlight {
  static constexpr int per_row = NF * (NS > 1 ? (VEC == VEC_S ? NS / 2 : NS) : 1);
  static constexpr int budget = (VEC == VEC_NONE) ? 16 : 12;
  static constexpr int value = per_row * 4 <= budget ? 4 : (per_row * 2 <= budget ? 2 : 1);
};

// NS: 0 = plain product (nothing summed), 1 = run-time loop over p.Ns, 2..4 = that many joint states, unrolled
template <int NF, int VEC, int NS>
__global__ void __launch_bounds__(NTHR, (NF <= 2 ? 4 : (NF == 3 ? 3 : 2))) contract_kernel(const ContractParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NS1 = NF + 1;  // table slots: the operands, then the output
  int32_t* jOff = reinterpret_cast<int32_t*>(smem_raw);  // [NF+1][H]
  int32_t* sOff = jOff + NS1 * p.H;                       // [NF][Ns]
  const int L0 = p.L[0], L1 = p.L[1], L2 = p.L[2];
  int32_t* tOff0 = sOff + NF * p.Ns;                      // [NF+1][L0]
  int32_t* tOff1 = tOff0 + NS1 * L0;
  int32_t* tOff2 = tOff1 + NS1 * L1;
  const int tid = threadIdx.x;
  constexpr int W = (VEC == VEC_LO) ? 2 : 1;  // outputs per thread per row
  constexpr int JU = RowsInFlight<NF, VEC, NS>::value;
  constexpr bool SUM = NS != 0;

  // ---- one-time tables
  const int jb = p.nlo, t0b = jb + p.nj, t1b = t0b + p.nt[0], t2b = t1b + p.nt[1], sb = t2b + p.nt[2];
  fill_table<NS1>(jOff, p.H, jb, p.nj, p);
  if (SUM) fill_table<NF>(sOff, p.Ns, sb, p.ns, p);
  // Tiles are dealt round-robin over p.waves waves of resident CTAs (all CTAs sweep one moving window of memory:
  // contiguous per-CTA chunks open ~1200 DRAM streams at once and reach 54% of HBM peak instead of 77%); the three
  // tile digits advance by the grid size as a mixed-radix add, so there are no divisions in the loop.
  // Schedules tried on B200 (tools/probe/stream_probe2.cu for the bare traffic mix, this kernel at 2^24 entries):
  //   bare loop: 1 wave static 0.85 of the copy peak, 4 waves 0.93, atomic tile counter 0.92, one CTA per tile 0.95;
  //   here     : a CTA per tile / 4-8 waves lose to the per-CTA prologue (offset tables), a per-CTA atomic counter
  //              needs a barrier per tile (+1..3% on plain streams, -4..7% with a gathered or third operand), per-warp
  //              counters serialise on the atomics.  2 waves for plain 1-2 operand streams, 1 wave otherwise.
  uint32_t i0, i1, i2, s0, s1, s2;
  {
    uint32_t t = blockIdx.x;
    i0 = t % (uint32_t)p.L[0]; t /= (uint32_t)p.L[0];
    i1 = t % (uint32_t)p.L[1]; i2 = t / (uint32_t)p.L[1];
    t = gridDim.x;
    s0 = t % (uint32_t)p.L[0]; t /= (uint32_t)p.L[0];
    s1 = t % (uint32_t)p.L[1]; s2 = t / (uint32_t)p.L[1];
  }
  fill_table<NS1>(tOff0, p.L[0], t0b, p.nt[0], p);
  fill_table<NS1>(tOff1, p.L[1], t1b, p.nt[1], p);
  fill_table<NS1>(tOff2, p.L[2], t2b, p.nt[2], p);
  // ---- this thread's column
  const int V = p.V, G = p.G;
  const int col = tid % V, g = tid / V;
  int32_t loOff[NS1];
#pragma unroll
  for (int k = 0; k < NS1; ++k) loOff[k] = 0;
  add_offsets<NS1>(loOff, (uint32_t)col * W, 0, p.nlo, p);
  __syncthreads();
  const bool active = g < G;  // threads beyond G * V columns have no rows (non power-of-two column counts)
  const int ns_rt = (NS == 1) ? p.Ns : NS;
  int vec128[NF], vstep[NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) { vec128[k] = p.vec128[k]; vstep[k] = p.vstep[k]; }

  for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int32_t base[NS1];
#pragma unroll
    for (int k = 0; k < NS1; ++k)
      base[k] = loOff[k] + tOff0[k * L0 + i0] + tOff1[k * L1 + i1] + tOff2[k * L2 + i2];
    i0 += s0;
    i1 += s1;
    if (i0 >= (uint32_t)p.L[0]) { i0 -= (uint32_t)p.L[0]; ++i1; }
    i2 += s2;
    if (i1 >= (uint32_t)p.L[1]) { i1 -= (uint32_t)p.L[1]; ++i2; }
    double* ocol = p.out + base[NF];
    // Pull the NEXT tile's rows of the leading operand from HBM into L2 with one TMA bulk prefetch per row
    // (cp.async.bulk.prefetch.L2, issued by H * pf_ns threads): the demand loads one tile later then pay L2
    // latency instead of DRAM latency, which is what bounds these streaming kernels (long_scoreboard stalls).
    if (p.pf_bytes && tid < p.H * p.pf_ns && tile + gridDim.x < p.n_tiles) {
      const int row = tid % p.H, ss = tid / p.H;
      const int lk = p.lead;
      int32_t o = tOff0[lk * p.L[0] + i0] + tOff1[lk * p.L[1] + i1] + tOff2[lk * p.L[2] + i2] + jOff[lk * p.H + row];
      if (SUM && p.pf_ns > 1) o += sOff[lk * p.Ns + ss];
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.f[lk] + o), "r"(p.pf_bytes) : "memory");
    }

    for (int j0 = g; active && j0 < p.H; j0 += G * JU) {
      int32_t off[JU][NF], ooff[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        const int j = (j0 + u * G < p.H) ? j0 + u * G : j0;  // clamp: redundant load, store skipped below
#pragma unroll
        for (int k = 0; k < NF; ++k) off[u][k] = base[k] + jOff[k * p.H + j];
        ooff[u] = jOff[NF * p.H + j];
      }
      if constexpr (VEC == VEC_LO) {
        double2 acc[JU];
        if constexpr (!SUM) {
          double2 v[JU][NF];
#pragma unroll
          for (int k = 0; k < NF; ++k) {
            bool done = false;
            if constexpr (JU == 4) {
              if (p.rowvec[k]) {
                // gathered operand whose dim order differs from the output's: a warp load touches 32 distinct
                // 128-byte lines whatever we do (the lanes differ in dims that are far apart in this operand), so
                // make every touched sector count: 32 useful bytes per 32-byte sector instead of 8
                const double2* bx = reinterpret_cast<const double2*>(p.f[k] + off[0][k]);
                const double2 x01 = __ldg(bx), x23 = __ldg(bx + 1);
                double2 y01 = x01, y23 = x23;
                if (vstep[k] != 0) {
                  const double2* by = reinterpret_cast<const double2*>(p.f[k] + off[0][k] + vstep[k]);
                  y01 = __ldg(by); y23 = __ldg(by + 1);
                }
                v[0][k] = make_double2(x01.x, y01.x); v[1][k] = make_double2(x01.y, y01.y);
                v[2][k] = make_double2(x23.x, y23.x); v[3][k] = make_double2(x23.y, y23.y);
                done = true;
              }
            }
            if (!done) {
#pragma unroll
              for (int u = 0; u < JU; ++u) v[u][k] = ld2(p.f[k], off[u][k], vec128[k], vstep[k]);
            }
          }
#pragma unroll
          for (int u = 0; u < JU; ++u) {
            acc[u] = v[u][0];
#pragma unroll
            for (int k = 1; k < NF; ++k) { acc[u].x *= v[u][k].x; acc[u].y *= v[u][k].y; }
          }
        } else if constexpr (NS > 1) {
          double2 v[NS][JU][NF];
#pragma unroll
          for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int u = 0; u < JU; ++u)
#pragma unroll
              for (int k = 0; k < NF; ++k) v[s][u][k] = ld2(p.f[k], off[u][k] + sOff[k * NS + s], vec128[k], vstep[k]);
#pragma unroll
          for (int u = 0; u < JU; ++u)
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              double2 pr = v[s][u][0];
#pragma unroll
              for (int k = 1; k < NF; ++k) { pr.x *= v[s][u][k].x; pr.y *= v[s][u][k].y; }
              if (s == 0) acc[u] = pr; else { acc[u].x += pr.x; acc[u].y += pr.y; }
            }
        } else {
#pragma unroll 4
          for (int s = 0; s < ns_rt; ++s) {
#pragma unroll
            for (int u = 0; u < JU; ++u) {
              double2 pr = ld2(p.f[0], off[u][0] + sOff[s], vec128[0], vstep[0]);
#pragma unroll
              for (int k = 1; k < NF; ++k) {
                const double2 x = ld2(p.f[k], off[u][k] + sOff[k * ns_rt + s], vec128[k], vstep[k]);
                pr.x *= x.x; pr.y *= x.y;
              }
              if (s == 0) acc[u] = pr; else { acc[u].x += pr.x; acc[u].y += pr.y; }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < JU; ++u)
          if (j0 + u * G < p.H) *reinterpret_cast<double2*>(ocol + ooff[u]) = acc[u];
      } else if constexpr (VEC == VEC_S) {
        // pairs run along the first summed dim (contiguous in the leading operand); Ns is even
        double acc[JU];
        if constexpr (NS > 1) {
          constexpr int NP = NS / 2;
          double2 v[NP][JU][NF];
#pragma unroll
          for (int s = 0; s < NP; ++s)
#pragma unroll
            for (int u = 0; u < JU; ++u)
#pragma unroll
              for (int k = 0; k < NF; ++k)
                v[s][u][k] = ld2(p.f[k], off[u][k] + sOff[k * NS + 2 * s], vec128[k], vstep[k]);
#pragma unroll
          for (int u = 0; u < JU; ++u)
#pragma unroll
            for (int s = 0; s < NP; ++s) {
              double2 pr = v[s][u][0];
#pragma unroll
              for (int k = 1; k < NF; ++k) { pr.x *= v[s][u][k].x; pr.y *= v[s][u][k].y; }
              acc[u] = (s == 0) ? pr.x : acc[u] + pr.x;
              acc[u] += pr.y;
            }
        } else {
#pragma unroll 4
          for (int s = 0; s < ns_rt; s += 2) {
#pragma unroll
            for (int u = 0; u < JU; ++u) {
              double2 pr = ld2(p.f[0], off[u][0] + sOff[s], vec128[0], vstep[0]);
#pragma unroll
              for (int k = 1; k < NF; ++k) {
                const double2 x = ld2(p.f[k], off[u][k] + sOff[k * ns_rt + s], vec128[k], vstep[k]);
                pr.x *= x.x; pr.y *= x.y;
              }
              acc[u] = (s == 0) ? pr.x : acc[u] + pr.x;
              acc[u] += pr.y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < JU; ++u)
          if (j0 + u * G < p.H) ocol[ooff[u]] = acc[u];
      } else {
        double acc[JU];
        if constexpr (!SUM) {
          double v[JU][NF];
#pragma unroll
          for (int u = 0; u < JU; ++u)
#pragma unroll
            for (int k = 0; k < NF; ++k) v[u][k] = __ldg(p.f[k] + off[u][k]);
#pragma unroll
          for (int u = 0; u < JU; ++u) {
            acc[u] = v[u][0];
#pragma unroll
            for (int k = 1; k < NF; ++k) acc[u] *= v[u][k];
          }
        } else {
#pragma unroll 4
          for (int s = 0; s < ns_rt; ++s) {
#pragma unroll
            for (int u = 0; u < JU; ++u) {
              double pr = __ldg(p.f[0] + off[u][0] + sOff[s]);
#pragma unroll
              for (int k = 1; k < NF; ++k) pr *= __ldg(p.f[k] + off[u][k] + sOff[k * ns_rt + s]);
              acc[u] = (s == 0) ? pr : acc[u] + pr;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < JU; ++u)
          if (j0 + u * G < p.H) ocol[ooff[u]] = acc[u];
      }
    }
  }
}

// ---- normalize!: two launches -- (1) per-CTA partial sums, the last CTA to finish folds them in a fixed order,
//      (2) x ./= total.  Deterministic (no fp atomics).  One-shot CTAs of 256 threads x 4 x 128 bit = 16 KB (the shape
//      that reaches 0.93 of the copy peak in tools/probe/stream_probe.cu); the scale pass walks the chunks in
//      DESCENDING order so that it starts on the bytes the reduce pass left in the 126 MB L2.
constexpr int NORM_THR = 256, NORM_VEC = 4;  // double2 per thread
constexpr long long NORM_CHUNK = (long long)NORM_THR * NORM_VEC;

__global__ void __launch_bounds__(NORM_THR) norm_reduce_kernel(const double* __restrict__ x, long long n, int p,
                                                               double* partial, unsigned int* ticket, double* total) {
  __shared__ double wsum[NORM_THR / 32];
  __shared__ bool last;
  const long long n2 = n / 2;
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const long long i = (long long)blockIdx.x * NORM_CHUNK + threadIdx.x;
  double2 v[NORM_VEC];
#pragma unroll
  for (int u = 0; u < NORM_VEC; ++u) v[u] = (i + u * NORM_THR < n2) ? __ldg(x2 + i + u * NORM_THR) : make_double2(0.0, 0.0);
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < NORM_VEC; ++u) acc += (p == 1) ? (fabs(v[u].x) + fabs(v[u].y)) : (v[u].x * v[u].x + v[u].y * v[u].y);
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double t = x[n - 1];
    acc += (p == 1) ? fabs(t) : t * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < NORM_THR / 32; ++w) s += wsum[w];
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    // fixed-order fold of the partials by this CTA: strided per-thread sums, then a fixed tree
    double s = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += NORM_THR) s += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tsum = 0.0;
      for (int w = 0; w < NORM_THR / 32; ++w) tsum += wsum[w];
      *total = tsum;
      *ticket = 0u;
    }
  }
}

__global__ void __launch_bounds__(NORM_THR) norm_scale_kernel(double* __restrict__ x, long long n, const double* total) {
  const double t = *total;
  const long long n2 = n / 2;
  double2* x2 = reinterpret_cast<double2*>(x);
  const long long i = (long long)(gridDim.x - 1 - blockIdx.x) * NORM_CHUNK + threadIdx.x;
  double2 v[NORM_VEC];
#pragma unroll
  for (int u = 0; u < NORM_VEC; ++u) if (i + u * NORM_THR < n2) v[u] = x2[i + u * NORM_THR];
#pragma unroll
  for (int u = 0; u < NORM_VEC; ++u)
    if (i + u * NORM_THR < n2) {
      v[u].x /= t; v[u].y /= t;  // true division, as `./=`
      x2[i + u * NORM_THR] = v[u];
    }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] /= t;
}

// ------------------------------------------------------------------------------------------ host
struct Operand {
  const double* ptr;
  std::vector<int32_t> dims, card;
};

static int find_dim(const std::vector<int32_t>& dims, int32_t d) {
  for (size_t i = 0; i < dims.size(); ++i) if (dims[i] == d) return (int)i;
  return -1;
}

template <int NF, int VEC, int NS>
static int32_t launch_contract_v(const ContractParams& p, int sm_count, size_t smem, cudaStream_t s) {
  static thread_local int occ = 0;  // resident CTAs per SM of this instantiation (per host thread)
  if (occ == 0) {
    BN_CUDA_TRY(cudaFuncSetAttribute(contract_kernel<NF, VEC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    BN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, contract_kernel<NF, VEC, NS>, NTHR, 16 * 1024));
    if (occ < 1) occ = 1;
  }
  BN_REQUIRE(smem <= 64 * 1024, BN_UNSUPPORTED, "contraction offset tables need %zu bytes of shared memory", smem);
  // tiles round-robin over p.waves waves of resident CTAs (see the kernel comment)
  const int grid = (int)std::min<long long>(p.n_tiles, (long long)sm_count * occ * p.waves);
  contract_kernel<NF, VEC, NS><<<grid, NTHR, smem, s>>>(p);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

template <int NF, int VEC>
static int32_t launch_contract_ns(const ContractParams& p, int ns_mode, int sm_count, size_t smem, cudaStream_t s) {
  if constexpr (VEC == VEC_S) {  // Ns is even here; 3 cannot occur
    if (ns_mode == 2) return launch_contract_v<NF, VEC, 2>(p, sm_count, smem, s);
    if constexpr (NF <= 3) {
      if (ns_mode == 4) return launch_contract_v<NF, VEC, 4>(p, sm_count, smem, s);
    }
    return launch_contract_v<NF, VEC, 1>(p, sm_count, smem, s);
  }
  if (ns_mode == 0) return launch_contract_v<NF, VEC, 0>(p, sm_count, smem, s);
  if constexpr (NF <= 3) {
    if (ns_mode == 2) return launch_contract_v<NF, VEC, 2>(p, sm_count, smem, s);
    if (ns_mode == 3) return launch_contract_v<NF, VEC, 3>(p, sm_count, smem, s);
    if (ns_mode == 4) return launch_contract_v<NF, VEC, 4>(p, sm_count, smem, s);
  }
  return launch_contract_v<NF, VEC, 1>(p, sm_count, smem, s);
}

template <int NF>
static int32_t launch_contract_nf(const ContractParams& p, int vec, int ns_mode, int sm_count, size_t smem, cudaStream_t s) {
  if (vec == VEC_S) return launch_contract_ns<NF, VEC_S>(p, ns_mode, sm_count, smem, s);
  if (vec == VEC_LO) return launch_contract_ns<NF, VEC_LO>(p, ns_mode, sm_count, smem, s);
  return launch_contract_ns<NF, VEC_NONE>(p, ns_mode, sm_count, smem, s);
}

// out[out_dims] = sum_{sum_dims} prod_k ops[k]; out_dims/out_card give the output layout (column-major),
// sum_dims/sum_card the reduced variables in accumulation order (first fastest).
static int32_t contract(int device, const std::vector<Operand>& ops, const std::vector<int32_t>& out_dims,
                        const std::vector<int32_t>& out_card, const std::vector<int32_t>& sum_dims,
                        const std::vector<int32_t>& sum_card, double* out) {
  const int nf = (int)ops.size();
  BN_REQUIRE(nf >= 1 && nf <= MAXF, BN_UNSUPPORTED, "contraction of %d factors (max %d)", nf, MAXF);
  const int no = (int)out_dims.size(), ns = (int)sum_dims.size();
  BN_REQUIRE(no + ns <= MAXD, BN_UNSUPPORTED, "contraction over %d variables (max %d)", no + ns, MAXD);
  ContractParams p{};
  p.nf = nf;
  p.ns = ns;
  p.out = out;
  long long Nout = 1, Ns = 1;
  for (int c : out_card) Nout *= c;
  for (int c : sum_card) Ns *= c;
  BN_REQUIRE(Ns <= MAX_NS, BN_UNSUPPORTED, "summing out %lld joint states at once (max %d)", Ns, MAX_NS);
  BN_REQUIRE(Nout < (1ll << 31), BN_UNSUPPORTED, "factor with %lld entries (max 2^31 - 1)", Nout);
  const DevProps& dp = dev_props(device);

  // strides of every operand along the output dims and the summed dims; the largest operand leads
  std::vector<std::vector<long long>> st(nf, std::vector<long long>(no + ns, 0));
  int lead = 0;
  long long lead_len = -1;
  for (int k = 0; k < nf; ++k) {
    p.f[k] = ops[k].ptr;
    long long acc = 1;
    for (size_t i = 0; i < ops[k].dims.size(); ++i) {
      int pos = find_dim(out_dims, ops[k].dims[i]);
      if (pos < 0) {
        pos = find_dim(sum_dims, ops[k].dims[i]);
        BN_REQUIRE(pos >= 0, BN_BAD_ARG, "operand dim %d is neither kept nor summed", ops[k].dims[i]);
        pos += no;
      }
      st[k][pos] = acc;
      acc *= ops[k].card[i];
    }
    BN_REQUIRE(acc < (1ll << 31), BN_UNSUPPORTED, "factor with %lld entries (max 2^31 - 1)", acc);
    if (acc > lead_len) { lead_len = acc; lead = k; }
  }
  // vector dim: the first summed dim if the largest operand is contiguous there, else the first output dim
  int vec = VEC_NONE;
  if (ns >= 1 && (sum_card[0] % 2) == 0 && st[lead][no] == 1) vec = VEC_S;
  else if (no >= 1 && (out_card[0] % 2) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0) vec = VEC_LO;
  const int W = vec == VEC_LO ? 2 : 1;

  // lo = longest prefix of the output dims whose vector columns fit one CTA pass
  int nlo = 0;
  long long Rlo = 1;
  while (nlo < no && Rlo * out_card[nlo] / W <= NTHR) { Rlo *= out_card[nlo]; ++nlo; }
  if (nlo == 0 && no > 0) {  // first dim alone is wider than a CTA pass: cannot happen for card <= 255 with W >= 1
    Rlo = out_card[0]; nlo = 1;
  }
  BN_REQUIRE(Rlo / W <= NTHR, BN_UNSUPPORTED, "leading output dim too large (%lld)", Rlo);
  const int V = (int)std::max<long long>(1, Rlo / W);
  const int G = NTHR / V;
  // loop-dim order [lo | j | t | s] as positions into out_dims.  j: first the output dims the leading operand
  // lacks (see the kernel comment), then the dims right after lo, until every row group has ~4 rows per tile.
  std::vector<int> order;
  for (int i = 0; i < nlo; ++i) order.push_back(i);
  std::vector<char> taken(no, 0);
  long long H = 1;
  int nj = 0;
  static const int h_rows = getenv("BN_FACTOR_H") ? atoi(getenv("BN_FACTOR_H")) : 4;
  const long long h_target = (long long)h_rows * G;
  // Row-vector gather (plain products of <= 3 operands, one row group): if a large non-leading operand has its two
  // fastest dims outside lo, make exactly those dims the tile's 4 rows so that a thread's 4 row values are 4
  // consecutive doubles of that operand (see rowvec in the kernel).  Skipped when the leading operand lacks some
  // output dim (then those dims go into the rows instead, which saves an HBM re-read of the leading operand).
  int rowvec_op = -1;
  static const bool no_rowvec = getenv("BN_NO_ROWVEC") != nullptr;
  if (vec == VEC_LO && ns == 0 && nf <= 3 && G == 1 && !no_rowvec) {
    bool lead_has_all = true;
    for (int i = nlo; i < no; ++i) lead_has_all = lead_has_all && st[lead][i] != 0;
    long long best_len = 1 << 15;  // smaller operands live in L1/L2 anyway
    for (int k = 0; k < nf && lead_has_all; ++k) {
      if (k == lead || ops[k].dims.empty()) continue;
      long long len = 1;
      for (int c : ops[k].card) len *= c;
      if (len < best_len || (reinterpret_cast<uintptr_t>(p.f[k]) % 32) != 0) continue;
      // its fastest dims as positions in the output
      const int p0 = find_dim(out_dims, ops[k].dims[0]);
      const int p1 = ops[k].dims.size() > 1 ? find_dim(out_dims, ops[k].dims[1]) : -1;
      if (p0 >= nlo && out_card[p0] == 4) { rowvec_op = k; best_len = len; taken.assign(no, 0); taken[p0] = 1; }
      else if (p0 >= nlo && p1 >= nlo && out_card[p0] * out_card[p1] == 4) {
        rowvec_op = k; best_len = len; taken.assign(no, 0); taken[p0] = 1; taken[p1] = 1;
      }
    }
    if (rowvec_op >= 0) {
      const int p0 = find_dim(out_dims, ops[rowvec_op].dims[0]);
      order.push_back(p0); H *= out_card[p0]; ++nj;
      if (H < 4) { const int p1 = find_dim(out_dims, ops[rowvec_op].dims[1]); order.push_back(p1); H *= out_card[p1]; ++nj; }
    }
  }
  // Sector rule for gathered operands: the 4 doubles of a 32-byte sector of the largest non-leading operand differ in
  // its fastest dims.  Any of those dims that is not a lo dim (lanes / warps of the same CTA) becomes a row dim, so a
  // sector is consumed by one thread back to back (L1 hit) instead of by 2-4 tiles far apart in time (2-4x the
  // L2 -> L1 traffic for that operand).
  if (rowvec_op < 0 && nf >= 2) {
    int big = -1;
    long long big_len = 1 << 15;
    for (int k = 0; k < nf; ++k) {
      if (k == lead) continue;
      long long len = 1;
      for (int c : ops[k].card) len *= c;
      if (len >= big_len) { big_len = len; big = k; }
    }
    if (big >= 0) {
      long long run = 1;
      for (size_t i = 0; i < ops[big].dims.size() && run < 4; ++i) {
        const int pos = find_dim(out_dims, ops[big].dims[i]);
        run *= ops[big].card[i];
        if (pos >= nlo && !taken[pos] && H * out_card[pos] <= 16) {
          taken[pos] = 1; order.push_back(pos); H *= out_card[pos]; ++nj;
        }
      }
    }
  }
  for (int pass = 0; pass < 2 && rowvec_op < 0; ++pass)
    for (int i = nlo; i < no; ++i) {
      if (taken[i] || (pass == 0 && st[lead][i] != 0)) continue;
      if (H >= h_target || H * out_card[i] > 1024) continue;
      taken[i] = 1; order.push_back(i); H *= out_card[i]; ++nj;
    }
  // t = the rest in ascending order, in <= 3 digit groups of <= 256 values
  std::vector<int> rest;
  for (int i = nlo; i < no; ++i) if (!taken[i]) rest.push_back(i);
  size_t ri = 0;
  long long n_tiles = 1;
  for (int l = 0; l < TLEV; ++l) {
    p.nt[l] = 0; p.L[l] = 1;
    while (ri < rest.size() && (long long)p.L[l] * out_card[rest[ri]] <= TLEV_MAX) {
      p.L[l] *= out_card[rest[ri]]; ++p.nt[l]; order.push_back(rest[ri]); ++ri;
    }
    n_tiles *= p.L[l];
  }
  BN_REQUIRE(ri == rest.size(), BN_UNSUPPORTED, "output too large for the tile tables (%lld entries)", Nout);
  BN_REQUIRE(n_tiles < (1ll << 31), BN_UNSUPPORTED, "too many tiles");
  p.nlo = nlo; p.nj = nj;
  p.H = (int)H; p.n_tiles = n_tiles; p.Ns = (int)Ns; p.V = V; p.G = G;
  std::vector<long long> ost(no, 1);  // column-major strides of the output
  for (int i = 1; i < no; ++i) ost[i] = ost[i - 1] * out_card[i - 1];
  for (int i = 0; i < no; ++i) {
    p.card[i] = out_card[order[i]];
    for (int k = 0; k < nf; ++k) p.fstride[k][i] = (int32_t)st[k][order[i]];
    p.fstride[nf][i] = (int32_t)ost[order[i]];
  }
  for (int i = 0; i < ns; ++i) {
    p.card[no + i] = sum_card[i];
    for (int k = 0; k < nf; ++k) p.fstride[k][no + i] = (int32_t)st[k][no + i];
    p.fstride[nf][no + i] = 0;
  }
  for (int i = 0; i < no + ns; ++i) p.magic[i] = (uint32_t)(((1u << 20) + (uint32_t)p.card[i] - 1u) / (uint32_t)p.card[i]);
  {  // prefix-product magics per dim group [lo | j | t0 | t1 | t2 | s]; lo and t groups hold < 1024 values each
    const int ends[6] = {nlo, nlo + nj, nlo + nj + p.nt[0], nlo + nj + p.nt[0] + p.nt[1], no, no + ns};
    int d = 0;
    for (int gI = 0; gI < 6; ++gI) {
      unsigned long long P = 1;
      for (; d < ends[gI]; ++d) {
        p.pmagic[d] = P <= 1024 ? (uint32_t)(((1ull << 20) + P - 1) / P) : 0u;
        P *= (unsigned long long)p.card[d];
      }
    }
  }
  const int vdim = vec == VEC_S ? no : 0;
  for (int k = 0; k < nf; ++k) {
    const long long sv = vec != VEC_NONE ? p.fstride[k][vdim] : 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(p.f[k]) % 16) == 0;
    p.vec128[k] = (vec != VEC_NONE && sv == 1 && aligned) ? 1 : 0;
    p.vstep[k] = (int32_t)sv;
    p.rowvec[k] = (k == rowvec_op) ? 1 : 0;
  }
  // L2 prefetch plan: the leading operand's bytes for one (row, summed state) must be one contiguous, 16-byte
  // aligned span: its strides over the lo dims are the running product of their cards (after the vector summed
  // dim in VEC_S mode, which then lies inside the span)
  p.lead = lead;
  p.pf_bytes = 0;
  p.pf_ns = 1;
  {
    long long run = (vec == VEC_S) ? sum_card[0] : 1;
    bool contig = true;
    for (int i = 0; i < nlo && contig; ++i) {
      contig = p.fstride[lead][i] == run;
      run *= p.card[i];
    }
    const long long span = run * 8;
    const long long ns_sep = (vec == VEC_S) ? Ns / sum_card[0] : Ns;  // summed states outside the span
    const bool aligned = (reinterpret_cast<uintptr_t>(p.f[lead]) % 16) == 0;
    if (contig && aligned && nlo > 0 && span % 16 == 0 && span >= 512 && (vec != VEC_S || ns_sep == 1) &&
        H * ns_sep <= NTHR && n_tiles > 1 && nf >= 3) {  // measured: +15% for the 3-operand fused eliminate,
      // -1..3% for 1-2 operand streams that already keep >= 128 B per thread in flight (profiles/r1c_factor_*.txt)
      p.pf_bytes = (int32_t)span;
      p.pf_ns = (int32_t)ns_sep;
    }
  }
  const size_t tl = (size_t)p.L[0] + p.L[1] + p.L[2];
  static const int waves_env = getenv("BN_FACTOR_WAVES") ? std::max(1, atoi(getenv("BN_FACTOR_WAVES"))) : 0;
  p.waves = waves_env ? waves_env : ((vec == VEC_LO && nf <= 2 && rowvec_op < 0) ? 2 : 1);
  const size_t smem = ((size_t)(nf + 1) * ((size_t)H + tl) + (size_t)nf * (size_t)Ns) * sizeof(int32_t);
  cudaStream_t s = tl_stream;
  const int ns_mode = ns == 0 ? 0 : ((Ns >= 2 && Ns <= 4) ? (int)Ns : 1);
  switch (nf) {
    case 1: return launch_contract_nf<1>(p, vec, ns_mode, dp.sm_count, smem, s);
    case 2: return launch_contract_nf<2>(p, vec, ns_mode, dp.sm_count, smem, s);
    case 3: return launch_contract_nf<3>(p, vec, ns_mode, dp.sm_count, smem, s);
    case 4: return launch_contract_nf<4>(p, vec, ns_mode, dp.sm_count, smem, s);
    case 5: return launch_contract_nf<5>(p, vec, ns_mode, dp.sm_count, smem, s);
    default: return launch_contract_nf<6>(p, vec, ns_mode, dp.sm_count, smem, s);
  }
}

// dims of phi1 * phi2 per the reference: larger first (strict <), union keeps phi1's order.
static int32_t join_dims(std::vector<int32_t>& d1, std::vector<int32_t>& c1, long long& n1,
                         const std::vector<int32_t>& d2in, const std::vector<int32_t>& c2in, long long n2in) {
  std::vector<int32_t> d2 = d2in, c2 = c2in;
  long long n2 = n2in;
  if (n1 < n2) { std::swap(d1, d2); std::swap(c1, c2); std::swap(n1, n2); }  // factors_dims.jl:189-191
  for (size_t j = 0; j < d2.size(); ++j) {
    int pos = find_dim(d1, d2[j]);
    if (pos >= 0) {
      BN_REQUIRE(c1[pos] == c2[j], BN_DIM_MISMATCH, "Common dimensions must have same size (dim %d: %d vs %d)",
                 d2[j], c1[pos], c2[j]);
    }
  }
  for (size_t j = 0; j < d2.size(); ++j)
    if (find_dim(d1, d2[j]) < 0) { d1.push_back(d2[j]); c1.push_back(c2[j]); n1 *= c2[j]; }
  return BN_OK;
}

static int32_t eliminate_impl(const std::vector<Factor*>& fs, const int32_t* sum_dims, int32_t n_sum, Factor** out) {
  BN_REQUIRE(!fs.empty(), BN_BAD_ARG, "no factors given");
  const int device = fs[0]->device;
  for (Factor* f : fs) BN_REQUIRE(f->device == device, BN_BAD_ARG, "factors live on different devices");
  // dims of reduce(*, factors) exactly as the reference's left fold would produce them
  std::vector<int32_t> d = fs[0]->dims, c = fs[0]->card;
  long long n = fs[0]->len;
  for (size_t k = 1; k < fs.size(); ++k) BN_TRY(join_dims(d, c, n, fs[k]->dims, fs[k]->card, fs[k]->len));
  std::vector<int32_t> od, oc, sd, sc;
  for (int j = 0; j < n_sum; ++j) {
    BN_REQUIRE(find_dim(d, sum_dims[j]) >= 0, BN_NOT_IN_FACTOR, "%d is not a valid dimension", sum_dims[j]);
    for (int j2 = 0; j2 < j; ++j2) BN_REQUIRE(sum_dims[j2] != sum_dims[j], BN_BAD_ARG, "Dimensions must be unique");
  }
  for (size_t i = 0; i < d.size(); ++i) {
    bool summed = false;
    for (int j = 0; j < n_sum; ++j) summed = summed || (sum_dims[j] == d[i]);
    if (summed) { sd.push_back(d[i]); sc.push_back(c[i]); } else { od.push_back(d[i]); oc.push_back(c[i]); }
  }
  Factor* o = nullptr;
  BN_TRY(new_factor(od, oc, device, &o));
  if (fs.size() == 1 && !sd.empty()) {  // sum(phi, dims) of one factor: the one-shot streaming kernel (factor_stream.cu)
    std::vector<char> flag(d.size(), 0);
    for (size_t i = 0; i < d.size(); ++i)
      for (int j = 0; j < n_sum; ++j) flag[i] = flag[i] || (sum_dims[j] == d[i]);
    bool handled = false;
    int32_t rc = sum_segmented(fs[0]->data.p, c, flag, o->data.p, &handled);
    if (rc != BN_OK) { delete o; return rc; }
    if (handled) { *out = o; return BN_OK; }
  }
  // operand order: largest first so the leading (register-offset, vectorised) operand is the big one;
  // the product value itself is order independent only up to commutativity, so keep the fold order
  // for the multiplication and merely choose which operand drives the vector mode.
  std::vector<Operand> ops;
  for (Factor* f : fs) ops.push_back(Operand{f->data.p, f->dims, f->card});
  int32_t rc = contract(device, ops, od, oc, sd, sc, o->data.p);
  if (rc != BN_OK) { delete o; return rc; }
  *out = o;
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

int32_t bn_factor_alloc(int32_t ndims, const int32_t* dims, const int32_t* card, int32_t device, bn_factor_t* out) {
  BN_REQUIRE(out && ndims >= 0 && (ndims == 0 || (dims && card)), BN_BAD_ARG, "bn_factor_alloc: bad arguments");
  BN_REQUIRE(ndims <= MAXD, BN_UNSUPPORTED, "factor with %d dims (max %d)", ndims, MAXD);
  std::vector<int32_t> d(dims, dims + ndims), c(card, card + ndims);
  for (int i = 0; i < ndims; ++i) {
    BN_REQUIRE(c[i] >= 1, BN_BAD_ARG, "dimension %d has length %d", d[i], c[i]);
    for (int j = 0; j < i; ++j) BN_REQUIRE(d[j] != d[i], BN_BAD_ARG, "Dimensions must be unique");  // errors.jl:10
  }
  int ndev = 0;
  BN_CUDA_TRY(cudaGetDeviceCount(&ndev));
  BN_REQUIRE(device >= 0 && device < ndev, BN_BAD_ARG, "device %d out of range", device);
  Factor* f = nullptr;
  BN_TRY(new_factor(d, c, device, &f));
  *out = register_factor(f);
  return BN_OK;
}

int32_t bn_factor_upload(int32_t ndims, const int32_t* dims, const int32_t* card, const double* data,
                         int32_t device, bn_factor_t* out) {
  BN_REQUIRE(data, BN_BAD_ARG, "bn_factor_upload: data is NULL");
  BN_TRY(bn_factor_alloc(ndims, dims, card, device, out));
  Factor* f = get_factor(*out);
  BN_CUDA_TRY(cudaMemcpyAsync(f->data.p, data, (size_t)f->len * 8, cudaMemcpyHostToDevice, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

int32_t bn_factor_from_cpd(bn_net_t h, int32_t node, bn_factor_t* out) {
  Net* net = get_net(h);
  BN_REQUIRE(net && out, BN_BAD_ARG, "bn_factor_from_cpd: bad arguments");
  BN_REQUIRE(node >= 0 && node < net->n, BN_BAD_ARG, "node %d out of range", node);
  // convert(Factor, cpd): dims = [name, parents...], same memory order as the flat CPT (factors_main.jl:62-67)
  std::vector<int32_t> d{node}, c{net->card[node]};
  for (int k = net->par_ptr[node]; k < net->par_ptr[node + 1]; ++k) {
    d.push_back(net->par_idx[k]);
    c.push_back(net->card[net->par_idx[k]]);
  }
  BN_REQUIRE((int)d.size() <= MAXD, BN_UNSUPPORTED, "node %d has too many parents", node);
  Factor* f = nullptr;
  BN_TRY(new_factor(d, c, net->device, &f));
  cudaError_t e = cudaMemcpyAsync(f->data.p, net->d_cpt.p + net->cpt_off[node], (size_t)f->len * 8,
                                  cudaMemcpyDeviceToDevice, tl_stream);
  if (e != cudaSuccess) { delete f; BN_CUDA_TRY(e); }
  *out = register_factor(f);
  return BN_OK;
}

}  // extern "C"

namespace bn {
int32_t factor_view_of_cpd(Net* net, int32_t node, bn_factor_t* out) {
  std::vector<int32_t> d{node}, c{net->card[node]};
  for (int k = net->par_ptr[node]; k < net->par_ptr[node + 1]; ++k) {
    d.push_back(net->par_idx[k]);
    c.push_back(net->card[net->par_idx[k]]);
  }
  BN_REQUIRE((int)d.size() <= MAXD, BN_UNSUPPORTED, "node %d has too many parents", node);
  const double* src = net->d_cpt.p + net->cpt_off[node];
  if (reinterpret_cast<uintptr_t>(src) % 16 != 0) {
    // tables are packed back to back, so one that follows an odd number of doubles is only 8-byte aligned; the
    // kernels would fall back to 64-bit loads on it, so take the (pool-allocated, aligned) copy instead
    Factor* c2 = nullptr;
    BN_TRY(new_factor(d, c, net->device, &c2));
    cudaError_t e = cudaMemcpyAsync(c2->data.p, src, (size_t)c2->len * 8, cudaMemcpyDeviceToDevice, tl_stream);
    if (e != cudaSuccess) { delete c2; BN_CUDA_TRY(e); }
    *out = register_factor(c2);
    return BN_OK;
  }
  Factor* f = new Factor();
  f->device = net->device;
  f->dims = d;
  f->card = c;
  f->len = net->cpt_off[node + 1] - net->cpt_off[node];
  f->data.borrow(const_cast<double*>(src), (size_t)f->len);
  *out = register_factor(f);
  return BN_OK;
}
}  // namespace bn

extern "C" {

int32_t bn_factor_relabel(bn_factor_t h, const int32_t* new_dims) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && (new_dims || f->dims.empty()), BN_BAD_ARG, "bn_factor_relabel: bad arguments");
  for (size_t i = 0; i < f->dims.size(); ++i)
    for (size_t j = 0; j < i; ++j) BN_REQUIRE(new_dims[i] != new_dims[j], BN_BAD_ARG, "Dimensions must be unique");
  for (size_t i = 0; i < f->dims.size(); ++i) f->dims[i] = new_dims[i];
  return BN_OK;
}

int32_t bn_factor_ndims(bn_factor_t h, int32_t* out_ndims) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out_ndims, BN_BAD_ARG, "bn_factor_ndims: bad arguments");
  *out_ndims = (int32_t)f->dims.size();
  return BN_OK;
}

int32_t bn_factor_shape(bn_factor_t h, int32_t* out_dims, int32_t* out_card) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_shape: unknown handle");
  for (size_t i = 0; i < f->dims.size(); ++i) {
    if (out_dims) out_dims[i] = f->dims[i];
    if (out_card) out_card[i] = f->card[i];
  }
  return BN_OK;
}

int32_t bn_factor_length(bn_factor_t h, int64_t* out_len) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out_len, BN_BAD_ARG, "bn_factor_length: bad arguments");
  *out_len = f->len;
  return BN_OK;
}

int32_t bn_factor_download(bn_factor_t h, double* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out, BN_BAD_ARG, "bn_factor_download: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(f->device));
  BN_CUDA_TRY(cudaMemcpyAsync(out, f->data.p, (size_t)f->len * 8, cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

int32_t bn_factor_dev_ptr(bn_factor_t h, void** out_ptr) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out_ptr, BN_BAD_ARG, "bn_factor_dev_ptr: bad arguments");
  *out_ptr = f->data.p;
  return BN_OK;
}

int32_t bn_factor_free(bn_factor_t h) {
  Factor* f = take_factor(h);
  BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_free: unknown handle");
  cudaSetDevice(f->device);
  delete f;
  return BN_OK;
}

int32_t bn_factor_product(bn_factor_t h1, bn_factor_t h2, bn_factor_t* out) {
  Factor* f1 = get_factor(h1);
  Factor* f2 = get_factor(h2);
  BN_REQUIRE(f1 && f2 && out, BN_BAD_ARG, "bn_factor_product: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(f1->device));
  if (f1->len < f2->len) std::swap(f1, f2);  // larger operand leads (factors_dims.jl:189-191)
  Factor* o = nullptr;
  BN_TRY(eliminate_impl({f1, f2}, nullptr, 0, &o));
  *out = register_factor(o);
  return BN_OK;
}

int32_t bn_factor_eliminate(const bn_factor_t* hs, int32_t n_factors, const int32_t* sum_dims, int32_t n_sum,
                            bn_factor_t* out) {
  BN_REQUIRE(hs && out && n_factors >= 1, BN_BAD_ARG, "bn_factor_eliminate: bad arguments");
  BN_REQUIRE(n_sum == 0 || sum_dims, BN_BAD_ARG, "bn_factor_eliminate: sum_dims is NULL");
  std::vector<Factor*> fs;
  for (int k = 0; k < n_factors; ++k) {
    Factor* f = get_factor(hs[k]);
    BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_eliminate: unknown handle #%d", k);
    fs.push_back(f);
  }
  BN_CUDA_TRY(cudaSetDevice(fs[0]->device));
  // put the largest operand first: it drives the coalesced / vectorised access pattern.  The fold's
  // dim order is recomputed from the ORIGINAL order inside eliminate_impl's join simulation, so do the
  // reordering only when it cannot change that order: i.e. when the largest is already first or the
  // reference itself would have swapped it to the front.
  Factor* o = nullptr;
  BN_TRY(eliminate_impl(fs, sum_dims, n_sum, &o));
  *out = register_factor(o);
  return BN_OK;
}

int32_t bn_factor_sum(bn_factor_t h, const int32_t* sum_dims, int32_t n_sum, bn_factor_t* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out, BN_BAD_ARG, "bn_factor_sum: bad arguments");
  BN_REQUIRE(n_sum == 0 || sum_dims, BN_BAD_ARG, "bn_factor_sum: sum_dims is NULL");
  BN_CUDA_TRY(cudaSetDevice(f->device));
  // More joint summed states than one launch takes (MAX_NS): sum in stages, the dims grouped in the factor's own order
  // (fastest first) so that every stage stays below the limit.  The stages re-associate the additions of the
  // reference's single pass (same terms, <= 1e-15 relative apart; sums within the limit stay bit-identical).
  long long ns_total = 1;
  std::vector<int32_t> ordered;  // the summed dims in the factor's dim order, with their cards
  std::vector<int32_t> ocard;
  for (size_t i = 0; i < f->dims.size(); ++i)
    for (int j = 0; j < n_sum; ++j)
      if (sum_dims[j] == f->dims[i]) { ordered.push_back(f->dims[i]); ocard.push_back(f->card[i]); ns_total *= f->card[i]; break; }
  Factor* o = nullptr;
  if (ns_total <= MAX_NS || (int)ordered.size() != n_sum) {  // (unknown / repeated dims: eliminate_impl reports them)
    BN_TRY(eliminate_impl({f}, sum_dims, n_sum, &o));
  } else {
    Factor* cur = f;
    size_t i = 0;
    while (i < ordered.size()) {
      long long grp = 1;
      size_t j = i;
      while (j < ordered.size() && grp * ocard[j] <= MAX_NS) { grp *= ocard[j]; ++j; }
      if (j == i) j = i + 1;  // a single dim of more than MAX_NS states cannot occur (card <= 255)
      Factor* nxt = nullptr;
      const int32_t rc = eliminate_impl({cur}, ordered.data() + i, (int32_t)(j - i), &nxt);
      if (cur != f) delete cur;
      if (rc != BN_OK) return rc;
      cur = nxt;
      i = j;
    }
    o = cur;
  }
  *out = register_factor(o);
  return BN_OK;
}

int32_t bn_factor_slice(bn_factor_t h, const int32_t* dims, const int32_t* states, int32_t n_assign, bn_factor_t* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out, BN_BAD_ARG, "bn_factor_slice: bad arguments");
  BN_REQUIRE(n_assign == 0 || (dims && states), BN_BAD_ARG, "bn_factor_slice: NULL assignment");
  BN_CUDA_TRY(cudaSetDevice(f->device));
  // _translate_index: src/Factors/factors_access.jl:48-72 (dims absent from phi are ignored)
  long long base = 0, st = 1;
  std::vector<int32_t> kd, kc;
  Operand op{nullptr, {}, {}};
  std::vector<long long> kstride;
  for (size_t i = 0; i < f->dims.size(); ++i) {
    int hit = -1;
    for (int j = 0; j < n_assign; ++j) if (dims[j] == f->dims[i]) hit = j;
    if (hit >= 0) {
      BN_REQUIRE(states[hit] >= 0 && states[hit] < f->card[i], BN_BOUNDS, "BoundsError: state %d for dim %d (size %d)",
                 states[hit], f->dims[i], f->card[i]);
      base += st * states[hit];
    } else {
      kd.push_back(f->dims[i]);
      kc.push_back(f->card[i]);
    }
    st *= f->card[i];
  }
  Factor* o = nullptr;
  BN_TRY(new_factor(kd, kc, f->device, &o));
  if (kd.size() == f->dims.size()) {
    cudaError_t e = cudaMemcpyAsync(o->data.p, f->data.p, (size_t)f->len * 8, cudaMemcpyDeviceToDevice, tl_stream);
    if (e != cudaSuccess) { delete o; BN_CUDA_TRY(e); }
  } else {
    // a strided gather = 1-operand contraction over the kept dims, sliced dims pinned via the base offset:
    // present the operand with its FULL dim list but give sliced dims a fresh id that is "summed" with card 1.
    Operand full{f->data.p + base, {}, {}};
    std::vector<int32_t> sd, sc;
    int32_t fresh = -1;
    for (size_t i = 0; i < f->dims.size(); ++i) {
      bool kept = find_dim(kd, f->dims[i]) >= 0;
      if (kept) { full.dims.push_back(f->dims[i]); full.card.push_back(f->card[i]); }
      else {
        // sliced dim: keeps its extent in the stride computation but only index 0 (the base) is visited
        while (find_dim(f->dims, fresh) >= 0) --fresh;
        full.dims.push_back(fresh);
        full.card.push_back(f->card[i]);
        sd.push_back(fresh);
        sc.push_back(1);
        --fresh;
      }
    }
    int32_t rc = contract(f->device, {full}, kd, kc, sd, sc, o->data.p);
    if (rc != BN_OK) { delete o; return rc; }
  }
  *out = register_factor(o);
  return BN_OK;
}

// permutedims(phi, perm): src/Factors/factors_main.jl:160-166.  out.dims[i] = phi.dims[perm[i]] (0-based positions):
// the tiled transpose of permute.cu; shapes its tiling does not cover go through a 1-operand contraction whose output
// dim order is the permuted one (a strided gather on the read side).
int32_t bn_factor_permutedims(bn_factor_t h, const int32_t* perm, int32_t n, bn_factor_t* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out && (perm || n == 0), BN_BAD_ARG, "bn_factor_permutedims: bad arguments");
  BN_REQUIRE(n == (int32_t)f->dims.size(), BN_BAD_ARG, "no valid permutation of dimensions (%d entries for %zu dims)", n,
             f->dims.size());
  std::vector<char> seen((size_t)n, 0);
  std::vector<int32_t> od, oc;
  bool identity = true;
  for (int i = 0; i < n; ++i) {
    BN_REQUIRE(perm[i] >= 0 && perm[i] < n && !seen[(size_t)perm[i]], BN_BAD_ARG, "no valid permutation of dimensions");
    seen[(size_t)perm[i]] = 1;
    identity = identity && perm[i] == i;
    od.push_back(f->dims[(size_t)perm[i]]);
    oc.push_back(f->card[(size_t)perm[i]]);
  }
  BN_CUDA_TRY(cudaSetDevice(f->device));
  Factor* o = nullptr;
  BN_TRY(new_factor(od, oc, f->device, &o));
  int32_t rc = BN_OK;
  if (identity || f->len <= 1) {
    cudaError_t e = cudaMemcpyAsync(o->data.p, f->data.p, (size_t)f->len * 8, cudaMemcpyDeviceToDevice, tl_stream);
    if (e != cudaSuccess) { delete o; BN_CUDA_TRY(e); }
  } else {
    static const bool no_tiled = getenv("BN_NO_TILED_PERMUTE") != nullptr;
    bool handled = false;
    if (!no_tiled) rc = permute_tiled(f->device, f->data.p, f->card, perm, o->data.p, &handled);
    if (rc == BN_OK && !handled)
      rc = contract(f->device, {Operand{f->data.p, f->dims, f->card}}, od, oc, {}, {}, o->data.p);
  }
  if (rc != BN_OK) { delete o; return rc; }
  *out = register_factor(o);
  return BN_OK;
}

// push!(phi, dim, length): src/Factors/factors_main.jl:148-158 over duplicate (auxiliary.jl:50-65) -- a new LAST
// dimension along which the whole table is repeated: a 1-operand contraction with an output dim the operand lacks
// (stride 0), so the operand is read once per tile and stored `length` times.
int32_t bn_factor_push(bn_factor_t h, int32_t dim, int32_t length, bn_factor_t* out) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out, BN_BAD_ARG, "bn_factor_push: bad arguments");
  BN_REQUIRE(find_dim(f->dims, dim) < 0, BN_BAD_ARG, "Dimension %d already exists", dim);
  BN_REQUIRE(length >= 1, BN_BAD_ARG, "bn_factor_push: length %d", length);
  BN_REQUIRE(f->len * (long long)length < (1ll << 31), BN_UNSUPPORTED, "factor with %lld entries (max 2^31 - 1)",
             f->len * (long long)length);
  std::vector<int32_t> od = f->dims, oc = f->card;
  od.push_back(dim);
  oc.push_back(length);
  BN_CUDA_TRY(cudaSetDevice(f->device));
  Factor* o = nullptr;
  BN_TRY(new_factor(od, oc, f->device, &o));
  bool handled = false;  // the table repeated `length` times: one load, `length` stores (factor_stream.cu)
  int32_t rc = push_duplicate(f->data.p, f->len, length, o->data.p, &handled);
  if (rc == BN_OK && !handled) rc = contract(f->device, {Operand{f->data.p, f->dims, f->card}}, od, oc, {}, {}, o->data.p);
  if (rc != BN_OK) { delete o; return rc; }
  *out = register_factor(o);
  return BN_OK;
}

// norm pass shared by normalize! and the split-factor entry points: total (device) = sum |x|^p, deterministic
static int32_t norm_total(Factor* f, int32_t pnorm, double** total_dev, int* grid_out) {
  const long long grid_ll = std::max<long long>(1, (f->len / 2 + NORM_CHUNK - 1) / NORM_CHUNK);
  BN_REQUIRE(grid_ll < (1ll << 31), BN_UNSUPPORTED, "factor too large to normalise");
  const int grid = (int)grid_ll;
  static thread_local DevBuf<double> scratch;      // [total | grid partials]
  static thread_local DevBuf<unsigned int> ticket;
  static thread_local int scratch_dev = -1;
  if (scratch_dev != f->device || scratch.n < (size_t)grid + 1) {
    BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));  // an earlier normalise may still be reading the old scratch
    BN_CUDA_TRY(scratch.alloc(std::max<size_t>((size_t)grid + 1, 16384)));
    BN_CUDA_TRY(ticket.alloc(1));
    BN_CUDA_TRY(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), tl_stream));
    scratch_dev = f->device;
  }
  if (pnorm) {
    norm_reduce_kernel<<<grid, NORM_THR, 0, tl_stream>>>(f->data.p, f->len, pnorm, scratch.p + 1, ticket.p, scratch.p);
    count_launch();
  }
  *total_dev = scratch.p;
  *grid_out = grid;
  return BN_OK;
}

int32_t bn_factor_normalize(bn_factor_t h, int32_t pnorm) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_normalize: unknown handle");
  BN_REQUIRE(pnorm == 1 || pnorm == 2, BN_BAD_ARG, "p = %d is not supported", pnorm);  // factors_dims.jl:51
  BN_REQUIRE(f->data.owns, BN_BAD_ARG, "bn_factor_normalize: the factor is a read-only view");
  BN_CUDA_TRY(cudaSetDevice(f->device));
  double* total = nullptr;
  int grid = 0;
  BN_TRY(norm_total(f, pnorm, &total, &grid));
  norm_scale_kernel<<<grid, NORM_THR, 0, tl_stream>>>(f->data.p, f->len, total);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

int32_t bn_factor_norm(bn_factor_t h, int32_t pnorm, double* out_total) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f && out_total, BN_BAD_ARG, "bn_factor_norm: bad arguments");
  BN_REQUIRE(pnorm == 1 || pnorm == 2, BN_BAD_ARG, "p = %d is not supported", pnorm);
  BN_CUDA_TRY(cudaSetDevice(f->device));
  double* total = nullptr;
  int grid = 0;
  BN_TRY(norm_total(f, pnorm, &total, &grid));
  BN_CUDA_TRY(cudaMemcpyAsync(out_total, total, sizeof(double), cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

int32_t bn_factor_div_scalar(bn_factor_t h, double total) {
  Factor* f = get_factor(h);
  BN_REQUIRE(f, BN_BAD_ARG, "bn_factor_div_scalar: unknown handle");
  BN_REQUIRE(f->data.owns, BN_BAD_ARG, "bn_factor_div_scalar: the factor is a read-only view");
  BN_CUDA_TRY(cudaSetDevice(f->device));
  double* tdev = nullptr;
  int grid = 0;
  BN_TRY(norm_total(f, 0, &tdev, &grid));  // scratch only, no reduction
  BN_CUDA_TRY(cudaMemcpyAsync(tdev, &total, sizeof(double), cudaMemcpyHostToDevice, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));  // `total` lives on this stack frame
  norm_scale_kernel<<<grid, NORM_THR, 0, tl_stream>>>(f->data.p, f->len, tdev);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

}  // extern "C"
