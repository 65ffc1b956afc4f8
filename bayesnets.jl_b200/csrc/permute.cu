// permute.cu -- permutedims(phi, perm) (src/Factors/factors_main.jl:160-166) as a tiled transpose.
//
// The reference calls Base.permutedims on the N-d array: a strided copy, cache-hostile on one side.  Through the
// contraction kernel (a gather on the read side) the same permutation reaches 0.26 of the HBM peak at 2^24 entries:
// a warp's 32 loads land in 32 different sectors.  Here a tile is the cross product of
//   A = the input's fastest dims (>= 32 contiguous doubles on the read side) and
//   B = the output's fastest dims (>= 32 contiguous doubles on the write side),
// grown with further dims to ~2048 entries; a CTA reads the tile in INPUT order (coalesced 256-byte warp loads), parks
// it in shared memory, and writes it in OUTPUT order (coalesced 256-byte warp stores).  What that takes:
//   * dims that are neighbours in both the input and the output are merged on the host first (a rotation of 24 binary
//     dims is 4 merged dims), dims of size 1 dropped;
//   * the per-entry address arithmetic is tile-invariant: each CTA tabulates, once, the input offset of the j-th entry
//     in read order and the (output offset, shared-memory slot) of the j-th entry in write order (mixed-radix
//     decomposition over the <= 16 tile dims), then walks tiles persistently with three table lookups per entry;
//   * the tile's base offsets come from the remaining (<= 32, merged) dims: lane d of warp 0 extracts digit d of the
//     tile index, a warp shuffle adds the contributions;
//   * shared-memory slots are XOR-swizzled, slot = j ^ fold4(j >> 4): 16 consecutive entries in read order stay a
//     permutation of one 128-byte line, and entries 2^m apart in read order (what 16 neighbouring lanes touch in write
//     order) fall into 16 different 8-byte banks for every m -- no padding, no conflicts for power-of-two dims.
// Pure data movement: the result is identical to the oracle's.  Shapes the tiling does not cover (a tile above 4096
// entries, more than 32 remaining dims) return "not handled" and bn_factor_permutedims falls back to the contraction
// kernel.  Algorithmic traffic: 16 bytes per entry (read once, written once).
#include <algorithm>

#include "common.cuh"

namespace bn {

namespace {

constexpr int PT_THR = 256;
constexpr int PT_MAXT = 16;    // tile dims (merged)
constexpr int PT_MAXR = 32;    // remaining dims (merged): one lane each
constexpr int PT_TILE_MAX = 4096;
constexpr int PT_TILE_AIM = 2048;
constexpr int PT_U = 8;        // loads in flight per thread per pass
// the prefetched tile is held in PASSES x PT_U registers per thread: PASSES = 1 for tiles <= 2048 entries, else 2

struct PermuteParams {
  const double* in;
  double* out;
  int nt, nr, T;
  uint32_t n_tiles;
  int32_t rcard[PT_MAXT], ris[PT_MAXT];                 // tile dims in read order: size, input stride
  int32_t wcard[PT_MAXT], wos[PT_MAXT], wrs[PT_MAXT];  // tile dims in write order: size, output stride, read-order stride
  uint32_t tcard[PT_MAXR], tpre[PT_MAXR];              // remaining dims: size, product of the earlier sizes
  long long tis[PT_MAXR], tos[PT_MAXR];                // their input / output strides
};

__device__ __forceinline__ uint32_t swz(uint32_t j) { return j ^ (((j >> 4) ^ (j >> 8) ^ (j >> 12)) & 15u); }

template <int PASSES, int MINB>
__global__ void __launch_bounds__(PT_THR, MINB) permute_kernel(const PermuteParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);              // [T16] (T rounded up to 16)
  const int T16 = (p.T + 15) & ~15;
  int32_t* rd_off = reinterpret_cast<int32_t*>(tile + T16);        // [T]
  int32_t* wr_off = rd_off + p.T;                                  // [T]
  uint16_t* wr_slot = reinterpret_cast<uint16_t*>(wr_off + p.T);   // [T]
  __shared__ long long s_base[2][2];  // by tile parity: warp 0 prepares the next tile while the others still store
  const int tid = threadIdx.x;
  for (int j = tid; j < p.T; j += PT_THR) {
    uint32_t r = (uint32_t)j;
    int32_t off = 0;
    for (int d = 0; d < p.nt; ++d) {
      const uint32_t q = r / (uint32_t)p.rcard[d];
      off += (int32_t)(r - q * (uint32_t)p.rcard[d]) * p.ris[d];
      r = q;
    }
    rd_off[j] = off;
    r = (uint32_t)j;
    int32_t oo = 0, rj = 0;
    for (int d = 0; d < p.nt; ++d) {
      const uint32_t q = r / (uint32_t)p.wcard[d];
      const int32_t dg = (int32_t)(r - q * (uint32_t)p.wcard[d]);
      oo += dg * p.wos[d];
      rj += dg * p.wrs[d];
      r = q;
    }
    wr_off[j] = oo;
    wr_slot[j] = (uint16_t)swz((uint32_t)rj);
  }
  // bases of tile t into s_base[par] (warp 0: lane d takes digit d of the tile index)
  auto tile_bases = [&](uint32_t t, int par) {
    if (tid < 32) {
      long long bi = 0, bo = 0;
      if (tid < p.nr) {
        const uint32_t dg = (t / p.tpre[tid]) % p.tcard[tid];
        bi = (long long)dg * p.tis[tid];
        bo = (long long)dg * p.tos[tid];
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        bi += __shfl_xor_sync(0xffffffffu, bi, s);
        bo += __shfl_xor_sync(0xffffffffu, bo, s);
      }
      if (tid == 0) { s_base[par][0] = bi; s_base[par][1] = bo; }
    }
  };
  // Software pipeline over this CTA's tiles: the loads of tile k+1 are issued (into registers) BEFORE tile k is
  // stored from shared memory, so the read latency of one tile hides behind the write phase of the other.
  // T <= PT_THR * PT_U * PASSES entries per tile.
  uint32_t t = blockIdx.x;
  if (t >= p.n_tiles) return;
  int par = 0;
  tile_bases(t, par);
  __syncthreads();  // tables and bases visible
  long long bo_cur = s_base[par][1];
  {
    const double* __restrict__ src = p.in + s_base[par][0];
    for (int j0 = tid; j0 < p.T; j0 += PT_THR * PT_U) {
      double v[PT_U];
#pragma unroll
      for (int u = 0; u < PT_U; ++u) {
        const int j = j0 + u * PT_THR;
        if (j < p.T) v[u] = __ldcs(src + rd_off[j]);
      }
#pragma unroll
      for (int u = 0; u < PT_U; ++u) {
        const int j = j0 + u * PT_THR;
        if (j < p.T) tile[swz((uint32_t)j)] = v[u];
      }
    }
  }
  for (;;) {
    const uint32_t tn = t + gridDim.x;
    const bool more = tn < p.n_tiles;
    if (more) tile_bases(tn, par ^ 1);
    __syncthreads();  // tile t resident in shared memory; bases of tile tn visible
    double v[PASSES][PT_U];
    long long bo_next = 0;
    if (more) {
      const double* __restrict__ src = p.in + s_base[par ^ 1][0];
      bo_next = s_base[par ^ 1][1];
#pragma unroll
      for (int w = 0; w < PASSES; ++w)
#pragma unroll
        for (int u = 0; u < PT_U; ++u) {
          const int j = tid + (w * PT_U + u) * PT_THR;
          if (j < p.T) v[w][u] = __ldcs(src + rd_off[j]);
        }
    }
    {
      double* __restrict__ dst = p.out + bo_cur;
      for (int j0 = tid; j0 < p.T; j0 += PT_THR * PT_U) {
#pragma unroll
        for (int u = 0; u < PT_U; ++u) {
          const int j = j0 + u * PT_THR;
          if (j < p.T) __stcs(dst + wr_off[j], tile[wr_slot[j]]);
        }
      }
    }
    if (!more) break;
    __syncthreads();  // every slot of tile t has been read
#pragma unroll
    for (int w = 0; w < PASSES; ++w)
#pragma unroll
      for (int u = 0; u < PT_U; ++u) {
        const int j = tid + (w * PT_U + u) * PT_THR;
        if (j < p.T) tile[swz((uint32_t)j)] = v[w][u];
      }
    t = tn;
    bo_cur = bo_next;
    par ^= 1;
  }
}

// The same kernel moving PAIRS: when the tile's first read dim and first write dim are both even (binary nets: always)
// entries 2jj, 2jj + 1 are neighbours in the input (read order) resp. in the output (write order), so both global
// sides are 128-bit accesses and the tables shrink to one entry per pair.  A pair's two slots differ only in bit 0
// (the swizzle touches the low 4 bits with a mask that depends on j >> 4), so the read phase parks a pair with one
// 128-bit shared store, swapped when the first slot is odd.
__global__ void __launch_bounds__(PT_THR, 4) permute_pairs_kernel(const PermuteParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);                // [T16]
  const int T16 = (p.T + 15) & ~15;
  const int P = p.T >> 1;                                            // pairs per tile (<= PT_THR * PT_U)
  int32_t* rd_off = reinterpret_cast<int32_t*>(tile + T16);          // [P] input offset of pair jj (read order)
  int32_t* wr_off = rd_off + P;                                      // [P] output offset of pair jj (write order)
  uint32_t* wr_slot = reinterpret_cast<uint32_t*>(wr_off + P);       // [P] its two slots, 16 bits each
  __shared__ long long s_base[2][2];
  const int tid = threadIdx.x;
  for (int jj = tid; jj < P; jj += PT_THR) {
    uint32_t r = 2u * (uint32_t)jj;
    int32_t off = 0;
    for (int d = 0; d < p.nt; ++d) {
      const uint32_t q = r / (uint32_t)p.rcard[d];
      off += (int32_t)(r - q * (uint32_t)p.rcard[d]) * p.ris[d];
      r = q;
    }
    rd_off[jj] = off;
    r = 2u * (uint32_t)jj;
    int32_t oo = 0, rj = 0;
    for (int d = 0; d < p.nt; ++d) {
      const uint32_t q = r / (uint32_t)p.wcard[d];
      const int32_t dg = (int32_t)(r - q * (uint32_t)p.wcard[d]);
      oo += dg * p.wos[d];
      rj += dg * p.wrs[d];
      r = q;
    }
    wr_off[jj] = oo;
    wr_slot[jj] = swz((uint32_t)rj) | (swz((uint32_t)(rj + p.wrs[0])) << 16);  // next value of the first write dim
  }
  auto tile_bases = [&](uint32_t t, int par) {
    if (tid < 32) {
      long long bi = 0, bo = 0;
      if (tid < p.nr) {
        const uint32_t dg = (t / p.tpre[tid]) % p.tcard[tid];
        bi = (long long)dg * p.tis[tid];
        bo = (long long)dg * p.tos[tid];
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        bi += __shfl_xor_sync(0xffffffffu, bi, s);
        bo += __shfl_xor_sync(0xffffffffu, bo, s);
      }
      if (tid == 0) { s_base[par][0] = bi; s_base[par][1] = bo; }
    }
  };
  auto park = [&](const double2 (&v)[PT_U]) {
#pragma unroll
    for (int u = 0; u < PT_U; ++u) {
      const int jj = tid + u * PT_THR;
      if (jj < P) {
        const uint32_t sl = swz(2u * (uint32_t)jj);
        *reinterpret_cast<double2*>(tile + (sl & ~1u)) = (sl & 1u) ? make_double2(v[u].y, v[u].x) : v[u];
      }
    }
  };
  uint32_t t = blockIdx.x;
  if (t >= p.n_tiles) return;
  int par = 0;
  tile_bases(t, par);
  __syncthreads();
  long long bo_cur = s_base[par][1];
  {
    const double* __restrict__ src = p.in + s_base[par][0];
    double2 v[PT_U];
#pragma unroll
    for (int u = 0; u < PT_U; ++u) {
      const int jj = tid + u * PT_THR;
      if (jj < P) v[u] = __ldcs(reinterpret_cast<const double2*>(src + rd_off[jj]));
    }
    park(v);
  }
  for (;;) {
    const uint32_t tn = t + gridDim.x;
    const bool more = tn < p.n_tiles;
    if (more) tile_bases(tn, par ^ 1);
    __syncthreads();  // tile t resident in shared memory; bases of tile tn visible
    double2 v[PT_U];
    long long bo_next = 0;
    if (more) {
      const double* __restrict__ src = p.in + s_base[par ^ 1][0];
      bo_next = s_base[par ^ 1][1];
#pragma unroll
      for (int u = 0; u < PT_U; ++u) {
        const int jj = tid + u * PT_THR;
        if (jj < P) v[u] = __ldcs(reinterpret_cast<const double2*>(src + rd_off[jj]));
      }
    }
    {
      double* __restrict__ dst = p.out + bo_cur;
#pragma unroll
      for (int u = 0; u < PT_U; ++u) {
        const int jj = tid + u * PT_THR;
        if (jj < P) {
          const uint32_t sl = wr_slot[jj];
          __stcs(reinterpret_cast<double2*>(dst + wr_off[jj]), make_double2(tile[sl & 0xffffu], tile[sl >> 16]));
        }
      }
    }
    if (!more) break;
    __syncthreads();  // every slot of tile t has been read
    park(v);
    t = tn;
    bo_cur = bo_next;
    par ^= 1;
  }
}

struct MDim {
  long long card, is, os;  // size, input stride, output stride
};

}  // namespace

// out (dims = in dims permuted by perm: out dim i = in dim perm[i]) <- in.  *handled = false when the shape is outside
// what the tiling covers (nothing launched).
int32_t permute_tiled(int device, const double* in, const std::vector<int32_t>& card, const int32_t* perm, double* out,
                      bool* handled) {
  *handled = false;
  const int n = (int)card.size();
  std::vector<long long> is((size_t)n, 1), os_in((size_t)n, 1);
  for (int i = 1; i < n; ++i) is[(size_t)i] = is[(size_t)i - 1] * card[(size_t)i - 1];
  long long N = 1, acc = 1;
  for (int i = 0; i < n; ++i) N *= card[(size_t)i];
  for (int i = 0; i < n; ++i) { os_in[(size_t)perm[i]] = acc; acc *= card[(size_t)perm[i]]; }
  if (N < 2 || N >= (1ll << 31)) return BN_OK;
  // size-1 dims dropped; the tile is chosen on the ORIGINAL dims (a merged dim could be larger than any tile)
  std::vector<MDim> d;
  for (int i = 0; i < n; ++i)
    if (card[(size_t)i] != 1) d.push_back({card[(size_t)i], is[(size_t)i], os_in[(size_t)i]});
  int m = (int)d.size();
  std::vector<int> by_out((size_t)m);
  for (int i = 0; i < m; ++i) by_out[(size_t)i] = i;
  std::sort(by_out.begin(), by_out.end(), [&](int a, int b) { return d[(size_t)a].os < d[(size_t)b].os; });
  // tile = input prefix (>= 32 entries) + output prefix (>= 32 entries), then alternately the next output / input
  // dim until ~2048 entries
  std::vector<char> in_tile((size_t)m, 0);
  long long T = 1;
  auto add = [&](int i) { in_tile[(size_t)i] = 1; T *= d[(size_t)i].card; };
  long long run = 1;
  for (int i = 0; i < m && run < 32; ++i) { run *= d[(size_t)i].card; add(i); }
  run = 1;
  for (int k = 0; k < m && run < 32; ++k) {
    const int i = by_out[(size_t)k];
    run *= d[(size_t)i].card;
    if (!in_tile[(size_t)i]) add(i);
  }
  if (T > PT_TILE_MAX) return BN_OK;
  static const long long aim = getenv("BN_PERMUTE_TILE") ? atoll(getenv("BN_PERMUTE_TILE")) : PT_TILE_AIM;
  for (bool from_out = true;; from_out = !from_out) {
    if (T >= aim) break;
    int pick = -1;
    for (int pass = 0; pass < 2 && pick < 0; ++pass) {
      const bool o = (pass == 0) ? from_out : !from_out;
      for (int k = 0; k < m; ++k) {
        const int i = o ? by_out[(size_t)k] : k;
        if (in_tile[(size_t)i]) continue;
        if (T * d[(size_t)i].card <= PT_TILE_MAX) pick = i;
        break;  // only the NEXT dim of that order keeps the runs contiguous
      }
    }
    if (pick < 0) break;
    add(pick);
  }
  // now merge input-neighbours that are output-neighbours in the same order and on the same side of the tile / rest
  // split (a rotation of 24 binary dims: 2 tile dims of 32 and 64, 2 remaining dims of 4 and 2048)
  {
    std::vector<MDim> md;
    std::vector<char> mt;
    for (int i = 0; i < m; ++i) {
      if (!md.empty() && mt.back() == in_tile[(size_t)i] && md.back().os * md.back().card == d[(size_t)i].os &&
          md.back().is * md.back().card == d[(size_t)i].is) {
        md.back().card *= d[(size_t)i].card;
        continue;
      }
      md.push_back(d[(size_t)i]);
      mt.push_back(in_tile[(size_t)i]);
    }
    d.swap(md);
    in_tile.swap(mt);
    m = (int)d.size();
    by_out.resize((size_t)m);
    for (int i = 0; i < m; ++i) by_out[(size_t)i] = i;
    std::sort(by_out.begin(), by_out.end(), [&](int a, int b) { return d[(size_t)a].os < d[(size_t)b].os; });
  }
  PermuteParams p{};
  p.in = in;
  p.out = out;
  p.T = (int)T;
  // tile dims in read order (ascending input stride = index order) and in write order
  std::vector<int> rd, wr;
  for (int i = 0; i < m; ++i) if (in_tile[(size_t)i]) rd.push_back(i);
  for (int k = 0; k < m; ++k) if (in_tile[(size_t)by_out[(size_t)k]]) wr.push_back(by_out[(size_t)k]);
  if ((int)rd.size() > PT_MAXT) return BN_OK;
  p.nt = (int)rd.size();
  std::vector<long long> rstride((size_t)m, 0);
  long long rs = 1;
  for (int k = 0; k < p.nt; ++k) {
    const int i = rd[(size_t)k];
    p.rcard[k] = (int32_t)d[(size_t)i].card;
    p.ris[k] = (int32_t)d[(size_t)i].is;
    rstride[(size_t)i] = rs;
    rs *= d[(size_t)i].card;
  }
  for (int k = 0; k < p.nt; ++k) {
    const int i = wr[(size_t)k];
    p.wcard[k] = (int32_t)d[(size_t)i].card;
    p.wos[k] = (int32_t)d[(size_t)i].os;
    p.wrs[k] = (int32_t)rstride[(size_t)i];
  }
  // remaining dims, output order (tiles that are neighbours in time write neighbouring output)
  long long pre = 1;
  p.nr = 0;
  for (int k = 0; k < m; ++k) {
    const int i = by_out[(size_t)k];
    if (in_tile[(size_t)i]) continue;
    if (p.nr == PT_MAXR) return BN_OK;
    p.tcard[p.nr] = (uint32_t)d[(size_t)i].card;
    p.tpre[p.nr] = (uint32_t)pre;
    p.tis[p.nr] = d[(size_t)i].is;
    p.tos[p.nr] = d[(size_t)i].os;
    pre *= d[(size_t)i].card;
    ++p.nr;
  }
  p.n_tiles = (uint32_t)pre;
  const size_t smem = (size_t)((T + 15) & ~15ll) * 8 + (size_t)T * 10 + 16;
  // resident CTAs per SM: measured at 2^24 entries (rotation of 24 binary dims, 2048-entry tiles): 4 CTAs 0.76 of the
  // HBM peak, 5 CTAs 0.74, 6 CTAs 0.63 -- more CTAs open more DRAM pages at once
  static const int ctas_env = getenv("BN_PERMUTE_CTAS") ? atoi(getenv("BN_PERMUTE_CTAS")) : 4;
  // 128-bit accesses on both global sides when the first read dim and the first write dim are even and contiguous
  static const bool no_pairs = getenv("BN_PERMUTE_NO_PAIRS") != nullptr;
  const bool pairs = !no_pairs && p.nt >= 1 && (p.rcard[0] % 2) == 0 && p.ris[0] == 1 && (p.wcard[0] % 2) == 0 && p.wos[0] == 1 &&
                     (reinterpret_cast<uintptr_t>(in) % 16) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0;
  void (*kern)(const PermuteParams) = pairs ? permute_pairs_kernel
                                            : (T <= PT_THR * PT_U ? permute_kernel<1, 4> : permute_kernel<2, 4>);
  BN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)PT_TILE_MAX * 18 + 256)));
  const DevProps& dp = dev_props(device);
  int per_sm = 0;
  BN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT_THR, smem));
  per_sm = std::max(1, std::min(per_sm, ctas_env));
  const int grid = (int)std::min<long long>((long long)p.n_tiles, (long long)dp.sm_count * per_sm);
  kern<<<grid, PT_THR, smem, tl_stream>>>(p);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return BN_OK;
}

}  // namespace bn
