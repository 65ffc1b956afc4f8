// stats.cu -- sufficient statistics of a discrete data table under a network structure.
//
// Replaces statistics(parent_list, ncategories, data) (src/DiscreteBayesNet/discrete_bayes_net.jl:151-172,221-231):
// for every node i a count table N_i[k, j] = #rows with x_i = k and parental instantiation j (first parent
// fastest).  The reference builds it per node with one sparse-matrix constructor over the whole data matrix; here
// ONE pass over the table counts all nodes at once.  SURVEY.md section 8(f) rank 4: the step after sampling in a
// learn-from-samples loop -- the input layout is exactly what bn_sample_dev writes (uint8, node-major,
// data[i * pitch + row]), so  rand(bn, n) -> statistics -> fit / bayesian_score  never leaves HBM.
//
// Output layout = the network's flat CPT layout (include/bn_cuda.h): counts[cpt_off[i] + k + card_i * j], which is
// the reference's r x q matrix in column-major order.
//
// Design (B200): persistent CTAs.  A tile of TR rows x all nodes is staged into shared memory with 16-byte
// asynchronous copies (cp.async), double-buffered so the next tile streams in while the current one is counted.  Counting is OWNER-COMPUTES, WITHOUT ATOMICS: inside a warp the 32 lanes are 32 different
// nodes, and a (row slice, node group) pair belongs to exactly one warp of the CTA, so a lane is the only thread of
// its CTA that ever touches its node's count table in its slice's histogram copy -- a plain load / add / store.
// Each lane takes four consecutive rows per step (one 32-bit load per column carries their four states).
// Measured on the C2 net (100 nodes, 2^24 rows, B200): with shared-memory atomics 3.98e9 rows/s (32 rows of one node
// per warp: a binary root has two bins for 32 lanes) and 3.79e9 (nodes across lanes, no same-address conflicts) --
// the atomic unit, not the conflicts, was the limit; owner-computes with a branchy inner loop 2.7e9 (72 instructions
// per lane-row, ncu), branch-free 4.25e9 (38 per lane-row; issue-bound at 25% occupancy, profiles/r1f_stats_*), four
// rows per step with 32-bit state loads 6.5e9.  At that point the kernel issues ~0.7 shared-memory wavefronts per clock
// per SM (state loads + histogram read-modify-writes, half of them bank-conflict replays): occupancy (25 -> 50%), a
// second histogram copy per slice for independent read-modify-writes and more row slices all measured within 3%.
// A lane keeps its node's program (table offset, first parents and strides) in registers for the whole tile.
// Column i of a tile starts at byte i * (TR + 16), i.e. in bank group 4 * (i % 8); lane i reads row r + 4 * (i / 8),
// which moves the four lanes of a bank group onto four different banks: the state loads are conflict free.  The
// histogram copies are summed and flushed once per CTA with 64-bit global atomics.  Algorithmic traffic: n_nodes
// bytes per row, read once.  Count tables too large for shared memory fall back to global atomics.
#include "common.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <type_traits>

namespace bn {

constexpr int STATS_REG_PARENTS = 3;  // parents held in registers; further ones are read from the program

struct StatsArgs {
  const uint8_t* data;   // [n_nodes][pitch]
  uint64_t n_rows, pitch;
  const uint32_t* prog;  // device copy: per node 4 + 2 * max(0, n_par - 3) words, see statistics_launch
  const uint32_t* prog_off;  // [n_nodes] word offset of each node's record
  uint32_t prog_words, n_nodes, n_bins, TR;
  uint32_t hist_in_smem, use_async;  // use_async: table and pitch 16-byte aligned
  uint32_t bins_pad, copies;                // words per histogram copy; copies (1 or 2) per row slice
  uint32_t n_groups, group_step, row_step;  // warp w: node groups w % group_step + k * group_step, row slice w / group_step
  unsigned long long* out;  // [n_bins]
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MORE: some node has more than STATS_REG_PARENTS parents (generic tail loop); SMEM: the histogram copies fit
template <bool MORE, bool SMEM>
__global__ void __launch_bounds__(1024, 1) statistics_kernel(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [prog | prog_off | hist u32 (n_bins, when it fits) | tile 0 | tile 1], tiles are [n_nodes][TR + 16] u8
  uint32_t* prog = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* prog_off = prog + a.prog_words;
  uint32_t* hist = prog_off + a.n_nodes;
  const uint32_t hist_words = SMEM ? a.bins_pad * a.copies * a.row_step : 0u;
  const uint32_t TRP = a.TR + 16u;
  const uint32_t tile_bytes = a.n_nodes * TRP;
  unsigned char* tiles = smem_raw + (((a.prog_words + a.n_nodes + hist_words) * 4u + 127u) & ~127u);
  const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, warp = tid >> 5;
  for (uint32_t w = tid; w < a.prog_words; w += T) prog[w] = a.prog[w];
  for (uint32_t w = tid; w < a.n_nodes; w += T) prog_off[w] = a.prog_off[w];
  for (uint32_t b = tid; b < hist_words; b += T) hist[b] = 0u;
  __syncthreads();

  const uint64_t n_tiles = (a.n_rows + a.TR - 1) / a.TR;
  // stage tile `t` into buffer `buf`: 16-byte asynchronous copies (cp.async / LDGSTS) spread over all threads.  A tile
  // is n_nodes column segments of only TR = 128..512 bytes; one TMA bulk copy per segment was measured to cap at
  // about one copy per 150 cycles per SM (1.6-1.9e9 copies/s chip-wide), which bounded the whole kernel, so the
  // segments are moved as independent 16-byte requests instead.  The ragged last tile and unaligned tables use
  // plain byte loads.
  auto stage = [&](uint64_t t, uint32_t buf) {
    unsigned char* dst = tiles + buf * tile_bytes;
    const uint64_t r0 = t * a.TR;
    if (a.use_async && (t + 1) * a.TR <= a.n_rows) {
      const uint32_t per_col = a.TR / 16u, total = a.n_nodes * per_col;
      for (uint32_t c = tid; c < total; c += T) {
        const uint32_t i = c / per_col, k = c - i * per_col;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_addr(dst + i * TRP + 16u * k)),
                     "l"(a.data + (uint64_t)i * a.pitch + r0 + 16u * k)
                     : "memory");
      }
    } else {
      const uint64_t rows = a.n_rows - r0 < a.TR ? a.n_rows - r0 : a.TR;
      for (uint32_t i = 0; i < a.n_nodes; ++i)
        for (uint32_t r = tid; r < rows; r += T) dst[i * TRP + r] = a.data[(uint64_t)i * a.pitch + r0 + r];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint64_t t = blockIdx.x;
  if (t < n_tiles) stage(t, 0);
  uint32_t buf = 0;
  const uint32_t skew = 4u * (lane >> 3);
  for (; t < n_tiles; t += gridDim.x, buf ^= 1u) {
    // one barrier per tile: it publishes tile t (every thread has waited for its own copies) and proves that every
    // thread is done with the other buffer, which is refilled right after it
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint64_t tn = t + gridDim.x;
    if (tn < n_tiles) stage(tn, buf ^ 1u);
    const unsigned char* tile = tiles + buf * tile_bytes;
    const uint64_t r0 = t * a.TR;
    const uint32_t rows = (uint32_t)(a.n_rows - r0 < a.TR ? a.n_rows - r0 : a.TR);
    for (uint32_t g = warp % a.group_step; g < a.n_groups; g += a.group_step) {
      const uint32_t node = g * 32u + lane;
      const bool live = node < a.n_nodes;
      // this lane's node program -> registers
      uint32_t base = 0, npar = 0, po[STATS_REG_PARENTS], ps[STATS_REG_PARENTS];
      const uint32_t* rec = prog;
      if (live) {
        rec = prog + prog_off[node];
        base = rec[0];
        npar = rec[1];
      }
#pragma unroll
      for (int k = 0; k < STATS_REG_PARENTS; ++k) {
        po[k] = (live && (uint32_t)k < npar) ? rec[2 + 2 * k] * TRP : 0u;
        ps[k] = (live && (uint32_t)k < npar) ? rec[3 + 2 * k] : 0u;  // stride 0: absent parents add nothing
      }
      // Branch-free inner loop: every lane always reads (in-bounds) states and always updates a bin; rows / lanes that
      // do not count (past the table end, 0xFF-marked, lanes beyond the last node, states out of range) are steered
      // to the spare bin n_bins of their copy (bins_pad > n_bins), which is never flushed.
      const uint32_t coff = live ? node * TRP : 0u;
      const uint32_t slice = warp / a.group_step;
      uint32_t* hA = hist + (a.copies * slice) * a.bins_pad;
      uint32_t* hB = hA + (a.copies - 1u) * a.bins_pad;
      const uint32_t spare = a.n_bins;
      // four consecutive rows per step: one 32-bit load per (node | parent) column carries their four states
      auto quad = [&](uint32_t Q, uint32_t (&idx)[4]) {
        uint32_t r = 4u * Q + skew;
        r = r >= a.TR ? r - a.TR : r;
        const uint32_t ws = *reinterpret_cast<const uint32_t*>(tile + coff + r);
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(tile + r);  // node 0's states carry the 0xFF marks
        uint32_t wp[STATS_REG_PARENTS];
#pragma unroll
        for (int k = 0; k < STATS_REG_PARENTS; ++k) wp[k] = *reinterpret_cast<const uint32_t*>(tile + po[k] + r);
        const bool quad_ok = live && 4u * Q < a.TR;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t v = base + ((ws >> (8 * b)) & 0xffu);
#pragma unroll
          for (int k = 0; k < STATS_REG_PARENTS; ++k) v += ps[k] * ((wp[k] >> (8 * b)) & 0xffu);
          if (MORE)
            for (uint32_t k = STATS_REG_PARENTS; k < npar; ++k) v += rec[3 + 2 * k] * (uint32_t)tile[rec[2 + 2 * k] * TRP + r + b];
          const bool ok = quad_ok && r + b < rows && ((w0 >> (8 * b)) & 0xffu) != 0xFFu && v < spare;
          idx[b] = ok ? v : spare;
        }
      };
      const uint32_t n_quads = a.TR / 4u;
      if (SMEM) {
        for (uint32_t Q = slice; Q < n_quads; Q += a.row_step) {
          uint32_t i[4];
          quad(Q, i);
          // this lane is the only thread of the CTA that touches node `node`'s bins in copies 2 * slice (+ 1); the
          // spare bins are shared by the warp's lanes, but nobody reads them.  Rows 0, 2 -> copy A, rows 1, 3 -> copy B.
          if (a.copies == 2u) {
            const uint32_t v0 = hA[i[0]], v1 = hB[i[1]];
            hA[i[0]] = v0 + 1u;
            hB[i[1]] = v1 + 1u;
            const uint32_t v2 = hA[i[2]], v3 = hB[i[3]];
            hA[i[2]] = v2 + 1u;
            hB[i[3]] = v3 + 1u;
          } else {  // one copy: the four read-modify-writes may hit the same bin, so they stay in program order
#pragma unroll
            for (int b = 0; b < 4; ++b) hA[i[b]] += 1u;
          }
        }
      } else {
        for (uint32_t Q = slice; Q < n_quads; Q += a.row_step) {
          uint32_t i[4];
          quad(Q, i);
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (i[b] != spare) atomicAdd(&a.out[i[b]], 1ull);
        }
      }
    }
  }
  __syncthreads();
  if (SMEM)
    for (uint32_t b = tid; b < a.n_bins; b += T) {
      unsigned long long c = 0;
      for (uint32_t k = 0; k < a.copies * a.row_step; ++k) c += hist[k * a.bins_pad + b];
      if (c) atomicAdd(&a.out[b], c);
    }
}

// The node-lane kernel above behind its launch logic: the fallback for families with more than 256 bins (a bin index
// must fit one byte for the row-lane kernel below) and for networks too wide for its tiles.  out_dev is already zeroed.
static int32_t statistics_launch_nodelanes(Net* net, const uint8_t* data_dev, uint64_t n_rows, uint64_t pitch,
                                           unsigned long long* out_dev) {
  const DevProps& dp = dev_props(net->device);
  cudaStream_t s = tl_stream;
  const int n = net->n;
  const int64_t n_bins = net->cpt_off[n];
  // per-node program: [table offset, n_par, (parent, stride) x n_par], stride includes the node's own cardinality
  std::vector<uint32_t> prog, prog_off((size_t)n);
  prog.reserve((size_t)n * 8);
  for (int i = 0; i < n; ++i) {
    const uint32_t npar = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
    prog_off[(size_t)i] = (uint32_t)prog.size();
    prog.push_back((uint32_t)net->cpt_off[i]);
    prog.push_back(npar);
    int64_t stride = net->card[i];
    for (int k = net->par_ptr[i]; k < net->par_ptr[i + 1]; ++k) {
      prog.push_back((uint32_t)net->par_idx[k]);
      prog.push_back((uint32_t)stride);
      stride *= net->card[net->par_idx[k]];
    }
    while (prog.size() - prog_off[(size_t)i] < 2 + 2 * STATS_REG_PARENTS) prog.push_back(0);  // register slots readable
  }
  while (prog.size() % 4) prog.push_back(0);
  // shape: one warp per (row slice, node group of 32 nodes); 2 histogram copies per row slice; two tiles of TR rows;
  // everything within 227 KB, aiming for >= 2 CTAs per SM
  StatsArgs a{};
  const size_t prog_b = (prog.size() + (size_t)n) * 4;
  const size_t budget = (size_t)dp.max_smem_optin - 2048;
  a.n_groups = (uint32_t)((n + 31) / 32);
  a.group_step = std::min<uint32_t>(a.n_groups, 32u);
  a.bins_pad = (uint32_t)((n_bins + 32) / 32 * 32 + 1);  // odd pitch: the copies of a bin sit in different banks
  const size_t min_tiles = 2 * (size_t)n * (32 + 16);
  static const uint32_t copies_env = getenv("BN_STATS_COPIES") ? (uint32_t)atoi(getenv("BN_STATS_COPIES")) : 0u;
  static const uint32_t slices_env = getenv("BN_STATS_SLICES") ? (uint32_t)atoi(getenv("BN_STATS_SLICES")) : 0u;
  const uint32_t copies = copies_env == 2u ? 2u : 1u;
  a.copies = copies;
  uint32_t slices = std::max<uint32_t>(1u, std::min<uint32_t>(4u, 8u / a.group_step));  // ~8 warps per CTA
  if (slices_env) slices = slices_env;
  while (slices > 1 && prog_b + (size_t)a.bins_pad * 4 * copies * slices + 2 * (size_t)n * (128 + 16) > budget / 2) slices /= 2;
  a.hist_in_smem = (prog_b + (size_t)a.bins_pad * 4 * copies * slices + min_tiles <= budget) ? 1u : 0u;
  a.row_step = slices;
  const size_t fixed = ((prog_b + (a.hist_in_smem ? (size_t)a.bins_pad * 4 * copies * slices : 0) + 127) / 128) * 128;
  BN_REQUIRE(fixed + min_tiles <= budget, BN_UNSUPPORTED, "network too large for the statistics kernel (%d nodes)", n);
  uint32_t TR = 512;
  while (TR > 32 && fixed + 2 * (size_t)n * (TR + 16) > budget / 2) TR /= 2;
  while (TR > 32 && fixed + 2 * (size_t)n * (TR + 16) > budget) TR /= 2;
  a.TR = TR;
  const uint32_t warps = a.group_step * slices;
  const int threads = (int)warps * 32;
  a.use_async = (pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(data_dev) % 16) == 0 && TR % 16 == 0) ? 1u : 0u;
  const size_t smem = fixed + 2 * (size_t)n * (TR + 16);
  // program -> per-call device staging from the stream-ordered pool (freed in stream order after the kernel); a copy
  // from pageable memory returns once the source has been staged, so the host vectors may go out of scope
  TmpBuf<unsigned char> stage;
  BN_CUDA_TRY(stage.alloc(prog_b + 16));
  BN_CUDA_TRY(cudaMemcpyAsync(stage.p, prog.data(), prog.size() * 4, cudaMemcpyHostToDevice, s));
  BN_CUDA_TRY(cudaMemcpyAsync(stage.p + prog.size() * 4, prog_off.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
  a.data = data_dev;
  a.n_rows = n_rows;
  a.pitch = pitch;
  a.prog = reinterpret_cast<const uint32_t*>(stage.p);
  a.prog_off = a.prog + prog.size();
  a.prog_words = (uint32_t)prog.size();
  a.n_nodes = (uint32_t)n;
  a.n_bins = (uint32_t)n_bins;
  a.out = out_dev;
  bool more = false;
  for (int i = 0; i < n; ++i) more = more || (net->par_ptr[i + 1] - net->par_ptr[i] > STATS_REG_PARENTS);
  void (*kern)(const StatsArgs) = a.hist_in_smem ? (more ? statistics_kernel<true, true> : statistics_kernel<false, true>)
                                                 : (more ? statistics_kernel<true, false> : statistics_kernel<false, false>);
  BN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin - 1024));
  int per_sm = (int)((size_t)dp.smem_per_sm / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm * threads > 2048) per_sm = 2048 / threads;
  const uint64_t n_tiles = (n_rows + TR - 1) / TR;
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)dp.sm_count * per_sm);
  BN_REQUIRE(n_rows / (uint64_t)grid < (1ull << 31), BN_UNSUPPORTED, "too many rows per CTA for the 32-bit CTA histogram");
  kern<<<grid, threads, smem, s>>>(a);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

// ------------------------------------------------------------------------------------------ row-lane kernel
// The kernel that serves every network whose families have <= 256 bins (all of BASELINE's: cards <= 4, <= 3 parents).
//
// What bounded the node-lane kernel above was shared memory: with 32 different NODES in a warp every histogram
// read-modify-write lands in 32 unrelated banks (52% of its wavefronts were conflict replays, profiles/r1f_stats_*)
// and every parent-state load is a gather.  This kernel turns the tile around: the 32 lanes of a warp are 32 ROW
// groups (8 consecutive rows each: one 64-bit load per column carries their states), a warp owns whole NODES, and
//   * a tile (256 rows x all nodes) arrives by ONE 2-D TMA copy (cp.async.bulk.tensor over a tensor map of the
//     node-major table; rows past the end are zero-filled by the hardware) issued by a producer warp into a two-stage
//     ring of full / empty mbarriers: consumer warps never meet at a CTA barrier, they only wait for data;
//   * state loads are consecutive 64-bit words: conflict free, one request per (node | parent) column per 256 rows;
//   * the bin index of 4 rows is computed at once on packed bytes: idx = s_i + sum_k m_k * s_parent_k as ONE IMAD per
//     parent on a 32-bit word holding 4 rows (no byte ever exceeds 255 because the family has <= 256 bins);
//   * every lane has PRIVATE counters in its own bank: 8-bit counters, 4 bins per 32-bit word, word (bin >> 2) of lane l
//     at hist[(row0_i + (bin >> 2)) * 32 + l] -- the bank is the lane, so the read-modify-writes are conflict free by
//     construction and need no atomics (a node belongs to one warp).  8 bits suffice because a lane adds at most 8 per
//     tile: every 31 tiles the warp folds its counters into 32-bit per-CTA totals (a skewed, conflict-free transpose
//     read, SIMD byte sums), which go to the 64-bit global table once at the end;
//   * a warp walks its nodes two at a time with their read-modify-write chains interleaved (the 8 updates of one node
//     must stay in order -- two rows may hit the same counter -- but two nodes never share one);
//   * nodes are dealt to the warps of the CTA by the host (longest processing time first), tiles to the CTAs round-robin.
// Rows past the table end and 0xFF-marked rows (rejection sampler gave up) take a per-tile slow path that zeroes their
// states and gives them an increment of 0; states >= cardinality (the _dev entry trusts the table) are masked to the
// node's power-of-two counter span, so they are miscounted but never written outside the counter area + its guard.
// Tables whose base or pitch is not 16-byte aligned cannot be described by a tensor map: plain loads, one buffer.
constexpr uint32_t ROWS_TR = 256;        // rows per tile = 8 per lane
constexpr uint32_t ROWS_REC = 12;        // words per node record
constexpr uint32_t ROWS_FLUSH = 31;      // tiles between folds: 31 * 8 = 248 <= 255
constexpr uint32_t ROWS_GUARD = 32;      // counter word-rows behind the last node: where masked out-of-range bins may land
constexpr uint32_t ROWS_MAX_WARPS = 28;  // consumer warps (+ 1 producer warp): 928 threads, <= 70 registers

struct RowArgs {
  const uint8_t* data;
  uint64_t n_rows, pitch;
  const uint32_t* prog;      // [3 * n_warps header | node records (ROWS_REC words, + 2 per extra parent)] -> smem
  const uint32_t* row_out;   // per counter word-row: global bin of its first bin | (valid bins - 1) << 30
  uint32_t prog_words, n_nodes, n_warps, n_wrows;
  uint32_t box_nodes, n_boxes;   // TMA: n_boxes copies of box_nodes columns x 256 rows per tile
  unsigned long long* out;
};

__device__ __forceinline__ uint32_t has_ff_byte(uint32_t v) {  // != 0 iff some byte of v is 0xFF
  const uint32_t x = ~v;
  return (x - 0x01010101u) & ~x & 0x80808080u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

template <bool MORE, bool TMA>
__global__ void __launch_bounds__((ROWS_MAX_WARPS + 1) * 32, 1)
statistics_rows_kernel(const RowArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* prog = reinterpret_cast<uint32_t*>(smem_raw);
  const uint32_t prog_pad = (a.prog_words + 31u) & ~31u;
  uint32_t* hist = prog + prog_pad;                       // [n_wrows + guard][32] packed 8-bit counters, column = lane
  uint32_t* tot = hist + (a.n_wrows + ROWS_GUARD) * 32u;  // [n_wrows * 4] 32-bit totals of this CTA
  const uint32_t cols_pad = TMA ? a.box_nodes * a.n_boxes : a.n_nodes;
  const uint32_t tile_bytes = cols_pad * ROWS_TR;         // dense [column][256 rows]
  unsigned char* tiles = reinterpret_cast<unsigned char*>(tot + ((a.n_wrows * 4u + 31u) & ~31u));
  __shared__ __align__(8) uint64_t bars[4];               // full[0..1], empty[0..1]
  const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, warp = tid >> 5;
  for (uint32_t w = tid; w < a.prog_words; w += T) prog[w] = a.prog[w];
  for (uint32_t w = tid; w < (a.n_wrows + ROWS_GUARD) * 32u + a.n_wrows * 4u; w += T) hist[w] = 0u;  // counters, guard, totals
  const uint32_t full0 = s_addr(&bars[0]), empty0 = s_addr(&bars[2]);
  if (TMA && tid == 0) {
    for (uint32_t st = 0; st < 2u; ++st) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8u * st));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8u * st), "r"(a.n_warps));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t n_tiles = (a.n_rows + ROWS_TR - 1) / ROWS_TR;

  if (TMA && warp == a.n_warps) {
    // ---- producer warp: one lane keeps the two-stage ring full
    if (lane == 0) {
      uint32_t it = 0;
      for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t st = it & 1u;
        if (it >= 2u) mbar_wait(empty0 + 8u * st, ((it >> 1) - 1u) & 1u);  // every consumer warp released tile it - 2
        const uint32_t bar = full0 + 8u * st;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
        for (uint32_t bx = 0; bx < a.n_boxes; ++bx) {
          const uint32_t dst = s_addr(tiles + st * tile_bytes + bx * a.box_nodes * ROWS_TR);
          const uint32_t c0 = (uint32_t)(t * ROWS_TR), c1 = bx * a.box_nodes;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
              "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(bar)
              : "memory");
        }
      }
    }
    return;
  }
  if (warp >= a.n_warps) return;  // (never: the launch has exactly n_warps consumer warps + the producer)

  const uint32_t rec0 = prog[3 * warp], n_my = prog[3 * warp + 1], my_row0 = prog[3 * warp + 2] & 0xffffu,
                 my_rows = prog[3 * warp + 2] >> 16;
  // fold this warp's 8-bit lane-private counters into the CTA's 32-bit totals and clear them.  Lane l takes word-row
  // r0 + l and walks its 32 lane columns starting at column l (skewed): every step touches 32 different banks.
  auto fold = [&]() {
    for (uint32_t base = 0; base < my_rows; base += 32u) {
      const uint32_t r = base + lane;
      if (r < my_rows) {
        uint32_t* row = hist + (my_row0 + r) * 32u;
        uint32_t lo = 0u, hi = 0u;  // 16-bit fields: bins 0, 2 | bins 1, 3 (32 * 255 < 65536)
#pragma unroll 8
        for (uint32_t c = 0; c < 32u; ++c) {
          const uint32_t col = (c + lane) & 31u;
          const uint32_t x = row[col];
          row[col] = 0u;
          lo += x & 0x00FF00FFu;
          hi += (x >> 8) & 0x00FF00FFu;
        }
        uint32_t* t4 = tot + (my_row0 + r) * 4u;
        t4[0] += lo & 0xFFFFu; t4[1] += hi & 0xFFFFu; t4[2] += lo >> 16; t4[3] += hi >> 16;
      }
    }
    __syncwarp();
  };

  // one tile: my 8 rows of every column start at `tile`; SLOW = some row of the tile does not count
  auto process = [&](const unsigned char* tile, auto slow_tag, uint32_t ok_lo, uint32_t ok_hi) {
    constexpr bool SLOW = decltype(slow_tag)::value;
    const uint32_t keep_lo = ok_lo * 0xFFu, keep_hi = ok_hi * 0xFFu;  // 0x01 -> 0xFF per counting row
    auto ld = [&](uint32_t col_off) {
      uint2 v = *reinterpret_cast<const uint2*>(tile + col_off);
      if (SLOW) { v.x &= keep_lo; v.y &= keep_hi; }  // a 0xFF state would carry into its neighbours' bytes below
      return v;
    };
    const uint32_t* rec = prog + rec0;
    for (uint32_t k = 0; k < n_my; k += 2u) {
      // records: [col offset, first counter word-row, byte mask of the node's power-of-two bin span, n_par, (parent col
      // offset, multiplier) x 3, ...]; uniform across the warp (broadcast loads).  Two nodes per trip.
      const uint4 A0 = *reinterpret_cast<const uint4*>(rec), A1 = *reinterpret_cast<const uint4*>(rec + 4),
                  A2 = *reinterpret_cast<const uint4*>(rec + 8);
      const uint32_t a_words = MORE ? ((ROWS_REC + (A0.w > 3u ? 2u * (A0.w - 3u) : 0u) + 3u) & ~3u) : ROWS_REC;
      const uint32_t* recb = rec + a_words;
      const uint4 B0 = *reinterpret_cast<const uint4*>(recb), B1 = *reinterpret_cast<const uint4*>(recb + 4),
                  B2 = *reinterpret_cast<const uint4*>(recb + 8);
      const uint32_t b_words = MORE ? ((ROWS_REC + (B0.w > 3u ? 2u * (B0.w - 3u) : 0u) + 3u) & ~3u) : ROWS_REC;
      uint2 ia = ld(A0.x), ib = ld(B0.x);
      {
        const uint2 pa = ld(A1.x), pb = ld(B1.x);  // absent parents: multiplier 0 (column 0 is always readable)
        ia.x += A1.y * pa.x; ia.y += A1.y * pa.y;
        ib.x += B1.y * pb.x; ib.y += B1.y * pb.y;
      }
      if ((A0.w | B0.w) > 1u) {
        const uint2 pa1 = ld(A1.z), pa2 = ld(A2.x), pb1 = ld(B1.z), pb2 = ld(B2.x);
        ia.x += A1.w * pa1.x + A2.y * pa2.x; ia.y += A1.w * pa1.y + A2.y * pa2.y;
        ib.x += B1.w * pb1.x + B2.y * pb2.x; ib.y += B1.w * pb1.y + B2.y * pb2.y;
      }
      if (MORE) {
        for (uint32_t e = 3; e < A0.w; ++e) {
          const uint2 pp = ld(rec[ROWS_REC + 2 * (e - 3)]);
          const uint32_t m = rec[ROWS_REC + 2 * (e - 3) + 1];
          ia.x += m * pp.x; ia.y += m * pp.y;
        }
        for (uint32_t e = 3; e < B0.w; ++e) {
          const uint2 pp = ld(recb[ROWS_REC + 2 * (e - 3)]);
          const uint32_t m = recb[ROWS_REC + 2 * (e - 3) + 1];
          ib.x += m * pp.x; ib.y += m * pp.y;
        }
      }
      // byte offset of every row's counter inside my lane column: (bin >> 2) * 128 + (bin & 3), two rows per 32-bit word
      // (16-bit fields); bins are first masked to the node's power-of-two span
      const uint32_t al = ia.x & A0.z, ah = ia.y & A0.z, bl = ib.x & B0.z, bh = ib.y & B0.z;
#define BN_F_EVEN(x) ((((x) & 0x00FC00FCu) << 5) | ((x) & 0x00030003u))
#define BN_F_ODD(x) ((((x) >> 3) & 0x1F801F80u) | (((x) >> 8) & 0x00030003u))
      const uint32_t a02 = BN_F_EVEN(al), a13 = BN_F_ODD(al), a46 = BN_F_EVEN(ah), a57 = BN_F_ODD(ah);
      const uint32_t b02 = BN_F_EVEN(bl), b13 = BN_F_ODD(bl), b46 = BN_F_EVEN(bh), b57 = BN_F_ODD(bh);
#undef BN_F_EVEN
#undef BN_F_ODD
      unsigned char* ha = reinterpret_cast<unsigned char*>(hist) + A0.y * 128u + 4u * lane;
      unsigned char* hb = reinterpret_cast<unsigned char*>(hist) + B0.y * 128u + 4u * lane;
      // lane-private bank: conflict free, no atomics.  The updates of one node stay in program order (two rows may hit
      // the same counter); the two nodes' chains are interleaved (they never share a counter).
#define BN_ROW2(FA, FB, HALF, J)                                                              \
      {                                                                                       \
        unsigned char* ca = ha + ((HALF) ? ((FA) >> 16) : ((FA) & 0xFFFFu));                  \
        unsigned char* cb = hb + ((HALF) ? ((FB) >> 16) : ((FB) & 0xFFFFu));                  \
        const uint32_t inc = SLOW ? ((((J) < 4 ? ok_lo : ok_hi) >> (8u * ((J) & 3u))) & 1u) : 1u; \
        const uint32_t va = *ca, vb = *cb;                                                    \
        *ca = (unsigned char)(va + inc);                                                      \
        *cb = (unsigned char)(vb + inc);                                                      \
      }
      BN_ROW2(a02, b02, 0, 0) BN_ROW2(a13, b13, 0, 1) BN_ROW2(a02, b02, 1, 2) BN_ROW2(a13, b13, 1, 3)
      BN_ROW2(a46, b46, 0, 4) BN_ROW2(a57, b57, 0, 5) BN_ROW2(a46, b46, 1, 6) BN_ROW2(a57, b57, 1, 7)
#undef BN_ROW2
      rec = recb + b_words;
    }
  };
  // rows of this tile that count: all (fast path) unless the tile is ragged or carries 0xFF marks in node 0's column
  auto run_tile = [&](const unsigned char* tile, uint32_t rows) {
    const uint2 w0 = *reinterpret_cast<const uint2*>(tile);
    const bool slow = __any_sync(0xffffffffu, (has_ff_byte(w0.x) | has_ff_byte(w0.y)) != 0u) || rows < ROWS_TR;
    if (!slow) {
      process(tile, std::false_type{}, 0x01010101u, 0x01010101u);
    } else {
      uint32_t ok_lo = 0u, ok_hi = 0u;
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) {
        const uint32_t st0 = ((j < 4u ? w0.x : w0.y) >> (8u * (j & 3u))) & 0xffu;
        const uint32_t ok = (8u * lane + j < rows && st0 != 0xFFu) ? 1u : 0u;
        if (j < 4u) ok_lo |= ok << (8u * j); else ok_hi |= ok << (8u * (j - 4u));
      }
      process(tile, std::true_type{}, ok_lo, ok_hi);
    }
  };

  uint32_t it = 0, since_fold = 0;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const uint64_t r0 = t * ROWS_TR;
    const uint32_t rows = (uint32_t)(a.n_rows - r0 < ROWS_TR ? a.n_rows - r0 : ROWS_TR);
    if (TMA) {
      const uint32_t st = it & 1u;
      mbar_wait(full0 + 8u * st, (it >> 1) & 1u);  // the tile has landed
      run_tile(tiles + st * tile_bytes + 8u * lane, rows);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8u * st) : "memory");
    } else {
      __syncthreads();  // everybody is done with the (single) buffer
      for (uint32_t i = 0; i < a.n_nodes; ++i)
        for (uint32_t r = tid; r < ROWS_TR; r += T)
          tiles[i * ROWS_TR + r] = r < rows ? a.data[(uint64_t)i * a.pitch + r0 + r] : (unsigned char)0;
      __syncthreads();
      run_tile(tiles + 8u * lane, rows);
    }
    if (++since_fold == ROWS_FLUSH) {
      fold();
      since_fold = 0;
    }
  }
  fold();
  // totals -> global table: each warp flushes the counter word-rows it owns
  for (uint32_t r = my_row0 + lane; r < my_row0 + my_rows; r += 32u) {
    const uint32_t ro = a.row_out[r], g = ro & 0x3FFFFFFFu, nv = (ro >> 30) + 1u;
    for (uint32_t f = 0; f < nv; ++f) {
      const uint32_t c = tot[r * 4u + f];
      if (c) atomicAdd(&a.out[g + f], (unsigned long long)c);
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// The plan of the row-lane kernel depends on the network structure only (and on whether tiles arrive by TMA): node
// groups that fit shared memory, the warp count of each, the node records and the counter-row map.  Built once per
// net, kept with a device copy of the records (read-only, so concurrent calls share it); planning per call cost
// ~0.1 ms of host time on a 0.8 ms kernel.
struct StatsGroup {
  int first = 0, last = 0, W = 0;
  uint32_t n_wrows = 0, prog_words = 0;
  std::vector<uint32_t> words;   // [prog | row_out]
  DevBuf<uint32_t> dev;
};
struct StatsPlan {
  bool ok = false, fatal = false;
  std::vector<StatsGroup> groups;
};
void stats_plan_free(StatsPlan* p) { delete p; }

static void build_stats_plan(Net* net, bool more, size_t budget, size_t tiles_b, StatsPlan* plan) {
  const int n = net->n;
  auto wrows_of = [&](int i) { return (uint32_t)((net->cpt_off[i + 1] - net->cpt_off[i] + 3) / 4); };
  auto rec_words_of = [&](int i) {
    const uint32_t np = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
    const uint32_t w = ROWS_REC + (np > 3 ? 2 * (np - 3) : 0);
    return more ? ((w + 3u) & ~3u) : ROWS_REC;
  };
  int first = 0;
  while (first < n) {
    size_t used = tiles_b + (ROWS_MAX_WARPS * 3 + 4) * 4 + 512 + ROWS_GUARD * 128 + 2 * ROWS_REC * 4;
    uint32_t wrows = 0;
    int last = first;
    while (last < n) {
      const size_t need = (size_t)wrows_of(last) * 144 + (size_t)rec_words_of(last) * 4 + 2 * ROWS_REC * 4;
      if (used + need > budget || wrows + wrows_of(last) + ROWS_GUARD > 65535u) break;
      used += need;
      wrows += wrows_of(last);
      ++last;
    }
    if (last == first) { plan->ok = false; plan->fatal = first != 0; return; }  // a single node does not fit beside the tiles
    // deal the group's nodes to warps, longest processing time first, and pick the warp count by a cost model fitted to
    // measurements on the C2 net (2^24 rows; W = 12 / 16 / 17 / 20 / 24 / 25 / 26 / 28 consumer warps: 0.905 / 0.877 /
    // 0.837 / 0.866 / 0.840 / 0.783 / 0.790 / 0.810 ms): time ~ longest warp + (work of all warps, dummy partners of odd
    // nodes and per-tile overheads included) / 12.  The longest warp dominates (every warp waits for it at the ring),
    // so counts that deal the nodes out evenly in PAIRS win: 100 nodes -> 25 warps of two pairs each.
    const int m = last - first;
    std::vector<int> order(m);
    for (int j = 0; j < m; ++j) order[j] = first + j;
    auto cost = [&](int i) { return 34 + 4 * std::min(3, net->par_ptr[i + 1] - net->par_ptr[i] > 1 ? 3 : 1); };
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost(x) > cost(y); });
    const int warps_env = getenv("BN_STATS_WARPS") ? atoi(getenv("BN_STATS_WARPS")) : 0;
    auto deal = [&](int W, std::vector<std::vector<int>>* out) {  // -> modelled time (nodes are walked in pairs)
      std::vector<int> load(W, 0);
      std::vector<std::vector<int>> mine(W);
      for (int i : order) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[w] += cost(i);
        mine[w].push_back(i);
      }
      int longest = 0, total = 0;
      for (int w = 0; w < W; ++w) {
        int l = 40;  // per-tile overhead of a warp
        for (size_t j = 0; j < mine[w].size(); j += 2) l += 2 * cost(mine[w][j]);  // an odd node drags a dummy partner
        longest = std::max(longest, l);
        total += l;
      }
      if (out) out->swap(mine);
      return (double)longest + (double)total / 12.0;
    };
    int best_w = 1;
    double best_t = 1e300;
    for (int W = std::min<int>(ROWS_MAX_WARPS, m); W >= 1; --W) {
      if (warps_env && W != std::min(warps_env, std::min<int>(ROWS_MAX_WARPS, m))) continue;
      const double tm = deal(W, nullptr);
      if (tm < best_t) { best_t = tm; best_w = W; }
    }
    const int W = best_w;
    if (getenv("BN_STATS_DEBUG")) fprintf(stderr, "[bn stats] nodes %d..%d: %d consumer warps (model %.1f)\n", first, last - 1, W, best_t);
    std::vector<std::vector<int>> mine;
    deal(W, &mine);
    uint32_t total_rows = 0;
    for (int i = first; i < last; ++i) total_rows += wrows_of(i);
    std::vector<uint32_t> prog((3 * (size_t)W + 3) / 4 * 4, 0), row_out;  // header, padded: records are read as 128-bit words
    uint32_t row = 0;
    for (int w = 0; w < W; ++w) {
      prog[3 * w] = (uint32_t)prog.size();
      const uint32_t row_begin = row;
      // pair nodes with the same shape (<= 1 parent | more): a pair takes the longer path together
      std::stable_sort(mine[w].begin(), mine[w].end(), [&](int x, int y) {
        return (net->par_ptr[x + 1] - net->par_ptr[x] > 1) > (net->par_ptr[y + 1] - net->par_ptr[y] > 1);
      });
      for (int i : mine[w]) {
        const uint32_t np = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
        const size_t at = prog.size();
        prog.resize(at + rec_words_of(i), 0);
        prog[at + 0] = (uint32_t)i * ROWS_TR;
        prog[at + 1] = row;
        {
          uint32_t span = 4;  // power of two >= the node's counter bytes (4 per word-row)
          while (span < 4 * wrows_of(i)) span <<= 1;
          prog[at + 2] = (span - 1) * 0x01010101u;
        }
        prog[at + 3] = np;
        int64_t stride = net->card[i];
        for (uint32_t e = 0; e < np; ++e) {
          const uint32_t col = (uint32_t)net->par_idx[net->par_ptr[i] + e] * ROWS_TR;
          const size_t slot = e < 3 ? at + 4 + 2 * e : at + ROWS_REC + 2 * (e - 3);
          prog[slot] = col;
          prog[slot + 1] = (uint32_t)stride;  // < 256: the family has <= 256 bins
          stride *= net->card[net->par_idx[net->par_ptr[i] + e]];
        }
        const uint32_t B = (uint32_t)(net->cpt_off[i + 1] - net->cpt_off[i]);
        for (uint32_t r = 0; r < wrows_of(i); ++r)
          row_out.push_back((uint32_t)(net->cpt_off[i] + 4 * r) | ((std::min(4u, B - 4 * r) - 1u) << 30));
        row += wrows_of(i);
      }
      uint32_t n_rec = (uint32_t)mine[w].size();
      if (n_rec % 2) {  // odd: a dummy partner that counts into the guard rows (column 0, mask 0, no parents)
        const size_t at = prog.size();
        prog.resize(at + ROWS_REC, 0);
        prog[at + 1] = total_rows;
        ++n_rec;
      }
      prog[3 * w + 1] = n_rec;
      prog[3 * w + 2] = row_begin | ((row - row_begin) << 16);
    }
    while (prog.size() % 4) prog.push_back(0);
    plan->groups.emplace_back();
    StatsGroup& g = plan->groups.back();
    g.first = first; g.last = last; g.W = W; g.n_wrows = row;
    g.prog_words = (uint32_t)prog.size();
    g.words = prog;
    g.words.insert(g.words.end(), row_out.begin(), row_out.end());
    first = last;
  }
  plan->ok = true;
}

// Plan + launch of the row-lane kernel, one launch per group of consecutive nodes whose counters fit shared memory
// together; returns BN_OK with *handled = false when the network does not qualify (the node-lane kernel takes over).
static int32_t statistics_launch_rows(Net* net, const uint8_t* data_dev, uint64_t n_rows, uint64_t pitch,
                                      unsigned long long* out_dev, bool* handled) {
  *handled = false;
  const DevProps& dp = dev_props(net->device);
  cudaStream_t s = tl_stream;
  const int n = net->n;
  const size_t budget = (size_t)dp.max_smem_optin - 1024;
  if (net->cpt_off[n] >= (int64_t)(1u << 30) || n > 4096) return BN_OK;
  for (int i = 0; i < n; ++i)
    if (net->cpt_off[i + 1] - net->cpt_off[i] > 256) return BN_OK;  // a bin index must fit one byte
  if (getenv("BN_STATS_NODELANES")) return BN_OK;
  bool more = false;
  for (int i = 0; i < n; ++i) more = more || (net->par_ptr[i + 1] - net->par_ptr[i] > 3);
  // TMA needs a 16-byte aligned table and pitch, and 32-bit coordinates
  const bool tma = pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(data_dev) % 16) == 0 && n_rows < (1ull << 32) &&
                   pitch < (1ull << 40) && encode_tiled() != nullptr && !getenv("BN_STATS_NO_TMA");
  const uint32_t n_boxes = (uint32_t)((n + 255) / 256), box_nodes = (uint32_t)((n + n_boxes - 1) / n_boxes);
  const size_t tile_b = (size_t)(tma ? box_nodes * n_boxes : (uint32_t)n) * ROWS_TR;
  const size_t tiles_b = tile_b * (tma ? 2 : 1);
  const uint64_t n_tiles = (n_rows + ROWS_TR - 1) / ROWS_TR;
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)dp.sm_count);
  BN_REQUIRE(n_rows / (uint64_t)grid < (1ull << 31), BN_UNSUPPORTED, "too many rows per CTA for the 32-bit CTA totals");
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma) {
    const cuuint64_t gdim[2] = {(cuuint64_t)n_rows, (cuuint64_t)n};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {ROWS_TR, box_nodes};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(data_dev), gdim, gstride, box,
                                      estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BN_REQUIRE(r == CUDA_SUCCESS, BN_CUDA_ERR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  StatsPlan* plan = nullptr;
  {
    std::lock_guard<std::mutex> lk(net->mu);
    StatsPlan*& slot = net->stats_plan[tma ? 1 : 0];
    if (!slot || getenv("BN_STATS_WARPS")) {  // (the sweep knob rebuilds the plan every call)
      if (slot) { stats_plan_free(slot); slot = nullptr; }
      struct Holder {  // the net gets the plan only when its device copies exist
        StatsPlan* p = new StatsPlan();
        ~Holder() { if (p) stats_plan_free(p); }
      } hold;
      build_stats_plan(net, more, budget, tiles_b, hold.p);
      for (StatsGroup& g : hold.p->groups) {
        BN_CUDA_TRY(g.dev.upload(g.words.data(), g.words.size(), s));
        BN_CUDA_TRY(cudaStreamSynchronize(s));  // the host vector is pageable; once per net
      }
      slot = hold.p;
      hold.p = nullptr;
    }
    plan = slot;
  }
  if (!plan->ok) return plan->fatal ? BN_UNSUPPORTED : BN_OK;
  for (StatsGroup& g : plan->groups) {
    RowArgs a{};
    a.data = data_dev;
    a.n_rows = n_rows;
    a.pitch = pitch;
    a.prog_words = g.prog_words;
    a.n_nodes = (uint32_t)n;
    a.n_warps = (uint32_t)g.W;
    a.n_wrows = g.n_wrows;
    a.box_nodes = box_nodes;
    a.n_boxes = n_boxes;
    a.out = out_dev;
    a.prog = g.dev.p;
    a.row_out = g.dev.p + g.prog_words;
    const size_t smem = (((size_t)a.prog_words + 31) & ~(size_t)31) * 4 + ((size_t)g.n_wrows + ROWS_GUARD) * 128 +
                        (((size_t)g.n_wrows * 4 + 31) & ~(size_t)31) * 4 + tiles_b;
    void (*kern)(const RowArgs, const CUtensorMap) =
        tma ? (more ? statistics_rows_kernel<true, true> : statistics_rows_kernel<false, true>)
            : (more ? statistics_rows_kernel<true, false> : statistics_rows_kernel<false, false>);
    BN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin - 1024));
    BN_REQUIRE(smem <= budget, BN_CUDA_ERR, "statistics plan exceeds shared memory (%zu bytes)", smem);
    kern<<<grid, (g.W + (tma ? 1 : 0)) * 32, smem, s>>>(a, tmap);
    count_launch();
    BN_CUDA_TRY(cudaGetLastError());
  }
  *handled = true;
  return BN_OK;
}

static int32_t statistics_launch(Net* net, const uint8_t* data_dev, uint64_t n_rows, uint64_t pitch,
                                 unsigned long long* out_dev) {
  const int n = net->n;
  const int64_t n_bins = net->cpt_off[n];
  BN_REQUIRE(n_bins < (int64_t)0x7fffffff, BN_UNSUPPORTED, "count table too large (%lld entries)", (long long)n_bins);
  BN_REQUIRE(pitch >= n_rows, BN_BAD_ARG, "pitch (%llu) smaller than n_rows (%llu)", (unsigned long long)pitch,
             (unsigned long long)n_rows);
  BN_CUDA_TRY(cudaMemsetAsync(out_dev, 0, (size_t)n_bins * sizeof(unsigned long long), tl_stream));
  if (n_rows == 0 || n == 0) return BN_OK;
  bool handled = false;
  BN_TRY(statistics_launch_rows(net, data_dev, n_rows, pitch, out_dev, &handled));
  if (handled) return BN_OK;
  return statistics_launch_nodelanes(net, data_dev, n_rows, pitch, out_dev);
}


// logpdf(bn, df) (src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-99 over src/bayes_nets.jl:322-328):
// the reference adds log P(x_i | pa_i) row by row, node by node.  Grouping equal terms, the same sum is the dot product
// of the table's sufficient statistics with the log CPTs, sum_b N[b] * log(cpt[b]) over the bins with N[b] > 0 -- so
// the n_rows x n_nodes part is the counting kernel above and what remains is a few thousand terms: one CTA, one term
// per thread per trip, folded in a fixed order (deterministic; no fp atomics).  A bin with N > 0 and cpt = 0 gives
// -Inf, as log(0) does in the reference's sum; bins never seen contribute nothing (the reference never evaluates them).
__global__ void __launch_bounds__(1024, 1) logl_kernel(const unsigned long long* __restrict__ counts,
                                                        const double* __restrict__ cpt, uint32_t n_bins,
                                                        double* __restrict__ out) {
  __shared__ double part[1024];
  double acc = 0.0;
  for (uint32_t b = threadIdx.x; b < n_bins; b += 1024) {
    const unsigned long long c = counts[b];
    if (c) acc += (double)c * log(cpt[b]);
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = part[0];
}

static int32_t logpdf_launch(Net* net, const uint8_t* data_dev, uint64_t n_rows, uint64_t pitch, double* out_dev) {
  const size_t n_bins = (size_t)net->cpt_off[net->n];
  TmpBuf<unsigned long long> d_counts;
  BN_CUDA_TRY(d_counts.alloc(n_bins + 1));
  BN_TRY(statistics_launch(net, data_dev, n_rows, pitch, d_counts.p));
  logl_kernel<<<1, 1024, 0, tl_stream>>>(d_counts.p, net->d_cpt.p, (uint32_t)n_bins, out_dev);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

// host table -> pitched device copy, states validated on the way (a state >= card would be counted out of bounds)
static int32_t upload_table(Net* net, const uint8_t* states, uint64_t n_rows, TmpBuf<uint8_t>& d_states, uint64_t* pitch) {
  for (int i = 0; i < net->n; ++i) {
    const uint8_t* col = states + (size_t)i * n_rows;
    const uint8_t c = (uint8_t)net->card[i];
    for (uint64_t r = 0; r < n_rows; ++r)
      BN_REQUIRE(col[r] < c, BN_BOUNDS, "row %llu: state %d of node %d outside 0:%d", (unsigned long long)r, (int)col[r], i,
                 (int)c - 1);
  }
  *pitch = (n_rows + 15) / 16 * 16;
  BN_CUDA_TRY(d_states.alloc((size_t)net->n * *pitch + 16));
  if (n_rows)
    BN_CUDA_TRY(cudaMemcpy2DAsync(d_states.p, *pitch, states, n_rows, n_rows, (size_t)net->n, cudaMemcpyHostToDevice, tl_stream));
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

int32_t bn_statistics_dev(bn_net_t h, const uint8_t* states_dev, uint64_t n_rows, uint64_t pitch, int64_t* out_counts_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net && out_counts_dev && (states_dev || n_rows == 0), BN_BAD_ARG, "bn_statistics_dev: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  return statistics_launch(net, states_dev, n_rows, pitch, reinterpret_cast<unsigned long long*>(out_counts_dev));
}

int32_t bn_statistics(bn_net_t h, const uint8_t* states, uint64_t n_rows, int64_t* out_counts) {
  Net* net = get_net(h);
  BN_REQUIRE(net && out_counts && (states || n_rows == 0), BN_BAD_ARG, "bn_statistics: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  const size_t n_bins = (size_t)net->cpt_off[net->n];
  uint64_t pitch = 0;
  TmpBuf<uint8_t> d_states;
  TmpBuf<unsigned long long> d_out;
  BN_TRY(upload_table(net, states, n_rows, d_states, &pitch));
  BN_CUDA_TRY(d_out.alloc(n_bins + 1));
  BN_TRY(statistics_launch(net, d_states.p, n_rows, pitch, d_out.p));
  BN_CUDA_TRY(cudaMemcpyAsync(out_counts, d_out.p, n_bins * sizeof(int64_t), cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

int32_t bn_logpdf_dev(bn_net_t h, const uint8_t* states_dev, uint64_t n_rows, uint64_t pitch, double* out_logl_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net && out_logl_dev && (states_dev || n_rows == 0), BN_BAD_ARG, "bn_logpdf_dev: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  return logpdf_launch(net, states_dev, n_rows, pitch, out_logl_dev);
}

int32_t bn_logpdf_counts_dev(bn_net_t h, const int64_t* counts_dev, double* out_logl_dev) {
  Net* net = get_net(h);
  BN_REQUIRE(net && counts_dev && out_logl_dev, BN_BAD_ARG, "bn_logpdf_counts_dev: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  logl_kernel<<<1, 1024, 0, tl_stream>>>(reinterpret_cast<const unsigned long long*>(counts_dev), net->d_cpt.p,
                                         (uint32_t)net->cpt_off[net->n], out_logl_dev);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
  return BN_OK;
}

int32_t bn_logpdf(bn_net_t h, const uint8_t* states, uint64_t n_rows, double* out_logl) {
  Net* net = get_net(h);
  BN_REQUIRE(net && out_logl && (states || n_rows == 0), BN_BAD_ARG, "bn_logpdf: bad arguments");
  BN_CUDA_TRY(cudaSetDevice(net->device));
  uint64_t pitch = 0;
  TmpBuf<uint8_t> d_states;
  TmpBuf<double> d_out;
  BN_TRY(upload_table(net, states, n_rows, d_states, &pitch));
  BN_CUDA_TRY(d_out.alloc(1));
  BN_TRY(logpdf_launch(net, d_states.p, n_rows, pitch, d_out.p));
  BN_CUDA_TRY(cudaMemcpyAsync(out_logl, d_out.p, sizeof(double), cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

}  // extern "C"
