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

static int32_t statistics_launch(Net* net, const uint8_t* data_dev, uint64_t n_rows, uint64_t pitch,
                                 unsigned long long* out_dev) {
  const DevProps& dp = dev_props(net->device);
  cudaStream_t s = tl_stream;
  const int n = net->n;
  const int64_t n_bins = net->cpt_off[n];
  BN_REQUIRE(n_bins < (int64_t)0x7fffffff, BN_UNSUPPORTED, "count table too large (%lld entries)", (long long)n_bins);
  BN_REQUIRE(pitch >= n_rows, BN_BAD_ARG, "pitch (%llu) smaller than n_rows (%llu)", (unsigned long long)pitch,
             (unsigned long long)n_rows);
  BN_CUDA_TRY(cudaMemsetAsync(out_dev, 0, (size_t)n_bins * sizeof(unsigned long long), s));
  if (n_rows == 0 || n == 0) return BN_OK;
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
