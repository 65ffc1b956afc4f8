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
// Design (B200): persistent CTAs, one table row per thread.  A tile of TR rows x all nodes is staged into shared
// memory by TMA bulk copies (one per node column: TR contiguous bytes), double-buffered on two mbarriers so the
// next tile streams in while the current one is counted.  Each thread walks a small per-node program (offset,
// parents, strides -- warp-uniform shared-memory broadcasts), reads its row's states as u8 from the tile and bumps
// a CTA-private u32 histogram with shared-memory atomics; the histogram is flushed once per CTA with 64-bit global
// atomics.  Algorithmic traffic: n_nodes bytes per row, read once (HBM-read-bound in principle; in practice the
// per-(row, node) instruction count and same-address atomic conflicts bound it).
#include "common.cuh"

namespace bn {

struct StatsArgs {
  const uint8_t* data;   // [n_nodes][pitch]
  uint64_t n_rows, pitch;
  const uint32_t* prog;  // device copy: per node  [cpt_off, card | n_par << 8, (parent, stride * card) ...]
  uint32_t prog_words, n_nodes, n_bins, TR;
  uint32_t hist_in_smem, use_tma;
  unsigned long long* out;  // [n_bins]
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024, 1) statistics_kernel(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2];
  // layout: [prog | hist u32 (n_bins, when it fits) | tile 0 | tile 1], tiles are [n_nodes][TR] u8
  uint32_t* prog = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* hist = prog + a.prog_words;
  const uint32_t hist_words = a.hist_in_smem ? a.n_bins : 0u;
  const uint32_t tile_bytes = a.n_nodes * a.TR;
  unsigned char* tiles = smem_raw + (((a.prog_words + hist_words) * 4u + 127u) & ~127u);
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  for (uint32_t w = tid; w < a.prog_words; w += T) prog[w] = a.prog[w];
  for (uint32_t b = tid; b < hist_words; b += T) hist[b] = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(&full_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(&full_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t n_tiles = (a.n_rows + a.TR - 1) / a.TR;
  // stage tile `t` into buffer `buf`: full tiles by TMA (one bulk copy per node column, issued by warp 0);
  // the ragged last tile and unaligned tables by plain loads of all threads (then a barrier instead of the mbarrier)
  auto tile_is_tma = [&](uint64_t t) { return a.use_tma && (t + 1) * a.TR <= a.n_rows; };
  auto stage = [&](uint64_t t, uint32_t buf) {
    unsigned char* dst = tiles + buf * tile_bytes;
    const uint64_t r0 = t * a.TR;
    if (tile_is_tma(t)) {
      if (tid < 32) {
        const uint32_t bar = s_addr(&full_bar[buf]);
        if (tid == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
        __syncwarp();
        for (uint32_t i = tid; i < a.n_nodes; i += 32)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s_addr(dst + i * a.TR)),
                       "l"(a.data + (uint64_t)i * a.pitch + r0), "r"(a.TR), "r"(bar)
                       : "memory");
      }
    } else {
      const uint64_t rows = a.n_rows - r0 < a.TR ? a.n_rows - r0 : a.TR;
      for (uint32_t i = 0; i < a.n_nodes; ++i)
        for (uint32_t r = tid; r < rows; r += T) dst[i * a.TR + r] = a.data[(uint64_t)i * a.pitch + r0 + r];
    }
  };
  uint32_t phase = 0u;  // bit b = parity the next wait on buffer b expects
  uint64_t t = blockIdx.x;
  if (t < n_tiles) stage(t, 0);
  uint32_t buf = 0;
  for (; t < n_tiles; t += gridDim.x, buf ^= 1u) {
    const uint64_t tn = t + gridDim.x;
    if (tn < n_tiles) stage(tn, buf ^ 1u);  // the other buffer was released by the barrier that ended the last tile
    if (tile_is_tma(t)) {
      const uint32_t bar = s_addr(&full_bar[buf]);
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"((phase >> buf) & 1u) : "memory");
      phase ^= 1u << buf;
    } else {
      __syncthreads();
    }
    const unsigned char* tile = tiles + buf * tile_bytes;
    const uint64_t r0 = t * a.TR;
    const uint64_t rows = a.n_rows - r0 < a.TR ? a.n_rows - r0 : a.TR;
    for (uint32_t r = tid; r < rows; r += T) {
      const unsigned char* st = tile + r;
      const uint32_t* pc = prog;
      if (st[0] == 0xFFu) continue;  // row the rejection sampler gave up on (bn_sample: out_states = 0xFF)
      for (uint32_t i = 0; i < a.n_nodes; ++i) {
        uint32_t idx = pc[0] + st[i * a.TR];
        const uint32_t npar = pc[1] >> 8;
        for (uint32_t k = 0; k < npar; ++k) idx += pc[3 + 2 * k] * (uint32_t)st[pc[2 + 2 * k] * a.TR];
        pc += 2 + 2 * npar;
        if (idx < a.n_bins) {  // memory safety only: a state >= card[i] is the caller's error (see bn_cuda.h)
          if (a.hist_in_smem) atomicAdd(&hist[idx], 1u);
          else atomicAdd(&a.out[idx], 1ull);
        }
      }
    }
    __syncthreads();  // everyone is done with this buffer (and, for plain staging, sees the next one complete)
  }
  if (a.hist_in_smem)
    for (uint32_t b = tid; b < a.n_bins; b += T) {
      const uint32_t c = hist[b];
      if (c) atomicAdd(&a.out[b], (unsigned long long)c);
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
  // per-node program
  std::vector<uint32_t> prog;
  prog.reserve((size_t)n * 6);
  for (int i = 0; i < n; ++i) {
    const uint32_t npar = (uint32_t)(net->par_ptr[i + 1] - net->par_ptr[i]);
    prog.push_back((uint32_t)net->cpt_off[i]);
    prog.push_back((uint32_t)net->card[i] | (npar << 8));
    int64_t stride = net->card[i];
    for (int k = net->par_ptr[i]; k < net->par_ptr[i + 1]; ++k) {
      prog.push_back((uint32_t)net->par_idx[k]);
      prog.push_back((uint32_t)stride);
      stride *= net->card[net->par_idx[k]];
    }
  }
  while (prog.size() % 4) prog.push_back(0);
  // shape: rows per tile = threads per CTA; two tiles + program + (histogram when it fits) within 227 KB
  StatsArgs a{};
  const size_t prog_b = prog.size() * 4;
  const size_t budget = (size_t)dp.max_smem_optin - 2048;
  a.hist_in_smem = ((size_t)n_bins * 4 + prog_b + 2 * (size_t)n * 64 <= budget) ? 1u : 0u;
  const size_t fixed = ((prog_b + (a.hist_in_smem ? (size_t)n_bins * 4 : 0) + 127) / 128) * 128;
  BN_REQUIRE(fixed + 2 * (size_t)n * 32 <= budget, BN_UNSUPPORTED, "network too large for the statistics kernel (%d nodes)", n);
  uint32_t TR = 1024;
  while (TR > 32 && fixed + 2 * (size_t)n * TR > budget / 2) TR /= 2;  // aim for two CTAs per SM
  while (TR > 32 && fixed + 2 * (size_t)n * TR > budget) TR /= 2;
  a.TR = TR;
  a.use_tma = (pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(data_dev) % 16) == 0 && TR % 16 == 0) ? 1u : 0u;
  const size_t smem = fixed + 2 * (size_t)n * TR;
  // program -> device scratch (reused across calls)
  if (net->scratch.n < prog_b + 16) BN_CUDA_TRY(net->scratch.alloc((prog_b + 16) * 2));
  BN_CUDA_TRY(cudaMemcpyAsync(net->scratch.p, prog.data(), prog_b, cudaMemcpyHostToDevice, s));
  BN_CUDA_TRY(cudaStreamSynchronize(s));  // `prog` is pageable host memory about to go out of scope
  a.data = data_dev;
  a.n_rows = n_rows;
  a.pitch = pitch;
  a.prog = reinterpret_cast<const uint32_t*>(net->scratch.p);
  a.prog_words = (uint32_t)prog.size();
  a.n_nodes = (uint32_t)n;
  a.n_bins = (uint32_t)n_bins;
  a.out = out_dev;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    BN_CUDA_TRY(cudaFuncSetAttribute(statistics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin - 1024));
    attr_set = true;
  }
  int per_sm = (int)((size_t)dp.smem_per_sm / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm * (int)TR > 2048) per_sm = 2048 / (int)TR;
  const uint64_t n_tiles = (n_rows + TR - 1) / TR;
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)dp.sm_count * per_sm);
  BN_REQUIRE(n_rows / (uint64_t)grid < (1ull << 31), BN_UNSUPPORTED, "too many rows per CTA for the 32-bit CTA histogram");
  statistics_kernel<<<grid, TR, smem, s>>>(a);
  count_launch();
  BN_CUDA_TRY(cudaGetLastError());
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
  // states must be valid (< card) or the kernel would count out of bounds: check on the host while copying
  for (int i = 0; i < net->n; ++i) {
    const uint8_t* col = states + (size_t)i * n_rows;
    const uint8_t c = (uint8_t)net->card[i];
    for (uint64_t r = 0; r < n_rows; ++r)
      BN_REQUIRE(col[r] < c, BN_BOUNDS, "row %llu: state %d of node %d outside 0:%d", (unsigned long long)r, (int)col[r], i,
                 (int)c - 1);
  }
  const uint64_t pitch = (n_rows + 15) / 16 * 16;
  DevBuf<uint8_t> d_states;
  DevBuf<unsigned long long> d_out;
  BN_CUDA_TRY(d_states.alloc((size_t)net->n * pitch + 16));
  BN_CUDA_TRY(d_out.alloc(n_bins + 1));
  if (n_rows)
    BN_CUDA_TRY(cudaMemcpy2DAsync(d_states.p, pitch, states, n_rows, n_rows, (size_t)net->n, cudaMemcpyHostToDevice, tl_stream));
  BN_TRY(statistics_launch(net, d_states.p, n_rows, pitch, d_out.p));
  BN_CUDA_TRY(cudaMemcpyAsync(out_counts, d_out.p, n_bins * sizeof(int64_t), cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

}  // extern "C"
