// Where do the microseconds between two factor ops go?  Times the same sum / fused-eliminate kernels (a) through the C
// ABI (pool allocation + free per op), (b) launched directly on preallocated outputs, (c) a plain device copy of the
// same bytes.  Links libbn_cuda.so (internal symbols of namespace bn are visible).
//   nvcc -O2 -std=c++17 -arch=sm_100a -o tools/probe/alloc_gap tools/probe/alloc_gap.cu -Iinclude -Lbayesnets.jl_b200 -lbn_cuda
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bn_cuda.h"

namespace bn {
struct StreamOperand { const double* ptr; const std::vector<int32_t>* dims; const std::vector<int32_t>* card; };
int32_t sum_segmented(const double* in, const std::vector<int32_t>& card, const std::vector<char>& summed, double* out, bool* handled);
int32_t stream_eliminate(const std::vector<StreamOperand>& ops, const std::vector<int32_t>& d, const std::vector<int32_t>& c,
                         const std::vector<char>& summed, double* out, bool* handled);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define BK(x) do { int32_t rc_ = (x); if (rc_ != 0) { printf("bn rc %d line %d: %s\n", rc_, __LINE__, bn_last_error()); exit(1); } } while (0)

int main() {
  const int k = 24, reps = 24, nrot = 3;
  const size_t N = 1ull << k;
  std::vector<int32_t> dims(k), card(k, 2);
  for (int i = 0; i < k; ++i) dims[i] = 1000 + i;
  std::vector<double> host(N);
  for (size_t i = 0; i < N; ++i) host[i] = (double)((i * 2654435761u) % 1000) / 1000.0 + 0.001;
  bn_factor_t A[nrot];
  const double* Ap[nrot];
  for (int r = 0; r < nrot; ++r) {
    BK(bn_factor_upload(k, dims.data(), card.data(), host.data(), 0, &A[r]));
    void* p; BK(bn_factor_dev_ptr(A[r], &p)); Ap[r] = (const double*)p;
  }
  double* outp[nrot];
  for (int r = 0; r < nrot; ++r) CK(cudaMalloc(&outp[r], N * 8));
  char* flush; CK(cudaMalloc(&flush, 512u << 20));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int sumdim_pos[3] = {0, k - 1, k / 2};
  for (int c = 0; c < 3; ++c) {
    const int32_t sd = dims[sumdim_pos[c]];
    std::vector<char> flag(k, 0); flag[sumdim_pos[c]] = 1;
    const double bytes = 8.0 * (N + N / 2);
    float ms_abi, ms_direct, ms_hold;
    // (a) C ABI, freeing results two ops later (what tools/factor_bench.py does)
    for (int pass = 0; pass < 2; ++pass) {
      bn_factor_t keep[2] = {0, 0};
      CK(cudaMemsetAsync(flush, 0, 512u << 20, 0));
      CK(cudaEventRecord(e0, 0));
      for (int it = 0; it < reps; ++it) {
        bn_factor_t o; BK(bn_factor_sum(A[it % nrot], &sd, 1, &o));
        if (keep[it % 2]) bn_factor_free(keep[it % 2]);
        keep[it % 2] = o;
      }
      CK(cudaEventRecord(e1, 0)); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms_abi, e0, e1));
      for (int i = 0; i < 2; ++i) if (keep[i]) bn_factor_free(keep[i]);
    }
    // (a2) C ABI, nothing freed inside the timed region
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<bn_factor_t> all;
      CK(cudaMemsetAsync(flush, 0, 512u << 20, 0));
      CK(cudaEventRecord(e0, 0));
      for (int it = 0; it < reps; ++it) { bn_factor_t o; BK(bn_factor_sum(A[it % nrot], &sd, 1, &o)); all.push_back(o); }
      CK(cudaEventRecord(e1, 0)); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms_hold, e0, e1));
      for (bn_factor_t h : all) bn_factor_free(h);
    }
    // (b) the same kernel on preallocated outputs
    for (int pass = 0; pass < 2; ++pass) {
      CK(cudaMemsetAsync(flush, 0, 512u << 20, 0));
      CK(cudaEventRecord(e0, 0));
      for (int it = 0; it < reps; ++it) { bool h = false; BK(bn::sum_segmented(Ap[it % nrot], card, flag, outp[it % nrot], &h)); if (!h) { printf("not handled\n"); return 1; } }
      CK(cudaEventRecord(e1, 0)); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms_direct, e0, e1));
    }
    printf("sum dim %2d: C ABI alloc+free %.4f ms/op (%.0f GB/s) | C ABI alloc only %.4f (%.0f) | preallocated %.4f (%.0f)\n", sumdim_pos[c],
           ms_abi / reps, bytes / (ms_abi / reps) / 1e6, ms_hold / reps, bytes / (ms_hold / reps) / 1e6, ms_direct / reps, bytes / (ms_direct / reps) / 1e6);
  }
  // (c) device copy of N/2 + N/2... same bytes as the sum: read N, write N/2 is not a copy; use copy of 0.75 N doubles
  float ms;
  for (int pass = 0; pass < 2; ++pass) {
    CK(cudaMemsetAsync(flush, 0, 512u << 20, 0));
    CK(cudaEventRecord(e0, 0));
    for (int it = 0; it < reps; ++it) CK(cudaMemcpyAsync(outp[it % nrot], Ap[(it + 1) % nrot], (N * 3 / 4) * 8, cudaMemcpyDeviceToDevice, 0));
    CK(cudaEventRecord(e1, 0)); CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("cudaMemcpyAsync D2D of the same bytes: %.4f ms/op (%.0f GB/s)\n", ms / reps, 8.0 * (N + N / 2) / (ms / reps) / 1e6);
  return 0;
}
