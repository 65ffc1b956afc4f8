// Follow-up to stream_probe.cu: out[i] = a[i] * b[i & 2047] (the P1 traffic mix, 16N bytes) with 16 KB tiles
// (256 threads x 4 x 128 bit, contiguous) under different tile schedules:
//   oneshot      : grid = n_tiles
//   oneshot+D    : same, every CTA first spins D cycles (a stand-in for a per-CTA prologue)
//   rr           : persistent grid, tiles round-robin (contract_kernel's persistent mode)
//   dyn          : persistent grid, tiles from an atomic counter
//   rr2          : persistent, two tiles' loads in flight per iteration (8 x 128 bit per thread)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void do_tile(const double2* __restrict__ a, const double2* __restrict__ b, double2* __restrict__ out, long long tile) {
  const long long i = tile * 1024 + threadIdx.x;
  double2 v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) v[u] = __ldg(a + i + u * 256);
#pragma unroll
  for (int u = 0; u < 4; ++u) { const double2 s = __ldg(b + ((i + u * 256) & 2047)); v[u].x *= s.x; v[u].y *= s.y; }
#pragma unroll
  for (int u = 0; u < 4; ++u) out[i + u * 256] = v[u];
}

__global__ void __launch_bounds__(256, 4) k_oneshot(const double2* a, const double2* b, double2* out, long long nt, int delay) {
  if (delay) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
  do_tile(a, b, out, blockIdx.x);
}
__global__ void __launch_bounds__(256, 4) k_rr(const double2* a, const double2* b, double2* out, long long nt, int delay) {
  if (delay) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
  for (long long t = blockIdx.x; t < nt; t += gridDim.x) do_tile(a, b, out, t);
}
__global__ void __launch_bounds__(256, 4) k_rr2(const double2* __restrict__ a, const double2* __restrict__ b, double2* __restrict__ out, long long nt, int delay) {
  for (long long t = blockIdx.x; t < nt; t += 2 * gridDim.x) {
    const long long i0 = t * 1024 + threadIdx.x;
    const long long t1 = t + gridDim.x < nt ? t + gridDim.x : t;
    const long long i1 = t1 * 1024 + threadIdx.x;
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = __ldg(a + i0 + u * 256); v[4 + u] = __ldg(a + i1 + u * 256); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 s = __ldg(b + ((i0 + u * 256) & 2047)); v[u].x *= s.x; v[u].y *= s.y;
      const double2 r = __ldg(b + ((i1 + u * 256) & 2047)); v[4 + u].x *= r.x; v[4 + u].y *= r.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { out[i0 + u * 256] = v[u]; if (t1 != t) out[i1 + u * 256] = v[4 + u]; }
  }
}
__global__ void __launch_bounds__(256, 4) k_dyn(const double2* a, const double2* b, double2* out, long long nt, unsigned int* ctr) {
  __shared__ unsigned int cur;
  for (;;) {
    if (threadIdx.x == 0) cur = atomicAdd(ctr, 1u);
    __syncthreads();
    const unsigned int t = cur;
    __syncthreads();
    if (t >= nt) break;
    do_tile(a, b, out, t);
  }
}

int main(int argc, char** argv) {
  const int kk = argc > 1 ? atoi(argv[1]) : 24;
  const long long N = 1ll << kk, n2 = N / 2, nt = n2 / 1024;
  double *A[3], *B[3], *O[3];
  unsigned int* ctr;
  for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&A[i], N * 8)); CK(cudaMalloc(&B[i], N * 8)); CK(cudaMalloc(&O[i], N * 8)); CK(cudaMemset(A[i], 0, N * 8)); CK(cudaMemset(B[i], 0, N * 8)); }
  CK(cudaMalloc(&ctr, 4 * 64));
  const double bytes = 16.0 * N, peak = 6546.6;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    for (int r = 0; r < 6; ++r) launch(r);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < 24; ++r) launch(r);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 24;
    printf("%-22s %.4f ms %7.1f GB/s %.3f\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / peak);
  };
#define ARGS(r) (const double2*)A[(r) % 3], (const double2*)B[(r) % 3], (double2*)O[(r) % 3], nt
  timeit("oneshot", [&](int r) { k_oneshot<<<(int)nt, 256>>>(ARGS(r), 0); });
  for (int d : {250, 500, 1000, 2000}) {
    char nm[64]; snprintf(nm, 64, "oneshot+%d", d);
    timeit(nm, [&](int r) { k_oneshot<<<(int)nt, 256>>>(ARGS(r), d); });
  }
  timeit("rr 592", [&](int r) { k_rr<<<592, 256>>>(ARGS(r), 0); });
  timeit("rr 592 +2000", [&](int r) { k_rr<<<592, 256>>>(ARGS(r), 2000); });
  timeit("rr 1184 (2 waves)", [&](int r) { k_rr<<<1184, 256>>>(ARGS(r), 0); });
  timeit("rr 2368 (4 waves)", [&](int r) { k_rr<<<2368, 256>>>(ARGS(r), 0); });
  timeit("rr2 592", [&](int r) { k_rr2<<<592, 256>>>(ARGS(r), 0); });
  timeit("dyn 592", [&](int r) { cudaMemsetAsync(ctr, 0, 4); k_dyn<<<592, 256>>>(ARGS(r), ctr); });
  return 0;
}
