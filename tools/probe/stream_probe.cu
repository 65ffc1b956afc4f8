// What does a trivially simple kernel reach on this B200 for the traffic mixes of the factor ops?
//   read   : sum of N doubles (pure read)                8N bytes
//   sum2   : out[i] = a[2i] + a[2i+1]  (S1)              8(N + N/2)
//   sumhi  : out[i] = a[i] + a[i + N/2] (S2)             8(N + N/2)
//   copy   : out[i] = a[i]                               16N
//   scale  : out[i] = a[i] * b[i & 4095]  (P1-like)      16N
// each as (p) persistent grid-stride with 4 x 128-bit loads in flight and (n) one-shot CTAs of 256 thr x 4 x 16 B.
// Operands rotate over 3 buffer sets (>> L2).  nvcc -arch=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE, bool PERSIST>
__global__ void __launch_bounds__(256) k(const double2* __restrict__ a, const double2* __restrict__ b, double2* __restrict__ out,
                                         long long n2, double* sink) {
  // n2 = N/2 double2 elements of input
  const long long step = PERSIST ? (long long)gridDim.x * 256 : 256;
  const long long chunk = 4;
  long long i = PERSIST ? (long long)blockIdx.x * 256 + threadIdx.x : (long long)blockIdx.x * 1024 + threadIdx.x;
  double acc = 0;
  for (; i < n2; i += (PERSIST ? chunk * step : n2)) {
    if (MODE == 0) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * step < n2) ? __ldg(a + i + u * step) : make_double2(0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y;
    } else if (MODE == 1) {  // sum2: 4 double2 in -> 4 doubles out; pair two loads per double2 store
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * step < n2) ? __ldg(a + i + u * step) : make_double2(0, 0);
      double* o = reinterpret_cast<double*>(out);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i + u * step < n2) o[i + u * step] = v[u].x + v[u].y;
    } else if (MODE == 2) {  // sumhi: i over N/4 double2 outputs
      double2 v[4], w[4];
      const long long h = n2 / 2;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        v[u] = (i + u * step < h) ? __ldg(a + i + u * step) : make_double2(0, 0);
        w[u] = (i + u * step < h) ? __ldg(a + h + i + u * step) : make_double2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) if (i + u * step < h) out[i + u * step] = make_double2(v[u].x + w[u].x, v[u].y + w[u].y);
      if (PERSIST) i -= 2 * step;  // only 2 per iteration
      if (!PERSIST && i + 2 * step < h) {  // second half of the one-shot chunk
#pragma unroll
        for (int u = 2; u < 4; ++u) {
          v[u] = (i + u * step < h) ? __ldg(a + i + u * step) : make_double2(0, 0);
          w[u] = (i + u * step < h) ? __ldg(a + h + i + u * step) : make_double2(0, 0);
        }
#pragma unroll
        for (int u = 2; u < 4; ++u) if (i + u * step < h) out[i + u * step] = make_double2(v[u].x + w[u].x, v[u].y + w[u].y);
      }
      if (i >= h) break;
    } else {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * step < n2) ? __ldg(a + i + u * step) : make_double2(0, 0);
      if (MODE == 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { const double2 s = __ldg(b + ((i + u * step) & 2047)); v[u].x *= s.x; v[u].y *= s.y; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i + u * step < n2) out[i + u * step] = v[u];
    }
  }
  if (MODE == 0 && acc == 1.2345e-300) *sink = acc;
}

template <int MODE, bool PERSIST>
float run(double** A, double** B, double** O, long long N, int grid_p, double* sink, int reps) {
  const long long n2 = N / 2;
  long long work2 = (MODE == 2) ? n2 / 2 : n2;
  int grid = PERSIST ? grid_p : (int)((work2 + 1023) / 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 6; ++r) k<MODE, PERSIST><<<grid, 256>>>((double2*)A[r % 3], (double2*)B[r % 3], (double2*)O[r % 3], n2, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k<MODE, PERSIST><<<grid, 256>>>((double2*)A[r % 3], (double2*)B[r % 3], (double2*)O[r % 3], n2, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main(int argc, char** argv) {
  const int kk = argc > 1 ? atoi(argv[1]) : 24;
  const long long N = 1ll << kk;
  double *A[3], *B[3], *O[3], *sink;
  for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&A[i], N * 8)); CK(cudaMalloc(&B[i], N * 8)); CK(cudaMalloc(&O[i], N * 8)); CK(cudaMemset(A[i], 0, N * 8)); CK(cudaMemset(B[i], 0, N * 8)); }
  CK(cudaMalloc(&sink, 8));
  const char* names[] = {"read", "sum2", "sumhi", "copy", "scale"};
  const double bytes[] = {8.0 * N, 12.0 * N, 12.0 * N, 16.0 * N, 16.0 * N};
  const double peak = 6546.6;
  for (int gp : {148 * 4, 148 * 8}) {
    float t;
    t = run<0, true>(A, B, O, N, gp, sink, 24); printf("%-6s persist grid %4d  %.4f ms %7.1f GB/s %.3f\n", names[0], gp, t, bytes[0] / t / 1e6, bytes[0] / t / 1e6 / peak);
    t = run<1, true>(A, B, O, N, gp, sink, 24); printf("%-6s persist grid %4d  %.4f ms %7.1f GB/s %.3f\n", names[1], gp, t, bytes[1] / t / 1e6, bytes[1] / t / 1e6 / peak);
    t = run<2, true>(A, B, O, N, gp, sink, 24); printf("%-6s persist grid %4d  %.4f ms %7.1f GB/s %.3f\n", names[2], gp, t, bytes[2] / t / 1e6, bytes[2] / t / 1e6 / peak);
    t = run<3, true>(A, B, O, N, gp, sink, 24); printf("%-6s persist grid %4d  %.4f ms %7.1f GB/s %.3f\n", names[3], gp, t, bytes[3] / t / 1e6, bytes[3] / t / 1e6 / peak);
    t = run<4, true>(A, B, O, N, gp, sink, 24); printf("%-6s persist grid %4d  %.4f ms %7.1f GB/s %.3f\n", names[4], gp, t, bytes[4] / t / 1e6, bytes[4] / t / 1e6 / peak);
  }
  float t;
  t = run<0, false>(A, B, O, N, 0, sink, 24); printf("%-6s one-shot            %.4f ms %7.1f GB/s %.3f\n", names[0], t, bytes[0] / t / 1e6, bytes[0] / t / 1e6 / peak);
  t = run<1, false>(A, B, O, N, 0, sink, 24); printf("%-6s one-shot            %.4f ms %7.1f GB/s %.3f\n", names[1], t, bytes[1] / t / 1e6, bytes[1] / t / 1e6 / peak);
  t = run<2, false>(A, B, O, N, 0, sink, 24); printf("%-6s one-shot            %.4f ms %7.1f GB/s %.3f\n", names[2], t, bytes[2] / t / 1e6, bytes[2] / t / 1e6 / peak);
  t = run<3, false>(A, B, O, N, 0, sink, 24); printf("%-6s one-shot            %.4f ms %7.1f GB/s %.3f\n", names[3], t, bytes[3] / t / 1e6, bytes[3] / t / 1e6 / peak);
  t = run<4, false>(A, B, O, N, 0, sink, 24); printf("%-6s one-shot            %.4f ms %7.1f GB/s %.3f\n", names[4], t, bytes[4] / t / 1e6, bytes[4] / t / 1e6 / peak);
  return 0;
}
