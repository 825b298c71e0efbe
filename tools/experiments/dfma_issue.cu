// Experiment: does a DFMA occupy the scheduler's issue slot for two cycles, or only the FP64 pipe?
// Per iteration and warp: 8 independent DFMA + K independent integer instructions (LOP3/IADD mix), 12 warps per SM.
// pipe-only model: cycles/iter/scheduler = 3 warps * max(16, 8 + K); issue-blocking model: 3 * (16 + K).
#include <cstdio>
template <int K>
__global__ void k(double* out, long long* cyc, int iters) {
  double a[8];
  unsigned x[8];
  for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; x[i] = threadIdx.x * 2654435761u + i; }
  const double m = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      a[i] = fma(a[i], m, y);
#pragma unroll
      for (int j = 0; j < K; j++) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(i + j) & 7]) : "r"(it));
      }
    }
  }
  long long t1 = clock64();
  double s = 0; unsigned xs = 0;
  for (int i = 0; i < 8; i++) { s += a[i]; xs ^= x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + xs;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K> void run(int warps) {
  double* out; long long* c; cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&c, 8);
  const int iters = 4096;
  k<K><<<1, 32 * warps>>>(out, c, iters); cudaDeviceSynchronize();
  k<K><<<1, 32 * warps>>>(out, c, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters / (warps / 4.0);
  printf("K %d (int ops per DFMA) warps %2d: %.1f cycles per iteration per warp-slot (8 DFMA + %d int): pipe-only model %d, issue-blocking model %d\n",
         K, warps, per, 8 * K, 16 > 8 + 8 * K ? 16 : 8 + 8 * K, 16 + 8 * K);
  cudaFree(out); cudaFree(c);
}
int main() {
  run<0>(12); run<1>(12); run<2>(12); run<3>(12); run<4>(12);
  run<0>(4); run<1>(4); run<2>(4);
  return 0;
}
