// Experiment: dependent-chain latency and single-warp throughput of DFMA/DMUL on this GPU (cycles per instruction).
#include <cstdio>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters) {
  double a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double x = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps) {
  double* out; long long* c; cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&c, 8);
  const int iters = 4096;
  k<ILP><<<1, 32 * warps>>>(out, c, iters); cudaDeviceSynchronize();
  k<ILP><<<1, 32 * warps>>>(out, c, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("warps/SM %2d ILP %2d: %.2f cycles per DFMA per warp, %.2f DFMA warp-instr/cycle/SM\n", warps, ILP, (double)h / (iters * ILP), (double)iters * ILP * warps / h);
  cudaFree(out); cudaFree(c);
}
int main() {
  run<1>(1); run<2>(1); run<4>(1); run<8>(1); run<16>(1);
  run<1>(4); run<8>(4); run<8>(8); run<8>(11); run<8>(12); run<8>(16); run<4>(16); run<2>(32);
  return 0;
}
