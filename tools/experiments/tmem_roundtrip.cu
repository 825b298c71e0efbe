// Experiment (not part of the product): can TMEM serve as per-lane scratch for the lineage slots?
// 16 warps store n 16-byte slots per lane with tcgen05.st, read them back with tcgen05.ld, check the values and
// time dependent LDTM round trips.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_roundtrip tmem_roundtrip.cu
#include <cstdint>
#include <cstdio>
#include <vector>
__global__ void __launch_bounds__(512, 1) k(double2* out, const double2* in, int n, long long* cyc) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tbase_s;
  const uint32_t myaddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 128);
  for (int s = 0; s < n; s++) {
    double2 v = in[(s * 16 + warp) * 32 + lane];
    uint32_t r0 = __double2loint(v.x), r1 = __double2hiint(v.x), r2 = __double2loint(v.y), r3 = __double2hiint(v.y);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(myaddr + 4 * s), "r"(r0), "r"(r1), "r"(r2), "r"(r3));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  double2 acc = make_double2(0., 0.);
  long long t0 = clock64();
  for (int rep = 0; rep < 64; rep++)
    for (int s = n - 1; s >= 0; s--) {
      uint32_t r0, r1, r2, r3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(myaddr + 4 * s));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc.x += __hiloint2double(r1, r0) * (s + 1);
      acc.y += __hiloint2double(r3, r2) * (s + 1);
    }
  long long t1 = clock64();
  out[(blockIdx.x * 16 + warp) * 32 + lane] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
}
int main() {
  const int n = 32, nb = 148 * 2;
  std::vector<double2> in(n * 16 * 32);
  for (size_t i = 0; i < in.size(); i++) in[i] = make_double2(1.0 + i * 1e-3, -2.0 - i * 1e-3);
  double2 *din, *dout; long long* dc;
  cudaMalloc(&din, in.size() * sizeof(double2)); cudaMalloc(&dout, nb * 512 * sizeof(double2)); cudaMalloc(&dc, nb * 8);
  cudaMemcpy(din, in.data(), in.size() * sizeof(double2), cudaMemcpyHostToDevice);
  k<<<nb, 512>>>(dout, din, n, dc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double2> out(nb * 512); std::vector<long long> cyc(nb);
  cudaMemcpy(out.data(), dout, out.size() * sizeof(double2), cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc.data(), dc, nb * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int b = 0; b < nb; b++)
    for (int w = 0; w < 16; w++)
      for (int l = 0; l < 32; l++) {
        double ex = 0, ey = 0;
        for (int s = 0; s < n; s++) { double2 v = in[(s * 16 + w) * 32 + l]; ex += v.x * (s + 1); ey += v.y * (s + 1); }
        ex *= 64; ey *= 64;
        double2 g = out[(b * 16 + w) * 32 + l];
        if (fabs(g.x - ex) > 1e-6 * fabs(ex) || fabs(g.y - ey) > 1e-6 * fabs(ey)) bad++;
      }
  printf("mismatches: %d of %d; cycles per dependent LDTM round trip (16 warps/SM): %.1f\n", bad, nb * 512, (double)cyc[0] / (64.0 * n));
  return bad != 0;
}
