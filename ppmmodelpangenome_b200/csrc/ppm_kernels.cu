// CUDA kernels of the PPM likelihood engine (sm_100a; FP64 CUDA cores, no tensor cores: the work is a
// masked product-reduction, not a contraction).
//
// Arithmetic (DESIGN.md "The path as arithmetic"; reference source/functions.cpp:245-372, objects.cpp:400-445):
//  * pure-loss categories 0/1 collapse to a SUM of per-branch log factors selected by the pattern's
//    "all-zero subtree" bits (no per-lineage state at all);
//  * the Mobile category keeps, per alive lineage s, a = m0 + pi1*(m1-m0) and d = m1-m0, so that the
//    integrand of likRep2 is  prod_s (a_s - d_s*K_s*Y_j)  -- one DFMA + one DMUL per (slice, lineage),
//    no transcendental in the loop -- and a merge is  m0 = prod_c (a_c - k0_c d_c),  m1 = prod_c (a_c + k1_c d_c);
//  * every exp/log lives in per-parameter-vector tables written by the prepass kernels.
// Mapping: one lane per (pattern, parameter vector) pair; the 32 lanes of a warp hold 32 patterns of the
// same parameter vector, so the sweep program (topology only) is executed warp-uniformly: no divergence, no
// masks; per-lineage partials live in shared memory as slots[slot][lane].
#include "ppm_device_fns.cuh"

#include <cmath>
#include <cstdio>

namespace ppm {

// ------------------------------------------------------------------------------------------------ kernels
__global__ void prepass_nodes_kernel(ModelDev m, TablesDev t) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < t.P) prepass_nodes_body(m, t, p);
}
__global__ void prepass_tables_kernel(ModelDev m, TablesDev t) {
  prepass_tables_body(m, t, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void prepass_i2_kernel(ModelDev m, TablesDev t, const double* i2_override) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < t.P) prepass_i2_body(m, t, i2_override, p);
}

template <int R, bool SMEM_SLOTS>
__global__ void __launch_bounds__(256, 2) sweep_kernel(SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ModelDev& m = a.m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  // shared memory: per warp { words[n_words][32] u32 | zw[n_slots] u32 | slots[n_slots][32] double2 }
  const size_t slot_bytes = SMEM_SLOTS ? (size_t)m.n_slots * 32 * sizeof(double2) : 0;
  const size_t head_bytes = (((size_t)m.n_words * 32 + m.n_slots) * 4 + 15) & ~(size_t)15;
  unsigned char* wbase = smem_raw + (size_t)warp * (head_bytes + slot_bytes);
  uint32_t* sw = reinterpret_cast<uint32_t*>(wbase);
  uint32_t* zw = sw + m.n_words * 32;
  SlotStore<SMEM_SLOTS> slots;
  if (SMEM_SLOTS) slots.base = reinterpret_cast<double2*>(wbase + head_bytes);
  else slots.base = a.gslots + ((size_t)blockIdx.x * wpc + warp) * (size_t)m.n_slots * 32;
  // work item = (parameter vector, batch of 32 patterns); warps stride over the items
  for (long long item = (long long)blockIdx.x * wpc + warp; item < a.n_items; item += (long long)gridDim.x * wpc) {
    double contrib = sweep_item<R>(a, item, lane, sw, zw, slots);
    if (a.mode == 1) continue;
    // fixed-order warp reduction -> partial[p][batch]; reduce_kernel sums the batches in a fixed order too
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    if (lane == 0) a.partial[item] = contrib;
    __syncwarp();
  }
}

__global__ void reduce_kernel(const double* partial, int nx, double* out) {
  // one CTA per parameter vector; fixed summation order => bitwise reproducible
  __shared__ double sh[256];
  const double* row = partial + (size_t)blockIdx.x * nx;
  double s = 0.;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) s += row[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

__global__ void finalize_kernel(const double* i2, double* ll, int P) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P && i2[p] < 0.) ll[p] = -1e9 * 1. + i2[p];  // functions.cpp:362-364
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_prepass_nodes(const ModelDev& m, const TablesDev& t, cudaStream_t s) {
  prepass_nodes_kernel<<<(t.P + 63) / 64, 64, 0, s>>>(m, t);
}
void launch_prepass_tables(const ModelDev& m, const TablesDev& t, cudaStream_t s) {
  int n = m.n_entries + m.n_slices;
  dim3 g((n + 255) / 256, t.P);
  prepass_tables_kernel<<<g, 256, 0, s>>>(m, t);
}
void launch_prepass_i2(const ModelDev& m, const TablesDev& t, const double* i2_override, cudaStream_t s) {
  prepass_i2_kernel<<<(t.P + 63) / 64, 64, 0, s>>>(m, t, i2_override);
}

static size_t per_warp_smem(const ModelDev& m, bool smem_slots) {
  size_t head = (((size_t)m.n_words * 32 + m.n_slots) * 4 + 15) & ~(size_t)15;
  return head + (smem_slots ? (size_t)m.n_slots * 32 * sizeof(double2) : 0);
}

int sweep_grid_x(const ModelDev& m, long long n_items, int n_sm, int* warps_per_cta, size_t* smem_bytes, bool* smem_slots) {
  const size_t budget = 100 * 1024;  // two CTAs per SM
  bool ss = per_warp_smem(m, true) * 2 <= budget;
  size_t pw = per_warp_smem(m, ss);
  int wpc = (int)std::min<size_t>(8, budget / pw);
  if (wpc < 1) wpc = 1;
  if (n_items < wpc) wpc = (int)std::max<long long>(1, n_items);
  *warps_per_cta = wpc;
  *smem_bytes = pw * wpc;
  *smem_slots = ss;
  long long ctas = (n_items + wpc - 1) / wpc;
  // shared-memory variant: one item per warp; global-slot variant: a persistent grid bounds the scratch
  if (!ss) ctas = std::min<long long>(ctas, (long long)n_sm * 2);
  return (int)std::min<long long>(ctas, 2147483647LL);
}

template <int R, bool SS>
static void launch_sweep_t(const SweepArgs& a, int grid_x, int wpc, size_t smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sweep_kernel<R, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  sweep_kernel<R, SS><<<grid_x, wpc * 32, smem, s>>>(a);
}

void launch_sweep(const SweepArgs& a, int grid_x, int wpc, size_t smem, bool smem_slots, cudaStream_t s) {
  if (smem_slots) launch_sweep_t<8, true>(a, grid_x, wpc, smem, s);
  else launch_sweep_t<8, false>(a, grid_x, wpc, smem, s);
}
void launch_reduce(const double* partial, int P, int nx, double* out, cudaStream_t s) {
  reduce_kernel<<<P, 256, 0, s>>>(partial, nx, out);
}
void launch_finalize(const double* i2, double* ll, int P, cudaStream_t s) {
  finalize_kernel<<<(P + 127) / 128, 128, 0, s>>>(i2, ll, P);
}

// ------------------------------------------------------------------------------------------------ DFMA peak
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double run_fp64_peak(int device) {
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * (double)iters * blocks * threads;
    if (rep > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? best : -1;
}

}  // namespace ppm
