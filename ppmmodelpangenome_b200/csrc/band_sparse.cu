// Sparse evaluation of the Mobile integral on band-kernel trees (sm_100a, FP64 CUDA cores).
//
// A lineage whose subtree holds no present leaf multiplies every slice by exactly what it multiplies for the ALL-ZERO row,
// whose per-slice integrand Z_j the prepass has anyway (zero_row_*: Inference::getPobs2, functions.cpp:300-333).  Hence
//   integrand_j = Z_j * prod over the pattern's DIRTY lineages s alive at j of f_s(j) / f0_s(j),
// numerators from the item's lineage records (band_prune_kernel), denominators from the zero row's records under the same
// parameter vector, ONE division per slice.  A renormalised clean node carries the zero row's shift, so the exponent of the
// ratio is the difference of the two shift prefixes at the slice.  Cost: proportional to the dirty (lineage, tile) incidences,
// ~3 % of the dense terms for a gene present in 70 of 5,000 genomes.
//
// Which lineages are dirty depends on the PATTERN only, so the term lists are built once per handle, like the frontier lists
// of the class sums: sparse_lists_kernel (lane = pattern) walks the leaves and the ops in birth order, keeps one dirty bit per
// internal lineage, and emits for every dirty lineage and every tile it is alive in one 8-byte record (lineage, slices
// [s0, s1) of the tile) -- a count pass, a prefix sum, a fill pass.  Patterns present in more than an eighth of the genomes
// get no list: they have few clean lineages and go through the dense kernels (band_integral_kernel skips the listed ones).
//
// band_sparse_kernel: one warp per item, lanes = slices (R per lane, a tile = 32 R slices, as the dense kernel).  Per tile:
// Y', Z and the exponent difference per slice; then the tile's records 32 at a time -- lane l fetches record l and its two
// lineage records into a per-warp buffer, then each record is broadcast and multiplied into the lanes of its range.
#include <cub/cub.cuh>

#include "band_fns.cuh"

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace ppm {

namespace {

enum { SL_MAXT = 256 };  // tiles per tree the list builder supports (local counters per thread)

#define SPK(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

}  // namespace

// lane = pattern.  FILL = false: cnt_off[pat * n_tiles + tile] = records; FILL = true: recs written at cnt_off[...] + running index.
// dirty: scratch [warps][n_dw][32] one bit per internal lineage.
template <bool FILL>
__global__ void __launch_bounds__(128) sparse_lists_kernel(ModelDev m, BandDev b, long long count, uint32_t* cnt_off, uint2* recs,
                                                           uint32_t* dirty_scratch) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if ((pat & ~31ll) >= count) return;  // a warp with no pattern at all
  const int n_tiles = b.n_tiles, T = b.T, n_dw = (b.n_int + 31) / 32;
  uint32_t* dirty = dirty_scratch + (size_t)(pat >> 5) * n_dw * 32 + lane;
  const int base = pat < count ? bs_sparse_item(m, pat) : 0;  // 1: dirty = a PRESENT leaf below, 2: an ABSENT leaf below
  const bool live = base != 0;
  const uint32_t flip = base == 2 ? 0xffffffffu : 0u;
  uint32_t cur[SL_MAXT];
  for (int t = 0; t < n_tiles; t++) cur[t] = (FILL && live) ? cnt_off[(size_t)pat * n_tiles + t] : 0u;
  const long long pj = pat < m.n_pat_pad ? pat : 0;
  const uint32_t* words = m.words + pj;
  const size_t ws = (size_t)m.n_pat_pad;
  auto emit = [&](uint32_t ref, uint32_t span) {
    const int js0 = (int)(span & 0xffffu), js1 = (int)(span >> 16);
    if (js0 >= js1) return;
    for (int t = js0 / T; t <= (js1 - 1) / T; t++) {
      const int s0 = max(js0, t * T) - t * T, s1 = min(js1, (t + 1) * T) - t * T;
      if (FILL) recs[cur[t]] = make_uint2(ref, (uint32_t)s0 | ((uint32_t)s1 << 8));
      cur[t]++;
    }
  };
  // present leaves (also under leaf powers: turning the leaf counts into one power per slice in the tile prologue was measured
  // -- two more dependent gathers and an exp2 per slice in a latency-bound prologue: 1,000 genomes 170 -> 208 ms, 5,000 genomes
  // unchanged -- and dropped; a listed leaf is a ratio against the baseline row like any other dirty lineage)
  for (int w = 0; w < m.n_words; w++) {
    uint32_t bits = live ? (words[(size_t)w * ws] ^ flip) : 0u;
    while (bits) {
      const int pos = 32 * w + __ffs(bits) - 1;
      bits &= bits - 1;
      if (pos < m.n_leaves) emit((uint32_t)pos | (1u << 31), b.leaf_span[pos]);
    }
  }
  // internal lineages in birth order: dirty = a dirty child (leaf: present)
  uint32_t dacc = 0u;
  for (int i = 0; i < b.n_ops; i++) {
    const BandOpDev o = b.ops[i];
    auto child = [&](int ref, bool leaf) -> uint32_t {
      if (leaf) return live ? ((words[(size_t)(ref >> 5) * ws] ^ flip) >> (ref & 31)) & 1u : 0u;
      const uint32_t wd = (ref >> 5) == (i >> 5) ? dacc : dirty[(size_t)(ref >> 5) * 32];
      return (wd >> (ref & 31)) & 1u;
    };
    const uint32_t d = child(o.lref, o.flags & 1) | child(o.rref, o.flags & 2);
    dacc |= d << (i & 31);
    if ((i & 31) == 31 || i == b.n_ops - 1) { dirty[(size_t)(i >> 5) * 32] = dacc; dacc = 0u; }
    if (d && live) emit((uint32_t)i, b.op_span[i]);
  }
  if (!FILL && pat < count)
    for (int t = 0; t < n_tiles; t++) cnt_off[(size_t)pat * n_tiles + t] = live ? cur[t] : 0u;
}

namespace {
template <int R>
__device__ __forceinline__ void sp_rescale(double (&p)[R], int (&e)[R]) {
  int mn = __double2hiint(p[0]);
#pragma unroll
  for (int k = 1; k < R; k++) { const int h = __double2hiint(p[k]); mn = h < mn ? h : mn; }
  if (mn < ((1023 - 300) << 20)) {
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int be = biased_exp(p[k]);
      if (be != 0 && be < 1023 - 300) { p[k] *= pow2i(300); e[k] += 300; }
    }
  }
}

enum { SPW = 4 };  // warps per CTA
struct __align__(16) SpBuf { double2 n[32]; double2 d[32]; uint32_t rg[32]; };
}  // namespace

template <int R>
__global__ void __launch_bounds__(32 * SPW, 4) band_sparse_kernel(BandStreamArgs a) {
  __shared__ SpBuf bufs[SPW];
  const ModelDev& m = a.m;
  const BandDev& b = a.b;
  const int lane = threadIdx.x & 31;
  SpBuf& buf = bufs[threadIdx.x >> 5];
  const int T = 32 * R, G = b.G, n_int = b.n_int, n_tiles = b.n_tiles;
  const unsigned total = (unsigned)a.n_blk * 32u;

  unsigned u = 0;
  if (lane == 0) u = atomicAdd(a.counter2, 1u);
  u = __shfl_sync(0xffffffffu, u, 0);
  while (u < total) {
    unsigned u_next = 0;
    if (lane == 0) u_next = atomicAdd(a.counter2, 1u);  // its latency hides behind this item
    const int il = (int)u;
    const long long item = a.item0 + il;
    const int p = (int)(item / a.count_pad);
    const long long pat = item - (long long)p * a.count_pad;
    const double* sc = a.t.scal + (size_t)p * SC_COUNT;
    const bool served = sc[SC_BAND] != 0. && (a.force || !(sc[SC_I2] < 0.));
    const int base = (pat < a.count && served) ? bs_sparse_use(m, a.lists, n_tiles, a.sparse_max_len, pat) : 0;
    if (base != 0) {
      const int ones = base == 2 ? 1 : 0;  // baseline row: all-zero / all-ones
      const double2* nodes = a.nodes + (size_t)(il >> 5) * n_int * 32 + (il & 31);
      const int* xk = a.xk + (size_t)(il >> 5) * (n_int + 1) * 32 + (il & 31);
      const double2* nodes0 = a.nodes0 + ((size_t)2 * p + ones) * n_int * 32;
      const int* xk0 = a.xk0 + ((size_t)2 * p + ones) * (n_int + 1) * 32;
      const double* bY = a.bt.bY + (size_t)p * b.S_pad;
      const double2* bleaf = a.bt.bleaf + (size_t)p * m.n_leaves * 2;
      const double2* zpart = ones ? a.z1 + (size_t)p * b.S_pad : a.t.zpart + (size_t)p * m.n_slices;
      const uint32_t* off = a.lists.off + (size_t)pat * n_tiles;
      double Im = 0.; int Ie = 0x3fffffff;
      for (int tau = 0; tau < n_tiles; tau++) {
        double pn[R], pd[R], Y[R];
        int en[R], ed[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
          const int j = tau * T + 32 * k + lane;
          Y[k] = bY[j];
          const double2 z = j < m.n_slices ? zpart[j] : make_double2(0., 0.);  // Z_j: mantissa (x slice weight), exponent
          const int nb = b.nborn[j];
          pn[k] = z.x; pd[k] = 1.;
          en[k] = (int)z.y + xk[(size_t)nb * 32] - xk0[(size_t)nb * 32];
          ed[k] = 0;
        }
        const uint32_t beg = off[tau], end = off[tau + 1];
        for (uint32_t c = beg; c < end; c += 32) {
          __syncwarp();  // (the previous chunk's broadcasts are done)
          {
            double2 rn = make_double2(1., 0.), rd = make_double2(1., 0.);
            uint32_t rg = 0u;  // empty range: s0 = s1 = 0
            if (c + lane < end) {
              const uint2 r = a.lists.recs[c + lane];
              const uint32_t ref = r.x & 0x7fffffffu;
              if (r.x >> 31) { rn = bleaf[ref * 2 + 1 - ones]; rd = bleaf[ref * 2 + ones]; }  // a listed leaf differs from the baseline row
              else { rn = nodes[(size_t)ref * 32]; rd = nodes0[(size_t)ref * 32]; }
              rg = r.y;
            }
            buf.n[lane] = rn; buf.d[lane] = rd; buf.rg[lane] = rg;
          }
          __syncwarp();
          const int n = (int)min(32u, end - c);
#pragma unroll 2
          for (int i = 0; i < n; i++) {
            const double2 rn = buf.n[i], rd = buf.d[i];
            const uint32_t rg = buf.rg[i];
            const uint32_t s0 = rg & 0xffu, len = (rg >> 8) - s0;
#pragma unroll
            for (int k = 0; k < R; k++) {
              const double fn = fma(rn.y, Y[k], rn.x), fd = fma(rd.y, Y[k], rd.x);
              if ((uint32_t)(32 * k + lane) - s0 < len) { pn[k] *= fn; pd[k] *= fd; }
            }
            if ((i & 15) == 15) { sp_rescale<R>(pn, en); sp_rescale<R>(pd, ed); }
          }
          sp_rescale<R>(pn, en); sp_rescale<R>(pd, ed);
        }
#pragma unroll
        for (int k = 0; k < R; k++) scaled_add(Im, Ie, pn[k] / pd[k], en[k] - ed[k]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double vm = __shfl_down_sync(0xffffffffu, Im, o);
        const int ve = __shfl_down_sync(0xffffffffu, Ie, o);
        scaled_add(Im, Ie, vm, ve);
      }
      // the mixture kernel sums G unit integrals per item: this one is unit 0, the others are empty
      for (int w = lane; w < G; w += 32) a.unit_I[(size_t)il * G + w] = w == 0 ? make_double2(Im, (double)Ie) : make_double2(0., (double)0x3fffffff);
    }
    u = __shfl_sync(0xffffffffu, u_next, 0);
  }
}


unsigned long long sparse_lists_count(const ModelDev& m, const BandDev& b, long long count, uint32_t* cnt_off, cudaStream_t s) {
  if (b.n_tiles > SL_MAXT) throw std::runtime_error("sparse evaluation: too many tiles");
  const long long n = count * b.n_tiles;
  const int n_dw = (b.n_int + 31) / 32;
  uint32_t* dirty = nullptr;
  SPK(cudaMalloc(&dirty, sizeof(uint32_t) * (size_t)((count + 31) / 32) * n_dw * 32));
  sparse_lists_kernel<false><<<(unsigned)((count + 127) / 128), 128, 0, s>>>(m, b, count, cnt_off, nullptr, dirty);
  SPK(cudaMemsetAsync(cnt_off + n, 0, sizeof(uint32_t), s));
  size_t tmp_bytes = 0;
  void* tmp = nullptr;
  // exclusive prefix sums in place (+ one trailing element = the total); the sum must fit 32 bits
  unsigned long long total = 0;
  {
    std::vector<uint32_t> h((size_t)n);
    SPK(cudaMemcpyAsync(h.data(), cnt_off, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
    SPK(cudaStreamSynchronize(s));
    for (long long i = 0; i < n; i++) total += h[i];
  }
  cudaFree(dirty);
  if (total >= (1ull << 32)) return total;
  SPK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt_off, cnt_off, (int)(n + 1), s));
  SPK(cudaMalloc(&tmp, tmp_bytes));
  SPK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt_off, cnt_off, (int)(n + 1), s));
  SPK(cudaStreamSynchronize(s));
  cudaFree(tmp);
  SPK(cudaGetLastError());
  return total;
}

void sparse_lists_fill(const ModelDev& m, const BandDev& b, long long count, const uint32_t* off, uint2* recs, cudaStream_t s) {
  const int n_dw = (b.n_int + 31) / 32;
  uint32_t* dirty = nullptr;
  SPK(cudaMalloc(&dirty, sizeof(uint32_t) * (size_t)((count + 31) / 32) * n_dw * 32));
  sparse_lists_kernel<true><<<(unsigned)((count + 127) / 128), 128, 0, s>>>(m, b, count, const_cast<uint32_t*>(off), recs, dirty);
  SPK(cudaStreamSynchronize(s));
  cudaFree(dirty);
  SPK(cudaGetLastError());
}

void launch_band_sparse(const BandStreamArgs& a, int n_sm, cudaStream_t s) {
  const long long units = (long long)a.n_blk * 32;
  const int grid = (int)std::max<long long>(1, std::min<long long>((long long)n_sm * 4, (units + SPW - 1) / SPW));
  if (a.b.R == 2) band_sparse_kernel<2><<<grid, 32 * SPW, 0, s>>>(a);
  else band_sparse_kernel<4><<<grid, 32 * SPW, 0, s>>>(a);
}

}  // namespace ppm
