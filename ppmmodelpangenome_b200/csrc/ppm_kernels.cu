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

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace ppm {

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(512) prepass_nodes_kernel(ModelDev m, TablesDev t) {
  prepass_nodes_body(m, t, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void prepass_tables_kernel(ModelDev m, TablesDev t) {
  prepass_tables_body(m, t, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}
// Class sums.  The frontier of a pattern depends on the pattern only, so it is listed once per handle (frontier_* kernels);
// per evaluation  a01[pattern][p] = Stot[p] + sum over the frontier of F[node][p].
// (1) tiled: a CTA stages F[.][p0..p0+31] in shared memory (n_nodes x 512 B) and serves a range of patterns, one warp per
//     pattern, lanes = parameter vectors; (2) big trees: the same from L2; (3) no lists (> 65,535 nodes): walk per call.
__global__ void __launch_bounds__(256) class_sums_tiled_kernel(ModelDev m, TablesDev t, long long count, int pats_per_cta) {
  extern __shared__ __align__(16) unsigned char cst_raw[];
  double2* tile = reinterpret_cast<double2*>(cst_raw);  // [n_nodes][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int p0 = blockIdx.x * 32, p = p0 + lane;
  const bool pv = p < t.P;
  for (int v = warp; v < m.n_nodes; v += wpc) tile[v * 32 + lane] = pv ? t.clsT[(size_t)v * t.P + p] : make_double2(0., 0.);
  double s0 = 0., s1 = 0.;
  if (pv) { const double* sc = t.scal + (size_t)p * SC_COUNT; s0 = sc[SC_STOT0]; s1 = sc[SC_STOT1]; }
  __syncthreads();
  const long long j0 = (long long)blockIdx.y * pats_per_cta;
  const long long j1 = j0 + pats_per_cta < count ? j0 + pats_per_cta : count;
  for (long long pat = j0 + warp; pat < j1; pat += wpc) {
    const uint32_t o0 = m.fr_off[pat], o1 = m.fr_off[pat + 1];
    double a0 = s0, a1 = s1;
    for (uint32_t k = o0; k < o1; k++) {
      const double2 f = tile[(int)m.fr_nodes[k] * 32 + lane];
      a0 += f.x; a1 += f.y;
    }
    if (pv) t.a01[(size_t)pat * t.P + p] = make_double2(a0, a1);
  }
}
__global__ void __launch_bounds__(256) class_sums_list_kernel(ModelDev m, TablesDev t, long long count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const long long pat = (long long)blockIdx.x * wpc + warp;
  if (pat >= count) return;
  const uint32_t o0 = m.fr_off[pat], o1 = m.fr_off[pat + 1];
  for (int p = lane; p < t.P; p += 32) {
    const double* sc = t.scal + (size_t)p * SC_COUNT;
    double a0 = sc[SC_STOT0], a1 = sc[SC_STOT1];
    for (uint32_t k = o0; k < o1; k++) {
      const double2 f = t.clsT[(size_t)m.fr_nodes[k] * t.P + p];
      a0 += f.x; a1 += f.y;
    }
    t.a01[(size_t)pat * t.P + p] = make_double2(a0, a1);
  }
}
// one warp per pattern walks the node ops (class_frontier); MODE 0: count, 1: fill the lists, 2: no lists, sum right away
template <int MODE>
__global__ void __launch_bounds__(256) class_walk_kernel(ModelDev m, TablesDev t, long long count, uint32_t* nf_out, const uint32_t* off,
                                                         uint16_t* nodes_out) {
  extern __shared__ uint32_t cs_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int zwords = (m.n_nodes + 31) >> 5, fwords = (m.n_nodes + 1) >> 1;
  uint32_t* zbits = cs_smem + (size_t)warp * (zwords + fwords);
  uint16_t* frontier = reinterpret_cast<uint16_t*>(zbits + zwords);
  const long long pat = (long long)blockIdx.x * wpc + warp;
  if (pat >= count) return;
  const int nf = class_frontier(m, m.words + pat, (size_t)m.n_pat_pad, zbits, frontier);
  __syncwarp();
  if (MODE == 0) { if (lane == 0) nf_out[pat] = (uint32_t)nf; }
  else if (MODE == 1) { for (int k = lane; k < nf; k += 32) nodes_out[off[pat] + k] = frontier[k]; }
  else { for (int p = lane; p < t.P; p += 32) t.a01[(size_t)pat * t.P + p] = class_sums(t, p, frontier, nf); }
}

// Pattern rows (row-major, bit k = k-th leaf in preorder) -> device layout: word-major, bits in program leaf order.  One thread
// per (pattern, destination word), patterns fastest: the 32 source bits of a word come from anywhere in the thread's own row
// (L1), the store is coalesced.
__global__ void __launch_bounds__(256) permute_words_kernel(const uint32_t* __restrict__ rows, const int* __restrict__ inv, long long count,
                                                            int nw, long long n_pat_pad, uint32_t* __restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = blockIdx.y;
  if (j >= count) return;
  const uint32_t* row = rows + (size_t)j * nw;
  uint32_t v = 0u;
#pragma unroll 8
  for (int b = 0; b < 32; b++) {
    const int k = inv[w * 32 + b];
    if (k >= 0) v |= ((row[k >> 5] >> (k & 31)) & 1u) << b;
  }
  out[(size_t)w * n_pat_pad + j] = v;
}
void launch_permute_words(const uint32_t* rows, const int* inv, long long count, int nw, long long n_pat_pad, uint32_t* out, cudaStream_t s) {
  if (count <= 0) return;
  permute_words_kernel<<<dim3((unsigned)((count + 255) / 256), nw), 256, 0, s>>>(rows, inv, count, nw, n_pat_pad, out);
}

// The all-zero row as three launches: shift prefixes (one CTA per parameter vector), slices (a grid over (slices, parameter
// vectors): with few parameter vectors one CTA each left the GPU idle for 10 ms at 5,000 genomes), fixed-order sum.
__global__ void __launch_bounds__(256) zero_row_prefix_kernel(ModelDev m, TablesDev t) {
  zero_row_prefix_body(m, t, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(128) zero_row_slices_kernel(ModelDev m, TablesDev t) {
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j < m.n_slices) zero_row_slice(m, t, blockIdx.x, j);
}
// fixed-shape sum of the slice values (mantissa * 2^-exponent): 256 contiguous chunks in slice order, then a binary tree
__global__ void __launch_bounds__(256) zero_row_sum_kernel(ModelDev m, TablesDev t) {
  __shared__ double sm[256];
  __shared__ int se[256];
  const int p = blockIdx.x, tid = threadIdx.x;
  const double2* zp = t.zpart + (size_t)p * m.n_slices;
  const int chunk = (m.n_slices + 255) / 256;
  double Im = 0.; int Ie = 0x3fffffff;
  for (int j = tid * chunk; j < min((tid + 1) * chunk, m.n_slices); j++) scaled_add(Im, Ie, zp[j].x, (int)zp[j].y);
  sm[tid] = Im; se[tid] = Ie;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { scaled_add(Im, Ie, sm[tid + o], se[tid + o]); sm[tid] = Im; se[tid] = Ie; }
    __syncthreads();
  }
  if (tid == 0) {
    double* scw = t.scal + (size_t)p * SC_COUNT;
    scw[SC_IZ_M] = Im; scw[SC_IZ_E] = (double)Ie;
  }
}
__global__ void prepass_i2_kernel(ModelDev m, TablesDev t, const double* i2_override) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < t.P) prepass_i2_body(m, t, i2_override, p);
}

__device__ __forceinline__ void copy16(void* dst, const void* src, size_t bytes) {
  // cooperative CTA copy, 16 B per thread per step; all staged arrays are 16-byte multiples and 16-byte aligned
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(src);
  for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

static const size_t kRecRingBytes = 3 * 32 * sizeof(EntryRec);  // per-warp record ring of the global-memory variant

struct StageLayout {
  size_t recs, blk, ops, oprec, yy, wt, uu, leafab, pairq, cherry, leaf, lb, lce, total;
};
__host__ __device__ inline StageLayout stage_layout(const ModelDev& m) {
  StageLayout L;
  size_t o = 0;
  L.recs = o; o += align16((size_t)(m.n_entries + 1) * sizeof(EntryRec));
  L.blk = o;  o += align16((size_t)(m.n_blocks + 1) * sizeof(BlockDesc));
  L.ops = o;  o += align16((size_t)(m.n_ops + 1) * sizeof(SweepOpDev));
  L.oprec = o; o += align16((size_t)m.n_ops * OPREC_DOUBLES * sizeof(double));
  L.yy = o;   o += align16((size_t)m.n_blocks * m.R * sizeof(double));
  L.wt = o;   o += align16((size_t)m.n_blocks * m.R * sizeof(double));
  L.uu = o;   o += align16((size_t)m.n_blocks * m.R * sizeof(double));
  // (a leaf-power program never reads the leaf / leaf-pair tables: three per-slice count tables instead)
  L.leafab = o; o += m.lp ? 0 : align16((size_t)m.n_leaves * 2 * sizeof(double2));
  L.pairq = o; o += m.lp ? 0 : align16((size_t)m.n_pairs * 4 * sizeof(double4));
  L.lb = o; o += m.lp ? align16((size_t)m.n_blocks * m.R * sizeof(double)) : 0;
  L.lce = o; o += m.lp ? align16((size_t)m.n_blocks * m.R * sizeof(double2)) : 0;
  L.cherry = o; o += align16((size_t)m.n_cherries * 4 * sizeof(double2));
  L.leaf = o; o += 32;
  L.total = o;
  return L;
}
// global-memory variant: the small tables named by `mask` (DS_* bits) are staged in shared memory all the same
struct DeepStage { size_t pairq, yy, wt, uu, blk, cherry, leafab, oprec, ops, lb, lce, total; };
__host__ __device__ inline size_t deep_stage_bytes(const ModelDev& m, int bit) {
  const size_t nbR = (size_t)m.n_blocks * m.R;
  switch (bit) {
    case DS_PAIRQ: return m.lp ? 0 : align16((size_t)m.n_pairs * 4 * sizeof(double4));
    case DS_SLICES: return (m.lp ? 6 : 3) * align16(nbR * sizeof(double)) + align16((size_t)(m.n_blocks + 1) * sizeof(BlockDesc));
    case DS_CHERRY: return align16((size_t)m.n_cherries * 4 * sizeof(double2));
    case DS_LEAFAB: return m.lp ? 0 : align16((size_t)m.n_leaves * 2 * sizeof(double2));
    case DS_OPS: return align16((size_t)m.n_ops * OPREC_DOUBLES * sizeof(double)) + align16((size_t)(m.n_ops + 1) * sizeof(SweepOpDev));
  }
  return 0;
}
__host__ __device__ inline DeepStage deep_stage_layout(const ModelDev& m, int mask) {
  DeepStage D{};
  const size_t nbR = (size_t)m.n_blocks * m.R;
  size_t o = 0;
  if (mask & DS_PAIRQ) { D.pairq = o; o += deep_stage_bytes(m, DS_PAIRQ); }
  if (mask & DS_SLICES) {
    D.yy = o; o += align16(nbR * sizeof(double));
    D.wt = o; o += align16(nbR * sizeof(double));
    D.uu = o; o += align16(nbR * sizeof(double));
    if (m.lp) {
      D.lb = o; o += align16(nbR * sizeof(double));
      D.lce = o; o += align16(nbR * sizeof(double2));
    }
    D.blk = o; o += align16((size_t)(m.n_blocks + 1) * sizeof(BlockDesc));
  }
  if (mask & DS_CHERRY) { D.cherry = o; o += deep_stage_bytes(m, DS_CHERRY); }
  if (mask & DS_LEAFAB) { D.leafab = o; o += deep_stage_bytes(m, DS_LEAFAB); }
  if (mask & DS_OPS) {
    D.oprec = o; o += align16((size_t)m.n_ops * OPREC_DOUBLES * sizeof(double));
    D.ops = o; o += align16((size_t)(m.n_ops + 1) * sizeof(SweepOpDev));
  }
  D.total = o;
  return D;
}
// per item and warp: pattern words [n_words][32]
__host__ __device__ inline size_t warp_head_bytes(const ModelDev& m) { return m.word_stride ? 0 : align16((size_t)m.n_words * 32 * 4); }

template <int MODE>
__device__ __forceinline__ void init_slots(SlotsMem& s, double2* smem_slots, double2* global_slots, uint32_t, int, int, int, int) {
  s.base = (MODE == 2 || MODE == 4) ? global_slots : smem_slots;
}
template <int MODE>
__device__ __forceinline__ void init_slots(SlotsTmem& s, double2* smem_slots, double2*, uint32_t tmem_base, int warp, int wpc, int q, int nt) {
  const int cols_per_warp = (512 / ((wpc + 3) >> 2)) & ~3;
  s.taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cols_per_warp + q * nt * 4);
  s.nt = nt;
  s.ovf = smem_slots;
}

// CTA work unit = (parameter vector p, chunk of warps*items_per_warp*NI pattern batches).
// MODE 0: tables + program staged in shared memory, lineage partials in shared memory (<= 12 warps).
// MODE 1: tables + program staged in shared memory, lineage partials in TENSOR MEMORY (tcgen05.ld/st; 16 warps):
//         TMEM is idle in this FP64 kernel and is per-lane private like the slots, so moving them there frees the
//         ~17 KB of shared memory per warp that bounds the resident warps of MODE 0.
// MODE 2: everything read from global memory (L1/L2), partials in a global scratch, deep prefetch (big trees).
// Every warp sweeps NI batches (NI x 32 patterns) at once -- see sweep_items.
template <int MODE> struct SlotsOf { typedef SlotsMem type; };
template <> struct SlotsOf<1> { typedef SlotsTmem type; };
template <> struct SlotsOf<3> { typedef SlotsTmem type; };  // MODE 3 = MODE 1 compiled for 12 warps (168 registers)

template <int R, int NI, int MODE, bool DIST = false, bool WG = false, bool HOT = false>  // WG: pattern rows read in place (global-memory variant, big trees); HOT: see sweep_items
__global__ void __launch_bounds__((MODE == 0 || MODE == 3 || MODE == 4) ? 384 : 512, 1) sweep_kernel(SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ int next_group;  // dynamic hand-out of the unit's warp items: the 4 schedulers of an SM rarely hold equally many warps
  typedef typename SlotsOf<MODE>::type Slots;
  constexpr bool STAGED = MODE != 2 && MODE != 4;  // MODE 4 = MODE 2 compiled for 12 warps (168 registers)
  constexpr bool TMEM = MODE == 1 || MODE == 3;
  const ModelDev& m = a.m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const StageLayout L = stage_layout(m);
  const size_t head = warp_head_bytes(m);
  // lineage slots per item: MODE 0 all in shared memory; MODE 1 the first nt in TMEM, the rest in shared memory
  int nt = 0;
  if (TMEM) {
    const int cols_per_warp = (512 / ((wpc + 3) >> 2)) & ~3;
    nt = cols_per_warp / 4 / NI;
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const int smem_slots = MODE == 0 ? m.n_slots : (TMEM ? (m.n_slots > nt ? m.n_slots - nt : 0) : 0);
  const size_t per_item = head + (size_t)smem_slots * 32 * sizeof(double2);
  const size_t per_warp = per_item * NI + (STAGED ? 0 : kRecRingBytes);
  const DeepStage DS = deep_stage_layout(m, STAGED ? 0 : a.stage_mask);
  unsigned char* wbase = smem_raw + (STAGED ? L.total : DS.total) + (size_t)warp * per_warp;
  EntryRec* wrec = STAGED ? nullptr : reinterpret_cast<EntryRec*>(wbase + per_item * NI);
  ItemMem<Slots> mem[NI];
#pragma unroll
  for (int q = 0; q < NI; q++) {
    unsigned char* ib = wbase + (size_t)q * per_item;
    // every variant copies the item's pattern words to shared memory, except WG (big trees: 40 KB per warp at 10,000 genomes
    // would cost resident warps), which reads the device rows in place
    mem[q].sw_copy = WG ? nullptr : reinterpret_cast<uint32_t*>(ib);
    mem[q].sw = WG ? m.words : reinterpret_cast<const uint32_t*>(ib); mem[q].sws = WG ? (size_t)m.n_pat_pad : 32;
    init_slots<MODE>(mem[q].slots, reinterpret_cast<double2*>(ib + head),
                     a.gslots + (((size_t)blockIdx.x * wpc + warp) * NI + q) * (size_t)(m.n_slots + 1) * 32,
                     tmem_base_s, warp, wpc, q, nt);
  }
  const int nbR = m.n_blocks * R;

  // Persistent CTAs: the (parameter vector, batch) items form one line, cut into gridDim.x equal contiguous
  // ranges (every item costs the same); a range touches at most a few parameter vectors, staged one after the other.
  const long long groups = (a.n_batches + NI - 1) / NI;
  const long long total = groups * a.t.P;
  const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
  const long long g_begin = per_cta * blockIdx.x;
  const long long g_end = g_begin + per_cta < total ? g_begin + per_cta : total;
  for (long long g0 = g_begin; g0 < g_end;) {
    const int p = (int)(g0 / groups);
    const long long b_begin = g0 - (long long)p * groups;
    const long long b_end = (g_end - g0 < groups - b_begin) ? b_begin + (g_end - g0) : groups;
    g0 += b_end - b_begin;
    if (a.fact_filter == 3) {  // the band kernel has served every parameter vector it can: this launch takes the others
      if (a.t.scal[(size_t)p * SC_COUNT + SC_BAND] != 0.) continue;
    } else if (a.fact_filter) {  // (warp-uniform, CTA-uniform) this launch serves only one kind of parameter vector
      const bool fu = a.t.scal[(size_t)p * SC_COUNT + SC_FACTORED] != 0.;
      if ((a.fact_filter == 1) == fu) continue;
    }
    ItemTables T;
    __syncthreads();  // the previous unit's readers (tables, work counter) are done
    if (threadIdx.x == 0) next_group = 0;
    if (STAGED) {
      copy16(smem_raw + L.recs, a.t.recs + (size_t)p * (m.n_entries + 1), align16((size_t)(m.n_entries + 1) * sizeof(EntryRec)));
      copy16(smem_raw + L.blk, m.blk, align16((size_t)(m.n_blocks + 1) * sizeof(BlockDesc)));
      copy16(smem_raw + L.ops, m.sops, align16((size_t)(m.n_ops + 1) * sizeof(SweepOpDev)));
      copy16(smem_raw + L.oprec, a.t.oprec + (size_t)p * m.n_ops * OPREC_DOUBLES, align16((size_t)m.n_ops * OPREC_DOUBLES * sizeof(double)));
      copy16(smem_raw + L.yy, a.t.yy + (size_t)p * nbR, align16((size_t)nbR * sizeof(double)));
      copy16(smem_raw + L.wt, m.slice_wt_pad, align16((size_t)nbR * sizeof(double)));
      copy16(smem_raw + L.uu, a.t.uu + (size_t)p * nbR, align16((size_t)nbR * sizeof(double)));
      if (!m.lp) {
        copy16(smem_raw + L.leafab, a.t.leafab + (size_t)p * m.n_leaves * 2, align16((size_t)m.n_leaves * 2 * sizeof(double2)));
        copy16(smem_raw + L.pairq, a.t.pairq + (size_t)p * m.n_pairs * 4, align16((size_t)m.n_pairs * 4 * sizeof(double4)));
      } else {
        copy16(smem_raw + L.lb, a.t.slb + (size_t)p * nbR, align16((size_t)nbR * sizeof(double)));
        copy16(smem_raw + L.lce, a.t.slce + (size_t)p * nbR, align16((size_t)nbR * sizeof(double2)));
      }
      copy16(smem_raw + L.cherry, a.t.cherry + (size_t)p * m.n_cherries * 4, align16((size_t)m.n_cherries * 4 * sizeof(double2)));
      copy16(smem_raw + L.leaf, a.t.scal + (size_t)p * SC_COUNT + SC_LA0, 32);
      __syncthreads();
      T.recs = reinterpret_cast<const EntryRec*>(smem_raw + L.recs);
      T.blk = reinterpret_cast<const BlockDesc*>(smem_raw + L.blk);
      T.ops = reinterpret_cast<const SweepOpDev*>(smem_raw + L.ops);
      T.oprec = reinterpret_cast<const double*>(smem_raw + L.oprec);
      T.yy = reinterpret_cast<const double*>(smem_raw + L.yy);
      T.wt = reinterpret_cast<const double*>(smem_raw + L.wt);
      T.leaftab = reinterpret_cast<const double2*>(smem_raw + L.leaf);
      T.uu = reinterpret_cast<const double*>(smem_raw + L.uu);
      T.leafab = reinterpret_cast<const double2*>(smem_raw + L.leafab);
      T.pairq = reinterpret_cast<const double4*>(smem_raw + L.pairq);
      T.cherry = reinterpret_cast<const double2*>(smem_raw + L.cherry);
      T.lb = reinterpret_cast<const double*>(smem_raw + L.lb);
      T.lce = reinterpret_cast<const double2*>(smem_raw + L.lce);
    } else {
      T.recs = a.t.recs + (size_t)p * (m.n_entries + 1);
      T.blk = m.blk; T.ops = m.sops;
      T.oprec = a.t.oprec + (size_t)p * m.n_ops * OPREC_DOUBLES;
      T.yy = a.t.yy + (size_t)p * nbR;
      T.wt = m.slice_wt_pad;
      T.leaftab = reinterpret_cast<const double2*>(a.t.scal + (size_t)p * SC_COUNT + SC_LA0);
      T.uu = a.t.uu + (size_t)p * nbR;
      T.leafab = a.t.leafab + (size_t)p * m.n_leaves * 2;
      T.pairq = a.t.pairq + (size_t)p * m.n_pairs * 4;
      T.cherry = a.t.cherry + (size_t)p * m.n_cherries * 4;
      T.lb = a.t.slb + (size_t)p * nbR; T.lce = a.t.slce + (size_t)p * nbR;
      const int sm = a.stage_mask;
      if (sm & DS_PAIRQ) { copy16(smem_raw + DS.pairq, T.pairq, deep_stage_bytes(m, DS_PAIRQ)); T.pairq = reinterpret_cast<const double4*>(smem_raw + DS.pairq); }
      if (sm & DS_SLICES) {
        copy16(smem_raw + DS.yy, T.yy, align16((size_t)nbR * sizeof(double))); T.yy = reinterpret_cast<const double*>(smem_raw + DS.yy);
        copy16(smem_raw + DS.wt, T.wt, align16((size_t)nbR * sizeof(double))); T.wt = reinterpret_cast<const double*>(smem_raw + DS.wt);
        copy16(smem_raw + DS.uu, T.uu, align16((size_t)nbR * sizeof(double))); T.uu = reinterpret_cast<const double*>(smem_raw + DS.uu);
        if (m.lp) {
          copy16(smem_raw + DS.lb, T.lb, align16((size_t)nbR * sizeof(double))); T.lb = reinterpret_cast<const double*>(smem_raw + DS.lb);
          copy16(smem_raw + DS.lce, T.lce, align16((size_t)nbR * sizeof(double2))); T.lce = reinterpret_cast<const double2*>(smem_raw + DS.lce);
        }
        copy16(smem_raw + DS.blk, T.blk, align16((size_t)(m.n_blocks + 1) * sizeof(BlockDesc))); T.blk = reinterpret_cast<const BlockDesc*>(smem_raw + DS.blk);
      }
      if (sm & DS_CHERRY) { copy16(smem_raw + DS.cherry, T.cherry, deep_stage_bytes(m, DS_CHERRY)); T.cherry = reinterpret_cast<const double2*>(smem_raw + DS.cherry); }
      if (sm & DS_LEAFAB) { copy16(smem_raw + DS.leafab, T.leafab, deep_stage_bytes(m, DS_LEAFAB)); T.leafab = reinterpret_cast<const double2*>(smem_raw + DS.leafab); }
      if (sm & DS_OPS) {
        copy16(smem_raw + DS.oprec, T.oprec, align16((size_t)m.n_ops * OPREC_DOUBLES * sizeof(double))); T.oprec = reinterpret_cast<const double*>(smem_raw + DS.oprec);
        copy16(smem_raw + DS.ops, T.ops, align16((size_t)(m.n_ops + 1) * sizeof(SweepOpDev))); T.ops = reinterpret_cast<const SweepOpDev*>(smem_raw + DS.ops);
      }
      __syncthreads();
    }
    while (true) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&next_group, 1);
      g = __shfl_sync(0xffffffffu, g, 0);
      if (b_begin + g >= b_end) break;
      const long long batch0 = (b_begin + g) * NI;
      if (WG) {
#pragma unroll
        for (int q = 0; q < NI; q++) mem[q].sw = m.words + (a.first + (batch0 + q) * 32);
      }
      double contrib[NI];
      sweep_items<R, NI, !STAGED, Slots, DIST, WG, HOT>(a, T, p, batch0, lane, mem, contrib, wrec);
      // fixed-order warp reduction -> partial[p][batch]; reduce_kernel sums the batches in a fixed order too
#pragma unroll
      for (int q = 0; q < NI; q++) {
        double c = contrib[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && batch0 + q < a.n_batches) a.partial[(size_t)p * a.n_batches + batch0 + q] = c;
      }
      __syncwarp();
    }
  }
  if (TMEM) {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(512));
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
  // one CTA per parameter vector; big trees with few parameter vectors want the wide CTA (5,000 genomes, P = 1: 0.77 ms with 128 threads)
  const int threads = (m.n_nodes >= 2048 && t.P <= 296) ? 512 : 128;
  prepass_nodes_kernel<<<t.P, threads, 0, s>>>(m, t);
}
void launch_prepass_tables(const ModelDev& m, const TablesDev& t, cudaStream_t s) {
  int n = m.n_entries + 1 + m.n_blocks * m.R + m.n_ops + m.n_leaves + m.n_pairs + m.n_cherries;
  dim3 g((n + 255) / 256, t.P);
  prepass_tables_kernel<<<g, 256, 0, s>>>(m, t);
}
// per warp: one bit per node + up to n_nodes 16-bit frontier ids (2.1 bytes per node: 139 KB at the 65,535-node limit the
// engine accepts).  Above 48 KB the kernel needs the opt-in attribute, set per device and instantiation.
template <int MODE>
static void walk_geometry(const ModelDev& m, long long count, int& wpc, size_t& smem, dim3& g) {
  const size_t per_warp = ((size_t)((m.n_nodes + 31) / 32) + (size_t)((m.n_nodes + 1) / 2)) * sizeof(uint32_t);
  const size_t budget = 200 * 1024;
  wpc = (int)std::min<size_t>(8, std::max<size_t>(1, (per_warp * 8 <= 48 * 1024 ? 48 * 1024 : budget) / per_warp));
  smem = per_warp * wpc;
  if (smem > 48 * 1024) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaFuncSetAttribute(class_walk_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  g = dim3((unsigned)((count + wpc - 1) / wpc));
}
void launch_frontier_count(const ModelDev& m, long long count, uint32_t* nf_out, cudaStream_t s) {
  if (count <= 0) return;
  int wpc; size_t smem; dim3 g;
  walk_geometry<0>(m, count, wpc, smem, g);
  class_walk_kernel<0><<<g, wpc * 32, smem, s>>>(m, TablesDev{}, count, nf_out, nullptr, nullptr);
}
void launch_frontier_fill(const ModelDev& m, long long count, const uint32_t* off, uint16_t* nodes_out, cudaStream_t s) {
  if (count <= 0) return;
  int wpc; size_t smem; dim3 g;
  walk_geometry<1>(m, count, wpc, smem, g);
  class_walk_kernel<1><<<g, wpc * 32, smem, s>>>(m, TablesDev{}, count, nullptr, off, nodes_out);
}
void launch_class_sums(const ModelDev& m, const TablesDev& t, long long first, long long count, cudaStream_t s) {
  (void)first;
  if (count <= 0) return;
  if (!m.fr_off) {  // no lists: walk per call
    int wpc; size_t smem; dim3 g;
    walk_geometry<2>(m, count, wpc, smem, g);
    class_walk_kernel<2><<<g, wpc * 32, smem, s>>>(m, t, count, nullptr, nullptr, nullptr);
    return;
  }
  const size_t tile = (size_t)m.n_nodes * 32 * sizeof(double2);
  if (tile <= 200 * 1024) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaFuncSetAttribute(class_sums_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int chunks = (t.P + 31) / 32;
    long long groups = std::max<long long>(1, (2LL * 148 + chunks - 1) / chunks);  // about two CTAs per SM in total
    long long per = std::max<long long>(8, (count + groups - 1) / groups);
    groups = (count + per - 1) / per;
    class_sums_tiled_kernel<<<dim3((unsigned)chunks, (unsigned)groups), 256, tile, s>>>(m, t, count, (int)per);
  } else {
    class_sums_list_kernel<<<(unsigned)((count + 7) / 8), 256, 0, s>>>(m, t, count);
  }
}
void launch_zero_row(const ModelDev& m, const TablesDev& t, cudaStream_t s) {
  zero_row_prefix_kernel<<<t.P, 256, 0, s>>>(m, t);
  zero_row_slices_kernel<<<dim3(t.P, (m.n_slices + 127) / 128), 128, 0, s>>>(m, t);
  zero_row_sum_kernel<<<t.P, 256, 0, s>>>(m, t);
}
void launch_prepass_i2(const ModelDev& m, const TablesDev& t, const double* i2_override, cudaStream_t s) {
  prepass_i2_kernel<<<(t.P + 63) / 64, 64, 0, s>>>(m, t, i2_override);
}

static const int kItemsPerWarp = 1;  // NI: batches swept simultaneously by one warp (2 measured slower: 5 warps/SM)

static int max_warps() {  // tuning knob: PPM_MAX_WARPS caps the warps per CTA (<= 16: 512 threads x 128 registers)
  const char* e = std::getenv("PPM_MAX_WARPS");
  int w = e ? std::atoi(e) : 16;
  return w < 1 ? 1 : (w > 16 ? 16 : w);
}

static bool tmem_enabled() { const char* e = std::getenv("PPM_NO_TMEM"); return !(e && std::atoi(e) != 0); }

SweepPlan plan_sweep(const ModelDev& m, long long n_batches, int P, int n_sm, bool dist) {
  SweepPlan pl{};
  const int NI = kItemsPerWarp;
  const size_t budget = 227 * 1024 - 1024;
  const StageLayout L = stage_layout(m);
  const size_t head = warp_head_bytes(m);
  const size_t slot_row = 32 * sizeof(double2);
  int warps = 0;
  pl.mode = 2;
  // MODE 0: lineage slots in shared memory -- preferred while at least 10 warps fit (measured on B200: 11 warps x 147
  // registers beat 16 warps x 128 registers with the slots in tensor memory by ~1 %)
  const size_t per_warp_staged = NI * (head + (size_t)m.n_slots * slot_row);
  int w0 = 0;
  if (L.total + per_warp_staged <= budget) w0 = (int)std::min<size_t>(std::min(max_warps(), 12), (budget - L.total) / per_warp_staged);
  // MODE 1: lineage slots in tensor memory (first nt per warp) + shared-memory overflow; most warps that fit
  int w1 = 0; size_t pw1 = 0;
  if (tmem_enabled()) {
    const char* t12 = std::getenv("PPM_TMEM12");
    for (int w = std::min(max_warps(), ((t12 && std::atoi(t12)) || dist) ? 12 : 16); w >= 8 && !w1; w -= 4) {
      const int nt = ((512 / ((w + 3) / 4)) & ~3) / 4 / NI;
      const size_t ovf = m.n_slots > nt ? (size_t)(m.n_slots - nt) : 0;
      if (L.total + (size_t)w * NI * (head + ovf * slot_row) <= budget) { w1 = w; pw1 = NI * (head + ovf * slot_row); }
    }
  }
  if (m.deep || m.R != 8) { w0 = 0; w1 = 0; }  // (R = 16 kernels exist for the global-memory variant only)  // a program built for the global-memory variant (padded segments, rows read in place) stays there
  const long long groups = (n_batches + NI - 1) / NI;  // warp-level work items per parameter vector
  const long long total = groups * (long long)P;
  const char* force = std::getenv("PPM_FORCE_TMEM");
  const bool tiny = w0 >= 1 && groups <= w0 && (w0 >= 3 || w1);  // e.g. the all-zero row: one small CTA per parameter vector
  // (only where the big sweep is staged too: both run the same, non-deep, program)
  if (tiny) { warps = (int)groups; pl.mode = 0; pl.per_warp_bytes = per_warp_staged; }
  else if (dist && w0 >= 1 && (w0 >= 3 || w1)) { warps = w0; pl.mode = 0; pl.per_warp_bytes = per_warp_staged; }  // DIST kernels exist for modes 0, 3, 4
  else if (w1 && ((force && std::atoi(force)) || w0 < 10)) { warps = w1; pl.mode = w1 <= 12 ? 3 : 1; pl.per_warp_bytes = pw1; }
  else if (w0 >= 3) { warps = w0; pl.mode = 0; pl.per_warp_bytes = per_warp_staged; }
  pl.staged = pl.mode != 2;
  if (!pl.staged) {
    pl.per_warp_bytes = NI * head + kRecRingBytes;
    const char* d16 = std::getenv("PPM_DEEP16");
    warps = (int)std::max<size_t>(1, std::min<size_t>(std::min(max_warps(), (d16 && std::atoi(d16) && !dist && !m.word_stride) ? 16 : 12), budget / pl.per_warp_bytes));
    if (warps <= 12) pl.mode = 4;
    if (groups < warps) warps = (int)std::max<long long>(1, groups);
    // whatever small tables still fit are staged in shared memory per parameter vector (the records stay in global
    // memory and stream through the per-warp rings)
    const char* ns = std::getenv("PPM_NO_DEEP_STAGE");
    size_t left = budget - pl.per_warp_bytes * warps;
    if (!(ns && std::atoi(ns)))
      for (int bit : {DS_SLICES, DS_PAIRQ, DS_CHERRY, DS_LEAFAB, DS_OPS}) {
        const size_t need = deep_stage_bytes(m, bit);
        if (need <= left) { pl.stage_mask |= bit; left -= need; }
      }
  }
  pl.warps = warps;
  // persistent CTAs over equal contiguous ranges of the (parameter vector, batch) line; when a parameter vector has
  // fewer items than a CTA has warps, one (small) CTA per parameter vector instead
  if (groups <= warps) pl.grid_x = P;
  else pl.grid_x = (int)std::max<long long>(1, std::min<long long>(n_sm, (total + warps - 1) / warps));
  pl.items_per_warp = (int)((total + (long long)pl.grid_x * warps - 1) / ((long long)pl.grid_x * warps));  // informational
  pl.n_chunks = 1;
  if (std::getenv("PPM_PLAN_DEBUG"))
    fprintf(stderr, "[ppm plan] tables %zu B, head %zu B, slots %d, per warp %zu B, w0 %d, w1 %d -> mode %d, %d warps, grid %d, deep stage mask %d (budget %zu)\n",
            L.total, head, m.n_slots, pl.per_warp_bytes, w0, w1, pl.mode, warps, pl.grid_x, pl.stage_mask, budget);
  pl.smem_bytes = (pl.staged ? L.total : deep_stage_layout(m, pl.stage_mask).total) + pl.per_warp_bytes * warps;
  return pl;
}
bool sweep_is_staged(const ModelDev& m) { return plan_sweep(m, 1 << 20, 1, 148).staged; }
size_t sweep_scratch_elems(const ModelDev& m, const SweepPlan& pl) {
  return pl.staged ? 0 : (size_t)pl.grid_x * pl.warps * kItemsPerWarp * (m.n_slots + 1) * 32;
}

template <int R, int MODE, bool DIST = false, bool WG = false, bool HOT = false>
static void launch_sweep_t(const SweepArgs& a, const SweepPlan& pl, cudaStream_t s) {
  static bool configured[64] = {};  // the attribute is per device (group handles launch on several)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(sweep_kernel<R, kItemsPerWarp, MODE, DIST, WG, HOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);  // static shared memory takes a few bytes
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  sweep_kernel<R, kItemsPerWarp, MODE, DIST, WG, HOT><<<pl.grid_x, pl.warps * 32, pl.smem_bytes, s>>>(a);
}

template <int R>
static void launch_sweep_r(const SweepArgs& a, const SweepPlan& pl, cudaStream_t s) {
  if (pl.mode == 1) launch_sweep_t<R, 1>(a, pl, s);
  else if (pl.mode == 3) launch_sweep_t<R, 3>(a, pl, s);
  else if (pl.mode == 0) launch_sweep_t<R, 0>(a, pl, s);
  else if (pl.mode == 4) launch_sweep_t<R, 4>(a, pl, s);
  else launch_sweep_t<R, 2>(a, pl, s);
}

void launch_sweep(SweepArgs a, const SweepPlan& pl, cudaStream_t s) {
  a.n_chunks = pl.n_chunks; a.items_per_warp = pl.items_per_warp; a.stage_mask = pl.stage_mask;
  if (a.m.word_stride) {  // pattern rows in place: R = 8, mode 4 (the engine sets word_stride only then)
    if (a.dist) launch_sweep_t<8, 4, true, true>(a, pl, s); else launch_sweep_t<8, 4, false, true>(a, pl, s);
    return;
  }
  if (a.dist) {  // per-slice integrand output (expectedTime2): R = 8 programs, modes 0 / 3 / 4 (plan_sweep(dist = true))
    if (pl.mode == 0) launch_sweep_t<8, 0, true>(a, pl, s);
    else if (pl.mode == 3) launch_sweep_t<8, 3, true>(a, pl, s);
    else launch_sweep_t<8, 4, true>(a, pl, s);
    return;
  }
  if (a.m.lean && a.m.lp && a.m.R == 8 && a.mode == 0 && pl.mode != 2) {
    // the objective on a lean leaf-power program (the usual case: ultrametric tree): HOT kernel, every parameter vector
    // (leaf powers need no factored form: nothing is expanded); fact_filter 3 (the band kernel's leftovers) is kept
    if (pl.mode == 0) launch_sweep_t<8, 0, false, false, true>(a, pl, s);
    else if (pl.mode == 1) launch_sweep_t<8, 1, false, false, true>(a, pl, s);
    else if (pl.mode == 3) launch_sweep_t<8, 3, false, false, true>(a, pl, s);
    else launch_sweep_t<8, 4, false, false, true>(a, pl, s);
    return;
  }
  if (a.m.R == 16) {  // tuning knob (PPM_R=16): global-memory variant only
    if (pl.mode == 4) launch_sweep_t<16, 4>(a, pl, s); else launch_sweep_t<16, 2>(a, pl, s);
  } else launch_sweep_r<8>(a, pl, s);
}
void launch_reduce(const double* partial, int P, int nx, double* out, cudaStream_t s) {
  reduce_kernel<<<P, 256, 0, s>>>(partial, nx, out);
}
__global__ void gather_kernel(const double* table, const int* index, long long n, double* out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = table[index[j]];
}
void launch_gather(const double* table, const int* index, long long n, double* out, cudaStream_t s) {
  if (n > 0) gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(table, index, n, out);
}
// group evaluation: out[p] = sum over the n shards, in rank order, of gather[k][p]
__global__ void sum_shards_kernel(const double* gather, int n, int P, double* out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double s = 0.;
  for (int k = 0; k < n; k++) s += gather[(size_t)k * P + p];
  out[p] = s;
}
void launch_sum_shards(const double* gather, int n, int P, double* out, cudaStream_t s) {
  sum_shards_kernel<<<(P + 127) / 128, 128, 0, s>>>(gather, n, P, out);
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
