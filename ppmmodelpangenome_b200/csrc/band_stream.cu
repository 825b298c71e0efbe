// Band kernels, stream layout: the slice-parallel sweep for big trees as THREE kernels per batch of items (sm_100a, FP64
// CUDA cores).  An item is one (pattern, parameter vector); a batch is a contiguous range of items, 32-item blocks.
//
//   band_prune_kernel     LANE = ITEM.  Each thread prunes its item's Mobile category (objects.cpp:427-435, linear domain)
//                         op by op in birth order -- the op list is topology only, so the warp never diverges and no level
//                         barrier exists (a 10,000-level caterpillar costs what a balanced tree costs) -- and leaves one
//                         16-byte record (a, -d e^{lam (h - H/2)}) per internal lineage in a global scratch laid out
//                         [block][lineage][32 items] (coalesced), plus the running sum of its renormalisation shifts.
//   band_integral_kernel  LANES = SLICES of the Mobile integral (functions.cpp:245-279).  A warp owns one work unit = one
//                         item x one of the program's G descriptor lists (whole tiles of 32*R slices, balanced by the
//                         builder); units are handed out item-major from a global counter, so the lineage records of the
//                         items in flight stay in L2.  Per chunk of <= 32 records: entries -> cp.async gather (lineage
//                         records from the scratch, leaf / leaf-pair records from the per-parameter tables by pattern bit)
//                         into a per-warp ring, BS_D chunks ahead of their use; then one broadcast LDS.128 per record and
//                         R x (DFMA + DMUL) with R independent dependency chains (leaf pairs: one quadratic in u, 3 FP64 per
//                         slice for two terms).  No CTA-level synchronisation at all: warps are independent.
//   band_mixture_kernel   lane = item: sum of the item's G unit integrals in list order (exponent-aligned), 3-way mixture
//                         (functions.cpp:345-358), x count, fixed-order warp sum -> partial[p][32-pattern block].
// Results do not depend on the dynamic unit hand-out: every sum runs in a fixed order.
#include "band_fns.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace ppm {

// Per warp: a ring of BS_NB record slots (64 records x 32 B or 128 x 16 B + 512 B of lane masks / tile tables) and a ring of BS_NB block
// slots (a descriptor's entries + code, band_codes.h).  Everything the walk needs from global memory arrives by cp.async:
// iteration di issues the block of descriptor di + 4 and -- from the block that has arrived by then -- the record gather of
// descriptor di + 2, waits for the group issued two iterations earlier and multiplies in descriptor di.  The software-pipelined
// FULL loops read one group of records past a slot (BS_PAD).
enum { BS_NB = 3, BS_SLOT_BYTES = kBandStreamChunk * 32 + 512, BS_MASK_OFF = kBandStreamChunk * 32, BS_PAD = 256,
       BS_RING_BYTES = BS_NB * BS_SLOT_BYTES + BS_PAD, BS_WARPS = 4, BS_NOP = 15 };
// a slot holds 64 records of 32 bytes or 128 of 16 (leaf-power programs), then 512 bytes of lane masks
static_assert(kBandStreamChunkLP * 16 == BS_MASK_OFF && kBandStreamChunkLP * 4 <= 512, "a leaf-power chunk fills a slot exactly");
static __host__ __device__ inline int bs_chunk(bool lp) { return lp ? (int)kBandStreamChunkLP : (int)kBandStreamChunk; }
static __host__ __device__ inline int bs_blk_bytes(bool lp) { return (bs_chunk(lp) + 4) * 4; }

static __host__ __device__ inline int bs_pw_words(int n_words) { return (n_words + 1 + 3) & ~3; }
// (a leaf-power program reads no pattern words)
static __host__ __device__ inline size_t bs_warp_bytes(int n_words, bool lp) { return (size_t)BS_RING_BYTES + BS_NB * bs_blk_bytes(lp) + (lp ? 0 : 4 * (size_t)bs_pw_words(n_words)); }

__device__ __forceinline__ void bs_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

__device__ __forceinline__ bool bs_served(const double* sc, int force) {
  return sc[SC_BAND] != 0. && (force || !(sc[SC_I2] < 0.));  // i2 < 0: penalty branch, no pattern work (functions.cpp:362-364)
}

// ------------------------------------------------------------------------------------------------ prune (lane = item)
// One thread walks its item's ops in birth order.  The walk is a chain of global loads (a child's record, a leaf's pattern
// word) in front of ~60 cycles of FP64 each, so it is software-pipelined: the operands of op i + BP_PF are requested while op i is
// computed -- pattern words always, a child's record when it was written at least BP_PF ops ago; younger children are
// forwarded from registers (a caterpillar's chain never touches memory).
// loads whose destination registers are the pipeline slot itself (plain C++ loads went through a temporary quad and were
// copied out at once -- a copy that waits for the load)
__device__ __forceinline__ void bp_ldg(double2& d, const double2* p) {
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(d.x), "=d"(d.y) : "l"(p) : "memory");
}
__device__ __forceinline__ void bp_ldg(uint32_t& d, const uint32_t* p) {
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(d) : "l"(p) : "memory");
}
enum { BP_PF = 3, BP_OC = 2049 };  // request distance; op records staged per CTA in chunks of BP_OC (+ 8 of lookahead)
static_assert(BP_OC % BP_PF == 0, "slots keep their phase across chunks");
__global__ void __launch_bounds__(32 * BS_WARPS) band_prune_kernel(BandStreamArgs a) {
  __shared__ uint2 ops_s[BP_OC + 8];  // op records as raw words: lref | rref << 16, dst | flags << 16
  // branch constants of the warp's parameter vector, 32 ops at a time, double-buffered by cp.async (4 x 16 bytes per op)
  __shared__ __align__(16) double2 cst_s[BS_WARPS][2][32 * 4];
  const int lane = threadIdx.x & 31;
  const int blk = blockIdx.x * BS_WARPS + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.counter = 0u;  // the integral kernel of this batch follows on the same stream
  const long long item = a.item0 + 32ll * (blk < a.n_blk ? blk : 0);
  const int p = (int)(item / a.count_pad);
  const long long pat = item - (long long)p * a.count_pad + lane;  // < n_pat_pad: the word rows are zero-padded
  const double* sc = a.t.scal + (size_t)p * SC_COUNT;
  const bool active = blk < a.n_blk && bs_served(sc, a.force);  // (idle warps still stage op records and meet the barriers)
  const double pi1 = sc[SC_PI1];
  const double2 leaf0 = make_double2(sc[SC_LA0], sc[SC_LD0]), leaf1 = make_double2(sc[SC_LA1], sc[SC_LD1]);
  const int n_ops = a.b.n_ops, n_int = a.b.n_int, last = n_ops - 1;
  const uint2* ops2 = reinterpret_cast<const uint2*>(a.b.ops);
  const double2* bop = reinterpret_cast<const double2*>(a.bt.bop + (size_t)p * n_ops * BOP_DOUBLES);
  const uint32_t* words = a.m.words + pat;
  const size_t wstride = (size_t)a.m.n_pat_pad;
  double2* nodes = a.nodes + (size_t)(active ? blk : 0) * n_int * 32 + lane;
  int* xk = a.xk + (size_t)(active ? blk : 0) * (n_int + 1) * 32 + lane;
  if (active) xk[0] = 0;
  int run = 0;
  // leaf powers: the present leaves that have died before each op (and the sum of their heights) travel with the shift prefixes
  const bool lp = a.b.lp != 0, hs1 = a.b.has_s1 != 0;
  int* n1k = a.n1k + (size_t)(active ? blk : 0) * (n_int + 1) * 32 + lane;
  double* s1k = a.s1k + (size_t)(active && hs1 ? blk : 0) * (n_int + 1) * 32 + lane;
  int n1d = 0;
  double s1d = 0.;
  if (active && lp) { n1k[0] = 0; if (hs1) s1k[0] = 0.; }

  // circular buffer of BP_PF slots with compile-time indices (the loop body is unrolled BP_PF times): slot s holds the
  // requested operands of op j (j % BP_PF == s); once op j has read them the slot is refilled for op j + BP_PF.  Nothing that
  // is still in flight is ever moved between registers.
  uint2 so[BP_PF];                     // op record of the slot's op
  uint32_t wl[BP_PF], wr[BP_PF];       // requested operands: pattern word of a leaf child ...
  double2 nl[BP_PF], nr[BP_PF];        // ... or the record of an internal child written at least BP_PF ops before the slot's op
  double2 prev[BP_PF];                 // prev[k] = record of op j - 1 - k (register forwarding of the youngest records)
#pragma unroll
  for (int k = 0; k < BP_PF; k++) {
    prev[k] = make_double2(1., 0.);
    so[k] = make_uint2(0u, 0u);
    wl[k] = 0u; wr[k] = 0u; nl[k] = make_double2(1., 0.); nr[k] = make_double2(1., 0.);
  }
  auto request = [&](int j, const uint2 oj, uint32_t& w_l, uint32_t& w_r, double2& n_l, double2& n_r) {
    const int lref = (int)(oj.x & 0xffffu), rref = (int)(oj.x >> 16), flags = (int)(oj.y >> 16);
    if (flags & 1) bp_ldg(w_l, words + (size_t)(lref >> 5) * wstride);
    else if (lref < j - BP_PF) bp_ldg(n_l, nodes + (size_t)lref * 32);
    if (flags & 2) bp_ldg(w_r, words + (size_t)(rref >> 5) * wstride);
    else if (rref < j - BP_PF) bp_ldg(n_r, nodes + (size_t)rref * 32);
  };
  double2* cst = &cst_s[threadIdx.x >> 5][0][0];
  auto stage_consts = [&](int chunk) {  // ops [32 chunk, 32 chunk + 32) of this warp's parameter vector -> buffer chunk & 1
    const int j = 32 * chunk + lane;
    if (j < n_ops) {
      double2* dst = cst + (chunk & 1) * 128 + 4 * lane;
      const double2* src = bop + 4 * (size_t)j;
      bs_cp16(dst, src); bs_cp16(dst + 1, src + 1); bs_cp16(dst + 2, src + 2);
      if (lp) bs_cp16(dst + 3, src + 3);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (active) stage_consts(0);
  for (int c0 = 0; c0 < n_ops; c0 += BP_OC) {
    __syncthreads();
    for (int x = threadIdx.x; x < BP_OC + 8; x += blockDim.x) ops_s[x] = ops2[c0 + x < last ? c0 + x : last];
    __syncthreads();
    if (!active) continue;
    if (c0 == 0) {
#pragma unroll
      for (int k = 0; k < BP_PF; k++) {
        const int j = k < last ? k : last;
        so[k] = ops_s[j];
        request(k, so[k], wl[k], wr[k], nl[k], nr[k]);
      }
    }
    const int c1 = c0 + BP_OC < n_ops ? c0 + BP_OC : n_ops;  // (BP_OC is a multiple of BP_PF: slots keep their phase)
    for (int i = c0; i < c1; i += BP_PF) {
#pragma unroll
      for (int sl = 0; sl < BP_PF; sl++) {
        const int j = i + sl;
        if (j < c1) {
          if ((j & 31) == 0) {  // this chunk's constants have landed; the next chunk's are requested
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            stage_consts((j >> 5) + 1);
          }
          const double2* cj = cst + ((j >> 5) & 1) * 128 + 4 * (j & 31);
          const double2 k01 = cj[0], k23 = cj[1], k4 = cj[2];
          // operands of op j out of the slot
          const int lref = (int)(so[sl].x & 0xffffu), rref = (int)(so[sl].x >> 16), flags = (int)(so[sl].y >> 16);
          double2 Lc, Rc;
          if (flags & 1) {
            const bool b1 = ((wl[sl] >> (lref & 31)) & 1u) != 0u;
            Lc = b1 ? leaf1 : leaf0;
            if (lp) { n1d += b1 ? 1 : 0; if (hs1) s1d += b1 ? k4.y : 0.; }
          } else {
            Lc = nl[sl];
#pragma unroll
            for (int k = 0; k < BP_PF; k++) if (lref == j - 1 - k) Lc = prev[k];
          }
          if (flags & 2) {
            const bool b1 = ((wr[sl] >> (rref & 31)) & 1u) != 0u;
            Rc = b1 ? leaf1 : leaf0;
            if (lp) { n1d += b1 ? 1 : 0; if (hs1) s1d += b1 ? cj[3].x : 0.; }
          } else {
            Rc = nr[sl];
#pragma unroll
            for (int k = 0; k < BP_PF; k++) if (rref == j - 1 - k) Rc = prev[k];
          }
          const double occ[5] = {k01.x, k01.y, k23.x, k23.y, k4.x};
          // refill the slot for op j + BP_PF (records up to j - 1 are written; the request takes those up to j - 1 only)
          {
            so[sl] = ops_s[j + BP_PF - c0];  // (clamped to the last op when staged)
            request(j + BP_PF, so[sl], wl[sl], wr[sl], nl[sl], nr[sl]);
          }
          int k;
          const double2 rec = band_merge(Lc.x, Lc.y, Rc.x, Rc.y, occ, pi1, k);
          nodes[(size_t)j * 32] = rec;  // stream layout: op j writes record j (birth order)
          run += k;
          xk[(size_t)(j + 1) * 32] = run;
          if (lp) { n1k[(size_t)(j + 1) * 32] = n1d; if (hs1) s1k[(size_t)(j + 1) * 32] = s1d; }
#pragma unroll
          for (int q = BP_PF - 1; q > 0; q--) prev[q] = prev[q - 1];
          prev[0] = rec;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ integral (lanes = slices)
__device__ __forceinline__ void bs_cp8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void bs_cp4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// t *= f on the lanes of the mask only (a predicated DMUL: no select instructions)
__device__ __forceinline__ void bs_pmul(double& t, double f, uint32_t mask, uint32_t lanebit) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 x;\n\tand.b32 x, %2, %3;\n\tsetp.ne.u32 p, x, 0;\n\t@p mul.f64 %0, %0, %1;\n\t}" : "+d"(t) : "d"(f), "r"(mask), "r"(lanebit));
}
template <int R>
__device__ __forceinline__ void bs_rescale(double (&prod)[R], int (&ex)[R]) {
  int mn = __double2hiint(prod[0]);
#pragma unroll
  for (int k = 1; k < R; k++) { const int h = __double2hiint(prod[k]); mn = h < mn ? h : mn; }
  if (mn < ((1023 - 300) << 20)) {  // rare: some product of this lane has drifted below 2^-300 (or is an exact zero)
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int be = biased_exp(prod[k]);
      if (be != 0 && be < 1023 - 300) { prod[k] *= pow2i(300); ex[k] += 300; }
    }
  }
}

// FULL chunks: every record multiplies all R registers of the lane.  Eight records per iteration in two register sets (A, B):
// set B is loaded while A is multiplied in and vice versa (no register rotation); the reads run one group past the chunk.
template <int R>
__device__ __forceinline__ void bs_pairs_full(const double4* rb, int cnt, const double (&U)[R], double (&prod)[R], int (&ex)[R]) {
  double4 A0 = rb[0], A1 = rb[1], A2 = rb[2], A3 = rb[3];
#pragma unroll 1
  for (int i = 0; i < cnt; i += 8) {
    const double4 B0 = rb[i + 4], B1 = rb[i + 5], B2 = rb[i + 6], B3 = rb[i + 7];
#pragma unroll
    for (int k = 0; k < R; k++) {
      prod[k] *= fma(fma(A0.z, U[k], A0.y), U[k], A0.x);
      prod[k] *= fma(fma(A1.z, U[k], A1.y), U[k], A1.x);
      prod[k] *= fma(fma(A2.z, U[k], A2.y), U[k], A2.x);
      prod[k] *= fma(fma(A3.z, U[k], A3.y), U[k], A3.x);
    }
    A0 = rb[i + 8]; A1 = rb[i + 9]; A2 = rb[i + 10]; A3 = rb[i + 11];
#pragma unroll
    for (int k = 0; k < R; k++) {
      prod[k] *= fma(fma(B0.z, U[k], B0.y), U[k], B0.x);
      prod[k] *= fma(fma(B1.z, U[k], B1.y), U[k], B1.x);
      prod[k] *= fma(fma(B2.z, U[k], B2.y), U[k], B2.x);
      prod[k] *= fma(fma(B3.z, U[k], B3.y), U[k], B3.x);
    }
    if ((i & 8) && i + 8 < cnt) bs_rescale<R>(prod, ex);  // every 16 records (the chunk's end checks too)
  }
}
template <int R>
__device__ __forceinline__ void bs_recs_full(const double2* rb, int cnt, const double (&Y)[R], double (&prod)[R], int (&ex)[R]) {
  double2 A0 = rb[0], A1 = rb[1], A2 = rb[2], A3 = rb[3];
#pragma unroll 1
  for (int i = 0; i < cnt; i += 8) {
    const double2 B0 = rb[i + 4], B1 = rb[i + 5], B2 = rb[i + 6], B3 = rb[i + 7];
#pragma unroll
    for (int k = 0; k < R; k++) {
      prod[k] *= fma(A0.y, Y[k], A0.x);
      prod[k] *= fma(A1.y, Y[k], A1.x);
      prod[k] *= fma(A2.y, Y[k], A2.x);
      prod[k] *= fma(A3.y, Y[k], A3.x);
    }
    A0 = rb[i + 8]; A1 = rb[i + 9]; A2 = rb[i + 10]; A3 = rb[i + 11];
#pragma unroll
    for (int k = 0; k < R; k++) {
      prod[k] *= fma(B0.y, Y[k], B0.x);
      prod[k] *= fma(B1.y, Y[k], B1.x);
      prod[k] *= fma(B2.y, Y[k], B2.x);
      prod[k] *= fma(B3.y, Y[k], B3.x);
    }
    if ((i & 8) && i + 8 < cnt) bs_rescale<R>(prod, ex);
  }
}
// per-register chunks: the records multiply one register of the lane; four independent chains
__device__ __forceinline__ double bs_pairs_one(const double4* rb, int cnt, double u) {
  double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
#pragma unroll 2
  for (int i = 0; i < cnt; i += 4) {
    const double4 q0 = rb[i], q1 = rb[i + 1], q2 = rb[i + 2], q3 = rb[i + 3];
    t0 *= fma(fma(q0.z, u, q0.y), u, q0.x);
    t1 *= fma(fma(q1.z, u, q1.y), u, q1.x);
    t2 *= fma(fma(q2.z, u, q2.y), u, q2.x);
    t3 *= fma(fma(q3.z, u, q3.y), u, q3.x);
  }
  return (t0 * t1) * (t2 * t3);
}
__device__ __forceinline__ double bs_recs_one(const double2* rb, int cnt, double y) {
  double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
#pragma unroll 2
  for (int i = 0; i < cnt; i += 4) {
    const double2 r0 = rb[i], r1 = rb[i + 1], r2 = rb[i + 2], r3 = rb[i + 3];
    t0 *= fma(r0.y, y, r0.x);
    t1 *= fma(r1.y, y, r1.x);
    t2 *= fma(r2.y, y, r2.x);
    t3 *= fma(r3.y, y, r3.x);
  }
  return (t0 * t1) * (t2 * t3);
}
__device__ __forceinline__ double bs_recs_masked(const double2* rb, const uint32_t* mb, int cnt, double y, uint32_t lanebit) {
  double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
#pragma unroll 2
  for (int i = 0; i < cnt; i += 4) {
    const double2 r0 = rb[i], r1 = rb[i + 1], r2 = rb[i + 2], r3 = rb[i + 3];
    const uint4 mk = *reinterpret_cast<const uint4*>(mb + i);
    bs_pmul(t0, fma(r0.y, y, r0.x), mk.x, lanebit);
    bs_pmul(t1, fma(r1.y, y, r1.x), mk.y, lanebit);
    bs_pmul(t2, fma(r2.y, y, r2.x), mk.z, lanebit);
    bs_pmul(t3, fma(r3.y, y, r3.x), mk.w, lanebit);
  }
  return (t0 * t1) * (t2 * t3);
}

// LP: the instantiation for leaf-power programs (band_program.h): no leaf or leaf-pair records, hence no pattern words, no
// leaf tables, no u table -- the paths that gather and multiply them are compiled out
template <int R, bool LP>
// (6 CTAs per SM at 80 registers, which the leaf-power instantiation's smaller shared memory allows, measured: configs[3] 556 -> 615 ms,
// configs[4] 112 -> 149 ms -- more items in flight than L2 holds records for)
__global__ void __launch_bounds__(32 * BS_WARPS, 5) band_integral_kernel(BandStreamArgs a) {
  static_assert(R <= 4, "a tile-begin block carries four shift indices per lane");
  extern __shared__ __align__(16) unsigned char bs_smem[];
  const ModelDev& m = a.m;
  const BandDev& b = a.b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wring = bs_smem + (size_t)warp * bs_warp_bytes(m.n_words, LP);
  unsigned char* bring = wring + BS_RING_BYTES;
  constexpr int CH = LP ? (int)kBandStreamChunkLP : (int)kBandStreamChunk;  // records per chunk
  constexpr int NPL = CH / 32;                                               // ... and per lane
  constexpr int BLK_BYTES = (CH + 4) * 4;
  uint32_t* pw = reinterpret_cast<uint32_t*>(bring + BS_NB * BLK_BYTES);
  const uint32_t lanebit = 1u << lane;
  const int T = 32 * R, G = b.G, n_int = b.n_int;
  const unsigned total = (unsigned)a.n_blk * 32u * (unsigned)G;

  unsigned u = 0;
  if (lane == 0) u = atomicAdd(a.counter, 1u);
  u = __shfl_sync(0xffffffffu, u, 0);
  while (u < total) {
    unsigned u_next = 0;
    if (lane == 0) u_next = atomicAdd(a.counter, 1u);  // its latency hides behind this unit
    const int il = (int)(u / (unsigned)G), w = (int)(u - (unsigned)il * (unsigned)G);
    const long long item = a.item0 + il;
    const int p = (int)(item / a.count_pad);
    const long long pat = item - (long long)p * a.count_pad;
    const double* sc = a.t.scal + (size_t)p * SC_COUNT;
    // (sparse evaluation: the items with term lists belong to band_sparse_kernel)
    // (slice_out: the baseline rows of the sparse evaluation -- only copy 0 of the all-ones row is integrated)
    const bool mine = a.slice_out ? pat == 32 : (pat < a.count && !(a.sparse && bs_sparse_use(m, a.lists, b.n_tiles, a.sparse_max_len, pat) != 0));
    if (mine && bs_served(sc, a.force)) {
      const double2* nodes = a.nodes + (size_t)(il >> 5) * n_int * 32 + (il & 31);
      const int* xk = a.xk + (size_t)(il >> 5) * (n_int + 1) * 32 + (il & 31);
      const double* bY = a.bt.bY + (size_t)p * b.S_pad;
      const double* bU = a.bt.bU + (size_t)p * b.S_pad;
      const double2* bleaf = a.bt.bleaf + (size_t)p * m.n_leaves * 2;
      const double4* pairq = a.t.pairq + (size_t)p * m.n_pairs * 4;
      constexpr bool lp = LP;
      const int* n1k = a.n1k + (size_t)(il >> 5) * (n_int + 1) * 32 + (il & 31);
      const double* s1k = a.s1k + (size_t)(b.has_s1 ? (il >> 5) : 0) * (n_int + 1) * 32 + (il & 31);
      const double* bLb = a.bt.bLb + (size_t)p * b.S_pad;
      const double* bLc = a.bt.bLc + (size_t)p * b.S_pad;
      const double* bLe = a.bt.bLe + (size_t)p * b.S_pad;
      int n1_0 = 0;
      double s1_0 = 0.;
      if (lp) { n1_0 = a.n1_0[pat]; if (b.has_s1) s1_0 = a.s1_0[pat]; }
      __syncwarp();  // (the previous unit's reads of pw are done)
      if (!lp) {     // (a leaf-power program gathers no leaf records: no pattern bits needed)
        for (int x = lane; x < m.n_words; x += 32) pw[x] = m.words[(size_t)x * m.n_pat_pad + pat];
        if (lane == 0) pw[m.n_words] = 0u;
      }

      const unsigned char* blocks = reinterpret_cast<const unsigned char*>(b.blocks) + (size_t)b.unit_blk_off[w] * BLK_BYTES;
      double prod[R], Y[R], U[R], lpacc[R];
      int ex[R];
#pragma unroll
      for (int k = 0; k < R; k++) { prod[k] = 1.; Y[k] = 0.; U[k] = 0.; ex[k] = 0; lpacc[k] = 0.; }
      double Im = 0.; int Ie = 0x3fffffff;
      int tau = 0;
      uint32_t c0 = BS_NOP, c1 = BS_NOP;  // codes of descriptors di (compute stage) and di + 1
      int s0 = 0, s2 = 2, s4 = 1;         // ring slots of descriptors di, di + 2 (gather stage), di + 4 (block stage): index mod BS_NB
      static_assert(BS_NB == 3, "slot arithmetic below");
      for (int di = -4;; di++) {
        // the group issued two iterations ago has landed: records of descriptor di, block of descriptor di + 2
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        // ---- block stage: the block of descriptor di + 4
#pragma unroll
        for (int x = lane; x < BLK_BYTES / 16; x += 32) bs_cp16(bring + s4 * BLK_BYTES + 16 * x, blocks + (size_t)(di + 4) * BLK_BYTES + 16 * x);
        // ---- gather stage: the records of descriptor di + 2 into their slot (records 2 lane and 2 lane + 1)
        uint32_t c2 = BS_NOP;
        if (di >= -2) {
          const uint32_t* blk = reinterpret_cast<const uint32_t*>(bring + s2 * BLK_BYTES);
          c2 = blk[CH];
          const uint2 e2 = *reinterpret_cast<const uint2*>(blk + 2 * lane);  // (tile begin: shift-prefix indices of the lane's slices)
          const uint32_t opG = c2 & 15u;
          unsigned char* slot = wring + s2 * BS_SLOT_BYTES;
          if (opG == BD_REC) {
            const bool masks = ((c2 >> 4) & 3u) == BV_M;
            // record h * 32 + lane; a short list (boundary lists hold ~60 records, a caterpillar's ~30) skips the empty rounds.
            // The compute loops read whole groups of 8 / 4 records: the pad entries up to the next multiple of 8 become identities.
            int cnt2 = (int)((c2 >> 9) & 127u);
            if (LP && cnt2 == 0) cnt2 = 128;
            const int cnt2r = (cnt2 + 7) & ~7;
#pragma unroll
            for (int h = 0; h < NPL; h++) {
              if (h * 32 >= cnt2r) break;
              const int x = h * 32 + lane;
              const uint32_t q = blk[x];
              const bool pad = q == kBandPadEntry;
              const uint32_t ref = pad ? 0u : (q & 0xffffu);
              const double2* src = nodes + (size_t)ref * 32;
              if (!LP && (q >> 31)) {
                const uint32_t bt1 = (pw[ref >> 5] >> (ref & 31u)) & 1u;
                src = bleaf + ref * 2 + bt1;
              }
              double2* dst = reinterpret_cast<double2*>(slot) + x;
              if (pad) *dst = make_double2(1., 0.);  // (never from global memory: one hot sector would serialise the whole GPU)
              else bs_cp16(dst, src);
              if (masks) {
                const uint32_t lo = (q >> 16) & 31u, len = ((q >> 21) & 63u) - lo;
                reinterpret_cast<uint32_t*>(slot + BS_MASK_OFF)[x] = pad ? 0u : ((len >= 32u ? 0xffffffffu : ((1u << len) - 1u)) << lo);
              }
            }
          } else if (!LP && opG == BD_PAIR) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const uint32_t q = h ? e2.y : e2.x;
              const bool pad = q == kBandPadEntry;
              const uint32_t ref = pad ? 0u : (q & 0xffffu);
              const uint32_t bits = (pw[ref >> 4] >> (2u * (ref & 15u))) & 3u;
              const double2* src = reinterpret_cast<const double2*>(pairq + ref * 4 + bits);
              double2* dst = reinterpret_cast<double2*>(slot) + 2 * (2 * lane + h);
              if (pad) { dst[0] = make_double2(1., 0.); dst[1] = make_double2(0., 0.); }
              else { bs_cp16(dst, src); bs_cp16(dst + 1, src + 1); }
            }
          } else if (opG == BD_TILE_BEGIN) {  // the tile's slice tables: Y' | u | exponent prefix, [k][lane] each
            const int tb = (int)(c2 >> 16) * T + lane;
#pragma unroll
            for (int k = 0; k < R; k++) {
              const uint32_t nb = ((k < 2 ? e2.x : e2.y) >> (16 * (k & 1))) & 0xffffu;
              bs_cp8(slot + (k * 32 + lane) * 8, bY + tb + 32 * k);
              if (!LP) bs_cp8(slot + 1024 + (k * 32 + lane) * 8, bU + tb + 32 * k);
              bs_cp4(slot + 2048 + (k * 32 + lane) * 4, xk + (size_t)nb * 32);
            }
          } else if (LP && opG == BD_TILE_LP) {  // leaf powers: count tables | dead present leaves (variant 0), height tables (variant 1)
            const int tb = (int)(c2 >> 16) * T + lane;
            const bool heights = ((c2 >> 4) & 3u) != 0u;
#pragma unroll
            for (int k = 0; k < R; k++) {
              const uint32_t nb = ((k < 2 ? e2.x : e2.y) >> (16 * (k & 1))) & 0xffffu;
              if (heights) {
                bs_cp8(slot + (k * 32 + lane) * 8, bLe + tb + 32 * k);
                bs_cp8(slot + 1024 + (k * 32 + lane) * 8, s1k + (size_t)nb * 32);
              } else {
                bs_cp8(slot + (k * 32 + lane) * 8, bLb + tb + 32 * k);
                bs_cp8(slot + 1024 + (k * 32 + lane) * 8, bLc + tb + 32 * k);
                bs_cp4(slot + 2048 + (k * 32 + lane) * 4, n1k + (size_t)nb * 32);
              }
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t code = c0;
        c0 = c1; c1 = c2;
        const unsigned char* buf = wring + s0 * BS_SLOT_BYTES;
        s0 = s0 == BS_NB - 1 ? 0 : s0 + 1; s2 = s2 == BS_NB - 1 ? 0 : s2 + 1; s4 = s4 == BS_NB - 1 ? 0 : s4 + 1;
        // ---- compute stage: descriptor di
        const int op = code & 15;
        if (op == BD_END) break;
        const int var = (code >> 4) & 3, kk = (code >> 6) & 7;
        int cnt = (code >> 9) & 127;
        if (LP && cnt == 0) cnt = 128;  // (7-bit field; only record chunks use it)
        if (op == BD_TILE_BEGIN) {
          tau = (int)(code >> 16);
#pragma unroll
          for (int k = 0; k < R; k++) {
            prod[k] = b.wt[tau * T + 32 * k + lane];  // the slice weights enter the products here (0 for padding slices)
            Y[k] = *reinterpret_cast<const double*>(buf + (k * 32 + lane) * 8);
            U[k] = LP ? 0. : *reinterpret_cast<const double*>(buf + 1024 + (k * 32 + lane) * 8);
            ex[k] = *reinterpret_cast<const int*>(buf + 2048 + (k * 32 + lane) * 4);
          }
          continue;
        }
        if (LP && op == BD_TILE_LP) {
          if (var) {  // heights first: S1 (g_1 - g_0)
#pragma unroll
            for (int k = 0; k < R; k++)
              lpacc[k] = *reinterpret_cast<const double*>(buf + (k * 32 + lane) * 8) * (s1_0 - *reinterpret_cast<const double*>(buf + 1024 + (k * 32 + lane) * 8));
          } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
              const int n1 = n1_0 - *reinterpret_cast<const int*>(buf + 2048 + (k * 32 + lane) * 4);
              const double s2 = fma((double)n1, *reinterpret_cast<const double*>(buf + 1024 + (k * 32 + lane) * 8),
                                    *reinterpret_cast<const double*>(buf + (k * 32 + lane) * 8)) + lpacc[k];
              int kx;
              prod[k] *= lp_pow2(s2, kx);
              ex[k] -= kx;
              lpacc[k] = 0.;
            }
          }
          continue;
        }
        if (op == BD_TILE_END) {
#pragma unroll
          for (int k = 0; k < R; k++) scaled_add(Im, Ie, prod[k], ex[k]);
          if (a.slice_out) {  // (baseline row of the sparse evaluation: its per-slice integrand)
#pragma unroll
            for (int k = 0; k < R; k++) a.z1[(size_t)p * b.S_pad + tau * T + 32 * k + lane] = make_double2(prod[k], (double)ex[k]);
          }
          continue;
        }
        if (!LP && op == BD_PAIR) {
          const double4* rb = reinterpret_cast<const double4*>(buf);
          if (var == BV_FULL) {
            bs_pairs_full<R>(rb, cnt, U, prod, ex);
          } else {
#pragma unroll
            for (int k = 0; k < R; k++)
              if (k == kk) prod[k] *= bs_pairs_one(rb, cnt, U[k]);
          }
        } else if (op == BD_REC) {  // 16-byte records (a, nbn), factor a + nbn * Y'
          const double2* rb = reinterpret_cast<const double2*>(buf);
          if (var == BV_FULL) {
            bs_recs_full<R>(rb, cnt, Y, prod, ex);
          } else if (var == BV_F) {
#pragma unroll
            for (int k = 0; k < R; k++)
              if (k == kk) prod[k] *= bs_recs_one(rb, cnt, Y[k]);
          } else {
            const uint32_t* mb = reinterpret_cast<const uint32_t*>(buf + BS_MASK_OFF);
#pragma unroll
            for (int k = 0; k < R; k++)
              if (k == kk) prod[k] *= bs_recs_masked(rb, mb, cnt, Y[k], lanebit);
          }
        } else {
          continue;  // BS_NOP
        }
        bs_rescale<R>(prod, ex);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double vm = __shfl_down_sync(0xffffffffu, Im, o);
        const int ve = __shfl_down_sync(0xffffffffu, Ie, o);
        scaled_add(Im, Ie, vm, ve);
      }
      if (lane == 0) a.unit_I[(size_t)il * G + w] = make_double2(Im, (double)Ie);
    }
    u = __shfl_sync(0xffffffffu, u_next, 0);
  }
}

// ------------------------------------------------------------------------------------------------ mixture (lane = item)
__global__ void __launch_bounds__(128) band_mixture_kernel(BandStreamArgs a) {
  const int lane = threadIdx.x & 31;
  const int blk = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (blk >= a.n_blk) return;
  const long long item = a.item0 + 32ll * blk;
  const int p = (int)(item / a.count_pad);
  const long long pat0 = item - (long long)p * a.count_pad;
  const long long pat = pat0 + lane;
  const double* sc = a.t.scal + (size_t)p * SC_COUNT;
  double v = 0.;
  if (pat < a.count && bs_served(sc, a.force)) {
    const double2* ui = a.unit_I + ((size_t)blk * 32 + lane) * a.b.G;
    double Im = 0.; int Ie = 0x3fffffff;
    for (int w = 0; w < a.b.G; w++) { const double2 x = ui[w]; scaled_add(Im, Ie, x.x, (int)x.y); }
    v = band_mixture(a.m, a.t, p, pat, Im, Ie, nullptr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) a.partial[(size_t)p * (a.count_pad / 32) + pat0 / 32] = v;
}

// leaf powers, once per handle: sum of the heights of every pattern's present leaves
__global__ void leaf_height_sums_kernel(ModelDev m, const double* leaf_h, long long count, double* s1_0) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pat >= count) return;
  double s = 0.;
  for (int w = 0; w < m.n_words; w++) {
    uint32_t x = m.words[(size_t)w * m.n_pat_pad + pat];
    while (x) { const int bq = __ffs((int)x) - 1; x &= x - 1u; s += leaf_h[32 * w + bq]; }
  }
  s1_0[pat] = s;
}
__global__ void leaf_height_sums_q_kernel(ModelDev m, const int* leaf_hq, long long count, int* s1q_0) {
  const long long pat = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pat >= count) return;
  int s = 0;
  for (int w = 0; w < m.n_words; w++) {
    uint32_t x = m.words[(size_t)w * m.n_pat_pad + pat];
    while (x) { const int bq = __ffs((int)x) - 1; x &= x - 1u; s += leaf_hq[32 * w + bq]; }
  }
  s1q_0[pat] = s;
}
void launch_leaf_height_sums_q(const ModelDev& m, const int* leaf_hq, long long count, int* s1q_0, cudaStream_t s) {
  if (count > 0) leaf_height_sums_q_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(m, leaf_hq, count, s1q_0);
}
void launch_leaf_height_sums(const ModelDev& m, const double* leaf_h, long long count, double* s1_0, cudaStream_t s) {
  if (count > 0) leaf_height_sums_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(m, leaf_h, count, s1_0);
}

// ------------------------------------------------------------------------------------------------ launchers
size_t band_stream_block_bytes(const BandDev& b) {  // scratch per 32-item block
  return (size_t)b.n_int * 32 * sizeof(double2) + (size_t)(b.n_int + 1) * 32 * sizeof(int) + (size_t)32 * b.G * sizeof(double2) +
         (b.lp ? (size_t)(b.n_int + 1) * 32 * sizeof(int) : 0) + (b.has_s1 ? (size_t)(b.n_int + 1) * 32 * sizeof(double) : 0);
}

template <int R, bool LP>
static void launch_integral_tl(const BandStreamArgs& a, int n_sm, cudaStream_t s) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(band_integral_kernel<R, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const size_t smem = (size_t)BS_WARPS * bs_warp_bytes(a.m.n_words, LP);
  int ctas = (int)std::min<size_t>(5, (227 * 1024) / (smem + 1024));
  if (const char* e = std::getenv("PPM_BAND_CTAS")) { const int v = std::atoi(e); if (v >= 1) ctas = std::min(ctas, v); }
  const long long units = (long long)a.n_blk * 32 * a.b.G;
  const int grid = (int)std::max<long long>(1, std::min<long long>((long long)n_sm * ctas, (units + BS_WARPS - 1) / BS_WARPS));
  band_integral_kernel<R, LP><<<grid, 32 * BS_WARPS, smem, s>>>(a);
}
template <int R>
static void launch_integral_t(const BandStreamArgs& a, int n_sm, cudaStream_t s) {
  if (a.b.lp) launch_integral_tl<R, true>(a, n_sm, s); else launch_integral_tl<R, false>(a, n_sm, s);
}

// Baseline rows of the sparse evaluation through the band kernels themselves.  base_words holds 64 pattern columns: 32 copies of
// the all-zero pattern, then 32 of the all-ones pattern (the prune kernel is lane = item), so that with count_pad = 64 block 2p is
// the zero row and block 2p + 1 the ones row of parameter vector p; the integral runs on copy 0 of the ones row only and stores
// its per-slice integrand (slice_out).
void launch_band_baseline_rows(BandStreamArgs a, const uint32_t* base_words, const int* base_n1, const double* base_s1, int n_sm, cudaStream_t s) {
  a.m.words = base_words; a.m.n_pat_pad = 64;
  a.n1_0 = base_n1; a.s1_0 = base_s1;
  a.n1k = const_cast<int*>(a.n1k0); a.s1k = const_cast<double*>(a.s1k0);
  a.count = 64; a.count_pad = 64; a.item0 = 0; a.n_blk = 2 * a.t.P;
  a.nodes = const_cast<double2*>(a.nodes0);
  a.xk = const_cast<int*>(a.xk0);
  a.sparse = 0; a.force = 1; a.slice_out = 1;
  band_prune_kernel<<<(a.n_blk + BS_WARPS - 1) / BS_WARPS, 32 * BS_WARPS, 0, s>>>(a);
  if (a.b.R == 2) launch_integral_t<2>(a, n_sm, s); else launch_integral_t<4>(a, n_sm, s);
}
void launch_band_prune(const BandStreamArgs& a, cudaStream_t s) {
  band_prune_kernel<<<(a.n_blk + BS_WARPS - 1) / BS_WARPS, 32 * BS_WARPS, 0, s>>>(a);
}
void launch_band_integral(const BandStreamArgs& a, int n_sm, cudaStream_t s, cudaEvent_t before_mixture) {  // + mixture
  switch (a.b.R) {
    case 2: launch_integral_t<2>(a, n_sm, s); break;
    default: launch_integral_t<4>(a, n_sm, s); break;
  }
  if (before_mixture) cudaStreamWaitEvent(s, before_mixture, 0);  // (the class sums of this evaluation, on their own stream)
  band_mixture_kernel<<<(a.n_blk + BS_WARPS - 1) / BS_WARPS, 128, 0, s>>>(a);
}

}  // namespace ppm
