// Band kernel: the slice-parallel sweep for big trees (sm_100a, FP64 CUDA cores).
//
// One (pattern, parameter vector) per CTA of G warps (G = 1: one warp, warp-level synchronisation only); LANES = SLICES
// of the Mobile integral (reference source/functions.cpp:245-279).  Per item:
//   phase A  level-parallel pruning of the Mobile category (objects.cpp:427-435, linear domain): every internal lineage
//            leaves ONE 16-byte record (a, -d e^{lam (h - H/2)}) in shared memory -- 16 KB at 1,000 genomes, 80 KB at
//            5,000 -- plus its renormalisation shift; nothing per-lineage ever goes to global memory;
//   phase B  each warp owns whole tiles of 32*R slices (lane l, register k = slice 32k + l) and walks its flat descriptor
//            list (band_program.h): a chunk of <= 32 lineage records is gathered lane-parallel into a per-warp buffer
//            while the previous chunk is multiplied in -- one broadcast LDS.128 per record, then R x (DFMA + DMUL) with R
//            independent dependency chains; leaf pairs enter as one quadratic in u (3 FP64 per slice for two terms);
//   end      integral = sum_j w_j prod_j 2^-ex_j (exponent-aligned), warp/CTA reduction, 3-way mixture on one thread.
// The previous mapping (lane = pattern, ppm_kernels.cu MODE 2/4) kept 16 KB of lineage partials PER LANE in a global scratch
// at 1,000 genomes: 94 GB of DRAM traffic per 63 ms launch.  Here DRAM sees the pattern words and the per-parameter tables.
#include "band_fns.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace ppm {

__global__ void prepass_band_kernel(ModelDev m, TablesDev t, BandDev b, BandTables bt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.n_ops + b.S_pad + m.n_leaves) prepass_band_body(m, t, b, bt, blockIdx.y, i);
}
void launch_prepass_band(const ModelDev& m, const TablesDev& t, const BandDev& b, const BandTables& bt, cudaStream_t s) {
  const int n = b.n_ops + b.S_pad + m.n_leaves;
  prepass_band_kernel<<<dim3((n + 255) / 256, t.P), 256, 0, s>>>(m, t, b, bt);
}

static __host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
// per-warp ring of record chunks: BAND_NB slots of 32 records x 32 B + 32 lane masks; a chunk is gathered BAND_D
// descriptors ahead of its use (cp.async for the per-parameter leaf tables, which come from L2)
enum { BAND_NB = 5, BAND_D = 3, BAND_SLOT_BYTES = 1024 + 128, BAND_WARP_BYTES = BAND_NB * BAND_SLOT_BYTES };
struct BandSmem { size_t node, xk, pw, warp, red, total; };
static __host__ __device__ inline BandSmem band_layout(int n_int, int n_words, int G) {
  BandSmem L;
  size_t o = 0;
  L.node = o; o += al16((size_t)(n_int > 0 ? n_int : 1) * sizeof(double2));
  L.xk = o;   o += al16((size_t)(n_int + 1) * sizeof(int));
  L.pw = o;   o += al16((size_t)n_words * sizeof(uint32_t));
  L.warp = o; o += (size_t)G * BAND_WARP_BYTES;
  L.red = o;  o += al16((size_t)G * 16);
  o += 256;  // the software-pipelined record loops read one group of records past a chunk
  L.total = o;
  return L;
}
size_t band_smem_bytes(const ModelDev& m, const BandDev& b) { return band_layout(b.n_int, m.n_words, b.G).total; }

template <bool MULTI> __device__ __forceinline__ void item_sync() { if (MULTI) __syncthreads(); else __syncwarp(); }

__device__ __forceinline__ void band_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// t *= f on the lanes of the mask only (a predicated DMUL: no select instructions)
__device__ __forceinline__ void band_pmul(double& t, double f, uint32_t mask, uint32_t lanebit) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 x;\n\tand.b32 x, %2, %3;\n\tsetp.ne.u32 p, x, 0;\n\t@p mul.f64 %0, %0, %1;\n\t}" : "+d"(t) : "d"(f), "r"(mask), "r"(lanebit));
}

template <int R>
__device__ __forceinline__ void band_rescale(double (&prod)[R], int (&ex)[R]) {
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

template <int R, bool MULTI>
__global__ void __launch_bounds__(MULTI ? 512 : 32) band_kernel(BandArgs a) {
  extern __shared__ __align__(16) unsigned char band_smem[];
  const ModelDev& m = a.m;
  const BandDev& b = a.b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, G = nthr >> 5;
  const BandSmem L = band_layout(b.n_int, m.n_words, G);
  double2* node = reinterpret_cast<double2*>(band_smem + L.node);
  int* xk = reinterpret_cast<int*>(band_smem + L.xk);
  uint32_t* pw = reinterpret_cast<uint32_t*>(band_smem + L.pw);
  unsigned char* wring = band_smem + L.warp + (size_t)warp * BAND_WARP_BYTES;
  double* red = reinterpret_cast<double*>(band_smem + L.red);
  const uint32_t lanebit = 1u << lane;
  const int T = 32 * R;
  const uint32_t* dl = b.desc + b.warp_off[warp];
  const uint32_t* el0 = b.entries + b.warp_ent_off[warp];

  const long long total = (long long)a.t.P * a.n_chunks;
  for (long long unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int p = (int)(unit / a.n_chunks);
    const long long c = unit - (long long)p * a.n_chunks;
    const double* sc = a.t.scal + (size_t)p * SC_COUNT;
    const bool served = sc[SC_BAND] != 0. && (a.force || !(sc[SC_I2] < 0.));  // i2 < 0: penalty branch, no pattern work (functions.cpp:362-364)
    if (!served) {
      if (tid == 0) a.partial[(size_t)p * a.n_chunks + c] = 0.;
      continue;
    }
    const double pi1 = sc[SC_PI1];
    const double2 leaf0 = make_double2(sc[SC_LA0], sc[SC_LD0]), leaf1 = make_double2(sc[SC_LA1], sc[SC_LD1]);
    const double* bop = a.bt.bop + (size_t)p * b.n_ops * BOP_DOUBLES;
    const double* bY = a.bt.bY + (size_t)p * b.S_pad;
    const double* bU = a.bt.bU + (size_t)p * b.S_pad;
    const double2* bleaf = a.bt.bleaf + (size_t)p * m.n_leaves * 2;
    const double4* pairq = a.t.pairq + (size_t)p * m.n_pairs * 4;
    double acc = 0.;
    const long long pat1 = (c + 1) * a.chunk < a.count ? (c + 1) * a.chunk : a.count;
    for (long long pat = c * a.chunk; pat < pat1; pat++) {
      for (int w = tid; w < m.n_words; w += nthr) pw[w] = m.words[(size_t)w * m.n_pat_pad + pat];
      if (tid == 0) xk[0] = 0;
      item_sync<MULTI>();
      // ---------------------------------------------------------------- phase A: pruning, level by level
      {
        // rounds of 32*G ops; the next round's topology record and branch constants are fetched (L2) while this one computes
        uint32_t rd = b.rounds[0];
        BandOpDev op = {};
        double2 c01 = {}, c23 = {}, c45 = {};
        auto fetch = [&](uint32_t r) {
          if (tid < (int)((r >> 16) & 0x3ffu)) {
            const int o = (int)(r & 0xffffu) + tid;
            op = b.ops[o];
            const double2* oc2 = reinterpret_cast<const double2*>(bop + (size_t)o * BOP_DOUBLES);
            c01 = oc2[0]; c23 = oc2[1]; c45 = oc2[2];
          }
        };
        fetch(rd);
        for (int r = 0; r < b.n_rounds; r++) {
          const uint32_t rn = b.rounds[r + 1];
          const BandOpDev o = op;
          const double oc[5] = {c01.x, c01.y, c23.x, c23.y, c45.x};
          const bool act = tid < (int)((rd >> 16) & 0x3ffu);
          fetch(rn);
          if (act) {
            double2 Lc, Rc;
            if (o.flags & 1) Lc = ((pw[o.lref >> 5] >> (o.lref & 31)) & 1u) ? leaf1 : leaf0; else Lc = node[o.lref];
            if (o.flags & 2) Rc = ((pw[o.rref >> 5] >> (o.rref & 31)) & 1u) ? leaf1 : leaf0; else Rc = node[o.rref];
            int k;
            node[o.dst] = band_merge(Lc.x, Lc.y, Rc.x, Rc.y, oc, pi1, k);
            xk[o.dst + 1] = k;
          }
          if (rd >> 30) item_sync<MULTI>();
          rd = rn;
        }
      }
      // shifts -> prefix sums in birth order (xk[i] = sum of the shifts of the first i nodes)
      if (warp == 0) {
        int run = 0;
        for (int base = 1; base <= b.n_int; base += 32) {
          const int i = base + lane;
          int v = i <= b.n_int ? xk[i] : 0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
          if (i <= b.n_int) xk[i] = v + run;
          run += __shfl_sync(0xffffffffu, v, 31);
        }
      }
      item_sync<MULTI>();
      // ---------------------------------------------------------------- phase B: this warp's tiles
      double prod[R], Y[R], U[R];
      int ex[R];
#pragma unroll
      for (int k = 0; k < R; k++) { prod[k] = 1.; Y[k] = 0.; U[k] = 0.; ex[k] = 0; }
      double Im = 0.; int Ie = 0x3fffffff;
      int tau = 0;
      // descriptor window: lane l holds descriptor wbase + l; the next 32 are prefetched
      int wbase = 0;
      uint32_t dw = dl[lane], dn = dl[32 + lane];
      const uint32_t* el = el0;
      uint32_t q0 = 0u, q1 = 0u;  // this lane's entries of the chunks gathered in this and in the next iteration
      int sG = 0, sC = 2;         // ring slots: gather stage (descriptor di+3), compute stage (descriptor di)
      for (int di = -(BAND_D + 2);; di++) {
        {
          const int lo = di < 0 ? 0 : di;
          if (lo + BAND_D + 2 - wbase >= 32) {  // slide the window to start at descriptor `lo`
            const int src = lane + (lo - wbase);
            const uint32_t c_a = __shfl_sync(0xffffffffu, dw, src & 31), c_b = __shfl_sync(0xffffffffu, dn, src & 31);
            dw = src < 32 ? c_a : c_b;
            wbase = lo;
            dn = dl[wbase + 32 + lane];
          }
        }
        // ---- E stage: this lane's entry of the chunk BAND_D + 2 descriptors ahead
        const uint32_t cE = __shfl_sync(0xffffffffu, dw, di + BAND_D + 2 - wbase);
        uint32_t e_new = 0u;
        if (((cE & 15u) | 2u) == 6u || (cE & 15u) == BD_REC) {  // BD_REC (2) or BD_PAIR (4)
          e_new = el[lane];
          el += (cE >> 9) & 63u;
        }
        // ---- G stage: gather the chunk BAND_D descriptors ahead into its ring slot
        if (di + BAND_D >= 0) {
          const uint32_t cG = __shfl_sync(0xffffffffu, dw, di + BAND_D - wbase);
          const uint32_t opG = cG & 15u;
          if (opG == BD_REC || opG == BD_PAIR) {
            const bool valid = lane < (int)((cG >> 9) & 63u);
            const uint32_t ref = q0 & 0xffffu;
            unsigned char* slot = wring + sG * BAND_SLOT_BYTES;
            if (opG == BD_REC) {
              double2* dst = reinterpret_cast<double2*>(slot) + lane;
              if (valid && (q0 >> 31)) {
                const uint32_t bt1 = (pw[ref >> 5] >> (ref & 31u)) & 1u;
                band_cp16(dst, bleaf + ref * 2 + bt1);
              } else {
                *dst = valid ? node[ref] : make_double2(1., 0.);
              }
              if (((cG >> 4) & 3u) == BV_M) {
                const uint32_t lo = (q0 >> 16) & 31u, len = ((q0 >> 21) & 63u) - lo;
                reinterpret_cast<uint32_t*>(slot + 1024)[lane] = valid ? ((len >= 32u ? 0xffffffffu : ((1u << len) - 1u)) << lo) : 0u;
              }
            } else {
              double4* dst = reinterpret_cast<double4*>(slot) + lane;
              if (valid) {
                const uint32_t bits = (pw[ref >> 4] >> (2u * (ref & 15u))) & 3u;
                const double4* src = pairq + ref * 4 + bits;
                band_cp16(dst, src);
                band_cp16(reinterpret_cast<unsigned char*>(dst) + 16, reinterpret_cast<const unsigned char*>(src) + 16);
              } else {
                *dst = make_double4(1., 0., 0., 0.);
              }
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        q0 = q1; q1 = e_new;
        sG = sG == BAND_NB - 1 ? 0 : sG + 1;
        if (di < 0) continue;
        static_assert(BAND_D == 3, "wait_group literal below");
        asm volatile("cp.async.wait_group 3;" ::: "memory");
        __syncwarp();
        // ---- C stage: the current descriptor
        const uint32_t code = __shfl_sync(0xffffffffu, dw, di - wbase);
        const int op = code & 15;
        if (op == BD_END) break;
        const unsigned char* buf = wring + sC * BAND_SLOT_BYTES;
        sC = sC == BAND_NB - 1 ? 0 : sC + 1;
        const int var = (code >> 4) & 3, kk = (code >> 6) & 7, cnt = (code >> 9) & 63;
        if (op == BD_TILE_BEGIN) {
          tau = (int)(code >> 16);
#pragma unroll
          for (int k = 0; k < R; k++) {
            const int j = tau * T + 32 * k + lane;
            prod[k] = 1.; Y[k] = bY[j]; U[k] = bU[j];
            ex[k] = xk[b.nborn[j]];
          }
          continue;
        }
        if (op == BD_TILE_END) {
#pragma unroll
          for (int k = 0; k < R; k++) {
            const int j = tau * T + 32 * k + lane;
            scaled_add(Im, Ie, prod[k] * b.wt[j], ex[k]);
          }
          continue;
        }
        if (op == BD_PAIR) {
          const double4* rb = reinterpret_cast<const double4*>(buf);
          if (var == BV_FULL) {
            double4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = rb[u];
#pragma unroll 1
            for (int i = 0; i < cnt; i += 4) {
              double4 qn[4];
#pragma unroll
              for (int u = 0; u < 4; u++) qn[u] = rb[i + 4 + u];  // next group (past the chunk: read and dropped)
#pragma unroll
              for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int k = 0; k < R; k++) prod[k] *= fma(fma(q[u].z, U[k], q[u].y), U[k], q[u].x);
              }
#pragma unroll
              for (int u = 0; u < 4; u++) q[u] = qn[u];
              if (i == 12 && cnt > 16) band_rescale<R>(prod, ex);
            }
          } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
              if (k == kk) {
                double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
                double4 q0r = rb[0], q1r = rb[1], q2r = rb[2], q3r = rb[3];
#pragma unroll 1
                for (int i = 0; i < cnt; i += 4) {
                  const double4 n0 = rb[i + 4], n1 = rb[i + 5], n2 = rb[i + 6], n3 = rb[i + 7];
                  t0 *= fma(fma(q0r.z, U[k], q0r.y), U[k], q0r.x);
                  t1 *= fma(fma(q1r.z, U[k], q1r.y), U[k], q1r.x);
                  t2 *= fma(fma(q2r.z, U[k], q2r.y), U[k], q2r.x);
                  t3 *= fma(fma(q3r.z, U[k], q3r.y), U[k], q3r.x);
                  q0r = n0; q1r = n1; q2r = n2; q3r = n3;
                }
                prod[k] *= (t0 * t1) * (t2 * t3);
              }
            }
          }
        } else {  // BD_REC: 16-byte records (a, nbn), factor a + nbn * Y'
          const double2* rb = reinterpret_cast<const double2*>(buf);
          if (var == BV_FULL) {
            double2 r[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r[u] = rb[u];
#pragma unroll 1
            for (int i = 0; i < cnt; i += 4) {
              double2 rn[4];
#pragma unroll
              for (int u = 0; u < 4; u++) rn[u] = rb[i + 4 + u];
#pragma unroll
              for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int k = 0; k < R; k++) prod[k] *= fma(r[u].y, Y[k], r[u].x);
              }
#pragma unroll
              for (int u = 0; u < 4; u++) r[u] = rn[u];
              if (i == 12 && cnt > 16) band_rescale<R>(prod, ex);
            }
          } else if (var == BV_F) {
#pragma unroll
            for (int k = 0; k < R; k++) {
              if (k == kk) {
                double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
                double2 r0 = rb[0], r1 = rb[1], r2 = rb[2], r3 = rb[3];
#pragma unroll 1
                for (int i = 0; i < cnt; i += 4) {
                  const double2 n0 = rb[i + 4], n1 = rb[i + 5], n2 = rb[i + 6], n3 = rb[i + 7];
                  t0 *= fma(r0.y, Y[k], r0.x);
                  t1 *= fma(r1.y, Y[k], r1.x);
                  t2 *= fma(r2.y, Y[k], r2.x);
                  t3 *= fma(r3.y, Y[k], r3.x);
                  r0 = n0; r1 = n1; r2 = n2; r3 = n3;
                }
                prod[k] *= (t0 * t1) * (t2 * t3);
              }
            }
          } else {
            const uint32_t* mb = reinterpret_cast<const uint32_t*>(buf + 1024);
#pragma unroll
            for (int k = 0; k < R; k++) {
              if (k == kk) {
                double t0 = 1., t1 = 1., t2 = 1., t3 = 1.;
                double2 r0 = rb[0], r1 = rb[1], r2 = rb[2], r3 = rb[3];
                uint4 mk = *reinterpret_cast<const uint4*>(mb);
#pragma unroll 1
                for (int i = 0; i < cnt; i += 4) {
                  const double2 n0 = rb[i + 4], n1 = rb[i + 5], n2 = rb[i + 6], n3 = rb[i + 7];
                  const uint4 mn = *reinterpret_cast<const uint4*>(mb + i + 4);
                  band_pmul(t0, fma(r0.y, Y[k], r0.x), mk.x, lanebit);
                  band_pmul(t1, fma(r1.y, Y[k], r1.x), mk.y, lanebit);
                  band_pmul(t2, fma(r2.y, Y[k], r2.x), mk.z, lanebit);
                  band_pmul(t3, fma(r3.y, Y[k], r3.x), mk.w, lanebit);
                  r0 = n0; r1 = n1; r2 = n2; r3 = n3; mk = mn;
                }
                prod[k] *= (t0 * t1) * (t2 * t3);
              }
            }
          }
        }
        band_rescale<R>(prod, ex);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      // ---------------------------------------------------------------- reduction and mixture
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double vm = __shfl_down_sync(0xffffffffu, Im, o);
        const int ve = __shfl_down_sync(0xffffffffu, Ie, o);
        scaled_add(Im, Ie, vm, ve);
      }
      if (MULTI) {
        if (lane == 0) { red[2 * warp] = Im; red[2 * warp + 1] = (double)Ie; }
        __syncthreads();
        if (tid == 0) {
          Im = red[0]; Ie = (int)red[1];
          for (int w = 1; w < G; w++) scaled_add(Im, Ie, red[2 * w], (int)red[2 * w + 1]);
        }
      }
      if (tid == 0) acc += band_mixture(m, a.t, p, pat, Im, Ie, nullptr);
      item_sync<MULTI>();  // the node records / pattern words are rewritten by the next item
    }
    if (tid == 0) a.partial[(size_t)p * a.n_chunks + c] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static const size_t kBandSmemMax = 227 * 1024;

BandPlan plan_band(const ModelDev& m, const BandDev& b, long long count, int P, int n_sm) {
  BandPlan pl{};
  pl.G = b.G;
  pl.smem_bytes = band_smem_bytes(m, b);
  int ctas = (int)std::min<size_t>(32, kBandSmemMax / (pl.smem_bytes + 1024));  // (+1 KB reserved per CTA by the driver)
  ctas = std::min(ctas, 64 / b.G);  // at most 64 warps per SM
  if (const char* e = std::getenv("PPM_BAND_CTAS")) { const int v = std::atoi(e); if (v >= 1) ctas = std::min(ctas, v); }
  pl.ctas_per_sm = std::max(ctas, 1);
  const long long slots = (long long)n_sm * pl.ctas_per_sm;
  // units = (parameter vector, pattern chunk).  The chunk size depends on the shard only (not on P), so a pattern's place
  // in the summation order -- and with it the result, bit for bit -- is the same whatever else shares the launch.
  long long chunk = std::max<long long>(1, std::min<long long>(64, (count + slots * 8 - 1) / (slots * 8)));
  pl.chunk = chunk;
  pl.n_chunks = (int)std::max<long long>(1, (count + chunk - 1) / chunk);
  pl.grid_x = (int)std::max<long long>(1, std::min<long long>(slots, (long long)pl.n_chunks * P));
  if (std::getenv("PPM_PLAN_DEBUG"))
    fprintf(stderr, "[ppm band plan] R %d, G %d, smem %zu B, %d CTAs/SM, grid %d, %d chunks of %lld patterns\n", b.R, b.G, pl.smem_bytes,
            pl.ctas_per_sm, pl.grid_x, pl.n_chunks, pl.chunk);
  return pl;
}

template <int R, bool MULTI>
static void launch_band_t(const BandArgs& a, const BandPlan& pl, cudaStream_t s) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(band_kernel<R, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBandSmemMax);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  band_kernel<R, MULTI><<<pl.grid_x, 32 * pl.G, pl.smem_bytes, s>>>(a);
}

void launch_band(BandArgs a, const BandPlan& pl, cudaStream_t s) {
  a.n_chunks = pl.n_chunks; a.chunk = pl.chunk;
  const bool multi = pl.G > 1;
  switch (a.b.R) {
    case 2: if (multi) launch_band_t<2, true>(a, pl, s); else launch_band_t<2, false>(a, pl, s); break;
    case 8: if (multi) launch_band_t<8, true>(a, pl, s); else launch_band_t<8, false>(a, pl, s); break;
    default: if (multi) launch_band_t<4, true>(a, pl, s); else launch_band_t<4, false>(a, pl, s); break;
  }
}

// out[p] = SC_BAND[p] ? sum(band_partial[p][.]) : sum(partial[p][.]); fixed summation order => bitwise reproducible
__global__ void reduce2_kernel(const double* partial, int nx, const double* band_partial, int nb, const double* scal, double* out) {
  __shared__ double sh[256];
  const bool band = scal[(size_t)blockIdx.x * SC_COUNT + SC_BAND] != 0.;
  const double* row = band ? band_partial + (size_t)blockIdx.x * nb : partial + (size_t)blockIdx.x * nx;
  const int n = band ? nb : nx;
  double s = 0.;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += row[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
void launch_reduce2(const double* partial, int nx, const double* band_partial, int nb, const double* scal, int P, double* out, cudaStream_t s) {
  reduce2_kernel<<<P, 256, 0, s>>>(partial, nx, band_partial, nb, scal, out);
}

}  // namespace ppm
