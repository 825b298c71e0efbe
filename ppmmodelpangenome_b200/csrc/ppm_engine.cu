// Engine object behind the C ABI of include/ppm_b200.h: owns the device copy of the model, the per-launch
// parameter tables, a stream and timing events.  No CPU fallback: every compute entry point needs an sm_100 GPU.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges are no-ops unless a profiler injects the NVTX library

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ppm_b200.h"
#include "band_program.h"
#include "host_model.h"
#include "ppm_kernels.cuh"

namespace {

thread_local std::string g_create_error;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if (count <= n) return;
    release();
    CK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void upload(const std::vector<T>& v, cudaStream_t s) {
    alloc(v.size());
    if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

// NVTX range for a scope (shows up in Nsight Systems / ncu --nvtx: "ppm:prepass", "ppm:sweep", ...)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace

// slices per block of the sweep program: 8; PPM_R=16 is a tuning knob of the global-memory variant (it halves the record
// and lineage-slot traffic per term but its register need costs as much: no gain measured)
static bool leaf_pow_enabled() { const char* e = std::getenv("PPM_NO_LEAFPOW"); return !(e && std::atoi(e) != 0); }
static int program_R(bool deep = false) {
  const char* e = std::getenv("PPM_R");
  return (deep && e && std::atoi(e) == 16) ? 16 : 8;
}

struct ppm_engine {
  // host model: owned by the first engine of a group, shared (read-only) by its peers
  std::shared_ptr<ppm::Tree> tree_own;
  std::shared_ptr<ppm::Patterns> pat_own;
  ppm::Tree& tree;
  ppm::Patterns& pat;
  ppm::Program prog;
  // group head only (ppm_create_group*): the engines of shards 1..n-1, one per GPU; the head itself is shard 0
  std::vector<std::unique_ptr<ppm_engine>> peers;
  cudaEvent_t ev_done = nullptr;  // "this shard's partial sums have reached the head"

  ppm_engine() : tree_own(std::make_shared<ppm::Tree>()), pat_own(std::make_shared<ppm::Patterns>()), tree(*tree_own), pat(*pat_own) {}
  ppm_engine(const std::shared_ptr<ppm::Tree>& t, const std::shared_ptr<ppm::Patterns>& p) : tree_own(t), pat_own(p), tree(*tree_own), pat(*pat_own) {}
  std::vector<ppm_engine*> members() {
    std::vector<ppm_engine*> v{this};
    for (auto& e : peers) v.push_back(e.get());
    return v;
  }
  bool is_group() const { return !peers.empty(); }
  int device = 0, n_sm = 148;
  int shard_rank = 0, shard_world = 1;
  long long first = 0, count = 0;  // this shard's pattern slice
  std::string err;
  cudaStream_t stream = nullptr;
  cudaStream_t side2 = nullptr;   // the class sums of an evaluation (next to the table prepass, the zero row and the first batch's prune)
  cudaStream_t side = nullptr;    // (until round 2: class sums run here, next to the table prepass and the zero row (they only need the node prepass)
  cudaEvent_t ev_nodes = nullptr, ev_class = nullptr;
  bool class_pending = false;
  bool hint_no_factored = false;  // set by the host-pointer entry points when no parameter vector has pi1 > 0.95
  bool hint_all_band = false;     // ... and when every parameter vector is served by the band kernel (SC_BAND)
  // band kernel (slice-parallel sweep of big trees, band_program.h)
  ppm::BandProgram band;
  bool use_band = false;
  ppm::BandDev bdev{};
  DevBuf<ppm::BandOpDev> d_bops;
  DevBuf<int> d_bop_nodes, d_blev_off, d_bwarp_off;
  DevBuf<uint16_t> d_bnborn;
  DevBuf<double> d_bwt, d_bslice_t, d_bop, d_bY, d_bU, d_partial_b;
  DevBuf<uint32_t> d_bdesc, d_bentries, d_brounds;
  DevBuf<int> d_bwarp_ent_off;
  DevBuf<double2> d_bleaf;
  // stream layout (band_stream.cu): per-batch scratch
  DevBuf<double2> d_bs_nodes[2], d_bs_unit[2], d_bs_ident;
  DevBuf<int> d_bs_xk[2];
  cudaEvent_t ev_bs_tables = nullptr, ev_bs_pruned[2] = {nullptr, nullptr}, ev_bs_done[2] = {nullptr, nullptr};
  DevBuf<unsigned int> d_bs_counter;
  DevBuf<uint32_t> d_bs_blocks, d_bs_op_span, d_bs_leaf_span, d_sp_off;
  DevBuf<uint2> d_sp_recs;
  bool sp_lists_ready = false, sp_lists_failed = false;
  unsigned long long sp_total = 0;
  DevBuf<double2> d_bs_nodes0, d_bs_z1, d_bs_unit_base;  // sparse evaluation: records / shift prefixes of the all-zero and all-ones
  DevBuf<int> d_bs_xk0;                                  // rows per parameter vector, per-slice integrand of the all-ones row
  DevBuf<uint32_t> d_bs_ones;
  // leaf powers (band_program.h: leaf_pow)
  DevBuf<int> d_blp_alive, d_bs_n1k[2], d_bs_n1k0, d_bs_base_n1, d_lp_alive;
  DevBuf<int> d_lp_leaf_hq, d_lp_s1q_0;
  DevBuf<double> d_lp_htot, d_slb;
  DevBuf<double2> d_slce;
  DevBuf<double> d_blp_htot, d_blp_op_h, d_blp_leaf_h, d_blp_s1_0, d_bLb, d_bLc, d_bLe, d_bs_s1k[2], d_bs_s1k0, d_bs_base_s1;
  bool sparse = false;            // ppm_set_sparse / PPM_SPARSE: the band kernels multiply dirty lineages only
  DevBuf<int> d_bs_unit_off;
  int bs_blk_cap = 0;             // 32-item blocks the scratch holds
  size_t bs_budget = 0;           // scratch budget in bytes (decided at the first band launch)
  size_t partial_b_cap = 0;
  int last_sweep_kind = 0;        // 0: lane-per-pattern sweep, 1: band kernel (+ fallback launch), for ppm_get_counters
  int last_sweep_mode = -1;       // SweepPlan::mode of the last lane-per-pattern launch
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  double n_launches = 0;

  // device model
  DevBuf<int> d_leaf_node, d_left, d_right, d_parent, d_entry_ref, d_entry_node, d_entry_block, d_lev_up_off, d_lev_up_nodes, d_lev_dn_off, d_lev_dn_nodes, d_lev_dn_chain;
  DevBuf<ppm::BlockDesc> d_blk;
  DevBuf<uint32_t> d_entry_mask;
  DevBuf<double> d_bl, d_height, d_slice_t, d_slice_wt, d_op_tb;
  DevBuf<ppm::NodeOpDev> d_ops;
  DevBuf<uint32_t> d_words, d_zero_words, d_fr_off;
  DevBuf<uint16_t> d_fr_nodes;
  DevBuf<int> d_counts, d_mrca, d_freq;
  long long n_pat_pad = 0;
  ppm::ModelDev md{};

  // per-launch tables
  int P_cap = 0;
  DevBuf<ppm::EntryRec> d_recs;
  DevBuf<double4> d_pairq;
  DevBuf<double2> d_cherry;
  DevBuf<ppm::SweepOpDev> d_sops;
  DevBuf<int> d_cherry_op;
  DevBuf<double2> d_leafab;
  DevBuf<double> d_params, d_yy, d_uu, d_oprec, d_scal, d_zeta, d_partial, d_ll, d_i2, d_comp, d_i2_override, d_et1, d_dist, d_gather;
  DevBuf<double4> d_cls;
  DevBuf<double2> d_kap, d_d1u, d_gslots, d_clsT, d_a01, d_zrow, d_zpart;
  DevBuf<int> d_zk, d_zxb;
  DevBuf<signed char> d_cat;
  size_t partial_cap = 0;

  ~ppm_engine() {
    peers.clear();
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) cudaGetLastError();  // never leave a sticky error behind
    if (ev_done) cudaEventDestroy(ev_done);
    if (ev_nodes) cudaEventDestroy(ev_nodes);
    if (ev_class) cudaEventDestroy(ev_class);
    if (ev_bs_tables) cudaEventDestroy(ev_bs_tables);
    for (int k = 0; k < 2; k++) {
      if (ev_bs_pruned[k]) cudaEventDestroy(ev_bs_pruned[k]);
      if (ev_bs_done[k]) cudaEventDestroy(ev_bs_done[k]);
    }
    if (side) cudaStreamDestroy(side);
    if (side2) cudaStreamDestroy(side2);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }

  void init_device(int dev) {
    device = -1;  // (the destructor must not select a device that was never validated)
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (dev < 0 || dev >= ndev) throw CudaError("no CUDA device with ordinal " + std::to_string(dev));
    CK(cudaSetDevice(dev));
    device = dev;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) throw CudaError(std::string("device '") + prop.name + "' is not sm_100 (this library carries sm_100a code only)");
    n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    {  // the side stream carries short latency-bound kernels (prune of the next batch, class sums) that must get their SMs
       // ahead of the persistent kernels of the main stream: highest priority
      int prio_lo = 0, prio_hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      CK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
      CK(cudaStreamCreateWithPriority(&side2, cudaStreamNonBlocking, prio_hi));
    }
    CK(cudaEventCreateWithFlags(&ev_nodes, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_class, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_bs_tables, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
      CK(cudaEventCreateWithFlags(&ev_bs_pruned[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_bs_done[k], cudaEventDisableTiming));
    }
  }

  void upload_model() {
    NvtxRange r("ppm:upload_model");
    upload_model_impl(false);
    const bool staged = ppm::sweep_is_staged(md);
    if (!staged) upload_model_impl(true);
    upload_patterns();
    setup_band(staged);
  }

  // Band kernel: chosen where the lane-per-pattern sweep would keep its lineage partials in global memory (or when forced),
  // unless the tree is so deep (caterpillar) that the level-by-level pruning of one item would dominate its integral.
  void setup_band(bool staged) {
    use_band = false;
    const char* no = std::getenv("PPM_NO_BAND");
    const char* force = std::getenv("PPM_FORCE_BAND");
    const bool forced = force && std::atoi(force) != 0;
    if ((no && std::atoi(no) != 0) || (staged && !forced) || count <= 0) return;
    // Measured crossover on B200 (20,000 genes, fraction of the FP64 peak, band | lane-per-pattern): 1,000 genomes 0.38 | 0.53,
    // 1,500 0.50 | 0.40, 2,000 0.54 | 0.39, 3,000 0.64 | 0.39, 5,000 0.73 | 0.30 -- below ~1,200 genomes a tile of 128 slices
    // sees too many births and deaths (per-chunk overhead), above it the lane-per-pattern sweep drowns in its global scratch.
    int min_leaves = 1200;
    // (a handle created for the sparse evaluation, PPM_SPARSE=1, takes the band kernels from 256 genomes: at 1,000 genomes the
    // dense band kernels only match the sweep, the term lists beat both by 3.5 x)
    if (const char* e = std::getenv("PPM_SPARSE")) { if (std::atoi(e) != 0) min_leaves = 256; }
    if (const char* e = std::getenv("PPM_BAND_MIN")) min_leaves = std::atoi(e);
    if (tree.n_leaves < min_leaves && !forced) return;
    const char* eR = std::getenv("PPM_BAND_R");
    const char* eG = std::getenv("PPM_BAND_G");
    const int fixR = eR ? std::atoi(eR) : 0, fixG = eG ? std::atoi(eG) : 0;
    const char* eV = std::getenv("PPM_BAND_V");
    const bool stream_layout = !(eV && std::atoi(eV) == 1);
    if (stream_layout) {
      // Stream layout (band_stream.cu): prune (lane = item) -> integral (warp = item x descriptor list) -> mixture.  G = work
      // units per item.  Units are handed out item-major, so about (resident warps / G) items are in flight; their lineage
      // records (16 B x internal nodes each) should stay in L2 (a third of it at most), which asks for G >= warps x bytes / 40 MB.
      const int R = fixR ? fixR : 4;
      const char* eL = std::getenv("PPM_NO_LEAFPOW");
      const bool leaf_pow = !(eL && std::atoi(eL) != 0);
      ppm::BandProgram probe = ppm::build_band_program(tree, prog, R, 1, true, leaf_pow);
      if (!probe.ok) { if (std::getenv("PPM_PLAN_DEBUG")) fprintf(stderr, "[ppm band] not used: %s\n", probe.why); return; }
      const double item_bytes = (probe.leaf_pow ? (probe.has_s1 ? 32.0 : 24.0) : 20.0) * probe.n_int;
      int G = (int)std::ceil(n_sm * 20.0 * item_bytes / 40e6);
      G = std::max(2, std::min(G, std::max(1, probe.n_tiles / 2)));
      if (fixG) G = fixG;
      G = std::max(1, std::min(G, probe.n_tiles));
      band = G == 1 ? std::move(probe) : ppm::build_band_program(tree, prog, R, G, true, leaf_pow);
      if (std::getenv("PPM_PLAN_DEBUG"))
        fprintf(stderr, "[ppm band] stream layout: R %d, %d units per item, %d tiles, balance %.3f; records full16 %lld full32 %lld f16 %lld f32 %lld m16 %lld; leaf powers %d (heights %d, max |h| %.3g)\n",
                band.R, band.G, band.n_tiles, band.cost_sum / (band.G * std::max(band.cost_max, 1.0)), band.n_full16, band.n_full32, band.n_f16,
                band.n_f32, band.n_m16, (int)band.leaf_pow, (int)band.has_s1, band.max_leaf_h);
    } else {
      // Candidates: slices per lane R in {4, 2}, warps per item G in {1, 2, 4, 8, 16}.  The kernel is latency-bound below ~16
      // resident warps per SM (measured: 1,000 genomes 110 ms at 8 warps, 79 ms at 20), and a CTA's warps own whole tiles, so
      // the score is (resident warps, capped at 16) x (tile balance over the G warps); ties go to the smaller G, then R = 4.
      double best = -1;
      int ctas_best = 0;
      bool any = false;
      for (int R : {4, 2, 8}) {
        if (fixR ? R != fixR : R != 4) continue;  // R = 4 measured best at 1,000 and 5,000 genomes (R = 2 doubles the chunks, R = 8 turns FULL records into per-register ones)
        for (int G : {1, 2, 4, 8, 16}) {
          if (fixG && G != fixG) continue;
          ppm::BandProgram cand = ppm::build_band_program(tree, prog, R, G);
          if (!cand.ok) { if (std::getenv("PPM_PLAN_DEBUG")) fprintf(stderr, "[ppm band] not used: %s\n", cand.why); return; }
          ppm::BandDev b{}; b.n_int = cand.n_int; b.G = G;
          const size_t sm = ppm::band_smem_bytes(md, b);
          if (sm + 1024 > 227 * 1024) continue;
          const int ctas = (int)std::min<size_t>(std::min<size_t>(32, 64 / G), (227 * 1024) / (sm + 1024));
          const double balance = cand.cost_sum / (G * std::max(cand.cost_max, 1.0));
          const double score = std::min(ctas * G, 16) / 16.0 * balance;
          if (score > best * 1.02) { best = score; band = std::move(cand); ctas_best = ctas; any = true; }
        }
      }
      if (!any) { if (std::getenv("PPM_PLAN_DEBUG")) fprintf(stderr, "[ppm band] not used: node records do not fit shared memory\n"); return; }
      {
        // modelled efficiency: phase A (latency-bound, one barrier per level) against phase B, with the resident warps of an SM
        const int bestG = band.G;
        const double A = 300.0 * band.n_levels + 60.0 * band.n_int / (32.0 * bestG);
        const double B = std::max(band.cost_max, 1.0) * 2.0;
        const double eff = std::min(1.0, (ctas_best * bestG / 4.0) * B / (A + B));
        if (std::getenv("PPM_PLAN_DEBUG"))
          fprintf(stderr, "[ppm band] R %d G %d (%d CTAs/SM): %d levels, %d tiles, balance %.3f, modelled efficiency %.2f; records full16 %lld full32 %lld f16 %lld f32 %lld m16 %lld\n",
                  band.R, bestG, ctas_best, band.n_levels, band.n_tiles, band.cost_sum / (bestG * std::max(band.cost_max, 1.0)), eff, band.n_full16, band.n_full32,
                  band.n_f16, band.n_f32, band.n_m16);
        if (eff < 0.7 && !forced) return;
      }
    }
    static_assert(sizeof(ppm::BandOpDev) == sizeof(ppm::BandOp), "band layouts");
    std::vector<ppm::BandOpDev> ops(band.ops.size());
    if (!ops.empty()) std::memcpy(ops.data(), band.ops.data(), ops.size() * sizeof(ppm::BandOp));
    d_bops.upload(ops, stream);
    d_bop_nodes.upload(band.op_nodes, stream); d_blev_off.upload(band.lev_off, stream); d_bwarp_off.upload(band.warp_off, stream);
    d_bnborn.upload(band.nborn, stream); d_bwt.upload(band.wt, stream); d_bslice_t.upload(band.slice_t, stream);
    d_bdesc.upload(band.desc, stream);
    d_bentries.upload(band.entries, stream);
    d_brounds.upload(band.rounds, stream); d_bwarp_ent_off.upload(band.warp_ent_off, stream);
    const std::vector<double2> ident{make_double2(1., 0.), make_double2(0., 0.)};
    if (band.stream) {
      d_bs_ident.upload(ident, stream);
      d_bs_counter.alloc(5);
      if (const char* e = std::getenv("PPM_SPARSE")) sparse = std::atoi(e) != 0;
      d_bs_blocks.upload(band.blocks, stream);
      d_bs_op_span.upload(band.op_span, stream); d_bs_leaf_span.upload(band.leaf_span, stream);
      d_bs_unit_off.upload(band.unit_blk_off, stream);
      if (band.leaf_pow) {
        d_blp_alive.upload(band.n_alive_leaves, stream); d_blp_htot.upload(band.htot, stream);
        d_blp_op_h.upload(band.op_leaf_h, stream); d_blp_leaf_h.upload(band.leaf_h, stream);
      }
    }
    CK(cudaStreamSynchronize(stream));
    bdev.R = band.R; bdev.G = band.G; bdev.T = band.T; bdev.n_tiles = band.n_tiles; bdev.S_pad = band.S_pad; bdev.n_int = band.n_int;
    bdev.n_ops = (int)band.ops.size(); bdev.n_levels = band.n_levels; bdev.n_desc = (int)band.desc.size(); bdev.n_entries = (int)band.entries.size();
    bdev.ops = d_bops.p; bdev.op_nodes = d_bop_nodes.p; bdev.lev_off = d_blev_off.p; bdev.nborn = d_bnborn.p; bdev.wt = d_bwt.p;
    bdev.slice_t = d_bslice_t.p; bdev.desc = d_bdesc.p; bdev.warp_off = d_bwarp_off.p; bdev.entries = d_bentries.p;
    bdev.rounds = d_brounds.p; bdev.n_rounds = band.n_rounds; bdev.warp_ent_off = d_bwarp_ent_off.p;
    bdev.blocks = d_bs_blocks.p; bdev.unit_blk_off = d_bs_unit_off.p;
    bdev.op_span = d_bs_op_span.p; bdev.leaf_span = d_bs_leaf_span.p;
    bdev.lp = band.leaf_pow ? 1 : 0; bdev.has_s1 = band.has_s1 ? 1 : 0; bdev.chunk = band.chunk;
    bdev.n_alive_leaves = d_blp_alive.p; bdev.htot = d_blp_htot.p; bdev.op_leaf_h = d_blp_op_h.p;
    md.lp_max_h = band.leaf_pow ? band.max_leaf_h : 0.;
    if (band.has_s1 && !prog.has_s1) {  // sum of the heights of every pattern's present leaves, once per handle
      d_blp_s1_0.alloc((size_t)std::max<long long>(n_pat_pad, 32));
      CK(cudaMemsetAsync(d_blp_s1_0.p, 0, sizeof(double) * (size_t)std::max<long long>(n_pat_pad, 32), stream));
      ppm::launch_leaf_height_sums(md, d_blp_leaf_h.p, count, d_blp_s1_0.p, stream);
      CK(cudaStreamSynchronize(stream));
    }
    use_band = true;
  }
  // Sparse evaluation: the term lists (pattern-only) of this shard, once per handle (band_sparse.cu).
  void build_sparse_lists() {
    NvtxRange r("ppm:sparse_lists");
    sp_lists_failed = true;
    if (band.op_span.empty() || count <= 0 || bdev.n_tiles > 256) return;
    d_sp_off.alloc((size_t)count * bdev.n_tiles + 1);
    sp_total = ppm::sparse_lists_count(md, bdev, count, d_sp_off.p, stream);
    if (sp_total >= (1ull << 32)) return;  // (offsets are 32 bits)
    d_sp_recs.alloc((size_t)std::max<unsigned long long>(sp_total, 1));
    ppm::sparse_lists_fill(md, bdev, count, d_sp_off.p, d_sp_recs.p, stream);
    std::vector<uint32_t> ones((size_t)pat.n_words * 64, 0u);  // word-major: 32 all-zero columns, then 32 all-ones columns
    for (int k = 0; k < tree.n_leaves; k++)
      for (int c = 32; c < 64; c++) ones[(size_t)(k >> 5) * 64 + c] |= 1u << (k & 31);
    d_bs_ones.upload(ones, stream);
    if (band.leaf_pow) {  // the baseline columns' present leaves and height sums
      double hall = 0.;
      for (double h : band.leaf_h) hall += h;
      std::vector<int> bn(64, 0);
      std::vector<double> bs(64, 0.);
      for (int c = 32; c < 64; c++) { bn[c] = tree.n_leaves; bs[c] = hall; }
      d_bs_base_n1.upload(bn, stream); d_bs_base_s1.upload(bs, stream);
    }
    CK(cudaStreamSynchronize(stream));
    sp_lists_ready = true; sp_lists_failed = false;
  }
  ppm::BandTables band_tables() const {
    ppm::BandTables bt{};
    bt.bop = d_bop.p; bt.bY = d_bY.p; bt.bU = d_bU.p; bt.bleaf = d_bleaf.p;
    bt.bLb = d_bLb.p; bt.bLc = d_bLc.p; bt.bLe = d_bLe.p;
    return bt;
  }

  void upload_model_impl(bool deep) {
    prog = ppm::build_program(tree, program_R(deep), deep, leaf_pow_enabled());
    // shard: contiguous, equal-sized slices of the deduplicated patterns (work per pattern is uniform)
    long long np = pat.n_patterns;
    first = np * shard_rank / shard_world;
    count = np * (shard_rank + 1) / shard_world - first;
    d_left.upload(tree.left, stream); d_right.upload(tree.right, stream); d_parent.upload(tree.parent, stream);
    d_bl.upload(tree.bl, stream); d_height.upload(tree.height, stream);
    d_lev_up_off.upload(tree.lev_up_off, stream); d_lev_up_nodes.upload(tree.lev_up_nodes, stream);
    d_lev_dn_off.upload(tree.lev_dn_off, stream); d_lev_dn_nodes.upload(tree.lev_dn_nodes, stream); d_lev_dn_chain.upload(tree.lev_dn_chain, stream);
    d_slice_t.upload(tree.slice_t, stream);
    std::vector<double> wt_pad((size_t)std::max(prog.n_blocks, 1) * prog.R, 0.);
    std::copy(tree.slice_wt.begin(), tree.slice_wt.end(), wt_pad.begin());
    d_slice_wt.upload(wt_pad, stream);
    static_assert(sizeof(ppm::NodeOpDev) == sizeof(ppm::NodeOp), "NodeOp layout");
    static_assert(sizeof(ppm::BlockDesc) == sizeof(ppm::BlockDescHost), "BlockDesc layout");
    std::vector<ppm::NodeOpDev> ops(prog.ops.size());
    if (!ops.empty()) std::memcpy(ops.data(), prog.ops.data(), ops.size() * sizeof(ppm::NodeOp));
    d_ops.upload(ops, stream);
    static_assert(sizeof(ppm::SweepOpDev) == sizeof(ppm::SweepOp), "SweepOp layout");
    std::vector<ppm::SweepOpDev> sops(prog.sops.size());
    if (!sops.empty()) std::memcpy(sops.data(), prog.sops.data(), sops.size() * sizeof(ppm::SweepOp));
    d_sops.upload(sops, stream);
    { std::vector<double> tb(prog.op_tb); if (tb.empty()) tb.assign(1, 0.); d_op_tb.upload(tb, stream); }
    { std::vector<int> co(prog.cherry_op); if (co.empty()) co.assign(1, 0); d_cherry_op.upload(co, stream); }
    std::vector<ppm::BlockDesc> blk(prog.blk.size());
    std::memcpy(blk.data(), prog.blk.data(), blk.size() * sizeof(ppm::BlockDesc));
    d_blk.upload(blk, stream);
    d_entry_ref.upload(prog.entry_ref, stream); d_entry_mask.upload(prog.entry_mask, stream);
    d_entry_node.upload(prog.entry_node, stream); d_entry_block.upload(prog.entry_block, stream);
    d_leaf_node.upload(prog.leaf_node_dev, stream);
    n_pat_pad = (count + 31) / 32 * 32;  // (the pattern words themselves: upload_patterns, once the program is final)
    std::vector<uint32_t> zeros((size_t)pat.n_words * 32, 0u);
    d_zero_words.upload(zeros, stream);
    std::vector<int> c(pat.counts.begin() + first, pat.counts.begin() + first + count);
    std::vector<int> mr(pat.mrca.begin() + first, pat.mrca.begin() + first + count);
    std::vector<int> fr(pat.freq.begin() + first, pat.freq.begin() + first + count);
    d_counts.upload(c, stream); d_mrca.upload(mr, stream); d_freq.upload(fr, stream);
    CK(cudaStreamSynchronize(stream));

    md.n_nodes = tree.n_nodes; md.n_leaves = tree.n_leaves; md.n_words = pat.n_words; md.n_slices = tree.n_slices;
    md.n_blocks = prog.n_blocks; md.n_slots = prog.n_slots; md.n_entries = (int)prog.entry_ref.size();
    md.n_ops = (int)prog.ops.size(); md.R = prog.R; md.deep = prog.deep ? 1 : 0; md.H = tree.H; md.n_obs = pat.n_obs;
    md.left = d_left.p; md.right = d_right.p; md.parent = d_parent.p; md.bl = d_bl.p; md.height = d_height.p;
    md.n_lev_up = (int)tree.lev_up_off.size() - 1; md.n_lev_dn = (int)tree.lev_dn_off.size() - 1;
    md.lev_up_off = d_lev_up_off.p; md.lev_up_nodes = d_lev_up_nodes.p; md.lev_dn_off = d_lev_dn_off.p; md.lev_dn_nodes = d_lev_dn_nodes.p; md.lev_dn_chain = d_lev_dn_chain.p;
    md.slice_t = d_slice_t.p; md.slice_wt_pad = d_slice_wt.p;
    md.blk = d_blk.p; md.ops = d_ops.p; md.sops = d_sops.p; md.op_tb = d_op_tb.p; md.cherry_op = d_cherry_op.p; md.n_cherries = prog.n_cherries;
    md.n_pairs = prog.n_pairs; md.leaf_node = d_leaf_node.p;
    md.entry_ref = d_entry_ref.p; md.entry_mask = d_entry_mask.p; md.entry_node = d_entry_node.p; md.entry_block = d_entry_block.p;
    md.words = d_words.p; md.n_pat = count; md.n_pat_pad = n_pat_pad;
    // big trees (global-memory variant): a shared-memory copy of the pattern words would cost resident warps (40 KB per
    // warp at 10,000 genomes: 5 warps; measured 11.3 -> 15.6 TFLOP/s), so the sweep reads the rows in place; the leaf-pair
    // records then carry byte offsets into the rows, which must fit 31 bits.  Smaller trees keep the copy (5,000 genomes:
    // 11.9 TFLOP/s with the copy and 10 warps, 9.9 in place).
    md.lean = (prog.lean && prog.R == 8 && !std::getenv("PPM_NO_LEAN")) ? 1 : 0;
    md.word_stride = 0;
    if (deep && prog.R == 8 && (size_t)pat.n_words * 128 > 24 * 1024 && (size_t)pat.n_words * (size_t)n_pat_pad * 4 < (1ull << 31) &&
        !std::getenv("PPM_WORDS_SMEM"))
      md.word_stride = (int)n_pat_pad;
    md.counts = d_counts.p; md.mrca = d_mrca.p; md.freq = d_freq.p; md.zero_words = d_zero_words.p;
    md.lp = prog.lp ? 1 : 0; md.has_s1 = prog.has_s1 ? 1 : 0; md.lp_hunit = prog.hunit;
    if (prog.lp) { d_lp_alive.upload(prog.n_alive_leaves, stream); d_lp_htot.upload(prog.htot, stream); }
    md.lp_alive = d_lp_alive.p; md.lp_htot = d_lp_htot.p;
  }

  // Patterns of this shard on the device: word-major (lanes = consecutive patterns read consecutive addresses), bits in
  // PROGRAM leaf order (host_model.h: Program::leaf_perm).  The rows go up as they are (row-major, preorder leaf order); a
  // kernel permutes the bits and transposes (on the host this cost 4 s per million patterns of 5,000 genomes).
  void upload_patterns() {
    NvtxRange r("ppm:upload_patterns");
    const int nw = pat.n_words;
    d_words.alloc((size_t)nw * std::max<long long>(n_pat_pad, 32));
    CK(cudaMemsetAsync(d_words.p, 0, sizeof(uint32_t) * (size_t)nw * std::max<long long>(n_pat_pad, 32), stream));
    if (count > 0) {
      std::vector<int> inv((size_t)nw * 32, -1);  // device bit position -> preorder leaf index
      for (int k = 0; k < tree.n_leaves; k++) inv[prog.leaf_perm[k]] = k;
      DevBuf<int> d_inv;
      DevBuf<uint32_t> d_rows;
      d_inv.upload(inv, stream);
      d_rows.alloc((size_t)count * nw);
      CK(cudaMemcpyAsync(d_rows.p, pat.words.data() + (size_t)first * nw, sizeof(uint32_t) * (size_t)count * nw, cudaMemcpyHostToDevice, stream));
      ppm::launch_permute_words(d_rows.p, d_inv.p, count, nw, n_pat_pad, d_words.p, stream);
      CK(cudaStreamSynchronize(stream));
      CK(cudaGetLastError());
    }
    md.words = d_words.p;
    md.lp_s1_0 = nullptr; md.lp_s1q_0 = nullptr;
    if (prog.has_s1) {  // leaf powers: sum of the heights of every pattern's present leaves, once per handle
      d_blp_leaf_h.upload(prog.leaf_h, stream);
      d_blp_s1_0.alloc((size_t)std::max<long long>(n_pat_pad, 32));
      CK(cudaMemsetAsync(d_blp_s1_0.p, 0, sizeof(double) * (size_t)std::max<long long>(n_pat_pad, 32), stream));
      ppm::launch_leaf_height_sums(md, d_blp_leaf_h.p, count, d_blp_s1_0.p, stream);
      CK(cudaStreamSynchronize(stream));
      md.lp_s1_0 = d_blp_s1_0.p;
      std::vector<int> hq((size_t)nw * 32, 0);
      std::copy(prog.leaf_hq.begin(), prog.leaf_hq.end(), hq.begin());
      d_lp_leaf_hq.upload(hq, stream);
      d_lp_s1q_0.alloc((size_t)std::max<long long>(n_pat_pad, 32));
      CK(cudaMemsetAsync(d_lp_s1q_0.p, 0, sizeof(int) * (size_t)std::max<long long>(n_pat_pad, 32), stream));
      ppm::launch_leaf_height_sums_q(md, d_lp_leaf_hq.p, count, d_lp_s1q_0.p, stream);
      CK(cudaStreamSynchronize(stream));
      md.lp_s1q_0 = d_lp_s1q_0.p;
    }
    // frontier lists of the class sums: pattern-only, so once per handle (count, prefix sums on the host, fill)
    md.fr_off = nullptr; md.fr_nodes = nullptr;
    if (tree.n_nodes <= 65535 && count > 0) {
      DevBuf<uint32_t> d_nf;
      d_nf.alloc((size_t)count);
      ppm::launch_frontier_count(md, count, d_nf.p, stream);
      CK(cudaGetLastError());  // (a failed launch would leave nf uninitialised)
      std::vector<uint32_t> nf((size_t)count), off((size_t)count + 1, 0u);
      CK(cudaMemcpyAsync(nf.data(), d_nf.p, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      unsigned long long total = 0;
      for (long long j = 0; j < count; j++) { off[j] = (uint32_t)total; total += nf[j]; }
      off[count] = (uint32_t)total;
      if (total < (1ull << 32)) {
        d_fr_off.upload(off, stream);
        d_fr_nodes.alloc((size_t)std::max<unsigned long long>(total, 1));
        ppm::launch_frontier_fill(md, count, d_fr_off.p, d_fr_nodes.p, stream);
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        md.fr_off = d_fr_off.p; md.fr_nodes = d_fr_nodes.p;
      }
    }
  }

  ppm::TablesDev tables(int P, const double* d_par) {
    if (P > P_cap) {
      size_t nn = tree.n_nodes;
      d_cls.alloc((size_t)P * nn); d_kap.alloc((size_t)P * nn); d_d1u.alloc((size_t)P * nn);
      d_recs.alloc((size_t)P * (prog.entry_ref.size() + 1)); d_yy.alloc((size_t)P * std::max(prog.n_blocks, 1) * prog.R);
      d_oprec.alloc((size_t)P * std::max<size_t>(prog.ops.size(), 1) * ppm::OPREC_DOUBLES);
      d_uu.alloc((size_t)P * std::max(prog.n_blocks, 1) * prog.R);
      d_leafab.alloc((size_t)P * tree.n_leaves * 2); d_pairq.alloc((size_t)P * std::max<size_t>(prog.n_pairs, 1) * 4);
      d_cherry.alloc((size_t)P * std::max<size_t>(prog.n_cherries, 1) * 4);
      d_clsT.alloc((size_t)P * nn); d_a01.alloc((size_t)P * std::max<long long>(n_pat_pad, 32));
      d_scal.alloc((size_t)P * ppm::SC_COUNT); d_zeta.alloc((size_t)P * 6 * nn);
      d_params.alloc((size_t)P * 8); d_ll.alloc(P); d_i2.alloc(P); d_i2_override.alloc(P);
      d_zrow.alloc((size_t)P * nn); d_zk.alloc((size_t)P * nn); d_zxb.alloc((size_t)P * (std::max(prog.n_blocks, 1) + 1));
      d_zpart.alloc((size_t)P * std::max(tree.n_slices, 1));
      if (prog.lp) { const size_t nbr = (size_t)P * std::max(prog.n_blocks, 1) * prog.R; d_slb.alloc(nbr); d_slce.alloc(nbr); }
      if (use_band) {
        d_bop.alloc((size_t)P * std::max<size_t>(band.ops.size(), 1) * ppm::BOP_DOUBLES);
        d_bY.alloc((size_t)P * band.S_pad); d_bU.alloc((size_t)P * band.S_pad);
        d_bleaf.alloc((size_t)P * tree.n_leaves * 2);
        if (band.leaf_pow) { d_bLb.alloc((size_t)P * band.S_pad); d_bLc.alloc((size_t)P * band.S_pad); d_bLe.alloc((size_t)P * band.S_pad); }
      }
      P_cap = P;
    }
    ppm::TablesDev t{};
    t.P = P; t.params = d_par; t.cls = d_cls.p; t.kap = d_kap.p; t.d1u = d_d1u.p; t.recs = d_recs.p; t.yy = d_yy.p; t.oprec = d_oprec.p; t.clsT = d_clsT.p; t.a01 = d_a01.p; t.uu = d_uu.p; t.leafab = d_leafab.p; t.pairq = d_pairq.p; t.cherry = d_cherry.p;
    t.scal = d_scal.p; t.zeta = d_zeta.p; t.et1 = nullptr;
    t.zrow = d_zrow.p; t.zk = d_zk.p; t.zxb = d_zxb.p; t.zpart = d_zpart.p;  // expectedTime1 tables only on request
    t.slb = d_slb.p; t.slce = d_slce.p;
    return t;
  }

  // prepass + zero row + i2 (functions.cpp:334-344).  After this, scal[p] is complete.
  // with_class_sums: a sweep follows; its class sums (which need the node prepass only) run on the side stream meanwhile
  void run_prepass(const ppm::TablesDev& t, const double* d_i2_ovr, cudaStream_t s, bool with_class_sums = false) {
    NvtxRange r("ppm:prepass");
    ppm::launch_prepass_nodes(md, t, s);
    if (with_class_sums && count > 0) {
      CK(cudaEventRecord(ev_nodes, s));
      CK(cudaStreamWaitEvent(side2, ev_nodes, 0));
      ppm::launch_class_sums(md, t, 0, count, side2);
      CK(cudaEventRecord(ev_class, side2));
      class_pending = true;
      n_launches += 1;
    }
    ppm::launch_prepass_tables(md, t, s);
    if (use_band && with_class_sums) { ppm::launch_prepass_band(md, t, bdev, band_tables(), s); n_launches += 1; }
    ppm::launch_zero_row(md, t, s);  // the all-zero row of computeI2: slices over threads, not a sweep item
    ppm::launch_prepass_i2(md, t, d_i2_ovr, s);
    n_launches += 6;
  }

  // mode 0: partial sums into d_out[P]; mode 2: categories/components for P == 1
  void run_sweep(const ppm::TablesDev& t, int mode, double* d_out, signed char* d_cat_out, double* d_comp_out, cudaStream_t s,
                 double* d_dist_out = nullptr) {
    NvtxRange r(use_band && mode == 0 && d_dist_out == nullptr ? "ppm:sweep(band)" : "ppm:sweep");
    ppm::SweepArgs a{};
    a.dist = d_dist_out;
    a.m = md; a.t = t; a.mode = mode; a.first = 0; a.count = count;
    a.n_batches = (count + 31) / 32; a.n_items = a.n_batches * t.P;
    a.cat = d_cat_out; a.comp = d_comp_out; a.force = (mode == 2);
    a.no_factored = hint_no_factored ? 1 : 0;
    size_t need = (size_t)std::max<long long>(a.n_items, 1);
    if (need > partial_cap) { d_partial.alloc(need); partial_cap = need; }
    a.partial = d_partial.p;
    const bool band_now = use_band && mode == 0 && d_dist_out == nullptr && a.n_items > 0;
    last_sweep_kind = band_now ? 1 : 0;
    if (band_now) {
      // band kernel for every parameter vector it can serve (SC_BAND), then the lane-per-pattern sweep for the others.
      // The class sums (4.4 ms at 5,000 genomes, on their own stream) are needed by the first MIXTURE only: the first batch's
      // prune (4 ms, latency-bound) and integral do not wait for them.
      bool class_wait = false;
      if (class_pending) { class_wait = band.stream; if (!class_wait) CK(cudaStreamWaitEvent(s, ev_class, 0)); class_pending = false; }
      else { ppm::launch_class_sums(md, t, 0, count, s); n_launches += 1; }
      int n_band_partials = 0;
      if (band.stream) {
        // batches of 32-item blocks through one scratch: prune -> integral -> mixture, back to back on the stream
        const long long n_blk_total = (long long)t.P * (n_pat_pad / 32);
        // scratch budget: every batch boundary drains the GPU once (~2 ms at 5,000 genomes), so batches should be big -- a
        // third of the free device memory, at most 48 GB, unless PPM_BAND_SCRATCH_MB says otherwise
        size_t budget = bs_budget;
        if (budget == 0) {
          size_t free_b = 0, total_b = 0;
          CK(cudaMemGetInfo(&free_b, &total_b));
          size_t have = 0;
          for (int k = 0; k < 2; k++) have += d_bs_nodes[k].n * sizeof(double2) + d_bs_xk[k].n * sizeof(int) + d_bs_unit[k].n * sizeof(double2);
          for (int k = 0; k < 2; k++) have += d_bs_n1k[k].n * sizeof(int) + d_bs_s1k[k].n * sizeof(double);
          budget = bs_budget = std::min<size_t>((size_t)48 << 30, std::max<size_t>((size_t)1 << 30, (free_b + have) / 3));
        }
        if (const char* e = std::getenv("PPM_BAND_SCRATCH_MB")) budget = (size_t)std::max(1, std::atoi(e)) << 20;
        // Two scratch buffers: the prune kernel of batch b + 1 (latency-bound: one thread walks 5,000 dependent ops) runs on
        // the side stream while the integral kernel of batch b fills the GPU.
        const size_t per_blk = ppm::band_stream_block_bytes(bdev);
        const int blk_batch = (int)std::max<long long>(1, std::min<long long>((n_blk_total + 1) / 2, (long long)(budget / 2 / per_blk)));
        if (blk_batch > bs_blk_cap) {
          for (int k = 0; k < 2; k++) {
            d_bs_nodes[k].alloc((size_t)blk_batch * bdev.n_int * 32);
            d_bs_xk[k].alloc((size_t)blk_batch * (bdev.n_int + 1) * 32);
            d_bs_unit[k].alloc((size_t)blk_batch * 32 * bdev.G);
            if (bdev.lp) d_bs_n1k[k].alloc((size_t)blk_batch * (bdev.n_int + 1) * 32);
            if (bdev.has_s1) d_bs_s1k[k].alloc((size_t)blk_batch * (bdev.n_int + 1) * 32);
          }
          bs_blk_cap = blk_batch;
        }
        if (sparse && !sp_lists_ready && !sp_lists_failed) build_sparse_lists();
        const bool sparse_now = sparse && sp_lists_ready;
        if (sparse_now) {
          d_bs_nodes0.alloc((size_t)2 * t.P * bdev.n_int * 32);
          d_bs_xk0.alloc((size_t)2 * t.P * (bdev.n_int + 1) * 32);
          if (bdev.lp) d_bs_n1k0.alloc((size_t)2 * t.P * (bdev.n_int + 1) * 32);
          if (bdev.has_s1) d_bs_s1k0.alloc((size_t)2 * t.P * (bdev.n_int + 1) * 32);
          d_bs_z1.alloc((size_t)t.P * bdev.S_pad);
          d_bs_unit_base.alloc((size_t)t.P * 64 * bdev.G);
        }
        n_band_partials = (int)(n_pat_pad / 32);
        const size_t needb = (size_t)t.P * n_band_partials;
        if (needb > partial_b_cap) { d_partial_b.alloc(needb); partial_b_cap = needb; }
        ppm::BandStreamArgs sa{};
        sa.m = md; sa.t = t; sa.b = bdev; sa.bt = band_tables(); sa.count = count; sa.count_pad = n_pat_pad; sa.force = 0;
        sa.ident = d_bs_ident.p; sa.partial = d_partial_b.p;
        if (const char* e = std::getenv("PPM_BAND_DBG")) sa.dbg = std::atoi(e);
        CK(cudaEventRecord(ev0, s));
        CK(cudaEventRecord(ev_bs_tables, s));            // the tables of this call are complete on s ...
        CK(cudaStreamWaitEvent(side, ev_bs_tables, 0));  // ... before the side stream prunes with them
        const long long n_batches = (n_blk_total + blk_batch - 1) / blk_batch;
        auto batch_args = [&](long long b) {
          ppm::BandStreamArgs x = sa;
          const int k = (int)(b & 1);
          x.nodes = d_bs_nodes[k].p; x.xk = d_bs_xk[k].p; x.unit_I = d_bs_unit[k].p; x.counter = d_bs_counter.p + k;
          x.n1k = d_bs_n1k[k].p; x.s1k = d_bs_s1k[k].p; x.n1_0 = md.freq; x.s1_0 = d_blp_s1_0.p;
          x.n1k0 = d_bs_n1k0.p; x.s1k0 = d_bs_s1k0.p;
          x.sparse = sparse_now ? 1 : 0; x.nodes0 = d_bs_nodes0.p; x.xk0 = d_bs_xk0.p; x.z1 = d_bs_z1.p;
          x.lists.off = d_sp_off.p; x.lists.recs = d_sp_recs.p; x.counter2 = d_bs_counter.p + 3 + k;
          // a sparse record costs about three dense ones (numerator and denominator, masked, every register)
          x.sparse_max_len = (unsigned int)std::max<long long>(64, (band.n_full16 + band.n_full32 + band.n_f16 + band.n_f32 + band.n_m16) / 3);
          if (const char* e = std::getenv("PPM_SPARSE_MAX_LEN")) x.sparse_max_len = (unsigned int)std::atoll(e);
          x.item0 = 32 * b * blk_batch;
          x.n_blk = (int)std::min<long long>(blk_batch, n_blk_total - b * blk_batch);
          return x;
        };
        if (sparse_now) {  // the baseline rows' records and the all-ones row's per-slice integrand, once per call, on the main
                           // stream next to the first batch's prune on the side stream
          ppm::BandStreamArgs z = sa;
          z.nodes0 = d_bs_nodes0.p; z.xk0 = d_bs_xk0.p; z.z1 = d_bs_z1.p; z.unit_I = d_bs_unit_base.p; z.counter = d_bs_counter.p + 2;
          z.n1k0 = d_bs_n1k0.p; z.s1k0 = d_bs_s1k0.p;
          ppm::launch_band_baseline_rows(z, d_bs_ones.p, d_bs_base_n1.p, d_bs_base_s1.p, n_sm, s);
          n_launches += 2;
        }
        ppm::launch_band_prune(batch_args(0), side);
        CK(cudaEventRecord(ev_bs_pruned[0], side));
        for (long long b = 0; b < n_batches; b++) {
          const int k = (int)(b & 1);
          if (b + 1 < n_batches) {
            if (b >= 1) CK(cudaStreamWaitEvent(side, ev_bs_done[k ^ 1], 0));  // integral b - 1 has released that buffer
            ppm::launch_band_prune(batch_args(b + 1), side);
            CK(cudaEventRecord(ev_bs_pruned[k ^ 1], side));
          }
          CK(cudaStreamWaitEvent(s, ev_bs_pruned[k], 0));
          if (sparse_now) {  // the listed items first (short, many), then the dense ones; the mixture joins them
            CK(cudaMemsetAsync(d_bs_counter.p + 3 + k, 0, sizeof(unsigned int), s));
            ppm::launch_band_sparse(batch_args(b), n_sm, s);
            n_launches += 1;
          }
          ppm::launch_band_integral(batch_args(b), n_sm, s, class_wait ? ev_class : nullptr);
          class_wait = false;
          CK(cudaEventRecord(ev_bs_done[k], s));
          n_launches += 3;
        }
      } else {
        ppm::BandArgs ba{};
        ba.m = md; ba.t = t; ba.b = bdev; ba.bt = band_tables(); ba.count = count; ba.force = 0;
        const ppm::BandPlan bpl = ppm::plan_band(md, bdev, count, t.P, n_sm);
        n_band_partials = bpl.n_chunks;
        const size_t needb = (size_t)t.P * bpl.n_chunks;
        if (needb > partial_b_cap) { d_partial_b.alloc(needb); partial_b_cap = needb; }
        ba.partial = d_partial_b.p;
        CK(cudaEventRecord(ev0, s));
        ppm::launch_band(ba, bpl, s);
        n_launches += 1;
      }
      if (!hint_all_band) {
        ppm::SweepPlan pl = ppm::plan_sweep(md, a.n_batches, t.P, n_sm, false);
        d_gslots.alloc(ppm::sweep_scratch_elems(md, pl)); a.gslots = d_gslots.p;
        a.fact_filter = 3;
        ppm::launch_sweep(a, pl, s);
        last_sweep_mode = pl.mode;
        n_launches += 1;
      }
      CK(cudaEventRecord(ev1, s));
      timed = true;
      ppm::launch_reduce2(d_partial.p, (int)a.n_batches, d_partial_b.p, n_band_partials, d_scal.p, t.P, d_out, s);
      n_launches += 1;
      CK(cudaGetLastError());
      return;
    }
    if (a.n_items > 0) {
      if (class_pending) { CK(cudaStreamWaitEvent(s, ev_class, 0)); class_pending = false; }
      else { ppm::launch_class_sums(md, t, 0, count, s); n_launches += 1; }
      ppm::SweepPlan pl = ppm::plan_sweep(md, a.n_batches, t.P, n_sm, d_dist_out != nullptr);
      last_sweep_mode = pl.mode;
      d_gslots.alloc(ppm::sweep_scratch_elems(md, pl)); a.gslots = d_gslots.p;
      CK(cudaEventRecord(ev0, s));
      ppm::launch_sweep(a, pl, s);
      CK(cudaEventRecord(ev1, s));
      timed = true;
      n_launches += 1;
    }
    ppm::launch_reduce(d_partial.p, t.P, (int)a.n_batches, d_out, s);
    n_launches += 1;
    CK(cudaGetLastError());
  }

  void extract_i2(const ppm::TablesDev& t, double* d_i2_out, double* d_parts_out, cudaStream_t s) {
    // strided copies out of scal
    if (d_i2_out)
      CK(cudaMemcpy2DAsync(d_i2_out, sizeof(double), d_scal.p + ppm::SC_I2, ppm::SC_COUNT * sizeof(double), sizeof(double), t.P,
                           cudaMemcpyDeviceToDevice, s));
    if (d_parts_out) {  // p0 p1 A B are contiguous in scal
      CK(cudaMemcpy2DAsync(d_parts_out, 4 * sizeof(double), d_scal.p + ppm::SC_P0, ppm::SC_COUNT * sizeof(double), 4 * sizeof(double),
                           t.P, cudaMemcpyDeviceToDevice, s));
    }
  }
};

// every parameter vector eligible for the band kernel (prepass: SC_BAND); known on the host for host-pointer calls
static bool all_band(const double* params, int P, double H) {
  for (int p = 0; p < P; p++) {
    const double g2 = params[8 * p + 4], l2 = params[8 * p + 5];
    const double lam = g2 + l2;
    const double pi1 = (g2 == 0.) ? 0. : g2 / lam;
    if (!(pi1 <= 0.95 && lam * H * 0.5 < 600.)) return false;
  }
  return true;
}
// pi1 = g2/(g2+l2) > 0.95 selects the factored leaf pairs (prepass: SC_FACTORED); known on the host for host-pointer calls
static bool none_factored(const double* params, int P) {
  for (int p = 0; p < P; p++) {
    const double g2 = params[8 * p + 4], l2 = params[8 * p + 5];
    const double pi1 = (g2 == 0.) ? 0. : g2 / (g2 + l2);
    if (!(pi1 <= 0.95)) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------- C ABI
#define GUARD_BEGIN try { if (h->device < 0) { h->err = "host-only handle (created with device < 0): no compute"; return PPM_ESTATE; }
#define GUARD_END(h)                                                  \
  }                                                                   \
  catch (const CudaError& e) { (h)->err = e.what(); return PPM_ECUDA; } \
  catch (const std::bad_alloc&) { (h)->err = "out of host memory"; return PPM_ENOMEM; } \
  catch (const std::exception& e) { (h)->err = e.what(); return PPM_EINVAL; }

// Unique patterns of a gene-row matrix.  With a device and enough rows the dedupe runs on the GPU (pattern_kernels.cu: hash,
// stable radix sort, verified runs); it is bit-identical to the host path, which remains for host-only handles, small
// inputs, dedupe = 0 and the (detected, never observed) case of a 64-bit hash collision.  PPM_GPU_DEDUPE_MIN: row threshold.
static ppm::Patterns make_patterns(const ppm::Tree& t, const uint32_t* rows, const int32_t* counts, int64_t n_rows, bool dedupe, int device) {
  long long min_rows = 50000;
  if (const char* e = std::getenv("PPM_GPU_DEDUPE_MIN")) min_rows = std::atoll(e);
  if (device >= 0 && dedupe && n_rows >= min_rows) {
    NvtxRange r("ppm:dedupe(gpu)");
    ppm::Patterns p;
    if (ppm::build_patterns_device(t, rows, counts, n_rows, device, p)) return p;
  }
  NvtxRange r("ppm:dedupe(host)");
  return ppm::build_patterns(t, rows, counts, n_rows, dedupe);
}
static ppm::Patterns make_patterns_from_file(const ppm::Tree& t, const char* path, bool dedupe, int device) {
  std::vector<uint32_t> rows;
  std::vector<int32_t> counts;
  const int64_t n = ppm::read_matrix_rows(t, path, rows, counts);
  return make_patterns(t, rows.data(), counts.empty() ? nullptr : counts.data(), n, dedupe, device);
}

static void finish_one(ppm_engine* e, int device, int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bad shard rank/world");
  e->shard_rank = rank; e->shard_world = world;
  if (device < 0) {  // host-only handle: model inspection, no compute
    e->device = -1;
    e->prog = ppm::build_program(e->tree, program_R(), false, leaf_pow_enabled());
    long long np = e->pat.n_patterns;
    e->first = np * rank / world;
    e->count = np * (rank + 1) / world - e->first;
  } else {
    // node ids travel as 16 bits through the frontier lists, the sweep records and the band program
    if (e->tree.n_nodes > 65535) throw std::runtime_error("trees with more than 32,768 genomes (65,535 nodes) are not supported on the device");
    e->init_device(device);
    e->upload_model();
  }
}
static int finish_create(ppm_handle* out, std::unique_ptr<ppm_engine>& e, int device, int rank, int world) {
  finish_one(e.get(), device, rank, world);
  *out = e.release();
  return PPM_OK;
}
// group: shard k of n on gpu_ids[k]; the head (shard 0) owns the host model, the peers share it
static int finish_create_group(ppm_handle* out, std::unique_ptr<ppm_engine>& e, const int32_t* gpu_ids, int n) {
  if (!gpu_ids || n < 1) throw std::runtime_error("group: need at least one GPU id");
  for (int k = 0; k < n; k++) if (gpu_ids[k] < 0) throw std::runtime_error("group: negative GPU id");
  finish_one(e.get(), gpu_ids[0], 0, n);
  for (int k = 1; k < n; k++) {
    std::unique_ptr<ppm_engine> peer(new ppm_engine(e->tree_own, e->pat_own));
    finish_one(peer.get(), gpu_ids[k], k, n);
    if (gpu_ids[k] != gpu_ids[0]) {  // direct NVLink stores into the head's gather buffer where the box allows it
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, gpu_ids[k], gpu_ids[0]) == cudaSuccess && can) {
        cudaSetDevice(gpu_ids[k]);
        cudaDeviceEnablePeerAccess(gpu_ids[0], 0);
      }
      cudaGetLastError();  // "already enabled" is fine; without peer access cudaMemcpyPeerAsync stages through the host
    }
    e->peers.push_back(std::move(peer));
  }
  *out = e.release();
  return PPM_OK;
}

extern "C" {

int ppm_create(ppm_handle* out, const ppm_tree* tree, const uint32_t* packed, const int32_t* counts, int64_t n_patterns,
               int dedupe, int device, int shard_rank, int shard_world) {
  if (!out || !tree || !packed) { g_create_error = "null argument"; return PPM_EINVAL; }
  *out = nullptr;
  std::unique_ptr<ppm_engine> e(new ppm_engine);
  try {
    e->tree = ppm::tree_from_arrays(tree->n_nodes, tree->left, tree->right, tree->branch_len);
    e->pat = make_patterns(e->tree, packed, counts, n_patterns, dedupe != 0, device);
    return finish_create(out, e, device, shard_rank, shard_world);
  } catch (const CudaError& x) { g_create_error = x.what(); return PPM_ECUDA; }
  catch (const std::bad_alloc&) { g_create_error = "out of host memory"; return PPM_ENOMEM; }
  catch (const std::exception& x) { g_create_error = x.what(); return PPM_EINVAL; }
}

int ppm_create_from_files(ppm_handle* out, const char* tree_path, const char* matrix_path, int dedupe, int device,
                          int shard_rank, int shard_world) {
  if (!out || !tree_path || !matrix_path) { g_create_error = "null argument"; return PPM_EINVAL; }
  *out = nullptr;
  std::unique_ptr<ppm_engine> e(new ppm_engine);
  try {
    try {
      e->tree = ppm::read_tree_file(tree_path);
      e->pat = make_patterns_from_file(e->tree, matrix_path, dedupe != 0, device);
    } catch (const std::runtime_error& x) {
      g_create_error = x.what();
      return (g_create_error.rfind("cannot open", 0) == 0) ? PPM_EIO : PPM_EINVAL;
    }
    return finish_create(out, e, device, shard_rank, shard_world);
  } catch (const CudaError& x) { g_create_error = x.what(); return PPM_ECUDA; }
  catch (const std::bad_alloc&) { g_create_error = "out of host memory"; return PPM_ENOMEM; }
  catch (const std::exception& x) { g_create_error = x.what(); return PPM_EINVAL; }
}

int ppm_create_group(ppm_handle* out, const ppm_tree* tree, const uint32_t* packed, const int32_t* counts, int64_t n_patterns,
                     int dedupe, const int32_t* gpu_ids, int32_t n_gpus) {
  if (!out || !tree || !packed) { g_create_error = "null argument"; return PPM_EINVAL; }
  *out = nullptr;
  std::unique_ptr<ppm_engine> e(new ppm_engine);
  try {
    e->tree = ppm::tree_from_arrays(tree->n_nodes, tree->left, tree->right, tree->branch_len);
    e->pat = make_patterns(e->tree, packed, counts, n_patterns, dedupe != 0, (gpu_ids && n_gpus > 0) ? gpu_ids[0] : -1);
    return finish_create_group(out, e, gpu_ids, n_gpus);
  } catch (const CudaError& x) { g_create_error = x.what(); return PPM_ECUDA; }
  catch (const std::bad_alloc&) { g_create_error = "out of host memory"; return PPM_ENOMEM; }
  catch (const std::exception& x) { g_create_error = x.what(); return PPM_EINVAL; }
}

int ppm_create_group_from_files(ppm_handle* out, const char* tree_path, const char* matrix_path, int dedupe,
                                const int32_t* gpu_ids, int32_t n_gpus) {
  if (!out || !tree_path || !matrix_path) { g_create_error = "null argument"; return PPM_EINVAL; }
  *out = nullptr;
  std::unique_ptr<ppm_engine> e(new ppm_engine);
  try {
    try {
      e->tree = ppm::read_tree_file(tree_path);
      e->pat = make_patterns_from_file(e->tree, matrix_path, dedupe != 0, (gpu_ids && n_gpus > 0) ? gpu_ids[0] : -1);
    } catch (const std::runtime_error& x) {
      g_create_error = x.what();
      return (g_create_error.rfind("cannot open", 0) == 0) ? PPM_EIO : PPM_EINVAL;
    }
    return finish_create_group(out, e, gpu_ids, n_gpus);
  } catch (const CudaError& x) { g_create_error = x.what(); return PPM_ECUDA; }
  catch (const std::bad_alloc&) { g_create_error = "out of host memory"; return PPM_ENOMEM; }
  catch (const std::exception& x) { g_create_error = x.what(); return PPM_EINVAL; }
}

void ppm_destroy(ppm_handle h) {
  if (!h) return;
  if (h->device >= 0) cudaSetDevice(h->device);
  delete h;
}

const char* ppm_last_error(ppm_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ppm_set_sparse(ppm_handle h, int on) {
  if (!h) return PPM_EINVAL;
  for (ppm_engine* e : h->members()) e->sparse = on != 0;
  return PPM_OK;
}

int ppm_write_ccc(ppm_handle h, const char* path) {
  if (!h || !path) return PPM_EINVAL;
  try {
    ppm::write_ccc_file(h->tree, h->pat, path);
    return PPM_OK;
  } catch (const std::exception& x) { h->err = x.what(); return PPM_EIO; }
}

int ppm_get_info(ppm_handle h, ppm_info* o) {
  if (!h || !o) return PPM_EINVAL;
  o->n_nodes = h->tree.n_nodes; o->n_leaves = h->tree.n_leaves;
  o->n_genes = h->pat.n_genes; o->n_patterns = h->pat.n_patterns; o->n_obs = h->pat.n_obs;
  o->n_words = h->pat.n_words; o->n_slices = h->tree.n_slices; o->n_terms = h->tree.n_terms;
  o->H = h->tree.H; o->L = h->tree.L; o->mean_genes = h->pat.mean_genes;
  o->device = h->device; o->shard_rank = h->shard_rank; o->shard_world = h->shard_world;
  o->shard_first = h->first; o->shard_count = h->count;
  o->group_size = (int32_t)h->peers.size() + 1;
  return PPM_OK;
}

int ppm_get_int_array(ppm_handle h, int which, int32_t* out, int64_t cap) {
  if (!h || !out) return PPM_EINVAL;
  const std::vector<int32_t>* v = nullptr;
  std::vector<int32_t> tmp;
  switch (which) {
    case 0: v = &h->pat.mrca; break;
    case 1: v = &h->pat.freq; break;
    case 2: v = &h->pat.counts; break;
    case 3: tmp.assign(h->tree.order.begin(), h->tree.order.end()); v = &tmp; break;
    case 4: v = &h->pat.gene_to_pattern; break;
    case 5: tmp.assign(h->tree.nsub.begin(), h->tree.nsub.end()); v = &tmp; break;
    case 6: tmp.assign(h->tree.alive.begin(), h->tree.alive.end()); v = &tmp; break;
    default: h->err = "unknown int array"; return PPM_EINVAL;
  }
  if ((int64_t)v->size() > cap) { h->err = "output buffer too small"; return PPM_EINVAL; }
  std::copy(v->begin(), v->end(), out);
  return PPM_OK;
}

int ppm_get_double_array(ppm_handle h, int which, double* out, int64_t cap) {
  if (!h || !out) return PPM_EINVAL;
  const std::vector<double>& v = which == 0 ? h->tree.height : h->tree.bl;
  if (which < 0 || which > 1) { h->err = "unknown double array"; return PPM_EINVAL; }
  if ((int64_t)v.size() > cap) { h->err = "output buffer too small"; return PPM_EINVAL; }
  std::copy(v.begin(), v.end(), out);
  return PPM_OK;
}

int ppm_loglik_device(ppm_handle h, const double* d_params, int32_t P, double* d_out_partial, double* d_out_i2, void* stream) {
  if (!h || !d_params || !d_out_partial || P < 1) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  // NULL is the legacy default stream, as everywhere in CUDA; PPM_STREAM_OWN selects the handle's private stream
  cudaStream_t s = (stream == PPM_STREAM_OWN) ? h->stream : (cudaStream_t)stream;
  h->n_launches = 0;
  ppm::TablesDev t = h->tables(P, d_params);
  h->run_prepass(t, nullptr, s, true);
  h->run_sweep(t, 0, d_out_partial, nullptr, nullptr, s);
  if (d_out_i2) h->extract_i2(t, d_out_i2, nullptr, s);
  return PPM_OK;
  GUARD_END(h)
}

int ppm_finalize_ll(const double* i2, double* ll, int32_t P) {
  if (!i2 || !ll) return PPM_EINVAL;
  for (int p = 0; p < P; p++)
    if (i2[p] < 0.) ll[p] = -1e9 * 1. + i2[p];
  return PPM_OK;
}

int ppm_finalize_ll_device(ppm_handle h, const double* d_i2, double* d_ll, int32_t P, void* stream) {
  if (!h || !d_i2 || !d_ll) return PPM_EINVAL;
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  ppm::launch_finalize(d_i2, d_ll, P, (stream == PPM_STREAM_OWN) ? h->stream : (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PPM_OK;
  GUARD_END(h)
}

// Group evaluation: every shard runs on its own GPU and stream (one host thread: all launches are asynchronous), each peer
// stores its P partial sums into the head's gather buffer over NVLink (cudaMemcpyPeerAsync on the peer's stream), the head
// waits on the peers' events, sums the shards in rank order (bitwise reproducible) and applies the i2 < 0 penalty.
static int group_loglik(ppm_handle h, const double* params, int32_t P, double* out_ll, double* out_i2) {
  GUARD_BEGIN
  const std::vector<ppm_engine*> mem = h->members();
  const int n = (int)mem.size();
  for (ppm_engine* e : mem) {
    CK(cudaSetDevice(e->device));
    e->tables(P, nullptr);
    CK(cudaMemcpyAsync(e->d_params.p, params, sizeof(double) * 8 * P, cudaMemcpyHostToDevice, e->stream));
    e->hint_no_factored = none_factored(params, P);
    e->hint_all_band = all_band(params, P, e->tree.H);
    int rc = ppm_loglik_device(e, e->d_params.p, P, e->d_ll.p, e->d_i2.p, (void*)e->stream);
    e->hint_no_factored = false; e->hint_all_band = false;
    if (rc) { if (e != h) h->err = "shard " + std::to_string(e->shard_rank) + ": " + e->err; return rc; }
  }
  CK(cudaSetDevice(h->device));
  h->d_gather.alloc((size_t)n * P);
  for (int k = 1; k < n; k++) {
    ppm_engine* e = mem[k];
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyPeerAsync(h->d_gather.p + (size_t)k * P, h->device, e->d_ll.p, e->device, sizeof(double) * P, e->stream));
    CK(cudaEventRecord(e->ev_done, e->stream));
  }
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_gather.p, h->d_ll.p, sizeof(double) * P, cudaMemcpyDeviceToDevice, h->stream));
  for (int k = 1; k < n; k++) CK(cudaStreamWaitEvent(h->stream, mem[k]->ev_done, 0));
  ppm::launch_sum_shards(h->d_gather.p, n, P, h->d_ll.p, h->stream);
  ppm::launch_finalize(h->d_i2.p, h->d_ll.p, P, h->stream);
  h->n_launches += 2;
  CK(cudaGetLastError());
  std::vector<double> i2(P);
  CK(cudaMemcpyAsync(out_ll, h->d_ll.p, sizeof(double) * P, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(i2.data(), h->d_i2.p, sizeof(double) * P, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int k = 1; k < n; k++) { CK(cudaSetDevice(mem[k]->device)); CK(cudaStreamSynchronize(mem[k]->stream)); }
  CK(cudaSetDevice(h->device));
  if (out_i2) std::copy(i2.begin(), i2.end(), out_i2);
  return PPM_OK;
  GUARD_END(h)
}

int ppm_loglik(ppm_handle h, const double* params, int32_t P, double* out_ll, double* out_i2) {
  if (!h || !params || !out_ll || P < 1) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  NvtxRange r("ppm_loglik");
  if (h->is_group()) return group_loglik(h, params, P, out_ll, out_i2);
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  h->tables(P, nullptr);
  CK(cudaMemcpyAsync(h->d_params.p, params, sizeof(double) * 8 * P, cudaMemcpyHostToDevice, h->stream));
  h->hint_no_factored = none_factored(params, P);
  h->hint_all_band = all_band(params, P, h->tree.H);
  int rc = ppm_loglik_device(h, h->d_params.p, P, h->d_ll.p, h->d_i2.p, (void*)h->stream);
  h->hint_no_factored = false; h->hint_all_band = false;
  if (rc) return rc;
  std::vector<double> i2(P);
  CK(cudaMemcpyAsync(out_ll, h->d_ll.p, sizeof(double) * P, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(i2.data(), h->d_i2.p, sizeof(double) * P, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->shard_world == 1) ppm_finalize_ll(i2.data(), out_ll, P);
  if (out_i2) std::copy(i2.begin(), i2.end(), out_i2);
  return PPM_OK;
  GUARD_END(h)
}

int ppm_compute_i2(ppm_handle h, const double* params, int32_t P, double* out_i2, double* out_parts) {
  if (!h || !params || !out_i2 || P < 1) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  h->n_launches = 0;
  h->tables(P, nullptr);
  CK(cudaMemcpyAsync(h->d_params.p, params, sizeof(double) * 8 * P, cudaMemcpyHostToDevice, h->stream));
  ppm::TablesDev t = h->tables(P, h->d_params.p);
  h->run_prepass(t, nullptr, h->stream);
  std::vector<double> sc((size_t)P * ppm::SC_COUNT);
  CK(cudaMemcpyAsync(sc.data(), h->d_scal.p, sc.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int p = 0; p < P; p++) {
    const double* s = &sc[(size_t)p * ppm::SC_COUNT];
    out_i2[p] = s[ppm::SC_I2];
    if (out_parts) { out_parts[4 * p] = s[ppm::SC_P0]; out_parts[4 * p + 1] = s[ppm::SC_P1]; out_parts[4 * p + 2] = s[ppm::SC_A]; out_parts[4 * p + 3] = s[ppm::SC_B]; }
  }
  return PPM_OK;
  GUARD_END(h)
}

// One shard: categories / components of its slice of the patterns into host buffers (either may be NULL)
static int eval_shard(ppm_handle h, const double* params8, double i2, signed char* out_cat, double* out_comp) {
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  h->n_launches = 0;
  h->tables(1, nullptr);
  CK(cudaMemcpyAsync(h->d_params.p, params8, sizeof(double) * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_i2_override.p, &i2, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  ppm::TablesDev t = h->tables(1, h->d_params.p);
  h->run_prepass(t, h->d_i2_override.p, h->stream, true);
  h->d_cat.alloc((size_t)std::max<long long>(h->count, 1));
  h->d_comp.alloc((size_t)std::max<long long>(h->count, 1) * 3);
  h->run_sweep(t, 2, h->d_ll.p, h->d_cat.p, h->d_comp.p, h->stream);
  if (out_comp && h->count) CK(cudaMemcpyAsync(out_comp, h->d_comp.p, sizeof(double) * 3 * h->count, cudaMemcpyDeviceToHost, h->stream));
  if (out_cat && h->count) CK(cudaMemcpyAsync(out_cat, h->d_cat.p, h->count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PPM_OK;
  GUARD_END(h)
}
// A plain handle evaluates its own shard; a group head evaluates every shard (contiguous slices in rank order), so the
// outputs cover all patterns.  Returns the number of patterns covered in *n_out.
static int eval_members(ppm_handle h, const double* params8, double i2, std::vector<signed char>* cat, std::vector<double>* comp) {
  long long total = 0;
  for (ppm_engine* e : h->members()) total += e->count;
  if (cat) cat->assign((size_t)std::max<long long>(total, 1), 0);
  if (comp) comp->assign((size_t)std::max<long long>(total, 1) * 3, 0.);
  long long off = 0;
  for (ppm_engine* e : h->members()) {
    int rc = eval_shard(e, params8, i2, cat ? cat->data() + off : nullptr, comp ? comp->data() + 3 * off : nullptr);
    if (rc) { if (e != h) h->err = "shard " + std::to_string(e->shard_rank) + ": " + e->err; return rc; }
    off += e->count;
  }
  if (h->device >= 0) cudaSetDevice(h->device);
  return PPM_OK;
}
static bool covers_all(ppm_handle h) { return h->shard_world == 1 || h->is_group(); }

int ppm_assign_cat(ppm_handle h, const double* params8, double i2, int8_t* out_cat) {
  if (!h || !params8 || !out_cat) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  if (!covers_all(h)) { h->err = "ppm_assign_cat needs an unsharded handle or a group"; return PPM_ESTATE; }
  std::vector<signed char> cat;
  int rc = eval_members(h, params8, i2, &cat, nullptr);
  if (rc) return rc;
  // scatter back to the genes of the input, in input order (inf_cat.txt has one value per input row)
  for (int64_t g = 0; g < h->pat.n_genes; g++) out_cat[g] = cat[h->pat.gene_to_pattern[g]];
  return PPM_OK;
}

int ppm_components(ppm_handle h, const double* params8, double i2, double* out) {
  if (!h || !params8 || !out) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  std::vector<double> comp;
  int rc = eval_members(h, params8, i2, nullptr, &comp);
  if (rc) return rc;
  long long total = 0;
  for (ppm_engine* e : h->members()) total += e->count;
  std::copy(comp.begin(), comp.begin() + 3 * total, out);
  return PPM_OK;
}

int ppm_proba_cat(ppm_handle h, const double* params8, double i2, double* out) {
  if (!h || !params8 || !out) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  if (!covers_all(h)) { h->err = "ppm_proba_cat needs an unsharded handle or a group"; return PPM_ESTATE; }
  std::vector<double> comp;
  int rc = eval_members(h, params8, i2, nullptr, &comp);
  if (rc) return rc;
  for (int64_t g = 0; g < h->pat.n_genes; g++) {
    const double* c = &comp[(size_t)h->pat.gene_to_pattern[g] * 3];
    double mx = std::max(c[0], std::max(c[1], c[2]));
    double den = mx + std::log(std::exp(c[0] - mx) + std::exp(c[1] - mx) + std::exp(c[2] - mx));
    for (int k = 0; k < 3; k++) out[3 * g + k] = c[k] - den;
  }
  return PPM_OK;
}

static int expected_time1_shard(ppm_handle h, const double* params8, double* out) {
  GUARD_BEGIN
  CK(cudaSetDevice(h->device));
  h->n_launches = 0;
  h->tables(1, nullptr);
  CK(cudaMemcpyAsync(h->d_params.p, params8, sizeof(double) * 8, cudaMemcpyHostToDevice, h->stream));
  ppm::TablesDev t = h->tables(1, h->d_params.p);
  h->d_et1.alloc((size_t)h->tree.n_nodes);
  t.et1 = h->d_et1.p;
  ppm::launch_prepass_nodes(h->md, t, h->stream);
  h->d_comp.alloc((size_t)std::max<long long>(h->count, 1) * 3);
  ppm::launch_gather(h->d_et1.p, h->d_mrca.p, h->count, h->d_comp.p, h->stream);
  h->n_launches += 2;
  CK(cudaGetLastError());
  if (h->count) CK(cudaMemcpyAsync(out, h->d_comp.p, sizeof(double) * h->count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PPM_OK;
  GUARD_END(h)
}
int ppm_expected_time1(ppm_handle h, const double* params8, double* out) {
  if (!h || !params8 || !out) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  long long off = 0;
  for (ppm_engine* e : h->members()) {
    int rc = expected_time1_shard(e, params8, out + off);
    if (rc) { if (e != h) h->err = "shard " + std::to_string(e->shard_rank) + ": " + e->err; return rc; }
    off += e->count;
  }
  if (h->device >= 0) cudaSetDevice(h->device);
  return PPM_OK;
}

static int expected_time2_shard(ppm_handle h, const double* params8, double* out_dist) {
  GUARD_BEGIN
  if (h->md.R != 8) { h->err = "ppm_expected_time2 needs the default program (PPM_R unset)"; return PPM_ESTATE; }
  CK(cudaSetDevice(h->device));
  h->n_launches = 0;
  h->tables(1, nullptr);
  CK(cudaMemcpyAsync(h->d_params.p, params8, sizeof(double) * 8, cudaMemcpyHostToDevice, h->stream));
  ppm::TablesDev t = h->tables(1, h->d_params.p);
  h->run_prepass(t, nullptr, h->stream);
  const size_t row = (size_t)h->tree.n_slices + 1;
  h->d_dist.alloc((size_t)std::max<long long>(h->count, 1) * row);
  h->d_cat.alloc((size_t)std::max<long long>(h->count, 1));
  h->d_comp.alloc((size_t)std::max<long long>(h->count, 1) * 3);
  h->run_sweep(t, 2, h->d_ll.p, h->d_cat.p, h->d_comp.p, h->stream, h->d_dist.p);
  if (h->count) CK(cudaMemcpyAsync(out_dist, h->d_dist.p, sizeof(double) * row * h->count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PPM_OK;
  GUARD_END(h)
}
int ppm_expected_time2(ppm_handle h, const double* params8, double* out_times, double* out_dist) {
  if (!h || !params8) { if (h) h->err = "bad argument"; return PPM_EINVAL; }
  if (out_times) std::copy(h->tree.slice_t.begin(), h->tree.slice_t.end(), out_times);  // host model: expectedTime2_times
  if (!out_dist) return PPM_OK;
  const size_t row = (size_t)h->tree.n_slices + 1;
  long long off = 0;
  for (ppm_engine* e : h->members()) {
    int rc = expected_time2_shard(e, params8, out_dist + (size_t)off * row);
    if (rc) { if (e != h) h->err = "shard " + std::to_string(e->shard_rank) + ": " + e->err; return rc; }
    off += e->count;
  }
  if (h->device >= 0) cudaSetDevice(h->device);
  return PPM_OK;
}

int ppm_expected_gains(ppm_handle h, double g2, double* out) {
  if (!h || !out) return PPM_EINVAL;
  *out = ppm::expected_gains(h->tree, g2);
  return PPM_OK;
}

int ppm_get_counters(ppm_handle h, double* out4) {
  if (!h || !out4) return PPM_EINVAL;
  try {
  out4[0] = h->n_launches;
  float ms = 0;
  if (h->timed) { CK(cudaEventSynchronize(h->ev1)); CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); }
  out4[1] = ms;
  const ppm::Tree& t = h->tree;
  out4[2] = 5.0 * (double)t.n_terms + 28.0 * (t.n_leaves - 1) + 2.0 * t.n_slices;
  out4[3] = (h->use_band && h->last_sweep_kind == 1) ? h->band.fp64_instr_per_eval : h->prog.fp64_instr_per_eval;
  return PPM_OK;
  GUARD_END(h)
}

int ppm_get_plan(ppm_handle h, int32_t* out16) {
  if (!h || !out16) return PPM_EINVAL;
  for (int k = 0; k < 16; k++) out16[k] = 0;
  out16[0] = h->use_band ? 1 : 0;
  out16[1] = h->use_band ? h->band.R : 0;
  out16[2] = h->use_band ? h->band.G : 0;
  out16[3] = h->last_sweep_kind;
  out16[4] = h->last_sweep_mode;
  out16[5] = h->md.word_stride ? 1 : 0;
  out16[6] = h->md.deep ? 1 : 0;
  out16[7] = h->md.lean ? 1 : 0;
  out16[8] = h->use_band ? h->band.n_levels : (int32_t)h->tree.lev_up_off.size() - 1;
  out16[9] = h->use_band ? h->band.n_tiles : 0;
  out16[13] = (h->use_band && h->band.stream) ? 1 : 0;
  out16[10] = h->prog.n_blocks;
  out16[11] = h->prog.n_slots;
  out16[12] = h->prog.leaf_u ? 1 : 0;
  out16[14] = h->prog.lp ? 1 : 0;
  out16[15] = (h->use_band && h->band.leaf_pow) ? 1 : 0;
  return PPM_OK;
}

int ppm_measure_fp64_peak(int device, double* out_tflops) {
  if (!out_tflops) return PPM_EINVAL;
  double v = ppm::run_fp64_peak(device);
  if (v <= 0) return PPM_ECUDA;
  *out_tflops = v;
  return PPM_OK;
}

}  // extern "C"
