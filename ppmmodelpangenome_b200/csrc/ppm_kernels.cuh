// Device-side data layout and kernel launchers of the PPM likelihood engine (sm_100a, FP64 CUDA cores).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "band_codes.h"

namespace ppm {

// Per-parameter-vector scalars written by the prepass (index into ParamScal::v)
enum {
  SC_PI1 = 0,   // g2/(g2+l2), 0 when g2 == 0 (functions.cpp:227-231)
  SC_PAD1,
  SC_LA0, SC_LD0,  // Mobile leaf partials for observation 0: a = m0 + pi1*(m1-m0), d = m1-m0 (16-byte aligned pair)
  SC_LA1, SC_LD1,  // ... observation 1
  SC_C0, SC_C1, SC_C2,  // log(N0/nObs), log(i1/(l1 nObs)), log(i2/nObs)  (functions.cpp:348,354,356)
  SC_I2, SC_P0, SC_P1, SC_A, SC_B,  // computeI2 (functions.cpp:334-344)
  SC_LOG1ME2,  // log(1 - s_10)
  SC_IZ_M, SC_IZ_E,  // zero-row Mobile integral as mantissa * 2^-exp
  SC_FACTORED,         // != 0: evaluate leaf pairs as two linear factors (pi1 close to 1: the expanded quadratic loses digits)
  SC_STOT0, SC_STOT1,  // sum over all branches of the "active" log factor S, categories 0 / 1
  SC_BAND,             // != 0: this parameter vector can be served by the band kernel (pi1 <= 0.95 and (g2+l2) H / 2 < 600)
  SC_COUNT = 22  // even: keeps the (a,d) pairs 16-byte aligned in every row
};

static_assert(SC_BAND < SC_COUNT && SC_COUNT % 2 == 0 && SC_LA0 % 2 == 0, "scal row layout");

struct NodeOpDev {  // == ppm::NodeOp
  int32_t node, lnode, rnode, lref, rref, dst, i0, pair;
};
struct __align__(16) SweepOpDev {  // == ppm::SweepOp (host_model.h): what the sweep reads per node op
  int32_t lslot, rslot, lbit, rbit, dst, i0, flags, pair;
  int32_t mself, ml, mr, comb;  // slice masks of the events folded into the op, and which child shares an update with the node (host_model.h)
  int32_t hlq, hrq, pad0, pad1; // leaf powers: quantised heights of leaf children
};
struct BlockDesc {  // == ppm::BlockDescHost; entry segments of a block, in order:
  // [e_slot0,e_pair0)   internal lineages alive in every slice (partials in slots)
  // [e_pair0,e_leafu0)  PAIRS of leaves alive in every slice, as one quadratic in u = 1 - exp(-(g2+l2) t)
  // [e_leafu0,e_leaf0)  single leaves alive in every slice, linear in u
  // [e_leaf0,e_pleaf0)  single leaves alive in every slice, block-relative K/Y form (trees with tall leaves only)
  // [e_pleaf0,e_pslot0) leaves merged inside the block | [e_pslot0,e_end) internal lineages born/merged inside it
  // node ops: [op0,opc) cherries read from the cherry table, [opc,op1) general merges
  int32_t slice0, nslices, op0, op1, e_slot0, e_pair0, e_leafu0, e_leaf0, e_pleaf0, e_pslot0, e_end, opc;
};
struct __align__(16) EntryRec {
  double K;       // pi1 * exp(-(g2+l2) * (t_block - h_lineage))   (per parameter vector)
  int32_t ref;    // slots: slot | next slot << 16;  K/Y leaves: word | next word << 16;  partial leaves and u-form
                  // leaves: bit position | next position << 16;  pairs: pair id  ("next" = prefetch hint)
  uint32_t mask;  // K/Y leaves: the leaf's bit inside its word;  pairs: position A | position B << 16;
                  // partial entries: code | live slices << 16;  otherwise unused
};
enum { OPREC_DOUBLES = 8 };  // per (parameter vector, node op): kap_l{x,y} kap_r{x,y} | K of the left child, the right child and the
                             // new node for the op's block (folded events) | unused

struct ModelDev {
  int n_nodes, n_leaves, n_words, n_slices, n_blocks, n_slots, n_entries, n_ops, R;
  int n_pairs;  // static leaf pairs (u-form quadratics)
  const int* leaf_node;  // [n_leaves] pattern bit position (program leaf order: pair p = bits 2p, 2p+1) -> node id
  int deep;  // program built for the global-memory variant: lineage segments padded to multiples of 4 records
  double H;
  long long n_obs;
  // tree (preorder ids)
  const int* left; const int* right; const int* parent;
  const double* bl; const double* height;
  // node levels for the prepass recursions: bottom-up (leaves = level 0) and top-down (root = level 0)
  int n_lev_up, n_lev_dn;
  const int* lev_up_off; const int* lev_up_nodes; const int* lev_dn_off; const int* lev_dn_nodes;
  const int* lev_dn_chain;  // [n_lev_dn + 1] levels l and l + 1 hold one internal node each (host_model.h)
  // slices
  const double* slice_t; const double* slice_wt_pad;  // [n_slices], [n_blocks*R] (0 for padding slices)
  // program
  const BlockDesc* blk;     // [n_blocks + 1]; the last one only carries the tail node ops
  const NodeOpDev* ops;
  const SweepOpDev* sops;   // [n_ops]
  const double* op_tb;      // [n_ops] start time of each op's block
  const int* cherry_op; int n_cherries;  // op index of each cherry (the first n_cherries static pairs)
  const int* entry_ref; const uint32_t* entry_mask; const int* entry_node; const int* entry_block;  // [n_entries]
  // patterns (this shard): words [n_words][n_pat_pad] (word-major), counts, mrca, freq
  const uint32_t* words; long long n_pat, n_pat_pad;
  int lean;         // the program has no K/Y-form leaf entries and no partial entry records: HOT kernels apply
  int word_stride;  // 0: every warp item copies its pattern words to shared memory [n_words][32]; else the sweep reads the device
                    // rows in place with this stride (= n_pat_pad; big trees, where the copy would cost resident warps)
  const int* counts; const int* mrca; const int* freq;
  // frontier of every pattern (roots of its maximal all-zero subtrees; class sums), computed once per handle:
  const uint32_t* fr_off;    // [n_pat + 1] offsets into fr_nodes, or null (trees with more than 65,535 nodes: walked per call)
  const uint16_t* fr_nodes;
  const uint32_t* zero_words;  // n_words * 32 zeros (the all-zero row of add0(), objects.cpp:274-281)
  // leaf-power sweep program (host_model.h: Program::lp): no leaf entries; per-slice count tables instead
  int lp, has_s1;
  const int* lp_alive; const double* lp_htot;  // [n_slices] leaves alive at each slice, sum of their heights
  const double* lp_s1_0;                       // [n_pat_pad] sum of the heights of every pattern's present leaves (has_s1; band kernels)
  const int* lp_s1q_0;                         // [n_pat_pad] the same in integer units of lp_hunit (lane-per-pattern sweep)
  double lp_hunit;
  double lp_max_h;  // leaf-power band program: largest |leaf height| (the prepass keeps parameter vectors with n (lam h)^2 > 1e-12 off it)
};

// Per-launch tables, one row per parameter vector
struct TablesDev {
  int P;
  const double* params;  // [P][8]
  double4* cls;          // [P][n_nodes]  {S0,T0,S1,T1} log-domain pure-loss factors of the branch above a node
  double2* kap;          // [P][n_nodes]  {pi1*E, (1-pi1)*E}, E = exp(-(g2+l2)*bl)
  double2* d1u;          // [P][n_nodes]  {log(V+U)-log V, log U}: Private "gained below the root" tables
  EntryRec* recs;        // [P][n_entries+1]
  double* yy;            // [P][n_blocks*R] exp(-(g2+l2)*(t_slice - t_block)), 0 for padding slices
  double* oprec;         // [P][n_ops][OPREC_DOUBLES]
  double* uu;            // [P][n_blocks*R] u_j = 1 - exp(-(g2+l2) t_j), 0 for padding slices
  double2* leafab;       // [P][n_leaves][2] leaf term = alpha + beta*u for observation 0 / 1
  double2* cherry;       // [P][n_cherries][4] (a, d) of a cherry node for the 4 observation pairs (left | right << 1)
  double4* pairq;        // [P][n_pairs][4] pair term = q.x + q.y*u + q.z*u^2 for the 4 bit combinations (A | B<<1)
  double2* clsT;         // [n_nodes][P] {cat0, cat1} frontier factors F_c = T_c - sum of S over the subtree of c
  double2* a01;          // [n_pat_pad][P] log p1(root) of the two pure-loss categories per (pattern, parameter vector)
  double* scal;          // [P][SC_COUNT]
  double* zeta;          // [P][6][n_nodes] scratch of the node prepass
  double2* zrow;         // [P][n_nodes] Mobile partials (a, d) of the ALL-ZERO row per node, mantissa-normalised ...
  int* zk;               // [P][n_nodes] ... and the renormalisation shift of each node (0 for leaves)
  int* zxb;              // [P][n_blocks+1] sum of the shifts of all nodes formed before each block
  double2* zpart;        // [P][n_slices] per-slice products of the all-zero row as (mantissa, exponent)
  double* slb; double2* slce;  // [P][n_blocks*R] leaf powers (sweep program): BandTables::bLb | (bLc, bLe)
  double* et1;           // [P][n_nodes] expected introduction time of a Private gene whose MRCA is the node (expectedTime1)
};

// sparse evaluation: 1 = term list against the all-zero row (present in at most an eighth of the genomes), 2 = against the
// all-ones row (absent in at most an eighth), 0 = dense (too few clean lineages either way)
__host__ __device__ inline int bs_sparse_item(const ModelDev& m, long long pat) {
  const long long f = m.freq[pat];
  return f * 8 <= m.n_leaves ? 1 : ((m.n_leaves - f) * 8 <= m.n_leaves ? 2 : 0);
}

// ---- band kernel (slice-parallel sweep for big trees; band_program.h) ------------------------------------------------
struct BandOpDev { uint16_t lref, rref, dst, flags; };   // == ppm::BandOp
enum { BOP_DOUBLES = 8 };  // per (parameter vector, band op): c0_l c1_l c0_r c1_r (child branch factors, see prepass_band_body),
                           // -exp(lam (h_node - H/2)), height of a leaf left child, of a leaf right child (leaf powers), unused
struct BandDev {
  int R, G, T, n_tiles, S_pad, n_int, n_ops, n_levels, n_rounds, n_desc, n_entries;
  const BandOpDev* ops; const int* op_nodes; const int* lev_off; const uint32_t* rounds;
  const uint16_t* nborn; const double* wt; const double* slice_t;
  const uint32_t* desc; const int* warp_off; const int* warp_ent_off; const uint32_t* entries;
  const uint32_t* blocks; const int* unit_blk_off;  // stream layout: descriptor blocks (band_program.h)
  const uint32_t* op_span; const uint32_t* leaf_span;  // alive slices js0 | js1 << 16 of every internal lineage (by record index) / leaf (by bit position)
  // leaf-power program (band_program.h): the lists hold internal lineages only
  int lp, has_s1;
  int chunk;  // stream layout: entries per chunk (band_codes.h)
  const int* n_alive_leaves; const double* htot;  // [S_pad] leaves alive at each slice and the sum of their heights
  const double* op_leaf_h;                        // [n_ops][2]
};
// sparse evaluation: which patterns have a term list.  Present in at most an eighth of the genomes: few dirty
// lineages (a present leaf below), the all-zero row serves the rest; the others have few clean lineages and stay dense.
struct SparseLists {       // pattern-only term lists, built once per handle (band_sparse.cu)
  const uint32_t* off;     // [count * n_tiles + 1] first record of every (pattern, tile)
  const uint2* recs;       // x: record index of an internal lineage, or leaf bit position | 1 << 31; y: s0 | s1 << 8 (slices [s0, s1) of the tile)
};
struct BandTables {     // one row per parameter vector, written by prepass_band
  double* bop;          // [P][n_ops][BOP_DOUBLES]
  double* bY;           // [P][S_pad]  pi1 * exp(-lam (t_j - H/2))
  double* bU;           // [P][S_pad]  u_j = 1 - exp(-lam t_j)
  double2* bleaf;       // [P][n_leaves][2]  single leaf at bit position pos observed 0 / 1: (a, -d exp(lam (h_leaf - H/2)))
  // leaf powers: log2 of the product of the alive leaves at slice j = bLb + n1 bLc + S1 bLe, n1 = present alive leaves, S1 = sum of their heights
  double* bLb;          // [P][S_pad]  n_alive log2 f_0(u_j) + g_0(j) sum of the alive leaves' heights
  double* bLc;          // [P][S_pad]  log2 f_1(u_j) - log2 f_0(u_j)
  double* bLe;          // [P][S_pad]  g_1(j) - g_0(j), g_x = d log2 f_x / d (leaf height) at height 0
};
struct BandArgs {
  ModelDev m; TablesDev t; BandDev b; BandTables bt;
  long long count;       // patterns of this shard
  int n_chunks;          // pattern chunks per parameter vector; unit = (parameter vector, chunk)
  long long chunk;       // patterns per chunk
  double* partial;       // [P][n_chunks]
  int force;
};
struct BandPlan { int grid_x, G, n_chunks; long long chunk; size_t smem_bytes; int ctas_per_sm; };
// stream layout (band_stream.cu): one batch = n_blk blocks of 32 consecutive items, item = p * count_pad + pattern
struct BandStreamArgs {
  ModelDev m; TablesDev t; BandDev b; BandTables bt;
  long long count, count_pad;  // patterns of this shard; padded to a multiple of 32 (== ModelDev::n_pat_pad)
  long long item0;             // first item of the batch (a multiple of 32)
  int n_blk;
  double2* nodes;              // scratch [n_blk][n_int][32]  lineage records (a, -d e^{lam (h - H/2)})
  int* xk;                     // scratch [n_blk][n_int + 1][32]  running sums of the renormalisation shifts, birth order
  int* n1k;                    // scratch [n_blk][n_int + 1][32]  leaf powers: present leaves that have died before each op ...
  double* s1k;                 // scratch [n_blk][n_int + 1][32]  ... and the sum of their heights (has_s1 only)
  const int* n1_0;             // [count_pad] present leaves of each pattern (ModelDev::freq; the baseline rows bring their own)
  const double* s1_0;          // [count_pad] sum of the heights of the present leaves (has_s1 only)
  double2* unit_I;             // scratch [n_blk * 32][G]  (mantissa, exponent) of every unit's share of the Mobile integral
  unsigned int* counter;       // unit hand-out
  const double2* ident;        // 32 bytes (1, 0, 0, 0): the record of a padding entry
  double* partial;             // [P][count_pad / 32]
  int force;
  int dbg;                     // experiments (PPM_BAND_DBG)
  // sparse evaluation: only the pattern's DIRTY lineages (a present leaf below) are multiplied, as a ratio
  // against the all-zero row whose per-slice integrand the prepass already has (TablesDev::zpart)
  int sparse;
  SparseLists lists;           // the handle's term lists
  unsigned int sparse_max_len; // a listed pattern is served by band_sparse_kernel only when its list holds at most this many records
                               // (a long list -- every ancestor of a present leaf on a caterpillar -- costs more than the dense walk)
  unsigned int* counter2;      // unit hand-out of the sparse kernel
  const double2* nodes0;       // [P][2][n_int][32]  lineage records of the all-zero / all-ones row per parameter vector (copy 0 of 32 is read)
  const int* xk0;              // [P][2][n_int + 1][32]  their shift prefixes
  const int* n1k0; const double* s1k0;  // [P][2][n_int + 1][32]  their dead-leaf counts / height sums (leaf powers)
  double2* z1;                 // [P][S_pad]  per-slice integrand (mantissa x weight, exponent) of the all-ones row; the zero row's is TablesDev::zpart
  int slice_out;               // != 0: the dense integral kernel serves copy 0 of every block only and stores its slices into z1
};

struct SweepArgs {
  ModelDev m;
  TablesDev t;
  int mode;              // 0: log-likelihood partial sums, 2: + categories/components (the all-zero row has its own kernel)
  long long first, count;  // pattern range
  long long n_batches, n_items;  // ceil(count/32) and P * n_batches
  int n_chunks, items_per_warp;  // CTA work unit = (parameter vector, chunk of warps*items_per_warp batches)
  double* partial;       // [P][n_batches]
  signed char* cat;      // [count] (mode 2, P == 1)
  double* comp;          // [count][3] (mode 2)
  double2* gslots;       // global slot storage when slots do not fit in shared memory
  int force;             // evaluate even when i2 < 0
  double* dist;          // [count][n_slices+1] per-slice Mobile integrand + the integral (expectedTime2_aux; DIST kernels only)
  int fact_filter;       // 0: every parameter vector; 1: only those with SC_FACTORED == 0 (HOT kernels); 2: only the factored ones;
                         // 3: only those the band kernel did not serve (SC_BAND == 0)
  int no_factored;       // host hint: no parameter vector of this launch is factored (skips the second launch)
  int stage_mask;        // global-memory variant: which small tables are staged in shared memory anyway (DS_* bits)
};
// small tables the global-memory variant stages per parameter vector when they fit (priority order)
enum { DS_PAIRQ = 1, DS_SLICES = 2, DS_CHERRY = 4, DS_LEAFAB = 8, DS_OPS = 16 };

// rows [count][nw] row-major in preorder leaf order -> out [nw][n_pat_pad] word-major in program leaf order; inv [nw * 32]:
// device bit position -> preorder leaf index (-1: no leaf)
void launch_permute_words(const uint32_t* rows, const int* inv, long long count, int nw, long long n_pat_pad, uint32_t* out, cudaStream_t s);
void launch_prepass_nodes(const ModelDev& m, const TablesDev& t, cudaStream_t s);
void launch_prepass_tables(const ModelDev& m, const TablesDev& t, cudaStream_t s);
void launch_class_sums(const ModelDev& m, const TablesDev& t, long long first, long long count, cudaStream_t s);
// once per handle: frontier sizes per pattern (nf_out) and, given their prefix sums, the lists themselves
void launch_frontier_count(const ModelDev& m, long long count, uint32_t* nf_out, cudaStream_t s);
void launch_frontier_fill(const ModelDev& m, long long count, const uint32_t* off, uint16_t* nodes_out, cudaStream_t s);
void launch_prepass_i2(const ModelDev& m, const TablesDev& t, const double* i2_override, cudaStream_t s);
// all-zero row of computeI2/getPobs2 (functions.cpp:300-333): slices spread over the threads; writes SC_IZ_M / SC_IZ_E
void launch_zero_row(const ModelDev& m, const TablesDev& t, cudaStream_t s);
// returns the grid size in x; smem_slots tells whether the shared-memory variant was used
struct SweepPlan { int grid_x, warps, items_per_warp, n_chunks, mode, stage_mask; size_t smem_bytes, per_warp_bytes; bool staged; };
SweepPlan plan_sweep(const ModelDev& m, long long n_batches, int P, int n_sm, bool dist = false);  // dist: a launch that writes SweepArgs::dist
bool sweep_is_staged(const ModelDev& m);  // depends on the tree/program sizes only
size_t sweep_scratch_elems(const ModelDev& m, const SweepPlan& pl);  // double2 elements of global slot scratch (0 if staged)
void launch_sweep(SweepArgs a, const SweepPlan& pl, cudaStream_t s);
void launch_reduce(const double* partial, int P, int nx, double* out, cudaStream_t s);
// band path: out[p] = SC_BAND[p] ? sum(band_partial[p][0..nb)) : sum(partial[p][0..nx))
void launch_reduce2(const double* partial, int nx, const double* band_partial, int nb, const double* scal, int P, double* out, cudaStream_t s);
void launch_prepass_band(const ModelDev& m, const TablesDev& t, const BandDev& b, const BandTables& bt, cudaStream_t s);
BandPlan plan_band(const ModelDev& m, const BandDev& b, long long count, int P, int n_sm);
size_t band_smem_bytes(const ModelDev& m, const BandDev& b);
void launch_band(BandArgs a, const BandPlan& pl, cudaStream_t s);
size_t band_stream_block_bytes(const BandDev& b);  // scratch per 32-item block
void launch_band_prune(const BandStreamArgs& a, cudaStream_t s);                 // one batch: lineage records into the scratch
void launch_band_integral(const BandStreamArgs& a, int n_sm, cudaStream_t s, cudaEvent_t before_mixture = nullptr);  // one batch: integral + mixture (which waits for the event)
// sparse evaluation: records of the all-zero and all-ones rows for every parameter vector into a.nodes0 / a.xk0 ([P][2][...]),
// per-slice integrand of the all-ones row into a.z1; base_words = [n_words][64] (32 zero columns, 32 all-ones columns);
// a.unit_I must hold P * 64 * G entries
// base_n1 / base_s1 [64]: present leaves / sum of their heights of the 64 baseline columns (leaf powers)
void launch_band_baseline_rows(BandStreamArgs a, const uint32_t* base_words, const int* base_n1, const double* base_s1, int n_sm, cudaStream_t s);
// leaf powers, once per handle: s1_0[pattern] = sum of leaf_h over the pattern's present leaves (leaf_h by device bit position)
void launch_leaf_height_sums(const ModelDev& m, const double* leaf_h, long long count, double* s1_0, cudaStream_t s);
void launch_leaf_height_sums_q(const ModelDev& m, const int* leaf_hq, long long count, int* s1q_0, cudaStream_t s);
// term lists: count pass (cnt [count * n_tiles + 1], turned into offsets in place; returns the number of records) and fill pass
unsigned long long sparse_lists_count(const ModelDev& m, const BandDev& b, long long count, uint32_t* cnt_off, cudaStream_t s);
void sparse_lists_fill(const ModelDev& m, const BandDev& b, long long count, const uint32_t* off, uint2* recs, cudaStream_t s);
void launch_band_sparse(const BandStreamArgs& a, int n_sm, cudaStream_t s);  // one batch: the items with term lists
void launch_finalize(const double* i2, double* ll, int P, cudaStream_t s);
void launch_sum_shards(const double* gather, int n, int P, double* out, cudaStream_t s);  // [n][P] -> [P], rank order
void launch_gather(const double* table, const int* index, long long n, double* out, cudaStream_t s);  // out[j] = table[index[j]]
double run_fp64_peak(int device);

}  // namespace ppm
