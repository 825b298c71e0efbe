// Host-side model of the PPM likelihood engine: Newick -> flat preorder arrays, 0/1 matrix -> bit-packed
// deduplicated patterns, and the topology-only "sweep program" the CUDA kernels execute.
// Mirrors what Inference::Inference builds (reference source/functions.cpp:76-156) -- see host_model.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ppm {

struct Tree {
  int n_nodes = 0, n_leaves = 0;
  std::vector<int> left, right, parent;  // preorder ids (objects.cpp:104-114); -1 = none
  std::vector<double> bl;                // branch above node (functions.cpp:95-96); bl[0] = 0
  std::vector<std::string> label;        // leaf labels, "" for internal nodes (objects.cpp:141-149)
  std::vector<double> height;            // H - depth (functions.cpp:106-109)
  std::vector<int> order, rank;          // nodes by increasing height (functions.cpp:112-114) and its inverse
  std::vector<int> last;                 // largest preorder id inside the subtree of each node
  std::vector<int> leaf_pos;             // node -> bit position (k-th leaf in preorder), -1 internal
  std::vector<int> leaf_node;            // bit position -> node id
  // level lists for level-synchronous recursions: up = by height in edges (leaves first), dn = by depth (root first)
  std::vector<int> lev_up_off, lev_up_nodes, lev_dn_off, lev_dn_nodes;
  // top-down lists: internal nodes first inside a level (leaves read their parent after the walk); lev_dn_chain[l] = 1 when
  // levels l and l + 1 hold exactly one internal node each (a caterpillar: one thread walks them without synchronising)
  std::vector<int> lev_dn_chain;
  int bfs_last = 0;                      // last node of a breadth-first walk (computeMRCA of an empty pattern)
  double H = 0, L = 0;                   // functions.cpp:108, objects.cpp:115-122
  // Mobile integral grid (functions.cpp:250-260)
  std::vector<int> nsub;                 // [n_nodes-1] slices of interval rank (0 where skipped)
  std::vector<int> slice_start;          // [n_nodes] prefix sums of nsub
  std::vector<int> alive;                // [n_nodes-1] |subtrees[rank]| (functions.cpp:130-141)
  std::vector<double> slice_t, slice_wt; // [S] midpoint times and weights w/nSub
  int n_slices = 0;
  int64_t n_terms = 0;                   // W
};

Tree parse_newick(const std::string& first_line);
Tree tree_from_arrays(int n_nodes, const int32_t* left, const int32_t* right, const double* bl);
Tree read_tree_file(const std::string& path);
// Inference::expectedGains (functions.cpp:593-615): expected number of Mobile gains along the tree, a function of the tree
// and g2 only.  The reference's double sum over (interval, alive lineage) telescopes per lineage: O(nodes).
double expected_gains(const Tree& t, double g2);

struct Patterns {
  int n_words = 0;
  int64_t n_genes = 0, n_patterns = 0, n_obs = 0;
  std::vector<uint32_t> words;        // [n_patterns][n_words] row-major, bit k = k-th leaf in preorder
  std::vector<int32_t> counts, mrca, freq;
  std::vector<int32_t> gene_to_pattern;
  double mean_genes = 0;
};

// rows: n_rows x n_words packed rows; counts may be null (all ones)
Patterns build_patterns(const Tree& t, const uint32_t* rows, const int32_t* counts, int64_t n_rows, bool dedupe);
Patterns read_matrix_file(const Tree& t, const std::string& path, bool dedupe);
// the parsing half of read_matrix_file: gene rows [n_genes][n_words] bit-packed (bit k = k-th leaf in preorder), the counts of a
// "ccc" header (empty otherwise); returns n_genes
int64_t read_matrix_rows(const Tree& t, const std::string& path, std::vector<uint32_t>& rows, std::vector<int32_t>& counts);
// unique patterns + counts in the reference's compressed ("ccc") matrix format
void write_ccc_file(const Tree& t, const Patterns& p, const std::string& path);
// build_patterns (dedupe = true) on the GPU (pattern_kernels.cu); false: two different rows share a hash, use the host path
bool build_patterns_device(const Tree& t, const uint32_t* rows, const int32_t* counts, int64_t n_rows, int device, Patterns& p);

// ---- sweep program ---------------------------------------------------------------------------------------
// The tree is swept once in height order.  Slices of the Mobile integral are grouped into blocks of at most
// R consecutive slices; before a block runs, every node born before the block's last slice is merged from
// its children ("node op"); then every lineage alive somewhere in the block contributes one "entry".
struct NodeOp {
  int32_t node;         // preorder id of the node being formed
  int32_t lnode, rnode; // preorder ids of its children (index the per-parameter node tables)
  int32_t lref, rref;   // >=0: slot holding the child's Mobile partials; <0: leaf, bit position = -(ref+1)
  int32_t dst;          // slot of the new node, -1 for the root (never alive in an interval)
  int32_t i0;           // first slice of the current block at which the node is alive (0..R); R = none
  int32_t pair;         // cherries of a u-form program (both children leaves): their static pair id, else -1
};
// What the sweep reads per node op (the NodeOp fields above serve the prepass and the class sums)
struct SweepOp {
  int32_t lslot, rslot;  // slot of an internal child (0 for a leaf child)
  int32_t lbit, rbit;    // leaf child: (pattern word << 5) | rotation that lands its bit on bit 4; 0 for an internal child
                         // cherry: (pattern word << 5) | rotation that lands its 2-bit observation pair on bits 4..5
  int32_t dst, i0;       // as NodeOp
  int32_t flags;         // bit 0: left child is a leaf, bit 1: right child is a leaf
  int32_t pair;          // as NodeOp
  // Events folded into the op (slice masks of the op's block, 0 = nothing to do): the new node's own alive slices when it
  // is born inside the block, and the alive slices of a child that dies with this op after being born in an EARLIER block
  // (a prefix of the block).  These lineages have no entry record in that block.
  int32_t mself, ml, mr;
  int32_t comb;  // 1 / 2: the left / right child's prefix and the new node's suffix are complementary and cover the block: one
                 // update with per-slice operands serves both (general ops only)
  int32_t hlq, hrq;  // leaf powers: heights of a leaf left / right child in units of Program::hunit (0 for an internal child)
  int32_t pad0, pad1;
};
struct BlockDescHost {  // entry segments: see BlockDesc in ppm_kernels.cuh
  int32_t slice0, nslices, op0, op1, e_slot0, e_pair0, e_leafu0, e_leaf0, e_pleaf0, e_pslot0, e_end;
  int32_t opc;  // [op0,opc): cherries served from the per-parameter cherry table, [opc,op1): general node ops
};
struct Program {
  int R = 8;
  int n_blocks = 0, n_slots = 0;
  bool deep = false;
  std::vector<BlockDescHost> blk;                 // [n_blocks+1]; blk[n_blocks] carries the node ops after the last block
  std::vector<NodeOp> ops;
  std::vector<SweepOp> sops;                      // [ops.size() + 1] (one padding record)
  std::vector<int32_t> cherry_op;                 // [n_cherries] op index of each cherry pair
  std::vector<double> op_tb;                      // [ops.size()] start time of the op's block (the K of its folded events)
  std::vector<int32_t> entry_ref;                 // see EntryRec::ref (ppm_kernels.cuh)
  std::vector<uint32_t> entry_mask;               // see EntryRec::mask
  std::vector<int32_t> entry_node, entry_block;   // node id / block of each entry (for the K table)
  // Device patterns are stored in PROGRAM leaf order: the two leaves of static pair p (u-form quadratics; cherries
  // are the first n_cherries pairs) own bits 2p, 2p+1, unpaired leaves follow.
  std::vector<int32_t> leaf_perm;                 // [n_leaves] preorder leaf index -> device bit position
  std::vector<int32_t> leaf_node_dev;             // [n_leaves] device bit position -> node id
  int n_pairs = 0, n_cherries = 0;
  bool leaf_u = false;                            // leaves are (numerically) at height 0: u-form entries are used
  int64_t n_full = 0, n_partial = 0;
  bool lean = false;                              // no K/Y-form leaf entries, no partial entry records (see sweep_items LEAN)
  // Leaf powers (see band_program.h): the program lists no leaf entries and folds no leaf events; the sweep keeps, per lane, the
  // number of present alive leaves (and the sum of their heights) and multiplies every slice by one power at the block's end.
  bool lp = false, has_s1 = false;
  double max_leaf_h = 0;
  std::vector<int32_t> n_alive_leaves;            // [n_slices]
  std::vector<double> htot;                       // [n_slices] sum of the alive leaves' heights
  std::vector<double> leaf_h;                     // [n_leaves] height by device bit position
  double hunit = 0;                               // the sweep counts leaf heights in integer units of max_leaf_h / 2^14
  std::vector<int32_t> leaf_hq;                   // [n_leaves] height by device bit position, in units of hunit
  // modelled FP64 instruction count per pattern-eval of the main kernel
  double fp64_instr_per_eval = 0;
};
// deep: pad every block's fully-alive lineage segment to a multiple of 4 records with padding records that refer
// to the constant row (slot index n_slots) -- needed by the global-memory sweep variant's four-deep prefetch
Program build_program(const Tree& t, int R, bool deep = false, bool leaf_pow = false);
// rows: n row-major bit-packed patterns in preorder leaf order (Patterns::words) -> the same rows in program leaf order
std::vector<uint32_t> permute_patterns(const Program& pg, const uint32_t* rows, long long n, int n_words);

}  // namespace ppm
