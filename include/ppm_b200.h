/*
 * ppm_b200.h -- C ABI of the B200-native PPM likelihood engine.
 *
 * Drop-in boundary for the objective function of JasmineGamblin/PPMmodelPangenome: the lambda handed to
 * nelder_mead (reference source/functions.cpp:404-412, :437-445) does
 *     tLik.clearLikValues(); return -logLik(par[0..5], {par[6],par[6],par[6],par[7]});
 * and the same engine state is read by computeI2 (functions.cpp:334-344; callers :397, :420, :453),
 * assignCat (:489-520) and nb0/nb1/nb2 (:522-541).  Each entry point below names what it replaces.
 *
 * Conventions: plain pointers and sizes only, no C++/torch types; every call returns 0 on success or a
 * PPM_E* code (message via ppm_last_error); no exception crosses the ABI; one handle owns one CUDA
 * device context, its streams and all device memory; calls on one handle are serialised by the caller
 * (the reference object is not re-entrant either).  There is NO CPU fallback: without a usable sm_100
 * device every compute entry point fails with PPM_ECUDA.
 *
 * Parameter vectors are 8 doubles in the reference's boundary order
 *     [N0, l0, i1, l1, g2, l2, s_10, s_01]            (functions.cpp:409; inference.cpp:43-44)
 * s_10 (false negative, the reference's error2) is used by all three categories, s_01 (false positive,
 * error1) by the Mobile category only (functions.cpp:356).
 */
#ifndef PPM_B200_H
#define PPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppm_engine* ppm_handle;

enum {
  PPM_OK = 0,
  PPM_EINVAL = 1,  /* bad argument / malformed input                                    */
  PPM_EIO = 2,     /* file could not be read                                             */
  PPM_ECUDA = 3,   /* CUDA error, no device, or the device is not sm_100                 */
  PPM_ENOMEM = 4,
  PPM_ESTATE = 5   /* call not valid in the handle's current state                       */
};

/* Flattened species tree in the reference's PREORDER numbering (RecTree::addNodeID, objects.cpp:104-114):
 * node 0 is the root, children have larger ids than their parent. */
typedef struct {
  int32_t n_nodes;          /* 2*n_leaves-1                                                          */
  const int32_t* left;      /* [n_nodes] id of the left child, -1 for a leaf   (objects.cpp:375)      */
  const int32_t* right;     /* [n_nodes] id of the right child, -1 for a leaf                         */
  const double* branch_len; /* [n_nodes] length of the branch above the node, [0] ignored (functions.cpp:95-96) */
} ppm_tree;

/* Model dimensions and constants, mirroring the private fields of Inference (functions.hpp:78-94). */
typedef struct {
  int32_t n_nodes, n_leaves;
  int64_t n_genes;       /* rows of the input matrix (nRep of an uncompressed input)   */
  int64_t n_patterns;    /* unique patterns held on the device (== n_genes without dedupe) */
  int64_t n_obs;         /* sum of multiplicities (nObs, functions.cpp:83)              */
  int32_t n_words;       /* 32-bit words per bit-packed pattern                          */
  int32_t n_slices;      /* S: midpoint-rule slices of the Mobile integral (functions.cpp:256) */
  int64_t n_terms;       /* W: sum over slices of alive lineages (functions.cpp:250-262) */
  double H, L, mean_genes; /* functions.cpp:108, :86, :123-127                           */
  int32_t device;        /* CUDA ordinal */
  int32_t shard_rank, shard_world;
  int64_t shard_first, shard_count; /* the slice of patterns this handle evaluates      */
  int32_t group_size;    /* GPUs behind this handle: 1, or n for a group head (ppm_create_group*) */
} ppm_info;

/* ---- construction ---------------------------------------------------------------------------------- */

/* Replaces Inference::Inference (functions.cpp:76-156) for callers that already hold flat data.
 * packed_patterns: n_patterns rows of n_words=ceil(n_leaves/32) uint32, row-major; bit k of a row is the
 * 0/1 observation of the k-th LEAF IN PREORDER (left-to-right in the Newick string).
 * counts: multiplicities (PAmatrix::counts, objects.cpp:175-187,210-213), NULL = all ones.
 * dedupe != 0 merges identical rows keeping first-occurrence order (PAmatrix::compress, objects.cpp:254-273).
 * device: CUDA ordinal; device < 0 creates a HOST-ONLY handle (model inspection through ppm_get_info /
 * ppm_get_*_array only; every compute call returns PPM_ESTATE).  shard_rank/shard_world: this handle evaluates the shard_rank-th of shard_world
 * contiguous slices of the (deduplicated) patterns; pass 0,1 for everything. */
int ppm_create(ppm_handle* out, const ppm_tree* tree, const uint32_t* packed_patterns, const int32_t* counts,
               int64_t n_patterns, int dedupe, int device, int shard_rank, int shard_world);

/* Same, reading the reference's input files: RecTree::readTree (objects.cpp:123-131, first line, binary,
 * every node ':length', internal labels skipped) and PAmatrix(path) (objects.cpp:164-214, incl. the "ccc"
 * compressed header). */
int ppm_create_from_files(ppm_handle* out, const char* tree_path, const char* matrix_path, int dedupe, int device,
                          int shard_rank, int shard_world);

/* Pattern-sharded evaluation on several GPUs of one box from ONE process: the handle owns one engine per
 * entry of gpu_ids (shard k of n_gpus on device gpu_ids[k]; ids may repeat), sharing one host model.  On such a GROUP
 * handle ppm_loglik evaluates all shards concurrently, each peer stores its P partial sums into the head's buffer over
 * NVLink (peer copy), the head adds them in rank order (bitwise reproducible) and applies the penalty: out_ll is the
 * complete objective, exactly as for an unsharded handle.  ppm_assign_cat / ppm_proba_cat / ppm_components /
 * ppm_expected_time1/2 cover every pattern.  ppm_compute_i2 and the *_device entry points act on shard 0 (device
 * gpu_ids[0]); with external plumbing (one process per GPU + NCCL) use ppm_create with shard_rank/shard_world instead. */
int ppm_create_group(ppm_handle* out, const ppm_tree* tree, const uint32_t* packed_patterns, const int32_t* counts,
                     int64_t n_patterns, int dedupe, const int32_t* gpu_ids, int32_t n_gpus);
int ppm_create_group_from_files(ppm_handle* out, const char* tree_path, const char* matrix_path, int dedupe,
                                const int32_t* gpu_ids, int32_t n_gpus);

void ppm_destroy(ppm_handle h);
const char* ppm_last_error(ppm_handle h); /* Sparse evaluation of the Mobile integral (nothing in the reference corresponds -- it is the reference's result by a
 * cheaper route): on trees served by the band kernels, a pattern's lineages without a present leaf below are not multiplied one
 * by one but taken from the all-zero row the prepass evaluates anyway (Inference::getPobs2, functions.cpp:300-333).  Off by default:
 * the dense evaluation is what the roofline numbers refer to.  Results agree with the dense evaluation to rounding (tests). */
int ppm_set_sparse(ppm_handle h, int on);
/* Writes the handle's unique patterns and their counts in the reference's compressed matrix format ("ccc,count,...": read by
 * PAmatrix::PAmatrix, source/objects.cpp:175-187; produced by compress.rep, simulation_aux.R:247-252), so that a later run -- of this
 * library or of the reference -- starts from the deduplicated matrix. */
int ppm_write_ccc(ppm_handle h, const char* path);
/* h may be NULL: error of the last failed create on this thread */
int ppm_get_info(ppm_handle h, ppm_info* out);

/* Integer model arrays for bit-exact checks against the reference's fields (functions.hpp:78-94):
 * which = 0 mrca[n_patterns] (preorder node id), 1 freq[n_patterns], 2 counts[n_patterns],
 *         3 order[n_nodes], 4 gene_to_pattern[n_genes], 5 nSub per interval rank [n_nodes-1],
 *         6 alive lineages per interval rank [n_nodes-1] (|subtrees[rank]|, functions.cpp:130-141). */
int ppm_get_int_array(ppm_handle h, int which, int32_t* out, int64_t capacity);
/* which = 0 height[n_nodes] (functions.cpp:106-109), 1 branch_len[n_nodes]. */
int ppm_get_double_array(ppm_handle h, int which, double* out, int64_t capacity);

/* ---- the objective ----------------------------------------------------------------------------------- */

/* Replaces clearLikValues()+Inference::logLik (functions.cpp:359-372, objective body :404-412) for P
 * parameter vectors at once: out_ll[p] = sum_patterns count*log-mixture, or -1e9+i2 when i2<0 (:362-364).
 * With shard_world>1 out_ll holds this shard's PARTIAL sum (and the penalty is NOT applied): add the
 * partials over shards, then call ppm_finalize_ll.  out_i2 may be NULL.  Host pointers. */
int ppm_loglik(ppm_handle h, const double* params, int32_t P, double* out_ll, double* out_i2);

/* Same with DEVICE pointers, asynchronous on `stream`: a cudaStream_t passed as void*.  NULL is the legacy default
 * stream, exactly as in the CUDA runtime (work is ordered against the caller's own default-stream work); pass
 * PPM_STREAM_OWN to use the handle's private non-blocking stream instead -- then the caller must order its reads and
 * writes of the buffers itself.  d_params [P][8], d_out_partial [P] per-shard partial sums (no penalty),
 * d_out_i2 [P].  This is the entry used with inputs resident in HBM and with an external all-reduce
 * (NCCL through torch.distributed) between it and ppm_finalize_ll_device. */
#define PPM_STREAM_OWN ((void*)(intptr_t)-1)
int ppm_loglik_device(ppm_handle h, const double* d_params, int32_t P, double* d_out_partial, double* d_out_i2,
                      void* stream);
/* ll[p] = i2[p] < 0 ? -1e9 + i2[p] : ll[p]   (functions.cpp:362-364), in place. */
int ppm_finalize_ll(const double* i2, double* ll, int32_t P);
int ppm_finalize_ll_device(ppm_handle h, const double* d_i2, double* d_ll, int32_t P, void* stream);

/* Replaces Inference::computeI2 (functions.cpp:334-344); out_p0p1AB [P][4] = p0, p1, A (getPobs1 :284-299),
 * B (getPobs2 :300-333), the inputs of nb0/nb1/nb2 (:522-541); may be NULL. */
int ppm_compute_i2(ppm_handle h, const double* params, int32_t P, double* out_i2, double* out_p0p1AB);

/* Replaces the per-pattern body of Inference::assignCat (functions.cpp:495-510): argmax over
 * {Persistent, Private (root + below-root), Mobile} with the reference's strict '>' tie-break, evaluated
 * at params8 with the given i2 (the reference passes its stored MLEi2).  out_cat has one value in {0,1,2}
 * per GENE of the input, in input order (shard_world must be 1). */
int ppm_assign_cat(ppm_handle h, const double* params8, double i2, int8_t* out_cat);

/* The three category log-terms of one parameter vector, per PATTERN of this shard:
 * out[3*r+0] = lik0r, out[3*r+1] = logSumExp(lik1r, lik1), out[3*r+2] = lik2 (functions.cpp:348-356, :497-505). */
int ppm_components(ppm_handle h, const double* params8, double i2, double* out /*[shard_count*3]*/);

/* Replaces the per-pattern body of Inference::computeProbaCat (functions.cpp:556-581): the log posterior
 * probabilities of the three categories, out[3*g+k] = lik_k - logSumExp(lik_0, lik_1, lik_2), one row per GENE of
 * the input in input order (shard_world must be 1).  The reference writes them as "p0 p1 p2\n" per row. */
int ppm_proba_cat(ppm_handle h, const double* params8, double i2, double* out /*[n_genes*3]*/);

/* Replaces Inference::expectedTime1 / expectedTime1_store (functions.cpp:637-708): the expected introduction time of
 * each PATTERN of this shard were it a Private gene, at l1 = params8[3], e = params8[6].  It depends on the pattern only
 * through its MRCA, so the engine fills a per-node table with a top-down recursion and gathers by mrca. */
int ppm_expected_time1(ppm_handle h, const double* params8, double* out /*[shard_count]*/);

/* Replaces Inference::expectedTime2_times (functions.cpp:709-725: out_times[n_slices], the slice mid-points; may be
 * NULL) and expectedTime2_aux + the denominator column of expectedTime2_store (:727-757): out_dist[r*(n_slices+1) + j]
 * = the Mobile integrand exp(f_aux) of pattern r at slice j, out_dist[r*(n_slices+1) + n_slices] = exp(likRep2) (the
 * integral).  out_dist may be NULL. */
int ppm_expected_time2(ppm_handle h, const double* params8, double* out_times, double* out_dist);

/* Replaces Inference::expectedGains (functions.cpp:593-615): expected number of Mobile gains along the tree for g2
 * (a function of the tree only; evaluated by the host model, no GPU needed). */
int ppm_expected_gains(ppm_handle h, double g2, double* out);

/* ---- measurement helpers ---------------------------------------------------------------------------- */

/* Counters of the last ppm_loglik*(): [0] kernels launched, [1] device ms of the main kernel (CUDA events
 * on the launching stream; valid after the stream is synchronised), [2] algorithmic flops per pattern-eval
 * F = 5W + 28(n-1) + 2S, [3] FP64 instructions per pattern-eval actually issued by the main
 * kernel's arithmetic (model count). */
int ppm_get_counters(ppm_handle h, double* out4);
/* Which kernels serve this handle (tests assert that every variant is exercised; nothing in the reference corresponds):
 * [0] band kernel in use (slice-parallel sweep of big trees), [1] its slices per lane R, [2] work units (stream layout)
 * or warps (one-CTA layout) per item G,
 * [3] kind of the last objective launch (0 lane-per-pattern sweep, 1 band kernel + fallback launch),
 * [4] variant of the last lane-per-pattern sweep (0 shared memory, 1/3 tensor memory, 2/4 global memory; -1 none yet),
 * [5] pattern rows read in place, [6] program built for the global-memory variant, [7] lean program (HOT kernels),
 * [8] pruning levels, [9] band tiles, [10] sweep blocks, [11] lineage slots, [12] leaves at height 0,
 * [13] band kernels in the stream layout (prune / integral / mixture, band_stream.cu),
 * [14] / [15] the lane-per-pattern sweep program / the band program use leaf powers (the alive leaves of a slice as one power of
 * the count of present alive leaves; ultrametric trees). */
int ppm_get_plan(ppm_handle h, int32_t* out16);
/* DFMA micro-benchmark: measured FP64 FMA throughput of the device in TFLOP/s (2 flop per FMA). */
int ppm_measure_fp64_peak(int device, double* out_tflops);

/* ---- input generation (benchmarks; host code, no GPU) ----------------------------------------------------- */

/* PPM gene simulator: the model of the reference's R scripts (simulation_aux.R: sim.N :92-118, sim.I1 :121-170, sim.I2
 * :173-238; run_simulations.R:63-78) on a flat preorder tree.  frac3 = fractions of Persistent / Private / Mobile genes
 * (genes are kept in that order), rates_L = {l0, l1, g2, l2} times the tree length L (the reference's parametrisation:
 * 1, 20, 50, 2500), s_10 / s_01 the observation error rates; all-zero genes are redrawn.  out_rows: n_genes rows of
 * ceil(n_leaves/32) uint32 in the ppm_create format; out_truth10 (may be NULL): N0 l0 i1 l1 g2 l2 s_10 s_01 H L of the
 * simulation.  Gene g uses its own random stream (seed, g): the result does not depend on n_threads (0 = all cores). */
int ppm_simulate_genes(const ppm_tree* tree, int64_t n_genes, uint64_t seed, const double* frac3, const double* rates_L,
                       double s_10, double s_01, int32_t n_threads, uint32_t* out_rows, double* out_truth10);

#ifdef __cplusplus
}
#endif
#endif /* PPM_B200_H */
