// Input pipeline on the device (kernel K2): gene rows -> unique patterns with multiplicities, in the order
// PAmatrix::compress produces (first occurrence; reference source/objects.cpp:254-273), plus the per-pattern integers the
// constructor of Inference derives from them: freq (functions.cpp:117-122) and mrca (computeMRCA, functions.cpp:46-65).
//
//   1. one warp per row: a position-aware 64-bit hash (sum over words of a mixed (word, position) value, so lanes can work
//      in any order), popcount, first and last set bit;
//   2. stable radix sort of (hash, row index): equal rows become one run, its first element the first occurrence;
//   3. run starts by a max-scan; every row is compared word by word with the head of its run -- a hash collision between
//      different rows is DETECTED, never trusted (the caller then falls back to the host path);
//   4. heads flagged in row order, exclusive scan = pattern id in first-occurrence order; counts by integer atomics;
//   5. mrca by binary lifting over the parent table (the LCA of the first and last present leaf, leaves in preorder).
// Integer work throughout: results are bit-identical to the host path (host_model.cpp: build_patterns), which the tests check.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_model.h"

namespace ppm {

namespace {

#define PK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
struct Buf {
  T* p = nullptr;
  explicit Buf(size_t n) { PK(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); }
  ~Buf() { if (p) cudaFree(p); }
  Buf(const Buf&) = delete;
  Buf& operator=(const Buf&) = delete;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__global__ void __launch_bounds__(256) row_stats_kernel(const uint32_t* __restrict__ rows, long long n_rows, int nw, uint64_t* hash,
                                                        int* freq, int* lo, int* hi) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_rows) return;
  const uint32_t* row = rows + (size_t)i * nw;
  uint64_t h = 0;
  int f = 0, l = 0x7fffffff, u = -1;
  for (int k = lane; k < nw; k += 32) {
    const uint32_t w = row[k];
    h += mix64(((uint64_t)(k + 1) << 32) | w);
    if (w) {
      f += __popc(w);
      l = min(l, 32 * k + __ffs(w) - 1);
      u = max(u, 32 * k + 31 - __clz(w));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
    l = min(l, __shfl_xor_sync(0xffffffffu, l, o));
    u = max(u, __shfl_xor_sync(0xffffffffu, u, o));
  }
  if (lane == 0) { hash[i] = mix64(h); freq[i] = f; lo[i] = f ? l : -1; hi[i] = u; }
}

__global__ void iota_kernel(int* v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int)i;
}
__global__ void run_heads_kernel(const uint64_t* key, long long n, int* start) {  // start[j] = j at a run's first element, else 0
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) start[j] = (j == 0 || key[j] != key[j - 1]) ? (int)j : 0;
}
// one warp per sorted position: the row equals the head of its run (else: collision); head_of[row] = head row
__global__ void __launch_bounds__(256) verify_runs_kernel(const uint32_t* __restrict__ rows, int nw, const int* idx, const int* run_start,
                                                          long long n, int* head_of, int* collision) {
  const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const int i = idx[j], r = idx[run_start[j]];
  bool same = true;
  if (i != r) {
    const uint32_t* a = rows + (size_t)i * nw;
    const uint32_t* b = rows + (size_t)r * nw;
    for (int k = lane; k < nw; k += 32) same = same && (a[k] == b[k]);
    same = __all_sync(0xffffffffu, same);
  }
  if (lane == 0) {
    head_of[i] = r;
    if (!same) atomicExch(collision, 1);
  }
}
__global__ void is_head_kernel(const int* head_of, long long n, int* flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = head_of[i] == (int)i ? 1 : 0;
}
__global__ void assign_kernel(const int* head_of, const int* pid, const int* counts_in, long long n, int* g2p, int* keep, int* counts_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = head_of[i], id = pid[r];
  g2p[i] = id;
  if (r == (int)i) keep[id] = (int)i;
  atomicAdd(&counts_out[id], counts_in ? counts_in[i] : 1);
}
// per unique pattern: freq and the MRCA of its present leaves (host_model.cpp: build_patterns, same walk)
__global__ void pattern_ints_kernel(const int* keep, const int* freq, const int* lo, const int* hi, long long n_pat, const int* up, int LOG,
                                    int n_nodes, const int* last, const int* parent, const int* leaf_node, int bfs_last, int* out_freq,
                                    int* out_mrca) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pat) return;
  const int i = keep[j], f = freq[i];
  out_freq[j] = f;
  int m;
  if (f == 0) m = bfs_last;
  else {
    int a = leaf_node[lo[i]];
    const int b = leaf_node[hi[i]];
    for (int k = LOG - 1; k >= 0; k--) {
      const int c = up[(size_t)k * n_nodes + a];
      if (last[c] < b) a = c;
    }
    m = (last[a] >= b) ? a : parent[a];
  }
  out_mrca[j] = m;
}

struct MaxOp {
  __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

template <typename T>
void h2d(Buf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  PK(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}
inline unsigned blocks_for(long long n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace

// rows: n_rows x n_words packed rows in host memory; counts may be null (all ones).  Fills `p` exactly as build_patterns
// (dedupe = true) would and returns true; returns false when two different rows share a hash (caller: host path).
bool build_patterns_device(const Tree& t, const uint32_t* rows, const int32_t* counts, int64_t n_rows, int device, Patterns& p) {
  const int nw = (t.n_leaves + 31) / 32;
  if (n_rows <= 0) throw std::runtime_error("matrix: no gene columns");
  if (n_rows >= (1ll << 31)) return false;
  if (t.n_leaves % 32) {  // bits beyond n_leaves must be clear
    const uint32_t mask = ~0u << (t.n_leaves % 32);
    for (int64_t i = 0; i < n_rows; i++)
      if (rows[(size_t)i * nw + nw - 1] & mask) throw std::runtime_error("patterns: bits set beyond n_leaves");
  }
  PK(cudaSetDevice(device));
  cudaStream_t s;
  PK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{s};
  const long long n = n_rows;

  Buf<uint32_t> d_rows((size_t)n * nw);
  PK(cudaMemcpyAsync(d_rows.p, rows, sizeof(uint32_t) * (size_t)n * nw, cudaMemcpyHostToDevice, s));
  Buf<uint64_t> d_key(n), d_key2(n);
  Buf<int> d_idx(n), d_idx2(n), d_freq(n), d_lo(n), d_hi(n), d_start(n), d_head_of(n), d_flag(n), d_pid(n), d_g2p(n), d_keep(n), d_cnt(n),
      d_coll(1);
  Buf<int> d_cin(counts ? n : 1);
  if (counts) PK(cudaMemcpyAsync(d_cin.p, counts, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  PK(cudaMemsetAsync(d_cnt.p, 0, sizeof(int) * n, s));
  PK(cudaMemsetAsync(d_coll.p, 0, sizeof(int), s));

  row_stats_kernel<<<blocks_for(n * 32, 256), 256, 0, s>>>(d_rows.p, n, nw, d_key.p, d_freq.p, d_lo.p, d_hi.p);
  iota_kernel<<<blocks_for(n, 256), 256, 0, s>>>(d_idx.p, n);
  {  // stable sort by hash: inside a run the row indices stay ascending
    size_t tmp_bytes = 0;
    PK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key.p, d_key2.p, d_idx.p, d_idx2.p, (int)n, 0, 64, s));
    Buf<unsigned char> tmp(tmp_bytes);
    PK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_key.p, d_key2.p, d_idx.p, d_idx2.p, (int)n, 0, 64, s));
    PK(cudaStreamSynchronize(s));
  }
  run_heads_kernel<<<blocks_for(n, 256), 256, 0, s>>>(d_key2.p, n, d_start.p);
  {
    size_t tmp_bytes = 0;
    PK(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, d_start.p, d_start.p, MaxOp(), (int)n, s));
    Buf<unsigned char> tmp(tmp_bytes);
    PK(cub::DeviceScan::InclusiveScan(tmp.p, tmp_bytes, d_start.p, d_start.p, MaxOp(), (int)n, s));
    PK(cudaStreamSynchronize(s));
  }
  verify_runs_kernel<<<blocks_for(n * 32, 256), 256, 0, s>>>(d_rows.p, nw, d_idx2.p, d_start.p, n, d_head_of.p, d_coll.p);
  is_head_kernel<<<blocks_for(n, 256), 256, 0, s>>>(d_head_of.p, n, d_flag.p);
  {
    size_t tmp_bytes = 0;
    PK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag.p, d_pid.p, (int)n, s));
    Buf<unsigned char> tmp(tmp_bytes);
    PK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, d_flag.p, d_pid.p, (int)n, s));
    PK(cudaStreamSynchronize(s));
  }
  assign_kernel<<<blocks_for(n, 256), 256, 0, s>>>(d_head_of.p, d_pid.p, counts ? d_cin.p : nullptr, n, d_g2p.p, d_keep.p, d_cnt.p);
  int collision = 0, last_pid = 0, last_flag = 0;
  PK(cudaMemcpyAsync(&collision, d_coll.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(&last_pid, d_pid.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(&last_flag, d_flag.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  PK(cudaStreamSynchronize(s));
  PK(cudaGetLastError());
  if (collision) return false;
  const long long n_pat = (long long)last_pid + last_flag;

  // tree tables for the MRCA walk
  int LOG = 1;
  while ((1 << LOG) < t.n_nodes) LOG++;
  std::vector<int> up((size_t)LOG * t.n_nodes);
  for (int v = 0; v < t.n_nodes; v++) up[v] = v ? t.parent[v] : 0;
  for (int k = 1; k < LOG; k++)
    for (int v = 0; v < t.n_nodes; v++) up[(size_t)k * t.n_nodes + v] = up[(size_t)(k - 1) * t.n_nodes + up[(size_t)(k - 1) * t.n_nodes + v]];
  Buf<int> d_up(up.size()), d_last(t.n_nodes), d_parent(t.n_nodes), d_leaf(t.n_leaves), d_pfreq(n_pat), d_pmrca(n_pat);
  h2d(d_up, up, s); h2d(d_last, t.last, s); h2d(d_parent, t.parent, s); h2d(d_leaf, t.leaf_node, s);
  pattern_ints_kernel<<<blocks_for(n_pat, 256), 256, 0, s>>>(d_keep.p, d_freq.p, d_lo.p, d_hi.p, n_pat, d_up.p, LOG, t.n_nodes, d_last.p, d_parent.p,
                                                             d_leaf.p, t.bfs_last, d_pfreq.p, d_pmrca.p);

  p = Patterns();
  p.n_words = nw; p.n_genes = n_rows; p.n_patterns = n_pat;
  p.gene_to_pattern.resize(n); p.counts.resize(n_pat); p.freq.resize(n_pat); p.mrca.resize(n_pat);
  std::vector<int> keep(n_pat);
  PK(cudaMemcpyAsync(p.gene_to_pattern.data(), d_g2p.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(p.counts.data(), d_cnt.p, sizeof(int) * n_pat, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(p.freq.data(), d_pfreq.p, sizeof(int) * n_pat, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(p.mrca.data(), d_pmrca.p, sizeof(int) * n_pat, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(keep.data(), d_keep.p, sizeof(int) * n_pat, cudaMemcpyDeviceToHost, s));
  PK(cudaStreamSynchronize(s));
  PK(cudaGetLastError());
  // the host keeps the unique rows (shards of group handles upload their slices from here)
  p.words.resize((size_t)n_pat * nw);
  for (long long j = 0; j < n_pat; j++) std::memcpy(&p.words[(size_t)j * nw], rows + (size_t)keep[j] * nw, sizeof(uint32_t) * nw);
  p.n_obs = 0;
  long double gene_sum = 0;
  for (long long j = 0; j < n_pat; j++) {
    p.n_obs += p.counts[j];                                  // functions.cpp:83
    gene_sum += (long double)p.freq[j] * p.counts[j];        // functions.cpp:123-126
  }
  p.mean_genes = (double)gene_sum / t.n_leaves;              // functions.cpp:127
  return true;
}

}  // namespace ppm
