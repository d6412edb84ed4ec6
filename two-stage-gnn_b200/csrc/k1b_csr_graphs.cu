// K1b -- CSR construction for block-diagonal packed batches: one CTA per graph, everything in
// shared memory.
//
// Same contract as tsg_csr_build(TSG_CSR_GCN, unit weights) -- PyG gcn_norm reached from
// Code/sag/network.py:34,38,42 and layers.py:18: drop input self loops, append one loop per node LAST
// in its row, val = deg^-1/2[src] * deg^-1/2[dst], stable (COO) order inside every row, both
// orientations -- but exploiting that a PyG `Batch` keeps every graph's edges contiguous
// (filter_adj preserves order, so this holds at every pooling level):
//   * per-graph edge ranges come from tsg_edge_ptr (binary search on the monotone "graph of edge");
//   * the CSR segment of graph g starts at (non-loop edges of graphs < g) + node_ptr[g], so the only
//     device-wide step is a scan over G counters;
//   * counting, slot claiming and the rank-in-row stable placement all run on shared-memory arrays
//     (counts, fill cursors, local rowptr, slot -> edge tables): global memory sees the edge list
//     twice (second time from L2) and each CSR array once.  Integer atomics only (order independent),
//     ranks make the result deterministic and bit-identical to the generic K1.
// Graphs that exceed the shared-memory budget make the caller fall back to the generic K1.
#include <stdlib.h>
#include "common.cuh"

namespace tsg {

// eptr[g] = first edge e with row[e] >= node_ptr[g]   (edges grouped by graph => monotone predicate)
__global__ void k_edge_ptr(const int64_t* __restrict__ row, int64_t E_cap, const int64_t* __restrict__ E_dev,
                           const int64_t* __restrict__ node_ptr, int G, int64_t* __restrict__ eptr) {
  const int64_t E = dev_count(E_cap, E_dev);
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g <= G; g += gridDim.x * blockDim.x) {
    const int64_t key = node_ptr[g];
    int64_t lo = 0, hi = E;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (row[mid] < key) lo = mid + 1; else hi = mid;
    }
    eptr[g] = g == G ? E : lo;
  }
}

// Edge endpoints come either as the batch's global int64 node ids (PyG edge_index) or as graph-LOCAL
// int32 ids (what a TU file stores, and what the compact feeder ships): same kernels, two loaders.
constexpr int CSRG_SMALL_NODES = 1024;     // 28.7 KB of node arrays per CTA
constexpr int CSRG_SMALL_EDGES = 3072;     // + 36 KB of per-edge state: 3 CTAs per SM, no global round trips


template <typename IdxT> struct EdgeIdx;
template <> struct EdgeIdx<int64_t> { static __device__ __forceinline__ int local(int64_t v, int64_t n0) { return (int)(v - n0); } };
template <> struct EdgeIdx<int32_t> { static __device__ __forceinline__ int local(int32_t v, int64_t) { return v; } };

// non-loop edges per graph (+ its node count): the CSR segment length of the graph
template <typename IdxT>
__global__ void __launch_bounds__(256)
k_graph_nnz(const IdxT* __restrict__ row, const IdxT* __restrict__ col, const int64_t* __restrict__ eptr,
            const int64_t* __restrict__ node_ptr, int G, int* __restrict__ seg_len) {
  __shared__ int sm[33];
  const int g = blockIdx.x;
  const int64_t e0 = eptr[g], e1 = eptr[g + 1];
  int c = 0;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) c += (row[e] != col[e]);   // equal ids <=> self loop in either index space
  int tot;
  block_excl_scan(c, sm, &tot);
  if (threadIdx.x == 0) seg_len[g] = tot + (int)(node_ptr[g + 1] - node_ptr[g]);
}

struct SegLen {
  const int* v;
  __device__ int operator()(int64_t i) const { return v[i]; }
};

// Shared memory is sized per LAUNCH, so one 5,748-node graph in a DD batch used to pin every CTA at 161 KB = one
// CTA per SM for 3,504 graphs of ~277 nodes.  The host launches the kernel twice when the largest graph is big:
// graphs with n in (n_lo, max_nodes] are this launch's, the rest exit at once.
template <typename IdxT>
__global__ void __launch_bounds__(256)
k_csr_graph(const IdxT* __restrict__ row, const IdxT* __restrict__ col, const int64_t* __restrict__ eptr,
            const int64_t* __restrict__ node_ptr, const int* __restrict__ seg_off, int G, int64_t E_cap,
            int* __restrict__ rowptr, int* __restrict__ colidx, float* __restrict__ val, int* __restrict__ eid,
            int* __restrict__ t_rowptr, int* __restrict__ t_colidx, float* __restrict__ t_val, int* __restrict__ t_eid,
            int* __restrict__ slot_d, int* __restrict__ slot_s, int max_nodes, int n_lo, int n_hi, int edge_cap,
            int sorted_fast_path) {
  extern __shared__ int sh[];
  __shared__ int scan_sm[33];
  const int g = blockIdx.x;
  const int64_t n0 = node_ptr[g];
  const int n = (int)(node_ptr[g + 1] - n0);
  if (n > n_hi) __trap();                            // caller broke the size contract: fail loudly
  if (n <= n_lo || n > max_nodes) return;            // the other launch's graph
  const int64_t e0 = eptr[g];
  const int m = (int)(eptr[g + 1] - e0);
  const int base = seg_off[g];
  int* cnt_d = sh;                       // [max_nodes]  in-degree (non-loop)
  int* cnt_s = cnt_d + max_nodes;        // [max_nodes]  out-degree
  int* rp_d = cnt_s + max_nodes;         // [max_nodes+1] local rowptr, dst-major
  int* rp_s = rp_d + max_nodes + 1;      // [max_nodes+1]
  int* fil_d = rp_s + max_nodes + 1;     // [max_nodes]  fill cursors
  int* fil_s = fil_d + max_nodes;        // [max_nodes]
  float* dis = reinterpret_cast<float*>(fil_s + max_nodes);   // [max_nodes]
  // Per-edge state.  A graph with at most edge_cap edges keeps it in shared memory: its endpoints packed as
  // (r << 16 | c) after the one global read, and the two slot -> local edge id tables (arbitrary order inside a row).
  // The rank loops below read those tables deg times per edge: from L2 they were 57 % of the kernel's stall samples
  // (profiles/r01c_k1b_notes.md).  Bigger graphs use their slice of a global scratch array and re-read the edge
  // list from L2.  Measured and dropped (profiles/r01c_k1b_notes.md): sorting every row's segment by one thread +
  // one thread per output slot (coalesced stores): 214 us instead of 193; a table-free stable counting sort with
  // per-warp 16-bit histograms and match.any ranks (6x fewer instructions): 253 us -- match.any is the bottleneck.
  int* sh_edge = reinterpret_cast<int*>(dis + max_nodes);
  const bool es = m <= edge_cap && max_nodes <= 65536;
  int* tmp_d = es ? sh_edge + edge_cap : slot_d + e0;
  int* tmp_s = es ? sh_edge + 2 * edge_cap : slot_s + e0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { cnt_d[i] = 0; cnt_s[i] = 0; fil_d[i] = 0; fil_s[i] = 0; }
  __syncthreads();
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int r = EdgeIdx<IdxT>::local(row[e0 + e], n0), c = EdgeIdx<IdxT>::local(col[e0 + e], n0);
    if ((unsigned)r >= (unsigned)n || (unsigned)c >= (unsigned)n) __trap();   // edge leaves its graph: not a packed batch
    if (es) sh_edge[e] = (int)(((unsigned)r << 16) | (unsigned)c);
    if (r != c) { atomicAdd(&cnt_d[c], 1); atomicAdd(&cnt_s[r], 1); }
  }
  __syncthreads();
  // local exclusive scans of (count + 1 self loop); also the normalisation 1/sqrt(in-degree + 1)
  for (int pass = 0; pass < 2; ++pass) {
    const int* cnt = pass == 0 ? cnt_d : cnt_s;
    int* rp = pass == 0 ? rp_d : rp_s;
    int carry = 0;
    for (int b0 = 0; b0 < n; b0 += blockDim.x) {
      const int i = b0 + threadIdx.x;
      const int v = i < n ? cnt[i] + 1 : 0;
      int tot;
      const int ex = block_excl_scan(v, scan_sm, &tot);
      if (i < n) rp[i] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) rp[n] = carry;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    dis[i] = __fdiv_rn(1.0f, __fsqrt_rn((float)(cnt_d[i] + 1)));
  __syncthreads();
  for (int i = threadIdx.x; i <= n; i += blockDim.x) {
    if (i < n || g == G - 1) {           // the last graph also writes the closing entry
      rowptr[n0 + i] = base + rp_d[i];
      if (t_rowptr) t_rowptr[n0 + i] = base + rp_s[i];
    }
  }
  auto endpoints = [&](int e, int& r, int& c) {
    if (es) { const unsigned p = (unsigned)sh_edge[e]; r = (int)(p >> 16); c = (int)(p & 0xffffu); }
    else { r = EdgeIdx<IdxT>::local(row[e0 + e], n0); c = EdgeIdx<IdxT>::local(col[e0 + e], n0); }
  };
  // Coalesced edge lists -- sorted by (row, col), no duplicates, no self loops, every edge with its reverse: what
  // PyG's TUDataset, the TU loader and tsg_dense_to_coo produce -- need no sorting at all: the
  // src-major row r is the run of row r in the list, and by symmetry the dst-major row r holds the same neighbours
  // in the same order (its k-th entry is the edge c_k -> r, found as the reverse of the k-th edge of the run), so both
  // orientations are the list itself with one loop appended per row, written with consecutive threads on consecutive
  // slots.  The three properties are VERIFIED per graph (packed keys strictly increasing; reverse edge found by
  // binary search in its row's run); a graph that fails any of them takes the general path below.
  int general = 1;
  if (es && sorted_fast_path) {
    int bad = 0;
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
      const unsigned p = (unsigned)sh_edge[e];
      const int r = (int)(p >> 16), c = (int)(p & 0xffffu);
      if (r == c || (e > 0 && (unsigned)sh_edge[e - 1] >= p)) { bad = 1; continue; }
      int lo = rp_s[c] - c;                              // run of row c (valid when the list is sorted and loop free)
      const int run_end = rp_s[c + 1] - (c + 1);
      int hi = run_end;
      const unsigned key = ((unsigned)c << 16) | (unsigned)r;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((unsigned)sh_edge[mid] < key) lo = mid + 1; else hi = mid;
      }
      if (lo < run_end && (unsigned)sh_edge[lo] == key) tmp_d[e] = lo;       // id of the reverse edge
      else bad = 1;
    }
    general = __syncthreads_or(bad);
    if (!general) {
      for (int e = threadIdx.x; e < m; e += blockDim.x) {
        const unsigned p = (unsigned)sh_edge[e];
        const int r = (int)(p >> 16), c = (int)(p & 0xffffu);
        const float v = __fmul_rn(__fmul_rn(dis[r], 1.0f), dis[c]);      // == (dis[c] * 1) * dis[r]: commutative
        const int pos = base + e + r;                     // rp[r] + (e - first edge of row r), r loops before it
        colidx[pos] = (int)n0 + c; val[pos] = v;
        if (eid) eid[pos] = (int)(e0 + tmp_d[e]);         // dst-major entry (r <- c) is the edge c -> r
        if (t_rowptr) {
          t_colidx[pos] = (int)n0 + c; t_val[pos] = v;
          if (t_eid) t_eid[pos] = (int)(e0 + e);
        }
      }
    }
  }
  if (general) {
  // slot claim (arbitrary order) ...
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    int r, c;
    endpoints(e, r, c);
    if (r == c) continue;
    tmp_d[rp_d[c] - c + atomicAdd(&fil_d[c], 1)] = e;        // rp - index = offset without the loops
    tmp_s[rp_s[r] - r + atomicAdd(&fil_s[r], 1)] = e;
  }
  __syncthreads();
  {
    // ... then the rank among the row's edge ids gives the stable (COO-order) position
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
      int r, c;
      endpoints(e, r, c);
      if (r == c) continue;
      const float v = __fmul_rn(__fmul_rn(dis[r], 1.0f), dis[c]);
      {
        const int s = rp_d[c] - c, len = cnt_d[c];
        int rank = 0;
        for (int q = 0; q < len; ++q) rank += (tmp_d[s + q] < e);
        const int p = base + rp_d[c] + rank;
        colidx[p] = (int)n0 + r; val[p] = v;
        if (eid) eid[p] = (int)(e0 + e);
      }
      if (t_rowptr) {
        const int s = rp_s[r] - r, len = cnt_s[r];
        int rank = 0;
        for (int q = 0; q < len; ++q) rank += (tmp_s[s + q] < e);
        const int p = base + rp_s[r] + rank;
        t_colidx[p] = (int)n0 + c; t_val[p] = v;
        if (t_eid) t_eid[p] = (int)(e0 + e);
      }
    }
  }
  }   // general path
  for (int i = threadIdx.x; i < n; i += blockDim.x) {        // appended self loops: last slot of the row
    const float v = __fmul_rn(__fmul_rn(dis[i], 1.0f), dis[i]);
    const int p = base + rp_d[i + 1] - 1;
    colidx[p] = (int)n0 + i; val[p] = v;
    if (eid) eid[p] = (int)(E_cap + n0 + i);
    if (t_rowptr) {
      const int q = base + rp_s[i + 1] - 1;
      t_colidx[q] = (int)n0 + i; t_val[q] = v;
      if (t_eid) t_eid[q] = (int)(E_cap + n0 + i);
    }
  }
}


static size_t graph_smem_bytes(int64_t max_nodes, int64_t edge_cap = 0) {
  return (size_t)(6 * max_nodes + 2) * 4 + (size_t)max_nodes * 4 + (size_t)edge_cap * 12 + 64;
}

}  // namespace tsg

using namespace tsg;

extern "C" int tsg_edge_ptr(const int64_t* row, int64_t num_edges, const int64_t* num_edges_dev,
                            const int64_t* node_ptr, int64_t num_graphs, int64_t* edge_ptr, void* stream) {
  TSG_REQUIRE(num_edges >= 0 && num_graphs >= 0 && node_ptr && edge_ptr && (num_edges == 0 || row), "edge_ptr: bad arguments");
  k_edge_ptr<<<grid_for(num_graphs + 1, 128), 128, 0, (cudaStream_t)stream>>>(row, num_edges, num_edges_dev, node_ptr,
                                                                             (int)num_graphs, edge_ptr);
  return check_launch("edge_ptr");
}

extern "C" size_t tsg_csr_build_graphs_workspace_bytes(int64_t num_graphs, int64_t num_edges_cap) {
  return 2 * ws_bytes((size_t)num_graphs + 2, 4) + ws_bytes(scan_ws_ints(num_graphs), 4) +
         2 * ws_bytes((size_t)num_edges_cap + 1, 4) + 512;
}

namespace tsg {
static int sorted_env() {             // TSG_CSRG_SORTED=0: every graph takes the general (slot claim + rank) path
  static const int v = [] { const char* e = getenv("TSG_CSRG_SORTED"); return e ? atoi(e) : 1; }();
  return v;
}

static int edge_cap_env() {           // TSG_CSRG_EDGE_CAP=0 restores the global slot tables (A/B measurements)
  static const int v = [] { const char* e = getenv("TSG_CSRG_EDGE_CAP"); int x = e ? atoi(e) : CSRG_SMALL_EDGES; return x < 0 ? 0 : (x > 8192 ? 8192 : x); }();
  return v;
}

template <typename IdxT>
static int csr_build_graphs_impl(const IdxT* row, const IdxT* col, const int64_t* edge_ptr,
                                 const int64_t* node_ptr, int64_t num_graphs, int64_t num_nodes,
                                 int64_t num_edges_cap, int64_t max_graph_nodes,
                                 int32_t* rowptr, int32_t* colidx, float* val, int32_t* eid,
                                 int32_t* t_rowptr, int32_t* t_colidx, float* t_val, int32_t* t_eid,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  TSG_REQUIRE(num_graphs > 0 && num_nodes > 0, "csr_build_graphs: empty batch");
  TSG_REQUIRE(num_nodes + num_edges_cap < (int64_t)0x7fffffff, "csr_build_graphs: sum n + sum E must stay below 2^31");
  TSG_REQUIRE((num_edges_cap == 0 || (row && col)) && edge_ptr && node_ptr && rowptr && colidx && val, "csr_build_graphs: null pointer");
  TSG_REQUIRE(!t_rowptr || (t_colidx && t_val), "csr_build_graphs: null transposed output");
  const size_t smem_max = graph_smem_bytes(max_graph_nodes);
  TSG_REQUIRE(smem_max <= 200 * 1024, "csr_build_graphs: a graph with %lld nodes needs %zu B of shared memory",
              (long long)max_graph_nodes, smem_max);
  if (workspace_bytes < tsg_csr_build_graphs_workspace_bytes(num_graphs, num_edges_cap)) { set_error("csr_build_graphs: workspace too small"); return TSG_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  int* seg_len = ws.take<int>(num_graphs + 2);
  int* seg_off = ws.take<int>(num_graphs + 2);
  int* scan_ws = ws.take<int>(scan_ws_ints(num_graphs));
  int* slot_d = ws.take<int>(num_edges_cap + 1);
  int* slot_s = ws.take<int>(num_edges_cap + 1);
  k_graph_nnz<IdxT><<<(int)num_graphs, 256, 0, st>>>(row, col, edge_ptr, node_ptr, (int)num_graphs, seg_len);
  int rc = exclusive_scan(SegLen{seg_len}, num_graphs, seg_off, scan_ws, st);
  if (rc) return rc;
  const int hi = (int)max_graph_nodes;
  const int split = hi > CSRG_SMALL_NODES ? CSRG_SMALL_NODES : hi;
  size_t smem_attr = graph_smem_bytes(split, edge_cap_env());
  if (smem_max > smem_attr) smem_attr = smem_max;
  if (smem_attr > 48 * 1024) cudaFuncSetAttribute(k_csr_graph<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attr);
  // size classes [0, split] and (split, hi]
  for (int cls = 0; cls < (hi > split ? 2 : 1); ++cls) {
    const int lo = cls == 0 ? -1 : split, cap = cls == 0 ? split : hi;
    const int ecap = cls == 0 ? edge_cap_env() : 0;      // big graphs: all shared memory goes to the node arrays
    k_csr_graph<IdxT><<<(int)num_graphs, 256, graph_smem_bytes(cap, ecap), st>>>(
        row, col, edge_ptr, node_ptr, seg_off, (int)num_graphs, num_edges_cap, rowptr, colidx, val, eid,
        t_rowptr, t_colidx, t_val, t_eid, slot_d, slot_s, cap, lo, hi, ecap, sorted_env());
  }
  return check_launch("csr_build_graphs");
}
}  // namespace tsg

/* returns TSG_EINVAL with a message if a graph does not fit the shared-memory budget (caller falls back) */
extern "C" int tsg_csr_build_graphs(const int64_t* row, const int64_t* col, const int64_t* edge_ptr,
                                    const int64_t* node_ptr, int64_t num_graphs, int64_t num_nodes,
                                    int64_t num_edges_cap, int64_t max_graph_nodes,
                                    int32_t* rowptr, int32_t* colidx, float* val, int32_t* eid,
                                    int32_t* t_rowptr, int32_t* t_colidx, float* t_val, int32_t* t_eid,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  return csr_build_graphs_impl<int64_t>(row, col, edge_ptr, node_ptr, num_graphs, num_nodes, num_edges_cap, max_graph_nodes,
                                        rowptr, colidx, val, eid, t_rowptr, t_colidx, t_val, t_eid, workspace, workspace_bytes, stream);
}

/* same, for graph-LOCAL int32 endpoints (row[e], col[e] in [0, n_g)) with the per-graph edge offsets given */
extern "C" int tsg_csr_build_graphs_local(const int32_t* row, const int32_t* col, const int64_t* edge_ptr,
                                          const int64_t* node_ptr, int64_t num_graphs, int64_t num_nodes,
                                          int64_t num_edges_cap, int64_t max_graph_nodes,
                                          int32_t* rowptr, int32_t* colidx, float* val, int32_t* eid,
                                          int32_t* t_rowptr, int32_t* t_colidx, float* t_val, int32_t* t_eid,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  return csr_build_graphs_impl<int32_t>(row, col, edge_ptr, node_ptr, num_graphs, num_nodes, num_edges_cap, max_graph_nodes,
                                        rowptr, colidx, val, eid, t_rowptr, t_colidx, t_val, t_eid, workspace, workspace_bytes, stream);
}
