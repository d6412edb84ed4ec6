// K0 -- device-side batch assembly ("batch packing & feeder").
//
// Replaces, for the throughput path, PyG `Batch.from_data_list` (via DataLoader,
// Code/sag/train.py:181) and the per-graph host marshaling of the reference's triplet loops
// (`one_triplet[...].to(device)`, Code/sag/train_triplet.py:207; `torch.Tensor([ndarray]).cuda()`,
// Code/sage+gat+diffpool/tripletnet.py:18-33 -- 115-125 ms per graph upstream): the corpus lives in HBM
// once (1,168 DD graphs = 1.3 MB of labels + 12.6 MB of edges), a step sends only the list of graph ids,
// and one kernel writes the packed `x` (one-hot expansion of the node labels, or a row gather of dense
// features), the globally re-indexed `edge_index` and nothing else.  Pure streaming writes: HBM bound.
#include "common.cuh"

namespace tsg {

__global__ void __launch_bounds__(256)
k_pack_batch(const int64_t* __restrict__ ids, const int64_t* __restrict__ out_node_ptr,
             const int64_t* __restrict__ out_edge_ptr, const int64_t* __restrict__ c_node_ptr,
             const int64_t* __restrict__ c_edge_ptr, const int* __restrict__ c_row, const int* __restrict__ c_col,
             const int* __restrict__ c_label, const float* __restrict__ c_x, int F,
             float* __restrict__ x_out, int64_t* __restrict__ row_out, int64_t* __restrict__ col_out) {
  const int b = blockIdx.x;
  const int64_t g = ids[b];
  const int64_t n0 = c_node_ptr[g], n = c_node_ptr[g + 1] - n0;
  const int64_t e0 = c_edge_ptr[g], m = c_edge_ptr[g + 1] - e0;
  const int64_t on = out_node_ptr[b], oe = out_edge_ptr[b];
  // features: one row per (node, column) element, coalesced along the feature axis
  const int64_t total = n * F;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t v = i / F; const int f = (int)(i - v * F);
    float val;
    if (c_label != nullptr) val = (c_label[n0 + v] == f) ? 1.f : 0.f;
    else val = c_x[(n0 + v) * F + f];
    x_out[(on + v) * F + f] = val;
  }
  for (int64_t e = threadIdx.x; e < m; e += blockDim.x) {
    row_out[oe + e] = (int64_t)c_row[e0 + e] + on;
    col_out[oe + e] = (int64_t)c_col[e0 + e] + on;
  }
}

// compact form: the batch keeps the corpus representation (labels, graph-local int32 endpoints); only a gather
__global__ void __launch_bounds__(256)
k_pack_batch_compact(const int64_t* __restrict__ ids, const int64_t* __restrict__ out_node_ptr,
                     const int64_t* __restrict__ out_edge_ptr, const int64_t* __restrict__ c_node_ptr,
                     const int64_t* __restrict__ c_edge_ptr, const int* __restrict__ c_row, const int* __restrict__ c_col,
                     const int* __restrict__ c_label, int* __restrict__ label_out, int* __restrict__ row_out,
                     int* __restrict__ col_out) {
  const int b = blockIdx.x;
  const int64_t g = ids[b];
  const int64_t n0 = c_node_ptr[g], n = c_node_ptr[g + 1] - n0;
  const int64_t e0 = c_edge_ptr[g], m = c_edge_ptr[g + 1] - e0;
  const int64_t on = out_node_ptr[b], oe = out_edge_ptr[b];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) label_out[on + i] = c_label[n0 + i];
  for (int64_t e = threadIdx.x; e < m; e += blockDim.x) {
    row_out[oe + e] = c_row[e0 + e];
    col_out[oe + e] = c_col[e0 + e];
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" int tsg_pack_batch(const int64_t* ids, const int64_t* out_node_ptr, const int64_t* out_edge_ptr,
                              int64_t B, const int64_t* c_node_ptr, const int64_t* c_edge_ptr,
                              const int32_t* c_row, const int32_t* c_col, const int32_t* c_label,
                              const float* c_x, int64_t F, float* x_out, int64_t* row_out, int64_t* col_out,
                              void* stream) {
  TSG_REQUIRE(B >= 0 && F > 0 && B < (int64_t)0x7fffffff, "pack_batch: bad shape");
  if (B == 0) return TSG_OK;
  TSG_REQUIRE(ids && out_node_ptr && out_edge_ptr && c_node_ptr && c_edge_ptr && c_row && c_col && x_out && row_out && col_out,
              "pack_batch: null pointer");
  TSG_REQUIRE((c_label != nullptr) != (c_x != nullptr), "pack_batch: pass exactly one of labels / dense features");
  k_pack_batch<<<(int)B, 256, 0, (cudaStream_t)stream>>>(ids, out_node_ptr, out_edge_ptr, c_node_ptr, c_edge_ptr,
                                                          c_row, c_col, c_label, c_x, (int)F, x_out, row_out, col_out);
  return check_launch("pack_batch");
}

extern "C" int tsg_pack_batch_compact(const int64_t* ids, const int64_t* out_node_ptr, const int64_t* out_edge_ptr,
                                      int64_t B, const int64_t* c_node_ptr, const int64_t* c_edge_ptr,
                                      const int32_t* c_row, const int32_t* c_col, const int32_t* c_label,
                                      int32_t* label_out, int32_t* row_out, int32_t* col_out, void* stream) {
  TSG_REQUIRE(B >= 0 && B < (int64_t)0x7fffffff, "pack_batch_compact: bad shape");
  if (B == 0) return TSG_OK;
  TSG_REQUIRE(ids && out_node_ptr && out_edge_ptr && c_node_ptr && c_edge_ptr && c_row && c_col && c_label && label_out &&
              row_out && col_out, "pack_batch_compact: null pointer");
  k_pack_batch_compact<<<(int)B, 256, 0, (cudaStream_t)stream>>>(ids, out_node_ptr, out_edge_ptr, c_node_ptr, c_edge_ptr,
                                                                  c_row, c_col, c_label, label_out, row_out, col_out);
  return check_launch("pack_batch_compact");
}
