// Dense wire format adapters for the dense directories (Code/sage+gat+diffpool, Code/eigengcn):
//   * tsg_dense_to_coo : zero-padded dense [B, R, C] matrices (adjacency `adj`, pooled adjacency
//     `adj_pool_k`, eigen-pooling operators `pool_adj_i_j`; wire format of
//     Code/sage+gat+diffpool/cross_val.py:163-184 and Code/eigengcn/graph_sampler.py:185-241)
//     -> COO triplets of the non-zeros in row-major order, ready for K1 (TSG_CSR_RAW).  The
//     reference multiplies these as dense [N,N] matrices (encoders.py:33, eigengcn/encoders.py:407):
//     6,849x more flops than the non-zeros need at N=1000.
//   * tsg_nodebn_{fwd,bwd} : `apply_bn` (encoders.py:134-138) on [B, N, F]: fresh BatchNorm1d(N) =
//     per-node statistics over (B, F), biased variance, eps 1e-5, no affine, always batch stats.
#include "common.cuh"

namespace tsg {

struct DenseArgs {
  const float* M; int B, R, C;
  const int64_t* nrows; const int64_t* ncols; const int64_t* row_off; const int64_t* col_off;
};

// warp per matrix row: count non-zeros among the valid columns
__global__ void __launch_bounds__(256) k_dense_count(DenseArgs a, int* cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = (int64_t)a.B * a.R;
  for (int64_t g = warp; g < total; g += nwarps) {
    const int b = (int)(g / a.R), r = (int)(g - (int64_t)b * a.R);
    int c = 0;
    if (r < (int)a.nrows[b]) {
      const int nc = (int)a.ncols[b];
      const float* p = a.M + ((int64_t)b * a.R + r) * a.C;
      for (int j = lane; j < nc; j += 32) c += (p[j] != 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[g] = c;
  }
}

struct CntF {
  const int* c;
  __device__ int operator()(int64_t i) const { return c[i]; }
};

__global__ void __launch_bounds__(256)
k_dense_fill(DenseArgs a, const int* off, int64_t* out_r, int64_t* out_c, float* out_w) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = (int64_t)a.B * a.R;
  for (int64_t g = warp; g < total; g += nwarps) {
    const int b = (int)(g / a.R), r = (int)(g - (int64_t)b * a.R);
    if (r >= (int)a.nrows[b]) continue;
    const int nc = (int)a.ncols[b];
    const float* p = a.M + ((int64_t)b * a.R + r) * a.C;
    int base = off[g];
    const int64_t gr = a.row_off[b] + r, gc0 = a.col_off[b];
    for (int j0 = 0; j0 < nc; j0 += 32) {
      const int j = j0 + lane;
      float v = j < nc ? p[j] : 0.f;
      unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
      if (v != 0.f) {
        int pos = base + __popc(m & ((1u << lane) - 1u));
        out_r[pos] = gr; out_c[pos] = gc0 + j; out_w[pos] = v;
      }
      base += __popc(m);
    }
  }
}

__global__ void k_store_count32(const int* src, int64_t* dst) { *dst = (int64_t)*src; }

// ---------------------------------- node-wise BN on [B, N, F] -------------------------------
__global__ void __launch_bounds__(256)
k_nodebn_fwd(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mean_out,
             float* __restrict__ rstd_out, int B, int N, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float cntf = (float)B * (float)F;
  for (int64_t i = warp; i < N; i += nwarps) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* p = x + ((int64_t)b * N + i) * F;
      for (int f = lane; f < F; f += 32) s += p[f];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / cntf;
    float q = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* p = x + ((int64_t)b * N + i) * F;
      for (int f = lane; f < F; f += 32) { float d = p[f] - mean; q += d * d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / cntf + 1e-5f);
    for (int b = 0; b < B; ++b) {
      const float* p = x + ((int64_t)b * N + i) * F;
      float* o = y + ((int64_t)b * N + i) * F;
      for (int f = lane; f < F; f += 32) o[f] = (p[f] - mean) * rstd;
    }
    if (lane == 0) { mean_out[i] = mean; rstd_out[i] = rstd; }
  }
}

__global__ void __launch_bounds__(256)
k_nodebn_bwd(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ rstd_in,
             float* __restrict__ dx, int B, int N, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float cntf = (float)B * (float)F;
  for (int64_t i = warp; i < N; i += nwarps) {
    float sg = 0.f, sgo = 0.f;
    for (int b = 0; b < B; ++b) {
      const int64_t o = ((int64_t)b * N + i) * F;
      for (int f = lane; f < F; f += 32) { float g = dy[o + f]; sg += g; sgo += g * y[o + f]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgo += __shfl_xor_sync(0xffffffffu, sgo, o);
    }
    const float mg = sg / cntf, mgo = sgo / cntf, rstd = rstd_in[i];
    for (int b = 0; b < B; ++b) {
      const int64_t o = ((int64_t)b * N + i) * F;
      for (int f = lane; f < F; f += 32) dx[o + f] = rstd * (dy[o + f] - mg - y[o + f] * mgo);
    }
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" size_t tsg_dense_to_coo_workspace_bytes(int64_t B, int64_t R) {
  size_t rows = (size_t)B * (size_t)R;
  return 2 * ws_bytes(rows + 2, 4) + ws_bytes(scan_ws_ints((int64_t)rows), 4) + 512;
}

extern "C" int tsg_dense_to_coo(const float* M, int64_t B, int64_t R, int64_t C,
                                const int64_t* nrows, const int64_t* ncols,
                                const int64_t* row_off, const int64_t* col_off,
                                int64_t* out_r, int64_t* out_c, float* out_w, int64_t capacity,
                                int64_t* out_count_dev, void* workspace, size_t workspace_bytes,
                                void* stream) {
  TSG_REQUIRE(B >= 0 && R > 0 && C > 0, "dense_to_coo: bad shape");
  TSG_REQUIRE(B * R < (int64_t)0x7fffffff, "dense_to_coo: too many rows");
  TSG_REQUIRE(out_count_dev, "dense_to_coo: null count");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) { cudaMemsetAsync(out_count_dev, 0, 8, st); return TSG_OK; }
  TSG_REQUIRE(M && nrows && ncols && row_off && col_off && out_r && out_c && out_w, "dense_to_coo: null pointer");
  if (workspace_bytes < tsg_dense_to_coo_workspace_bytes(B, R)) { set_error("dense_to_coo: workspace too small"); return TSG_EWORKSPACE; }
  (void)capacity;   // caller guarantees capacity >= number of non-zeros (<= B*R*C)
  Workspace ws(workspace, workspace_bytes);
  int64_t rows = B * R;
  int* cnt = ws.take<int>(rows + 2);
  int* off = ws.take<int>(rows + 2);
  int* scan_ws = ws.take<int>(scan_ws_ints(rows));
  DenseArgs a{M, (int)B, (int)R, (int)C, nrows, ncols, row_off, col_off};
  int grid = grid_for(rows, 8);
  k_dense_count<<<grid, 256, 0, st>>>(a, cnt);
  int rc = exclusive_scan(CntF{cnt}, rows, off, scan_ws, st);
  if (rc) return rc;
  k_store_count32<<<1, 1, 0, st>>>(off + rows, out_count_dev);
  k_dense_fill<<<grid, 256, 0, st>>>(a, off, out_r, out_c, out_w);
  return check_launch("dense_to_coo");
}

extern "C" int tsg_nodebn_fwd(const float* x, float* y, float* mean, float* rstd,
                              int64_t B, int64_t N, int64_t F, void* stream) {
  TSG_REQUIRE(B > 0 && N >= 0 && F > 0, "nodebn_fwd: bad shape");
  if (N == 0) return TSG_OK;
  TSG_REQUIRE(x && y && mean && rstd, "nodebn_fwd: null pointer");
  k_nodebn_fwd<<<grid_for(N, 8), 256, 0, (cudaStream_t)stream>>>(x, y, mean, rstd, (int)B, (int)N, (int)F);
  return check_launch("nodebn_fwd");
}

extern "C" int tsg_nodebn_bwd(const float* dy, const float* y, const float* rstd, float* dx,
                              int64_t B, int64_t N, int64_t F, void* stream) {
  TSG_REQUIRE(B > 0 && N >= 0 && F > 0, "nodebn_bwd: bad shape");
  if (N == 0) return TSG_OK;
  TSG_REQUIRE(dy && y && rstd && dx, "nodebn_bwd: null pointer");
  k_nodebn_bwd<<<grid_for(N, 8), 256, 0, (cudaStream_t)stream>>>(dy, y, rstd, dx, (int)B, (int)N, (int)F);
  return check_launch("nodebn_bwd");
}
