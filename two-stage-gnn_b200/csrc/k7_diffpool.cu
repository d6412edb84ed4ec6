// K7 -- DiffPool per-graph dense contractions on the packed layout.
//
// Replaces SoftPoolingGcnEncoder.forward lines 374-375 (Code/sage+gat+diffpool/encoders.py):
//     x   = S^T . Z          [B,K,N] x [B,N,D]
//     adj = S^T . adj . S    [B,K,N] x [B,N,N] x [B,N,K]
// which the reference evaluates as dense 1000x1000 padded batched matmuls (239 MFLOP per graph as
// written, 10.8 MFLOP on the real rows).  Here T = A.S is a K2 SpMM and the two remaining products
// are ONE per-graph contraction over the graph's real rows:
//     C_g [K, D+K] = S_g^T . [Z_g | T_g]        ("segment contraction", tsg_seg_contract)
// plus the per-graph row-local products (tsg_seg_linear) used by the post-pool tower
// (y_g = A'_g . x_g with a DENSE K x K weighted adjacency, encoders.py:378) and by the backward
// (dZ = S dX'^T...).
//
// Two implementations of the contraction sit behind the same entry point:
//   * SIMT fp32 (k_seg_contract_simt): 4x8 register micro-tiles, rows staged through shared memory;
//   * tcgen05 (k_seg_contract_tc, see k7_tc.cuh): 3xTF32 error-compensated split, fp32 accumulators
//     in TMEM, one CTA per graph, M = 128 x N <= 256 tile.  Selected by `use_tensor_cores`.
#include "common.cuh"

namespace tsg {

// ------------------------------------------------------------------------------------------
// segment contraction, SIMT:  C[g] (Kx x Ky) = sum_{r in graph g} X[r,:]^T Y[r,:]
// CTA per graph, 256 threads = 32 (x groups of 4) x 8 (y groups of 8) => 128 x 64 output tile.
// ------------------------------------------------------------------------------------------
constexpr int SC_ROWS = 32;

__global__ void __launch_bounds__(256)
k_seg_contract_simt(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ gptr,
                    int Kx, int Ky, float* __restrict__ C) {
  __shared__ __align__(16) float Xs[SC_ROWS][128 + 4];
  __shared__ __align__(16) float Ys[SC_ROWS][64 + 4];
  const int g = blockIdx.x;
  const int64_t lo = gptr[g], hi = gptr[g + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = tx * 4, y0 = ty * 8;
  float* Cg = C + (int64_t)g * Kx * Ky;
  for (int xc = 0; xc < Kx; xc += 128) {
    for (int yc = 0; yc < Ky; yc += 64) {
      float acc[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
      for (int64_t r0 = lo; r0 < hi; r0 += SC_ROWS) {
        const int rows = (int)min((int64_t)SC_ROWS, hi - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < SC_ROWS * 128; i += 256) {
          int r = i >> 7, c = i & 127;
          Xs[r][c] = (r < rows && xc + c < Kx) ? X[(r0 + r) * Kx + xc + c] : 0.f;
        }
        for (int i = threadIdx.x; i < SC_ROWS * 64; i += 256) {
          int r = i >> 6, c = i & 63;
          Ys[r][c] = (r < rows && yc + c < Ky) ? Y[(r0 + r) * Ky + yc + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < SC_ROWS; ++r) {
          float4 a = *reinterpret_cast<const float4*>(&Xs[r][x0]);
          float4 b0 = *reinterpret_cast<const float4*>(&Ys[r][y0]);
          float4 b1 = *reinterpret_cast<const float4*>(&Ys[r][y0 + 4]);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          int xr = xc + x0 + p, yr = yc + y0 + q;
          if (xr < Kx && yr < Ky) Cg[(int64_t)xr * Ky + yr] = acc[p][q];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// segment linear:  Y[r,:] = X[r,:] . Wg   for r in graph g, Wg = W[g] (Kin x M) or W[g]^T (M x Kin)
// CTA per graph; lanes: CG column groups of 4, 2 rows per thread (same scheme as k_linear_fwd).
// ------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(256)
k_seg_linear(const float* __restrict__ X, const float* __restrict__ W, const int64_t* __restrict__ gptr,
             int Kin, int M, int Mfull, int m0, int w_transposed, float* __restrict__ Y) {
  extern __shared__ __align__(16) float sl_smem[];
  constexpr int RS = 256 / CG;
  constexpr int R = RS * 2;
  const int Mp = CG * 4;
  const int pitch = Kin | 1;
  float* Ws = sl_smem;
  float* Xs = sl_smem + (size_t)Kin * Mp;
  const int g = blockIdx.x;
  const int64_t lo = gptr[g], hi = gptr[g + 1];
  // this launch produces output columns [m0, m0+M) of the Mfull-wide result
  const float* Wg = W + (int64_t)g * Kin * Mfull;
  const int cg = threadIdx.x % CG, rs = threadIdx.x / CG;
  for (int i = threadIdx.x; i < Kin * Mp; i += 256) {
    int k = i / Mp, m = i - k * Mp;
    float w = 0.f;
    if (m < M) w = w_transposed ? Wg[(int64_t)(m0 + m) * Kin + k] : Wg[(int64_t)k * Mfull + m0 + m];
    Ws[i] = w;
  }
  for (int64_t r0 = lo; r0 < hi; r0 += R) {
    const int rows = (int)min((int64_t)R, hi - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < R * Kin; i += 256) {
      int r = i / Kin, k = i - r * Kin;
      Xs[r * pitch + k] = r < rows ? X[(r0 + r) * Kin + k] : 0.f;
    }
    __syncthreads();
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const float* x0 = Xs + rs * pitch;
    const float* x1 = Xs + (rs + RS) * pitch;
    for (int k = 0; k < Kin; ++k) {
      float a0 = x0[k], a1 = x1[k];
      float4 w = *reinterpret_cast<const float4*>(Ws + k * Mp + cg * 4);
      acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
      acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
      acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
      acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int r = rs + j * RS;
      if (r < rows) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int m = cg * 4 + c;
          if (m < M) Y[(r0 + r) * Mfull + m0 + m] = acc[j][c];
        }
      }
    }
  }
}

}  // namespace tsg

#include "k7_tc.cuh"

using namespace tsg;

extern "C" int tsg_seg_contract(const float* X, const float* Y, const int64_t* graph_ptr, int64_t G,
                                int64_t Kx, int64_t Ky, float* C, int use_tensor_cores,
                                int32_t* status_dev, void* stream) {
  TSG_REQUIRE(G >= 0 && Kx > 0 && Ky > 0 && G < (int64_t)0x7fffffff, "seg_contract: bad shape");
  if (G == 0) return TSG_OK;
  TSG_REQUIRE(X && Y && graph_ptr && C, "seg_contract: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tensor_cores) {
    TSG_REQUIRE(status_dev, "seg_contract(tcgen05): status_dev (int32 on the device) is required");
    return launch_seg_contract_tc(X, Y, graph_ptr, (int)G, (int)Kx, (int)Ky, C, status_dev, st);
  }
  k_seg_contract_simt<<<(int)G, 256, 0, st>>>(X, Y, graph_ptr, (int)Kx, (int)Ky, C);
  return check_launch("seg_contract");
}

extern "C" int tsg_seg_linear(const float* X, const float* W, const int64_t* graph_ptr, int64_t G,
                              int64_t Kin, int64_t M, int w_transposed, float* Y, void* stream) {
  TSG_REQUIRE(G >= 0 && Kin > 0 && M > 0 && G < (int64_t)0x7fffffff, "seg_linear: bad shape");
  if (G == 0) return TSG_OK;
  TSG_REQUIRE(X && W && graph_ptr && Y, "seg_linear: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t m0 = 0; m0 < M; m0 += 128) {            // <= 128 output columns per launch
    int mc = (int)(M - m0 < 128 ? M - m0 : 128);
    int cg = 1; while (cg * 4 < mc && cg < 32) cg <<= 1;
    while (cg < 32 && (size_t)(256 / cg) * 2 * (size_t)(Kin | 1) * 4 > 64 * 1024) cg <<= 1;
    size_t smem = ((size_t)Kin * cg * 4 + (size_t)(256 / cg) * 2 * (Kin | 1)) * sizeof(float);
    TSG_REQUIRE(smem <= 227 * 1024, "seg_linear: Kin*M too large for shared memory");
#define TSG_GO(C_)                                                                                       \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_seg_linear<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    k_seg_linear<C_><<<(int)G, 256, smem, st>>>(X, W, graph_ptr, (int)Kin, mc, (int)M, (int)m0, w_transposed, Y)
    switch (cg) {
      case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
      case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
    }
#undef TSG_GO
  }
  return check_launch("seg_linear");
}
