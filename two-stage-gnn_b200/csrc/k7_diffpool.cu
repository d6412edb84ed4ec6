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
#include <stdlib.h>

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

// Wide variant (round 2) for M > 64 output columns per launch (all 32 lanes of a warp share their rows, so the x loads
// are warp broadcasts whatever the pitch).  The launch list of the DiffPool step (profiles/r02_launches_diffpool_step.csv)
// showed where k_seg_linear's time is: the backward of the tcgen05 contraction -- dS = [Z|T] dC^T (K_in = 196, M = 100)
// and d[Z|T] = S dC (K_in = 100, M = 196) -- 2.9 + 2.8 ms of a 24.5 ms step at 13 TFLOP/s: the 2 x 4 micro-tile issues 3
// shared-memory loads per 8 FMAs.  Here a thread owns 4 rows x 4 columns and walks k four at a time: 8 128-bit loads per
// 64 FMAs.  Per output the products are still added in k-ascending order with FMA (zero padding adds fma(0, 0, acc) =
// acc): bit-identical to k_seg_linear.  (A 64 x 128 tile kernel with K chunked through shared memory was measured
// first: slower, 25.9 vs 24.3 ms per step -- two barriers per 16-deep chunk with the loads exposed.)
__global__ void __launch_bounds__(256)
k_seg_linear_wide(const float* __restrict__ X, const float* __restrict__ W, const int64_t* __restrict__ gptr,
                  int Kin, int M, int Mfull, int m0, int w_transposed, float* __restrict__ Y) {
  extern __shared__ __align__(16) float sl_smem[];
  constexpr int CG = 32, RS = 256 / CG, R = RS * 4;              // 8 row slots x 4 rows = 32 rows per tile
  const int Mp = CG * 4;
  const int Kp = (Kin + 3) & ~3;
  float* Ws = sl_smem;                                            // [Kp][Mp]
  float* Xs = sl_smem + (size_t)Kp * Mp;                          // [R][Kp]
  const int g = blockIdx.x;
  const int64_t lo = gptr[g], hi = gptr[g + 1];
  const float* Wg = W + (int64_t)g * Kin * Mfull;
  const int cg = threadIdx.x % CG, rs = threadIdx.x / CG;
  for (int i = threadIdx.x; i < Kp * Mp; i += 256) {
    const int k = i / Mp, m = i - k * Mp;
    float w = 0.f;
    if (m < M && k < Kin) w = w_transposed ? Wg[(int64_t)(m0 + m) * Kin + k] : Wg[(int64_t)k * Mfull + m0 + m];
    Ws[i] = w;
  }
  for (int64_t r0 = lo; r0 < hi; r0 += R) {
    const int rows = (int)min((int64_t)R, hi - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < R * Kp; i += 256) {
      const int r = i / Kp, k = i - r * Kp;
      Xs[i] = (r < rows && k < Kin) ? X[(r0 + r) * Kin + k] : 0.f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    for (int k = 0; k < Kp; k += 4) {
      float4 xv[4], w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float4*>(Xs + (rs + j * RS) * Kp + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(Ws + (k + q) * Mp + cg * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = fmaf(xv[j].x, w[0].x, acc[j][0]); acc[j][1] = fmaf(xv[j].x, w[0].y, acc[j][1]);
        acc[j][2] = fmaf(xv[j].x, w[0].z, acc[j][2]); acc[j][3] = fmaf(xv[j].x, w[0].w, acc[j][3]);
        acc[j][0] = fmaf(xv[j].y, w[1].x, acc[j][0]); acc[j][1] = fmaf(xv[j].y, w[1].y, acc[j][1]);
        acc[j][2] = fmaf(xv[j].y, w[1].z, acc[j][2]); acc[j][3] = fmaf(xv[j].y, w[1].w, acc[j][3]);
        acc[j][0] = fmaf(xv[j].z, w[2].x, acc[j][0]); acc[j][1] = fmaf(xv[j].z, w[2].y, acc[j][1]);
        acc[j][2] = fmaf(xv[j].z, w[2].z, acc[j][2]); acc[j][3] = fmaf(xv[j].z, w[2].w, acc[j][3]);
        acc[j][0] = fmaf(xv[j].w, w[3].x, acc[j][0]); acc[j][1] = fmaf(xv[j].w, w[3].y, acc[j][1]);
        acc[j][2] = fmaf(xv[j].w, w[3].z, acc[j][2]); acc[j][3] = fmaf(xv[j].w, w[3].w, acc[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rs + j * RS;
      if (r < rows) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int m = cg * 4 + c;
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
  static const bool no_gemm = getenv("TSG_SEG_LINEAR_SIMPLE") != nullptr;      // A/B: round 1's 2 x 4 micro-tile kernel only
  for (int64_t m0 = 0; m0 < M; m0 += 128) {            // <= 128 output columns per launch
    int mc = (int)(M - m0 < 128 ? M - m0 : 128);
    if (!no_gemm && mc > 64) {                          // wide outputs: 4 x 4 micro-tile, k four at a time (bit-identical)
      const int Kp = (int)((Kin + 3) & ~(int64_t)3);
      const size_t smw = ((size_t)Kp * 128 + (size_t)32 * Kp) * sizeof(float);
      if (smw <= 227 * 1024) {
        static size_t configured = 0;
        if (smw > 48 * 1024 && smw > configured) {
          cudaFuncSetAttribute(k_seg_linear_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          configured = 227 * 1024;
        }
        k_seg_linear_wide<<<(int)G, 256, smw, st>>>(X, W, graph_ptr, (int)Kin, mc, (int)M, (int)m0, w_transposed, Y);
        continue;
      }
    }
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

extern "C" int tsg_seg_linear_tc(const float* X, const float* W, const int64_t* graph_ptr, int64_t G, int64_t Kin, int64_t M,
                                 int w_transposed, float* Y, int32_t* status_dev, void* stream) {
  TSG_REQUIRE(G >= 0 && Kin > 0 && M > 0 && G < (int64_t)0x7fffffff && Kin < (int64_t)0x7fffffff, "seg_linear_tc: bad shape");
  if (G == 0) return TSG_OK;
  TSG_REQUIRE(X && W && graph_ptr && Y && status_dev, "seg_linear_tc: null pointer");
  return launch_seg_linear_tc(X, W, graph_ptr, (int)G, 0, (int)Kin, (int)M, w_transposed, Y, status_dev, (cudaStream_t)stream);
}

// in-place softmax(Y[r, :] + bias) per row, M <= 256: a warp per row, the row in registers
__global__ void __launch_bounds__(256)
k_bias_softmax_rows(float* __restrict__ Y, const float* __restrict__ bias, int64_t N, int M) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += nwarps) {
    float v[8];
    float mx = -3.4e38f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < M ? Y[r * M + c] + (bias ? bias[c] : 0.f) : -3.4e38f;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = lane + 32 * i < M ? expf(v[i] - mx) : 0.f; se += v[i]; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < M) Y[r * M + c] = v[i] / se;
    }
  }
}

/* Y = X . W (W [Kin, M]; or W^T with W stored [M, Kin]) on tcgen05 (k_seg_linear_tc with one shared weight, a CTA per 128
 * rows), then, with softmax != 0, Y = softmax(Y + bias) per row in place -- DiffPool's assignment Linear + softmax
 * (encoders.py:366-369) at config-4 size is a 970 k x 228 x 100 product: 2.1 ms on the fp32 SIMT kernel. */
extern "C" int tsg_linear_tc(const float* X, const float* W, const float* bias, int64_t N, int64_t Kin, int64_t M,
                             int w_transposed, int softmax, float* Y, int32_t* status_dev, void* stream) {
  TSG_REQUIRE(N >= 0 && Kin > 0 && M > 0 && Kin < (int64_t)0x7fffffff && N / 128 < (int64_t)0x7fffff00, "linear_tc: bad shape");
  if (N == 0) return TSG_OK;
  TSG_REQUIRE(X && W && Y && status_dev, "linear_tc: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_seg_linear_tc(X, W, nullptr, 0, N, (int)Kin, (int)M, w_transposed, Y, status_dev, st);
  if (rc) return rc;
  if (softmax || bias) {
    TSG_REQUIRE(softmax, "linear_tc: a bias without the softmax epilogue is not implemented");
    const int64_t ctas = (N + 7) / 8;
    k_bias_softmax_rows<<<(int)(ctas < (int64_t)TSG_NUM_SMS * 16 ? ctas : (int64_t)TSG_NUM_SMS * 16), 256, 0, st>>>(Y, bias, N, (int)M);
    return check_launch("linear_tc(softmax)");
  }
  return TSG_OK;
}
