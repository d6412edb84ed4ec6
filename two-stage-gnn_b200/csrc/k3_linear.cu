// K3 -- tall-skinny row-local dense kernels: Y = X.W (+b) with the dense-directory epilogue
// (row L2-normalise -> ReLU -> node-wise BatchNorm), its backward, and dW = X^T.dY / db.
//
// Replaces, for N ~ 10^6 packed node rows and K, M <= 128..256 feature columns:
//   * `x @ weight` of PyG GCNConv (Code/sag/network.py:34 via GCNConv.forward) and its autograd
//     (dX = dH.W^T, dW = X^T.dH), which torch hands to cuBLAS SIMT sgemm kernels that cost more
//     than every aggregation kernel together (profiles/r01_launch_summary.md);
//   * `torch.matmul(y, self.weight) + self.bias` -> `F.normalize(y, p=2, dim=2)` of the dense
//     GraphConv (Code/sage+gat+diffpool/encoders.py:36-40; Code/eigengcn/encoders.py:34-39),
//     `self.act` (ReLU) and `apply_bn` (encoders.py:134-138: a FRESH BatchNorm1d(num_nodes) per
//     call => per-node statistics over the feature axis, biased variance, eps 1e-5, no affine, no
//     running state; with the scripts' batch size of one graph it is a per-node LayerNorm).
//
// These products are HBM-bound (AI = K.M/(2(K+M)) flop/B ~ 12 at 89x32): the design goal is to read X
// once, keep W in shared memory, and skip exact zeros (one-hot node-label features have ONE
// non-zero in 89 columns: the dot product stays bit-identical because the skipped terms are +0).
// Plain FP32 FFMA; no tensor cores (TF32 would break the 1e-5 parity bar; DiffPool's contractions
// use the 3xTF32 split on tcgen05 in k7).
#include "common.cuh"

namespace tsg {

constexpr int LIN_THREADS = 256;
constexpr int LIN_RPT = 2;                 // rows per thread
constexpr float NORM_EPS = 1e-12f;         // F.normalize eps
constexpr float BN_EPS = 1e-5f;            // BatchNorm1d eps

template <int CG>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int d = CG / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d);
  return v;
}

// epilogue on one row held as 4 columns per lane across CG lanes. `ok[c]` masks padded columns.
template <int CG>
__device__ __forceinline__ void row_epilogue(float (&u)[4], const bool (&ok)[4], int M, int flags,
                                             unsigned gmask) {
  if (flags & TSG_LIN_NORMALIZE) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (ok[c]) ss += u[c] * u[c];
    ss = group_sum<CG>(ss, gmask);
    float nrm = fmaxf(sqrtf(ss), NORM_EPS);
#pragma unroll
    for (int c = 0; c < 4; ++c) u[c] = u[c] / nrm;
  }
  if (flags & TSG_LIN_RELU) {
#pragma unroll
    for (int c = 0; c < 4; ++c) u[c] = fmaxf(u[c], 0.f);
  }
  if (flags & TSG_LIN_NODEBN) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (ok[c]) s += u[c];
    float mean = group_sum<CG>(s, gmask) / (float)M;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (ok[c]) { float d = u[c] - mean; q += d * d; }
    float var = group_sum<CG>(q, gmask) / (float)M;
    float rstd = 1.0f / sqrtf(var + BN_EPS);
#pragma unroll
    for (int c = 0; c < 4; ++c) u[c] = (u[c] - mean) * rstd;
  }
}

// Y[N,M] = epilogue(X[N,K] . Wop + bias), Wop = W[K,M] or W[M,K]^T.  Persistent CTAs: W staged in
// shared memory once, then a grid-stride loop over tiles of R = (256/CG)*2 rows.
template <int CG>
__global__ void __launch_bounds__(LIN_THREADS)
k_linear_fwd(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
             float* __restrict__ Y, int N, int K, int M, int ldw, int ldy, int w_transposed, int flags) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RS = LIN_THREADS / CG;
  constexpr int R = RS * LIN_RPT;
  const int Mp = CG * 4;
  const int pitch = K | 1;                       // odd pitch: rows of a warp hit distinct banks
  float* Ws = smem;                              // [K][Mp]
  float* Xs = smem + (size_t)K * Mp;             // [R][pitch]
  const int cg = threadIdx.x % CG, rs = threadIdx.x / CG;
  const unsigned gmask = CG == 32 ? 0xffffffffu : (((1u << CG) - 1u) << (((threadIdx.x & 31) / CG) * CG));

  for (int i = threadIdx.x; i < K * Mp; i += LIN_THREADS) {
    int k = i / Mp, m = i - k * Mp;
    float w = 0.f;
    if (m < M) w = w_transposed ? W[(size_t)m * ldw + k] : W[(size_t)k * ldw + m];
    Ws[i] = w;
  }
  bool ok[4]; float b4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ok[c] = cg * 4 + c < M;
    b4[c] = (bias != nullptr && ok[c]) ? bias[cg * 4 + c] : 0.f;
  }
  const int num_tiles = (N + R - 1) / R;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int r0 = tile * R;
    const int rows = min(R, N - r0);
    __syncthreads();                             // previous tile's Xs readers are done; Ws visible
    {
      const float* src = X + (size_t)r0 * K;
      const int total = rows * K;
      for (int i = threadIdx.x; i < R * K; i += LIN_THREADS) {
        int r = i / K, k = i - r * K;
        Xs[r * pitch + k] = i < total ? src[i] : 0.f;
      }
    }
    __syncthreads();
    float acc[LIN_RPT][4];
#pragma unroll
    for (int j = 0; j < LIN_RPT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    const float* x0 = Xs + (rs)*pitch;
    const float* x1 = Xs + (rs + RS) * pitch;
    for (int k = 0; k < K; ++k) {
      float a0 = x0[k], a1 = x1[k];
      if (!__any_sync(0xffffffffu, (a0 != 0.f) | (a1 != 0.f))) continue;   // exact zeros contribute +0
      float4 w = *reinterpret_cast<const float4*>(Ws + k * Mp + cg * 4);
      acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
      acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
      acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
      acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
    }
#pragma unroll
    for (int j = 0; j < LIN_RPT; ++j) {
      const int r = rs + j * RS;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] += b4[c];
      if (flags) row_epilogue<CG>(acc[j], ok, M, flags, gmask);
      if (r < rows) {
        float* dst = Y + (size_t)(r0 + r) * ldy + cg * 4;
        if ((M & 3) == 0 && (ldy & 3) == 0) {
          if (ok[0]) *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (ok[c]) dst[c] = acc[j][c];
        }
      }
    }
  }
}

// Backward of the dense-directory epilogue: dU from dO, recomputing u = agg.W + b (agg is read for
// dW anyway; recomputing saves writing + re-reading u, v and the statistics).  A.4 of SURVEY:
//   BN:  dR = rstd * (dO - mean(dO) - o * mean(dO*o));  ReLU: dV = dR * (v > 0)
//   normalise: dU = (dV - v * (v.dV)) / max(||u||, eps)   (||u|| < eps: dU = dV / eps)
template <int CG>
__global__ void __launch_bounds__(LIN_THREADS)
k_dense_epilogue_bwd(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ dO, float* __restrict__ dU, int N, int K, int M, int flags) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RS = LIN_THREADS / CG;
  constexpr int R = RS * LIN_RPT;
  const int Mp = CG * 4;
  const int pitch = K | 1;
  float* Ws = smem;
  float* Xs = smem + (size_t)K * Mp;
  const int cg = threadIdx.x % CG, rs = threadIdx.x / CG;
  const unsigned gmask = CG == 32 ? 0xffffffffu : (((1u << CG) - 1u) << (((threadIdx.x & 31) / CG) * CG));
  for (int i = threadIdx.x; i < K * Mp; i += LIN_THREADS) {
    int k = i / Mp, m = i - k * Mp;
    Ws[i] = m < M ? W[(size_t)k * M + m] : 0.f;
  }
  bool ok[4]; float b4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ok[c] = cg * 4 + c < M;
    b4[c] = (bias != nullptr && ok[c]) ? bias[cg * 4 + c] : 0.f;
  }
  const int num_tiles = (N + R - 1) / R;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int r0 = tile * R;
    const int rows = min(R, N - r0);
    __syncthreads();
    {
      const float* src = X + (size_t)r0 * K;
      const int total = rows * K;
      for (int i = threadIdx.x; i < R * K; i += LIN_THREADS) {
        int r = i / K, k = i - r * K;
        Xs[r * pitch + k] = i < total ? src[i] : 0.f;
      }
    }
    __syncthreads();
    float u[LIN_RPT][4];
#pragma unroll
    for (int j = 0; j < LIN_RPT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) u[j][c] = 0.f;
    const float* x0 = Xs + (rs)*pitch;
    const float* x1 = Xs + (rs + RS) * pitch;
    for (int k = 0; k < K; ++k) {
      float a0 = x0[k], a1 = x1[k];
      if (!__any_sync(0xffffffffu, (a0 != 0.f) | (a1 != 0.f))) continue;
      float4 w = *reinterpret_cast<const float4*>(Ws + k * Mp + cg * 4);
      u[0][0] = fmaf(a0, w.x, u[0][0]); u[0][1] = fmaf(a0, w.y, u[0][1]);
      u[0][2] = fmaf(a0, w.z, u[0][2]); u[0][3] = fmaf(a0, w.w, u[0][3]);
      u[1][0] = fmaf(a1, w.x, u[1][0]); u[1][1] = fmaf(a1, w.y, u[1][1]);
      u[1][2] = fmaf(a1, w.z, u[1][2]); u[1][3] = fmaf(a1, w.w, u[1][3]);
    }
#pragma unroll
    for (int j = 0; j < LIN_RPT; ++j) {
      const int r = rs + j * RS;
      const bool live = r < rows;
      float g[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        u[j][c] += b4[c];
        g[c] = (live && ok[c]) ? dO[(size_t)(r0 + r) * M + cg * 4 + c] : 0.f;
      }
      // forward recompute
      float nrm = 1.f, v[4], rl[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = u[j][c];
      if (flags & TSG_LIN_NORMALIZE) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (ok[c]) ss += v[c] * v[c];
        ss = group_sum<CG>(ss, gmask);
        nrm = fmaxf(sqrtf(ss), NORM_EPS);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = v[c] / nrm;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) rl[c] = (flags & TSG_LIN_RELU) ? fmaxf(v[c], 0.f) : v[c];
      if (flags & TSG_LIN_NODEBN) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (ok[c]) s += rl[c];
        float mean = group_sum<CG>(s, gmask) / (float)M;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (ok[c]) { float d = rl[c] - mean; q += d * d; }
        float var = group_sum<CG>(q, gmask) / (float)M;
        float rstd = 1.0f / sqrtf(var + BN_EPS);
        float sg = 0.f, sgo = 0.f, o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          o[c] = (rl[c] - mean) * rstd;
          if (ok[c]) { sg += g[c]; sgo += g[c] * o[c]; }
        }
        float mg = group_sum<CG>(sg, gmask) / (float)M;
        float mgo = group_sum<CG>(sgo, gmask) / (float)M;
#pragma unroll
        for (int c = 0; c < 4; ++c) g[c] = rstd * (g[c] - mg - o[c] * mgo);
      }
      if (flags & TSG_LIN_RELU) {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (!(v[c] > 0.f)) g[c] = 0.f;
      }
      if (flags & TSG_LIN_NORMALIZE) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (ok[c]) dot += v[c] * g[c];
        dot = group_sum<CG>(dot, gmask);
        // torch: y = x / clamp_min(||x||, eps).  Above the clamp d||x||/dx = x/||x|| = v;
        // at / below the clamp the denominator is the constant eps.
        const bool clamped = !(nrm > NORM_EPS);
#pragma unroll
        for (int c = 0; c < 4; ++c) g[c] = clamped ? g[c] / nrm : (g[c] - v[c] * dot) / nrm;
      }
      if (live) {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (ok[c]) dU[(size_t)(r0 + r) * M + cg * 4 + c] = g[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// dW[K,M] = X^T . dY   and   db[M] = colsum(dY)   -- deterministic two-stage reduction.
// A "unit" owns feature row k (k == K is the bias row with x == 1) x one 32-column chunk of M and
// keeps 32 accumulators in registers; row groups split the tile's rows; exact zeros are skipped.
// stage 1 writes part[cta][K+1][M]; stage 2 sums the CTAs in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int LBW_ROWS = 64;
constexpr int LBW_GRID = TSG_NUM_SMS * 2;

__global__ void __launch_bounds__(LIN_THREADS)
k_linear_bwd_weight(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ part,
                    int N, int K, int M, int units, int row_groups) {
  extern __shared__ __align__(16) float smem[];
  const int Mp = (M + 31) / 32 * 32;             // padded so every chunk is 32 wide
  const int mch = Mp / 32;
  const int pitch = K | 1;
  float* Xs = smem;                              // [ROWS][pitch]
  float* Ds = smem + LBW_ROWS * pitch;           // [ROWS][Mp]
  const int K1 = K + 1;
  const int slots = units * row_groups;          // active threads per pass
  const int passes = (K1 * mch + units - 1) / units;    // units per thread when K1*mch > units
  const int num_tiles = (N + LBW_ROWS - 1) / LBW_ROWS;
  float* mypart = part + (size_t)blockIdx.x * K1 * M;

  for (int pass = 0; pass < passes; ++pass) {
    const int t = threadIdx.x;
    const bool active = t < slots;
    const int unit = pass * units + (active ? t % units : 0);
    const int rg = active ? t / units : 0;
    const bool uok = active && unit < K1 * mch;
    const int k = uok ? unit / mch : 0;
    const int mc = uok ? unit % mch : 0;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int r0 = tile * LBW_ROWS;
      const int rows = min(LBW_ROWS, N - r0);
      __syncthreads();
      for (int i = threadIdx.x; i < LBW_ROWS * K; i += LIN_THREADS) {
        int r = i / K, kk = i - r * K;
        Xs[r * pitch + kk] = r < rows ? X[(size_t)(r0 + r) * K + kk] : 0.f;
      }
      for (int i = threadIdx.x; i < LBW_ROWS * Mp; i += LIN_THREADS) {
        int r = i / Mp, m = i - r * Mp;
        Ds[i] = (r < rows && m < M) ? dY[(size_t)(r0 + r) * M + m] : 0.f;
      }
      __syncthreads();
      if (uok) {
        for (int r = rg; r < rows; r += row_groups) {
          float x = k < K ? Xs[r * pitch + k] : 1.f;
          if (x == 0.f) continue;
          const float4* d = reinterpret_cast<const float4*>(Ds + r * Mp + mc * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = d[q];
            acc[4 * q + 0] = fmaf(x, v.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x, v.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x, v.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, v.w, acc[4 * q + 3]);
          }
        }
      }
    }
    // combine the row groups in a fixed order through shared memory, one group at a time
    __syncthreads();
    float* red = smem;                           // reuse: [units][32]
    for (int g = 0; g < row_groups; ++g) {
      if (uok && rg == g) {
        float* dst = red + (size_t)(t % units) * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) dst[c] = g == 0 ? acc[c] : dst[c] + acc[c];
      }
      __syncthreads();
    }
    if (uok && rg == 0) {
      const float* src = red + (size_t)(t % units) * 32;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        int m = mc * 32 + c;
        if (m < M) mypart[(size_t)k * M + m] = src[c];
      }
    }
    __syncthreads();
  }
}

// lanes per row: enough for M columns (4 per lane) and few enough rows per tile that the X tile
// (R = 256/cg*2 rows x K floats) stays below ~64 KB of shared memory
static int pick_cg(int64_t M, int64_t K) {
  int c = 1;
  while (c * 4 < M && c < 32) c <<= 1;
  while (c < 32 && (size_t)(LIN_THREADS / c) * LIN_RPT * (size_t)(K | 1) * 4 > 64 * 1024) c <<= 1;
  return c;
}

static size_t lin_smem_bytes(int cg, int64_t K) {
  int R = (LIN_THREADS / cg) * LIN_RPT;
  return ((size_t)K * cg * 4 + (size_t)R * (K | 1)) * sizeof(float);
}

}  // namespace tsg

using namespace tsg;

template <typename Kern>
static int set_smem(Kern kern, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) { set_error("%s: needs %zu B of shared memory (> 227 KB): K*M too large", what, bytes); return TSG_EINVAL; }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return TSG_ELAUNCH; }
  }
  return TSG_OK;
}

extern "C" int tsg_linear_fwd(const float* X, const float* W, const float* bias, float* Y,
                              int64_t N, int64_t K, int64_t M, int w_transposed, int flags, void* stream) {
  TSG_REQUIRE(N >= 0 && K > 0 && M > 0, "linear_fwd: bad shape");
  TSG_REQUIRE(M <= 128 || flags == 0, "linear_fwd: the fused epilogue needs out_feat <= 128 (got %lld)", (long long)M);
  TSG_REQUIRE(N < (int64_t)0x7fffffff / (K > M ? K : M), "linear_fwd: N*K overflows int32 tiles");
  if (N == 0) return TSG_OK;
  TSG_REQUIRE(X && W && Y, "linear_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // output columns are produced in chunks of <= 128 (one float4 per lane, 32 lanes per row)
  for (int64_t m0 = 0; m0 < M; m0 += 128) {
    int64_t mc = M - m0 < 128 ? M - m0 : 128;
    int cg = pick_cg(mc, K);
    size_t smem = lin_smem_bytes(cg, K);
    int R = (LIN_THREADS / cg) * LIN_RPT;
    int tiles = (int)((N + R - 1) / R);
    int grid = tiles < TSG_NUM_SMS * 4 ? tiles : TSG_NUM_SMS * 4;
    const float* Wc = w_transposed ? W + m0 * K : W + m0;
    int ldw = w_transposed ? (int)K : (int)M;
    const float* bc = bias ? bias + m0 : nullptr;
    float* Yc = Y + m0;
    int rc;
#define TSG_GO(C)                                                                                  \
    rc = set_smem(k_linear_fwd<C>, smem, "linear_fwd"); if (rc) return rc;                          \
    k_linear_fwd<C><<<grid, LIN_THREADS, smem, st>>>(X, Wc, bc, Yc, (int)N, (int)K, (int)mc, ldw, (int)M, w_transposed, flags)
    switch (cg) {
      case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
      case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
    }
#undef TSG_GO
  }
  return check_launch("linear_fwd");
}

extern "C" int tsg_dense_epilogue_bwd(const float* X, const float* W, const float* bias, const float* dO,
                                      float* dU, int64_t N, int64_t K, int64_t M, int flags, void* stream) {
  TSG_REQUIRE(N >= 0 && K > 0 && M > 0 && M <= 128, "dense_epilogue_bwd: bad shape");
  if (N == 0) return TSG_OK;
  TSG_REQUIRE(X && W && dO && dU, "dense_epilogue_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int cg = pick_cg(M, K);
  size_t smem = lin_smem_bytes(cg, K);
  int R = (LIN_THREADS / cg) * LIN_RPT;
  int tiles = (int)((N + R - 1) / R);
  int grid = tiles < TSG_NUM_SMS * 4 ? tiles : TSG_NUM_SMS * 4;
  int rc;
#define TSG_GO(C)                                                                                  \
  rc = set_smem(k_dense_epilogue_bwd<C>, smem, "dense_epilogue_bwd"); if (rc) return rc;            \
  k_dense_epilogue_bwd<C><<<grid, LIN_THREADS, smem, st>>>(X, W, bias, dO, dU, (int)N, (int)K, (int)M, flags)
  switch (cg) {
    case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
    case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
  }
#undef TSG_GO
  return check_launch("dense_epilogue_bwd");
}

extern "C" size_t tsg_linear_bwd_weight_workspace_bytes(int64_t K, int64_t M) {
  return ws_bytes((size_t)LBW_GRID * (size_t)(K + 1) * (size_t)M, 4) + 256;
}

extern "C" int tsg_linear_bwd_weight(const float* X, const float* dY, float* dW, float* db,
                                     int64_t N, int64_t K, int64_t M,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  TSG_REQUIRE(N >= 0 && K > 0 && M > 0, "linear_bwd_weight: bad shape");
  TSG_REQUIRE(X && dY && (dW || db), "linear_bwd_weight: null pointer");
  if (workspace_bytes < tsg_linear_bwd_weight_workspace_bytes(K, M)) { set_error("linear_bwd_weight: workspace too small"); return TSG_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  int Mp = (int)((M + 31) / 32 * 32), mch = Mp / 32;
  int total_units = (int)(K + 1) * mch;
  int units = total_units < LIN_THREADS ? total_units : LIN_THREADS;
  int row_groups = LIN_THREADS / units; if (row_groups < 1) row_groups = 1;
  if (row_groups > LBW_ROWS) row_groups = LBW_ROWS;
  size_t smem_tile = ((size_t)LBW_ROWS * ((int)K | 1) + (size_t)LBW_ROWS * Mp) * sizeof(float);
  size_t smem_red = (size_t)units * 32 * sizeof(float);
  size_t smem = smem_tile > smem_red ? smem_tile : smem_red;
  int rc = set_smem(k_linear_bwd_weight, smem, "linear_bwd_weight"); if (rc) return rc;
  int tiles = (int)((N + LBW_ROWS - 1) / LBW_ROWS);
  int grid = LBW_GRID;                       // fixed => the summation order never depends on the device
  (void)tiles;
  k_linear_bwd_weight<<<grid, LIN_THREADS, smem, st>>>(X, dY, part, (int)N, (int)K, (int)M, units, row_groups);
  int total = (int)((K + 1) * M);
  launch_partial_sum_final(part, dW, (int)(K * M), db, grid, total, st);
  return check_launch("linear_bwd_weight");
}
