// K6 -- per-graph readouts (segment max with arg-max, mean, sum), forward and backward.
//
// Replaces torch.cat([global_max_pool(x,batch), global_mean_pool(x,batch)],1) of
// Code/sag/network.py:36,40,44 (torch-scatter atomics + a batch.max().item() sync upstream) and
// the torch.max(x, dim=1) readouts of Code/sage+gat+diffpool/encoders.py:183,190,197,353,383.
//
// One warp per graph.  A row is covered by LPR lanes x VEC floats (128-bit loads when F%4==0);
// the 32/LPR lane groups stride over the graph's rows so a warp keeps 32/LPR independent
// row loads in flight; groups are combined with xor-shuffles in a fixed order (deterministic).
// Max ties resolve to the FIRST (lowest) row, the torch-scatter CPU rule; the backward routes
// the max gradient to that row only.  Streaming, HBM-bound: x is read once, dx written once.
#include "common.cuh"
#include <float.h>

namespace tsg {

constexpr int RO_THREADS = 128;   // 4 graphs per block
constexpr int RO_BWD_THREADS = 256;

template <int VEC, int LPR>
__global__ void __launch_bounds__(RO_THREADS)
k_readout_fwd(const float* __restrict__ x, const int64_t* __restrict__ gptr, int G, int F, int mode,
              float* __restrict__ out, int64_t out_stride, int* __restrict__ argmax) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int FV = F / VEC;
  const int warps_per_block = RO_THREADS / 32;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < G;
       g += gridDim.x * warps_per_block) {
    const int64_t lo = gptr[g], hi = gptr[g + 1];
    const int n = (int)(hi - lo);
    for (int fb = 0; fb < FV; fb += LPR) {           // warp-uniform trip count
      const int f = fb + l;
      const bool fok = f < FV;
      float mx[VEC], sm[VEC]; int am[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { mx[v] = -FLT_MAX; sm[v] = 0.f; am[v] = -1; }
      if (fok) {
        for (int i = sub; i < n; i += RPW) {
          float vals[VEC];
          if (VEC == 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(x) + (lo + i) * FV + f);
            vals[0] = t.x; vals[1 % VEC] = t.y; vals[2 % VEC] = t.z; vals[3 % VEC] = t.w;
          } else {
            vals[0] = __ldg(x + (lo + i) * FV + f);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            sm[v] += vals[v];
            if (vals[v] > mx[v] || am[v] < 0) { mx[v] = vals[v]; am[v] = i; }
          }
        }
      }
      // combine the RPW sub-groups (all 32 lanes participate)
#pragma unroll
      for (int d = LPR; d < 32; d <<= 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float omx = __shfl_xor_sync(0xffffffffu, mx[v], d);
          int oam = __shfl_xor_sync(0xffffffffu, am[v], d);
          float osm = __shfl_xor_sync(0xffffffffu, sm[v], d);
          // sum: fixed pairing => deterministic; both partners compute the same value
          sm[v] = (sub & (d / LPR)) ? osm + sm[v] : sm[v] + osm;
          bool take = oam >= 0 && (am[v] < 0 || omx > mx[v] || (omx == mx[v] && oam < am[v]));
          if (take) { mx[v] = omx; am[v] = oam; }
        }
      }
      if (fok && sub == 0) {
        int col = 0;
        if (mode & TSG_READOUT_MAX) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            out[(int64_t)g * out_stride + f * VEC + v] = am[v] >= 0 ? mx[v] : 0.f;
            if (argmax) argmax[(int64_t)g * F + f * VEC + v] = am[v] >= 0 ? (int)(lo + am[v]) : -1;
          }
          col = F;
        }
        if (mode & (TSG_READOUT_MEAN | TSG_READOUT_SUM)) {
          float inv = (mode & TSG_READOUT_MEAN) ? 1.f / (float)(n > 0 ? n : 1) : 1.f;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float r = (mode & TSG_READOUT_MEAN) ? sm[v] / (float)(n > 0 ? n : 1) : sm[v];
            (void)inv;
            out[(int64_t)g * out_stride + col + f * VEC + v] = r;
          }
        }
      }
    }
  }
}

// SAGPool gate + readout in one pass (K10 executor, F % 4 == 0): row i of graph g is
//     xo[lo + i] = x[perm[lo + i]] * tanh(score[perm[lo + i]])                 layers.py:21
// written once and reduced on the fly into [gmp || gap] (network.py:36,40,44) -- the gated rows are not read back.
// Same reduction order as k_readout_fwd<4, LPR> (sub-group s takes rows s, s + RPW, ... in order, then the fixed
// butterfly) and the same products as k_gate_gather_fwd<4>: out, argmax and xo are bit-identical to the two-kernel
// sequence.  GRO_U row batches per sub-group are loaded before any is reduced (perm -> score / x is a dependent chain).
constexpr int GRO_U = 4;

template <int LPR>
__global__ void __launch_bounds__(RO_THREADS)
k_gate_readout_fwd(const float4* __restrict__ x, const float* __restrict__ score, const int64_t* __restrict__ perm,
                   const int64_t* __restrict__ gptr, int G, int F4, float4* __restrict__ xo,
                   float* __restrict__ out, int64_t out_stride, int* __restrict__ argmax) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int F = F4 * 4;
  const int warps_per_block = RO_THREADS / 32;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < G; g += gridDim.x * warps_per_block) {
    const int64_t lo = gptr[g], hi = gptr[g + 1];
    const int n = (int)(hi - lo);
    for (int fb = 0; fb < F4; fb += LPR) {           // warp-uniform trip count
      const int f = fb + l;
      const bool fok = f < F4;
      float mx[4], sm[4]; int am[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) { mx[v] = -FLT_MAX; sm[v] = 0.f; am[v] = -1; }
      if (fok) {
        for (int ib = sub; ib < n; ib += RPW * GRO_U) {
          int64_t j[GRO_U]; float t[GRO_U]; float4 xv[GRO_U];
#pragma unroll
          for (int u = 0; u < GRO_U; ++u) {
            const int i = ib + u * RPW;
            j[u] = i < n ? __ldg(perm + lo + i) : -1;
          }
#pragma unroll
          for (int u = 0; u < GRO_U; ++u) {
            t[u] = 0.f; xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j[u] >= 0) { t[u] = __ldg(score + j[u]); xv[u] = __ldg(x + j[u] * F4 + f); }
          }
#pragma unroll
          for (int u = 0; u < GRO_U; ++u) {
            if (j[u] < 0) continue;
            const int i = ib + u * RPW;
            const float th = tanhf(t[u]);
            float4 w = xv[u];
            w.x *= th; w.y *= th; w.z *= th; w.w *= th;
            xo[(lo + i) * F4 + f] = w;
            const float vals[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              sm[v] += vals[v];
              if (vals[v] > mx[v] || am[v] < 0) { mx[v] = vals[v]; am[v] = i; }
            }
          }
        }
      }
#pragma unroll
      for (int d = LPR; d < 32; d <<= 1) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float omx = __shfl_xor_sync(0xffffffffu, mx[v], d);
          int oam = __shfl_xor_sync(0xffffffffu, am[v], d);
          float osm = __shfl_xor_sync(0xffffffffu, sm[v], d);
          sm[v] = (sub & (d / LPR)) ? osm + sm[v] : sm[v] + osm;
          bool take = oam >= 0 && (am[v] < 0 || omx > mx[v] || (omx == mx[v] && oam < am[v]));
          if (take) { mx[v] = omx; am[v] = oam; }
        }
      }
      if (fok && sub == 0) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          out[(int64_t)g * out_stride + f * 4 + v] = am[v] >= 0 ? mx[v] : 0.f;
          argmax[(int64_t)g * F + f * 4 + v] = am[v] >= 0 ? (int)(lo + am[v]) : -1;
          out[(int64_t)g * out_stride + F + f * 4 + v] = sm[v] / (float)(n > 0 ? n : 1);
        }
      }
    }
  }
}

template <int VEC, int LPR>
__global__ void __launch_bounds__(RO_BWD_THREADS)
k_readout_bwd(const float* __restrict__ dout, int64_t dstride, const int* __restrict__ argmax,
              const int64_t* __restrict__ gptr, int G, int F, int mode, float* __restrict__ dx) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int FV = F / VEC;
  // CTA per graph (a warp per graph left 3,504 warps streaming 62 MB: 65 us): the 8 warps of a CTA interleave
  // the graph's rows, every thread keeps the graph's gradient row for its feature column(s) in registers
  const int warps_per_block = RO_BWD_THREADS / 32;
  const int wsub = (threadIdx.x >> 5) * RPW + sub;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int64_t lo = gptr[g], hi = gptr[g + 1];
    const int n = (int)(hi - lo);
    for (int f = l; f < FV; f += LPR) {
      float gm[VEC], gs[VEC]; int am[VEC];
      int col = 0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { gm[v] = 0.f; gs[v] = 0.f; am[v] = -1; }
      if (mode & TSG_READOUT_MAX) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          gm[v] = dout[(int64_t)g * dstride + f * VEC + v];
          am[v] = argmax[(int64_t)g * F + f * VEC + v];
        }
        col = F;
      }
      if (mode & (TSG_READOUT_MEAN | TSG_READOUT_SUM)) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float d = dout[(int64_t)g * dstride + col + f * VEC + v];
          gs[v] = (mode & TSG_READOUT_MEAN) ? d / (float)(n > 0 ? n : 1) : d;
        }
      }
      for (int i = wsub; i < n; i += RPW * warps_per_block) {
        float r[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = gs[v] + ((am[v] == (int)(lo + i)) ? gm[v] : 0.f);
        if (VEC == 4) {
          float4* p = reinterpret_cast<float4*>(dx) + (lo + i) * FV + f;
          float4 o = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
          if (mode & TSG_READOUT_ACCUM) { const float4 e = *p; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
          *p = o;
        } else {
          float* p = dx + (lo + i) * FV + f;
          *p = (mode & TSG_READOUT_ACCUM) ? *p + r[0] : r[0];
        }
      }
    }
  }
}

}  // namespace tsg

using namespace tsg;

template <int VEC>
static int launch_fwd(int lpr, int grid, cudaStream_t st, const float* x, const int64_t* gptr, int G,
                      int F, int mode, float* out, int64_t os, int* am) {
#define TSG_GO(L) k_readout_fwd<VEC, L><<<grid, RO_THREADS, 0, st>>>(x, gptr, G, F, mode, out, os, am)
  switch (lpr) {
    case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
    case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
  }
#undef TSG_GO
  return check_launch("readout_fwd");
}

template <int VEC>
static int launch_bwd(int lpr, int grid, cudaStream_t st, const float* dout, int64_t ds, const int* am,
                      const int64_t* gptr, int G, int F, int mode, float* dx) {
#define TSG_GO(L) k_readout_bwd<VEC, L><<<grid, RO_BWD_THREADS, 0, st>>>(dout, ds, am, gptr, G, F, mode, dx)
  switch (lpr) {
    case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
    case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
  }
#undef TSG_GO
  return check_launch("readout_bwd");
}

static int pick_lpr(int64_t fv) { int l = 1; while (l < fv && l < 32) l <<= 1; return l; }

extern "C" int tsg_readout_fwd(const float* x, const int64_t* gptr, int64_t G, int64_t F, int mode,
                               float* out, int64_t out_stride, int32_t* argmax, void* stream) {
  TSG_REQUIRE(G >= 0 && F > 0 && G < (int64_t)0x7fffffff, "readout_fwd: bad shape");
  TSG_REQUIRE((mode & (TSG_READOUT_MAX | TSG_READOUT_MEAN | TSG_READOUT_SUM)) != 0, "readout_fwd: empty mode");
  TSG_REQUIRE(!((mode & TSG_READOUT_MEAN) && (mode & TSG_READOUT_SUM)), "readout_fwd: mean and sum are exclusive");
  if (G == 0) return TSG_OK;
  TSG_REQUIRE(gptr && out, "readout_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(G, RO_THREADS / 32);
  bool vec = F % 4 == 0 && (((uintptr_t)x) & 15) == 0;
  if (vec) return launch_fwd<4>(pick_lpr(F / 4), grid, st, x, gptr, (int)G, (int)F, mode, out, out_stride, argmax);
  return launch_fwd<1>(pick_lpr(F), grid, st, x, gptr, (int)G, (int)F, mode, out, out_stride, argmax);
}

extern "C" int tsg_readout_bwd(const float* dout, int64_t dout_stride, const int32_t* argmax,
                               const int64_t* gptr, int64_t G, int64_t N, int64_t F, int mode,
                               float* dx, void* stream) {
  TSG_REQUIRE(G >= 0 && F > 0 && N >= 0 && G < (int64_t)0x7fffffff, "readout_bwd: bad shape");
  if (G == 0 || N == 0) return TSG_OK;
  TSG_REQUIRE(dout && gptr && dx, "readout_bwd: null pointer");
  TSG_REQUIRE(!(mode & TSG_READOUT_MAX) || argmax, "readout_bwd: max mode needs argmax");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(G, 1);
  bool vec = F % 4 == 0 && (((uintptr_t)dx) & 15) == 0;
  if (vec) return launch_bwd<4>(pick_lpr(F / 4), grid, st, dout, dout_stride, argmax, gptr, (int)G, (int)F, mode, dx);
  return launch_bwd<1>(pick_lpr(F), grid, st, dout, dout_stride, argmax, gptr, (int)G, (int)F, mode, dx);
}

/* gate (x[perm] * tanh(score[perm])) + [max || mean] readout in one pass; see k_gate_readout_fwd.  feat % 4 == 0. */
extern "C" int tsg_gate_readout_fwd(const float* x, const float* score, const int64_t* perm, const int64_t* gptr,
                                    int64_t G, int64_t F, float* xo, float* out, int64_t out_stride, int32_t* argmax,
                                    void* stream) {
  TSG_REQUIRE(G >= 0 && F > 0 && F % 4 == 0 && G < (int64_t)0x7fffffff, "gate_readout_fwd: bad shape");
  if (G == 0) return TSG_OK;
  TSG_REQUIRE(x && score && perm && gptr && xo && out && argmax, "gate_readout_fwd: null pointer");
  TSG_REQUIRE(((((uintptr_t)x) | ((uintptr_t)xo)) & 15) == 0, "gate_readout_fwd: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(G, RO_THREADS / 32);
  const int F4 = (int)(F / 4);
#define TSG_GR(L) k_gate_readout_fwd<L><<<grid, RO_THREADS, 0, st>>>((const float4*)x, score, perm, gptr, (int)G, F4, (float4*)xo, out, out_stride, argmax)
  switch (pick_lpr(F4)) {
    case 1: TSG_GR(1); break; case 2: TSG_GR(2); break; case 4: TSG_GR(4); break;
    case 8: TSG_GR(8); break; case 16: TSG_GR(16); break; default: TSG_GR(32); break;
  }
#undef TSG_GR
  return check_launch("gate_readout_fwd");
}
