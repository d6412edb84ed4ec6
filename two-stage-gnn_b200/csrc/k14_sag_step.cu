// K14 -- the whole 2stg SAGPool training step behind ONE C-ABI call.
//
// Round 1 enqueued the encoder from C++ (K10) but the head MLP, the triplet loss and their backward went through
// Python: three autograd Functions for the Linear layers, aten ReLU / dropout / log_softmax, the triplet Function --
// 1.46 ms of host time per 1.69 ms step, the limit of 8-GPU scaling (eight processes share the box's cores and every
// rank waits for the slowest at the all-reduce).  Here the step is
//
//   encoder forward (K10)  ->  head forward (k_head_fwd)  ->  triplet loss (K9)  ->  triplet backward (K9)
//   ->  head backward (k_head_bwd, fixed-order parameter gradients)  ->  encoder backward (K10)
//
// enqueued by tsg_sag_triplet_step_compact without returning to the interpreter; Python keeps the sampler, one
// NCCL all-reduce over the flat gradient buffer and the optimiser.
//
// Head = Code/sag/network.py:48-52: lin1 (2H -> H) / ReLU / dropout(p) / lin2 (H -> H/2) / ReLU / lin3 (H/2 -> C) /
// log_softmax.  One warp per graph, weights in shared memory (rows padded to an odd pitch), sequential dot products.
// Dropout: a caller-provided keep-multiplier matrix (parity tests inject the oracle's) or a counter-based hash of
// (seed, graph, column) -- the reference draws from torch's unseeded-per-step generator, so no mask sequence can be
// "the reference's"; SURVEY A.2 (RNG).
#include "common.cuh"

namespace tsg {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_WARPS = HEAD_THREADS / 32;

struct HeadDims { int G, H, H2, C; };

__device__ __forceinline__ unsigned head_hash(unsigned long long seed, unsigned idx) {
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(idx + 1u);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return (unsigned)((x ^ (x >> 31)) >> 32);
}

// shared-memory image of the head parameters: W1 [H][2H+1], b1 [H], W2 [H2][H+1], b2 [H2], W3 [C][H2+1], b3 [C]
__host__ __device__ inline int head_param_floats(int H, int H2, int C) {
  return H * (2 * H + 1) + H + H2 * (H + 1) + H2 + C * (H2 + 1) + C;
}

__device__ void head_load_params(const float* const* hp, float* sm, HeadDims d) {
  const int p1 = 2 * d.H + 1, p2 = d.H + 1, p3 = d.H2 + 1;
  float* W1 = sm; float* b1 = W1 + d.H * p1; float* W2 = b1 + d.H; float* b2 = W2 + d.H2 * p2;
  float* W3 = b2 + d.H2; float* b3 = W3 + d.C * p3;
  for (int i = threadIdx.x; i < d.H * 2 * d.H; i += blockDim.x) W1[(i / (2 * d.H)) * p1 + i % (2 * d.H)] = hp[0][i];
  for (int i = threadIdx.x; i < d.H; i += blockDim.x) b1[i] = hp[1][i];
  for (int i = threadIdx.x; i < d.H2 * d.H; i += blockDim.x) W2[(i / d.H) * p2 + i % d.H] = hp[2][i];
  for (int i = threadIdx.x; i < d.H2; i += blockDim.x) b2[i] = hp[3][i];
  for (int i = threadIdx.x; i < d.C * d.H2; i += blockDim.x) W3[(i / d.H2) * p3 + i % d.H2] = hp[4][i];
  for (int i = threadIdx.x; i < d.C; i += blockDim.x) b3[i] = hp[5][i];
}

struct HeadPtrs { const float* hp[6]; };

// z [G, 2H] -> a1 [G, H] (after ReLU and dropout), m [G, H] (dropout keep multiplier), a2 [G, H2], emb [G, C]
__global__ void __launch_bounds__(HEAD_THREADS)
k_head_fwd(HeadPtrs P, const float* __restrict__ z, HeadDims d, float dropout_p, unsigned long long seed,
           const float* __restrict__ mask_in, float* __restrict__ a1, float* __restrict__ m, float* __restrict__ a2,
           float* __restrict__ emb) {
  extern __shared__ __align__(16) float hsm[];
  float* prm = hsm;
  float* stage = hsm + head_param_floats(d.H, d.H2, d.C);          // per warp: 2H + H + H2 + C floats
  head_load_params(P.hp, prm, d);
  __syncthreads();
  const int p1 = 2 * d.H + 1, p2 = d.H + 1, p3 = d.H2 + 1;
  const float* W1 = prm; const float* b1 = W1 + d.H * p1; const float* W2 = b1 + d.H; const float* b2 = W2 + d.H2 * p2;
  const float* W3 = b2 + d.H2; const float* b3 = W3 + d.C * p3;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sz = stage + (size_t)w * (2 * d.H + d.H + d.H2 + d.C);
  float* s1 = sz + 2 * d.H; float* s2 = s1 + d.H; float* s3 = s2 + d.H2;
  const float scale = dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f;
  for (int g = blockIdx.x * HEAD_WARPS + w; g < d.G; g += gridDim.x * HEAD_WARPS) {
    for (int i = lane; i < 2 * d.H; i += 32) sz[i] = z[(size_t)g * 2 * d.H + i];
    __syncwarp();
    for (int j = lane; j < d.H; j += 32) {                          // lin1 + ReLU + dropout      network.py:48-49
      float acc = 0.f;
      for (int k = 0; k < 2 * d.H; ++k) acc = fmaf(sz[k], W1[j * p1 + k], acc);
      acc = fmaxf(acc + b1[j], 0.f);
      float keep;
      if (mask_in) keep = mask_in[(size_t)g * d.H + j];
      else if (dropout_p > 0.f) keep = (head_hash(seed, (unsigned)(g * d.H + j)) >> 8) * (1.0f / 16777216.0f) >= dropout_p ? scale : 0.f;
      else keep = 1.f;
      acc *= keep;
      s1[j] = acc;
      a1[(size_t)g * d.H + j] = acc; m[(size_t)g * d.H + j] = keep;
    }
    __syncwarp();
    for (int j = lane; j < d.H2; j += 32) {                         // lin2 + ReLU                 network.py:50
      float acc = 0.f;
      for (int k = 0; k < d.H; ++k) acc = fmaf(s1[k], W2[j * p2 + k], acc);
      acc = fmaxf(acc + b2[j], 0.f);
      s2[j] = acc; a2[(size_t)g * d.H2 + j] = acc;
    }
    __syncwarp();
    float mx = -3.4e38f;
    for (int j = lane; j < d.C; j += 32) {                          // lin3                        network.py:51
      float acc = 0.f;
      for (int k = 0; k < d.H2; ++k) acc = fmaf(s2[k], W3[j * p3 + k], acc);
      acc += b3[j];
      s3[j] = acc; mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int j = lane; j < d.C; j += 32) se += expf(s3[j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = mx + logf(se);
    for (int j = lane; j < d.C; j += 32) emb[(size_t)g * d.C + j] = s3[j] - lse;      // log_softmax
    __syncwarp();
  }
}

// demb [G, C] -> dz [G, 2H]; parameter gradients as ONE partial row per CTA (fixed graph ranges, fixed order inside)
// layout of a partial row / of the packed head gradient: [dW1 H*2H | db1 H | dW2 H2*H | db2 H2 | dW3 C*H2 | db3 C]
__host__ __device__ inline int head_grad_floats(int H, int H2, int C) { return H * 2 * H + H + H2 * H + H2 + C * H2 + C; }

__global__ void __launch_bounds__(HEAD_THREADS)
k_head_bwd(HeadPtrs P, const float* __restrict__ z, HeadDims d, const float* __restrict__ a1, const float* __restrict__ m,
           const float* __restrict__ a2, const float* __restrict__ emb, const float* __restrict__ demb,
           float* __restrict__ dz, float* __restrict__ part) {
  extern __shared__ __align__(16) float hsm[];
  float* prm = hsm;
  const int per = 2 * d.H + d.H + d.H + d.H2 + d.H2 + d.C;         // staged per graph: z, a1, da1, a2, da2, do
  float* stage = hsm + head_param_floats(d.H, d.H2, d.C);          // HEAD_WARPS graphs
  float* acc = stage + (size_t)HEAD_WARPS * per;                   // this CTA's gradient row
  head_load_params(P.hp, prm, d);
  const int NG = head_grad_floats(d.H, d.H2, d.C);
  for (int i = threadIdx.x; i < NG; i += HEAD_THREADS) acc[i] = 0.f;
  __syncthreads();
  const int p1 = 2 * d.H + 1, p2 = d.H + 1, p3 = d.H2 + 1;
  const float* W1 = prm; const float* W2 = W1 + d.H * p1 + d.H; const float* W3 = W2 + d.H2 * p2 + d.H2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sz = stage + (size_t)w * per;
  float* s1 = sz + 2 * d.H; float* d1 = s1 + d.H; float* s2 = d1 + d.H; float* d2 = s2 + d.H2; float* dO = d2 + d.H2;
  const int gpc = (d.G + gridDim.x - 1) / gridDim.x;               // contiguous graph range of this CTA
  const int g_begin = blockIdx.x * gpc, g_end = min(d.G, g_begin + gpc);
  for (int g0 = g_begin; g0 < g_end; g0 += HEAD_WARPS) {
    const int g = g0 + w;
    const bool on = g < g_end;
    if (on) {
      for (int i = lane; i < 2 * d.H; i += 32) sz[i] = z[(size_t)g * 2 * d.H + i];
      for (int i = lane; i < d.H; i += 32) s1[i] = a1[(size_t)g * d.H + i];
      for (int i = lane; i < d.H2; i += 32) s2[i] = a2[(size_t)g * d.H2 + i];
      // log_softmax backward: do = demb - softmax * sum(demb)
      float sd = 0.f;
      for (int j = lane; j < d.C; j += 32) sd += demb[(size_t)g * d.C + j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
      for (int j = lane; j < d.C; j += 32) dO[j] = demb[(size_t)g * d.C + j] - expf(emb[(size_t)g * d.C + j]) * sd;
      __syncwarp();
      for (int k = lane; k < d.H2; k += 32) {                       // da2 = (do W3) * relu'
        float t = 0.f;
        for (int j = 0; j < d.C; ++j) t = fmaf(dO[j], W3[j * p3 + k], t);
        d2[k] = s2[k] > 0.f ? t : 0.f;
      }
      __syncwarp();
      for (int k = lane; k < d.H; k += 32) {                        // da1 = (da2 W2) * keep * relu'
        float t = 0.f;
        for (int j = 0; j < d.H2; ++j) t = fmaf(d2[j], W2[j * p2 + k], t);
        // a1 is stored AFTER dropout: a1 > 0 <=> pre-dropout activation > 0 and kept
        d1[k] = s1[k] > 0.f ? t * m[(size_t)g * d.H + k] : 0.f;
      }
      __syncwarp();
      for (int k = lane; k < 2 * d.H; k += 32) {                    // dz = da1 W1
        float t = 0.f;
        for (int j = 0; j < d.H; ++j) t = fmaf(d1[j], W1[j * p1 + k], t);
        dz[(size_t)g * 2 * d.H + k] = t;
      }
    }
    __syncthreads();
    // every thread owns a fixed set of gradient entries and adds this batch of graphs in graph order
    const int nb = min(HEAD_WARPS, g_end - g0);
    for (int i = threadIdx.x; i < NG; i += HEAD_THREADS) {
      float t = acc[i];
      int o = i;
      if (o < d.H * 2 * d.H) {                                       // dW1[j][k] += da1_pre[j] * z[k]; a1 holds post-dropout a1
        const int j = o / (2 * d.H), k = o % (2 * d.H);
        for (int b = 0; b < nb; ++b) { const float* s = stage + (size_t)b * per; t = fmaf(s[3 * d.H + j], s[k], t); }
      } else if ((o -= d.H * 2 * d.H) < d.H) {
        for (int b = 0; b < nb; ++b) t += (stage + (size_t)b * per)[3 * d.H + o];
      } else if ((o -= d.H) < d.H2 * d.H) {                          // dW2[j][k] += da2[j] * a1[k]
        const int j = o / d.H, k = o % d.H;
        for (int b = 0; b < nb; ++b) { const float* s = stage + (size_t)b * per; t = fmaf(s[4 * d.H + d.H2 + j], s[2 * d.H + k], t); }
      } else if ((o -= d.H2 * d.H) < d.H2) {
        for (int b = 0; b < nb; ++b) t += (stage + (size_t)b * per)[4 * d.H + d.H2 + o];
      } else if ((o -= d.H2) < d.C * d.H2) {                         // dW3[j][k] += do[j] * a2[k]
        const int j = o / d.H2, k = o % d.H2;
        for (int b = 0; b < nb; ++b) { const float* s = stage + (size_t)b * per; t = fmaf(s[4 * d.H + 2 * d.H2 + j], s[4 * d.H + k], t); }
      } else {
        o -= d.C * d.H2;
        for (int b = 0; b < nb; ++b) t += (stage + (size_t)b * per)[4 * d.H + 2 * d.H2 + o];
      }
      acc[i] = t;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NG; i += HEAD_THREADS) part[(size_t)blockIdx.x * NG + i] = acc[i];
}

// the packed head gradient row -> the six parameter gradient tensors
__global__ void __launch_bounds__(256)
k_head_unpack(const float* __restrict__ packed, float* g0, float* g1, float* g2, float* g3, float* g4, float* g5,
              int n0, int n1, int n2, int n3, int n4, int n5) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int o = i;
  if (o < n0) { g0[o] = packed[i]; return; }
  if ((o -= n0) < n1) { g1[o] = packed[i]; return; }
  if ((o -= n1) < n2) { g2[o] = packed[i]; return; }
  if ((o -= n2) < n3) { g3[o] = packed[i]; return; }
  if ((o -= n3) < n4) { g4[o] = packed[i]; return; }
  if ((o -= n4) < n5) { g5[o] = packed[i]; return; }
}

__global__ void k_store_one(float* p, float v) { *p = v; }

static int head_ctas(int G) {
  // one batch of HEAD_WARPS graphs per CTA up to two CTAs per SM (ncu, first cut: 110 CTAs of 4 batches = 12 % warps
  // active, 55 us for 3,504 graphs -- the kernel is latency bound, so it wants every SM busy)
  int c = (G + HEAD_WARPS - 1) / HEAD_WARPS;
  if (c > 2 * TSG_NUM_SMS) c = 2 * TSG_NUM_SMS;
  return c < 1 ? 1 : c;
}

}  // namespace tsg

using namespace tsg;

#define TSG_TRY(call)                 \
  do {                                \
    int _rc = (call);                 \
    if (_rc != TSG_OK) return _rc;    \
  } while (0)

static bool head_ok(const tsg_sag_shape* sh, const tsg_sag_head* hd) {
  if (!sh || !hd || hd->num_classes <= 0 || hd->num_triplets <= 0 || sh->hidden < 2 || sh->hidden % 2) return false;
  const int H = (int)sh->hidden, H2 = H / 2, C = (int)hd->num_classes;
  const size_t fwd = ((size_t)head_param_floats(H, H2, C) + (size_t)HEAD_WARPS * (3 * H + H2 + C)) * 4;
  const size_t bwd = ((size_t)head_param_floats(H, H2, C) + (size_t)HEAD_WARPS * (4 * H + 2 * H2 + C) + head_grad_floats(H, H2, C)) * 4;
  return fwd <= 200 * 1024 && bwd <= 200 * 1024 && hd->dropout_p >= 0.f && hd->dropout_p < 1.f;
}

struct StepWs {
  float *z, *a1, *m, *a2, *emb, *demb, *dz, *dp, *dn, *part, *packed, *one;
  void* trip_ws; size_t trip_bytes;
  size_t total;
};

static void step_layout(const tsg_sag_shape* sh, const tsg_sag_head* hd, void* base, StepWs* w) {
  char* b = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* { void* p = b ? (void*)(b + off) : nullptr; off += align_up(bytes, 256); return p; };
  const int64_t G = sh->num_graphs, H = sh->hidden, H2 = H / 2, C = hd->num_classes, T = hd->num_triplets;
  w->z = (float*)take(G * 2 * H * 4); w->a1 = (float*)take(G * H * 4); w->m = (float*)take(G * H * 4);
  w->a2 = (float*)take(G * H2 * 4); w->emb = (float*)take(G * C * 4); w->demb = (float*)take(G * C * 4);
  w->dz = (float*)take(G * 2 * H * 4); w->dp = (float*)take(T * 4); w->dn = (float*)take(T * 4);
  const size_t NG = head_grad_floats((int)H, (int)H2, (int)C);
  w->part = (float*)take((size_t)head_ctas((int)G) * NG * 4); w->packed = (float*)take(NG * 4);
  w->one = (float*)take(256);
  w->trip_bytes = tsg_triplet_workspace_bytes(T, G, C);
  w->trip_ws = take(w->trip_bytes);
  w->total = off;
}

extern "C" size_t tsg_sag_triplet_step_workspace_bytes(const tsg_sag_shape* sh, const tsg_sag_head* hd) {
  if (!head_ok(sh, hd)) return 0;
  StepWs w;
  step_layout(sh, hd, nullptr, &w);
  return w.total;
}

static void head_attr_once() {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_head_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_head_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
}

// encoder forward (K10) + head forward -> emb [G, C]
static int step_fwd(const tsg_sag_shape* sh, const tsg_sag_head* hd, const int32_t* label, const int32_t* local_row,
                    const int32_t* local_col, const int64_t* edge_ptr, const int64_t* level_ptr, const float* const* params,
                    const float* dropout_mask, float* emb, const StepWs& w, void* arena, size_t arena_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int H = (int)sh->hidden, H2 = H / 2, C = (int)hd->num_classes, G = (int)sh->num_graphs;
  HeadDims d{G, H, H2, C};
  HeadPtrs P;
  for (int i = 0; i < 6; ++i) P.hp[i] = params[12 + i];
  TSG_TRY(tsg_sag_encoder_fwd_compact(sh, label, local_row, local_col, edge_ptr, level_ptr, params, w.z, arena, arena_bytes, stream));
  const size_t sm_f = ((size_t)head_param_floats(H, H2, C) + (size_t)HEAD_WARPS * (3 * H + H2 + C)) * 4;
  head_attr_once();
  k_head_fwd<<<grid_for(G, HEAD_WARPS, 2), HEAD_THREADS, sm_f, st>>>(P, w.z, d, hd->dropout_p, (unsigned long long)hd->seed,
                                                                       dropout_mask, w.a1, w.m, w.a2, emb);
  return check_launch("sag_step(head fwd)");
}

// head backward (dz + fixed-order parameter gradients) + encoder backward (K10), from d(loss)/d(emb)
static int step_bwd(const tsg_sag_shape* sh, const tsg_sag_head* hd, const int32_t* label, const int64_t* level_ptr,
                    const float* const* params, const float* emb, const float* demb, float* const* grads, const StepWs& w,
                    void* arena, size_t arena_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int H = (int)sh->hidden, H2 = H / 2, C = (int)hd->num_classes, G = (int)sh->num_graphs;
  HeadDims d{G, H, H2, C};
  HeadPtrs P;
  for (int i = 0; i < 6; ++i) P.hp[i] = params[12 + i];
  const size_t sm_b = ((size_t)head_param_floats(H, H2, C) + (size_t)HEAD_WARPS * (4 * H + 2 * H2 + C) + head_grad_floats(H, H2, C)) * 4;
  head_attr_once();
  const int ctas = head_ctas(G);
  const int NG = head_grad_floats(H, H2, C);
  k_head_bwd<<<ctas, HEAD_THREADS, sm_b, st>>>(P, w.z, d, w.a1, w.m, w.a2, emb, demb, w.dz, w.part);
  TSG_LAUNCH_CHECK("sag_step(head bwd)");
  launch_partial_sum_final(w.part, w.packed, NG, nullptr, ctas, NG, st);
  k_head_unpack<<<(NG + 255) / 256, 256, 0, st>>>(w.packed, grads[12], grads[13], grads[14], grads[15], grads[16], grads[17],
                                                   H * 2 * H, H, H2 * H, H2, C * H2, C);
  TSG_LAUNCH_CHECK("sag_step(head grads)");
  return tsg_sag_encoder_bwd_compact(sh, label, level_ptr, params, w.dz, grads, arena, arena_bytes, stream);
}

extern "C" int tsg_sag_triplet_step_compact(const tsg_sag_shape* sh, const tsg_sag_head* hd, const int32_t* label,
                                            const int32_t* local_row, const int32_t* local_col, const int64_t* edge_ptr,
                                            const int64_t* level_ptr, const float* const* params, const int64_t* triplets,
                                            const float* dropout_mask, float* const* grads, float* loss, float* emb_out,
                                            void* arena, size_t arena_bytes, void* workspace, size_t workspace_bytes,
                                            void* stream) {
  TSG_REQUIRE(head_ok(sh, hd), "sag_triplet_step: bad head / shape (hidden must be even, head must fit shared memory)");
  TSG_REQUIRE(params && grads && triplets && loss && workspace, "sag_triplet_step: null pointer");
  StepWs w;
  step_layout(sh, hd, workspace, &w);
  if (workspace_bytes < w.total) { set_error("sag_triplet_step: workspace too small (%zu < %zu)", workspace_bytes, w.total); return TSG_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int C = (int)hd->num_classes, G = (int)sh->num_graphs;
  const int64_t T = hd->num_triplets;
  float* emb = emb_out ? emb_out : w.emb;
  TSG_TRY(step_fwd(sh, hd, label, local_row, local_col, edge_ptr, level_ptr, params, dropout_mask, emb, w, arena, arena_bytes, stream));
  // triplet loss forward + backward (K9), d(loss) = 1
  TSG_TRY(tsg_triplet_fwd(emb, triplets, T, G, C, hd->margin, hd->eps, w.dp, w.dn, loss, w.trip_ws, w.trip_bytes, stream));
  k_store_one<<<1, 1, 0, st>>>(w.one, 1.0f);
  TSG_TRY(tsg_triplet_bwd(emb, triplets, T, G, C, hd->margin, hd->eps, w.dp, w.dn, w.one, w.demb, w.trip_ws, w.trip_bytes, stream));
  return step_bwd(sh, hd, label, level_ptr, params, emb, w.demb, grads, w, arena, arena_bytes, stream);
}

/* The two halves of the step for losses evaluated OUTSIDE the call (the all-gather formulation: embeddings of every rank
 * are gathered, the global triplet loss is evaluated on the gathered matrix, this rank's slice of its gradient comes
 * back).  Same arena / workspace in both calls; num_triplets of `head` only sizes the workspace (>= 1). */
extern "C" int tsg_sag_step_fwd_compact(const tsg_sag_shape* sh, const tsg_sag_head* hd, const int32_t* label,
                                        const int32_t* local_row, const int32_t* local_col, const int64_t* edge_ptr,
                                        const int64_t* level_ptr, const float* const* params, const float* dropout_mask,
                                        float* emb, void* arena, size_t arena_bytes, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  TSG_REQUIRE(head_ok(sh, hd), "sag_step_fwd: bad head / shape");
  TSG_REQUIRE(params && emb && workspace, "sag_step_fwd: null pointer");
  StepWs w;
  step_layout(sh, hd, workspace, &w);
  if (workspace_bytes < w.total) { set_error("sag_step_fwd: workspace too small"); return TSG_EWORKSPACE; }
  return step_fwd(sh, hd, label, local_row, local_col, edge_ptr, level_ptr, params, dropout_mask, emb, w, arena, arena_bytes, stream);
}

extern "C" int tsg_sag_step_bwd_compact(const tsg_sag_shape* sh, const tsg_sag_head* hd, const int32_t* label,
                                        const int64_t* level_ptr, const float* const* params, const float* emb,
                                        const float* demb, float* const* grads, void* arena, size_t arena_bytes,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  TSG_REQUIRE(head_ok(sh, hd), "sag_step_bwd: bad head / shape");
  TSG_REQUIRE(params && emb && demb && grads && workspace, "sag_step_bwd: null pointer");
  StepWs w;
  step_layout(sh, hd, workspace, &w);
  if (workspace_bytes < w.total) { set_error("sag_step_bwd: workspace too small"); return TSG_EWORKSPACE; }
  return step_bwd(sh, hd, label, level_ptr, params, emb, demb, grads, w, arena, arena_bytes, stream);
}
