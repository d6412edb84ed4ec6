// K12 -- stage-2 evaluation heads on graph embeddings.
//
// (a) tsg_knn_predict: the 3-nearest-neighbour classifier of `evaluate`
//     (Code/sage+gat+diffpool/train_triplet.py:80-85: sklearn KNeighborsClassifier(n_neighbors=3).fit(train)
//     .predict(val); Euclidean metric, uniform weights).  Warp per query: every lane scans a strided slice of
//     the train rows keeping its k best (distance, index) pairs, the lanes' lists are merged through shared
//     memory, the k winners vote; distance ties -> lower train index, vote ties -> lower class id (the rule of
//     scipy.stats.mode that sklearn's predict applies).
// (b) tsg_mlp1_train: the stage-2 classifier of `evaluate_mlp` (train_triplet.py:148-165): an MLP
//     in -> 64 -> 32 -> 2 with LeakyReLU(0.01), trained with Adam(lr 1e-3) on ONE embedding per step, in order --
//     inherently sequential, ~30 tiny kernels per sample when driven from Python.  Here ONE CTA keeps the
//     weights and both Adam moments in shared memory and runs all N steps (forward, cross-entropy, backward,
//     Adam) back to back; the weight gradients are outer products formed inside the update.
//     Arithmetic: fp32, torch's Adam formulas (bias-corrected step size, denom = sqrt(v)/sqrt(bc2) + eps).
#include "common.cuh"
#include <float.h>

namespace tsg {

constexpr int KNN_MAXK = 8;

__global__ void __launch_bounds__(128)
k_knn_predict(const float* __restrict__ train, const int64_t* __restrict__ labels, const float* __restrict__ query,
              int M, int Q, int D, int k, int num_classes, int64_t* __restrict__ pred) {
  extern __shared__ float sm[];                        // per warp: query row [D] + 32*k (dist, idx) pairs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* qrow = sm + (size_t)warp * (D + 64 * KNN_MAXK);
  float* cd = qrow + D;                                // [32 * KNN_MAXK] distances
  int* ci = reinterpret_cast<int*>(cd + 32 * KNN_MAXK);   // [32 * KNN_MAXK] indices
  for (int q = blockIdx.x * wpb + warp; q < Q; q += gridDim.x * wpb) {
    for (int d = lane; d < D; d += 32) qrow[d] = query[(int64_t)q * D + d];
    __syncwarp();
    float bd[KNN_MAXK]; int bi[KNN_MAXK];
#pragma unroll
    for (int j = 0; j < KNN_MAXK; ++j) { bd[j] = FLT_MAX; bi[j] = 0x7fffffff; }
    for (int r = lane; r < M; r += 32) {
      const float* t = train + (int64_t)r * D;
      float s = 0.f;
      for (int d = 0; d < D; ++d) { const float df = t[d] - qrow[d]; s = fmaf(df, df, s); }
      // insert (s, r) into the sorted list of the k best (ties -> lower index; r ascends, so strict <)
      if (s < bd[k - 1]) {
        int j = k - 1;
        while (j > 0 && s < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
        bd[j] = s; bi[j] = r;
      }
    }
#pragma unroll
    for (int j = 0; j < KNN_MAXK; ++j) if (j < k) { cd[lane * KNN_MAXK + j] = bd[j]; ci[lane * KNN_MAXK + j] = bi[j]; }
    __syncwarp();
    if (lane == 0) {
      // k rounds of selection over the 32*k candidates by (distance, index)
      int votes[16];
      for (int c = 0; c < 16; ++c) votes[c] = 0;
      int head[32];
      for (int l = 0; l < 32; ++l) head[l] = 0;
      int best_cls = 0;
      for (int round = 0; round < k && round < M; ++round) {
        float md = FLT_MAX; int mi = 0x7fffffff, ml = -1;
        for (int l = 0; l < 32; ++l) {
          if (head[l] >= k) continue;
          const float d = cd[l * KNN_MAXK + head[l]]; const int i = ci[l * KNN_MAXK + head[l]];
          if (i == 0x7fffffff) continue;
          if (d < md || (d == md && i < mi)) { md = d; mi = i; ml = l; }
        }
        if (ml < 0) break;
        ++head[ml];
        const int c = (int)labels[mi];
        if (c >= 0 && c < 16) ++votes[c];
      }
      int bv = -1;
      for (int c = 0; c < num_classes && c < 16; ++c) if (votes[c] > bv) { bv = votes[c]; best_cls = c; }
      pred[q] = best_cls;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ sequential MLP trainer
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : slope * x; }

struct AdamCfg { float lr, b1, b2, eps; };

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamCfg& a, float bc1, float sqrt_bc2) {
  m = m + (1.f - a.b1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = a.b2 * v + (1.f - a.b2) * g * g;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / sqrt_bc2 + a.eps;
  p = p - (a.lr / bc1) * (m / denom);
}

__global__ void __launch_bounds__(256)
k_mlp1_train(const float* __restrict__ emb, const int64_t* __restrict__ labels, int N, int D, int H1, int H2, int C,
             float* __restrict__ params, float* __restrict__ am, float* __restrict__ av, long long step0,
             AdamCfg cfg, float slope, float* __restrict__ losses) {
  extern __shared__ float sm[];
  const int P = H1 * D + H1 + H2 * H1 + H2 + C * H2 + C;
  float* w = sm; float* m = w + P; float* v = m + P;
  float* x = v + P; float* z1 = x + D; float* h1 = z1 + H1; float* dz1 = h1 + H1;
  float* z2 = dz1 + H1; float* h2 = z2 + H2; float* dz2 = h2 + H2; float* dl = dz2 + H2;   // dl[C]
  const int oW1 = 0, ob1 = oW1 + H1 * D, oW2 = ob1 + H1, ob2 = oW2 + H2 * H1, oW3 = ob2 + H2, ob3 = oW3 + C * H2;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { w[i] = params[i]; m[i] = am[i]; v[i] = av[i]; }
  __syncthreads();
  for (int s = 0; s < N; ++s) {
    const long long t = step0 + s + 1;
    // torch evaluates the bias corrections in double on the host (1 - beta ** step) and rounds once
    const float bc1 = (float)(1.0 - pow((double)cfg.b1, (double)t));
    const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)cfg.b2, (double)t));
    for (int d = threadIdx.x; d < D; d += blockDim.x) x[d] = emb[(int64_t)s * D + d];
    __syncthreads();
    if (threadIdx.x < H1) {
      float a = w[ob1 + threadIdx.x];
      const float* wr = w + oW1 + threadIdx.x * D;
      for (int d = 0; d < D; ++d) a = fmaf(wr[d], x[d], a);
      z1[threadIdx.x] = a; h1[threadIdx.x] = leaky(a, slope);
    }
    __syncthreads();
    if (threadIdx.x < H2) {
      float a = w[ob2 + threadIdx.x];
      const float* wr = w + oW2 + threadIdx.x * H1;
      for (int d = 0; d < H1; ++d) a = fmaf(wr[d], h1[d], a);
      z2[threadIdx.x] = a; h2[threadIdx.x] = leaky(a, slope);
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // logits, softmax, cross-entropy (C is tiny)
      float mx = -FLT_MAX;
      for (int c = 0; c < C; ++c) {
        float a = w[ob3 + c];
        for (int d = 0; d < H2; ++d) a = fmaf(w[oW3 + c * H2 + d], h2[d], a);
        dl[c] = a; mx = fmaxf(mx, a);
      }
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(dl[c] - mx);
      const int y = (int)labels[s];
      if (losses) losses[s] = logf(se) + mx - dl[y];
      for (int c = 0; c < C; ++c) dl[c] = expf(dl[c] - mx) / se - (c == y ? 1.f : 0.f);
    }
    __syncthreads();
    if (threadIdx.x < H2) {                              // dz2 = (W3^T dl) * leaky'(z2)
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(w[oW3 + c * H2 + threadIdx.x], dl[c], a);
      dz2[threadIdx.x] = a * (z2[threadIdx.x] > 0.f ? 1.f : slope);
    }
    __syncthreads();
    if (threadIdx.x < H1) {                              // dz1 = (W2^T dz2) * leaky'(z1)
      float a = 0.f;
      for (int o = 0; o < H2; ++o) a = fmaf(w[oW2 + o * H1 + threadIdx.x], dz2[o], a);
      dz1[threadIdx.x] = a * (z1[threadIdx.x] > 0.f ? 1.f : slope);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) {  // gradients are outer products: formed in the update
      float g;
      if (i < ob1) { const int o = i / D, d = i - o * D; g = dz1[o] * x[d]; }
      else if (i < oW2) g = dz1[i - ob1];
      else if (i < ob2) { const int j = i - oW2, o = j / H1, d = j - o * H1; g = dz2[o] * h1[d]; }
      else if (i < oW3) g = dz2[i - ob2];
      else if (i < ob3) { const int j = i - oW3, o = j / H2, d = j - o * H2; g = dl[o] * h2[d]; }
      else g = dl[i - ob3];
      adam_update(w[i], m[i], v[i], g, cfg, bc1, sqrt_bc2);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < P; i += blockDim.x) { params[i] = w[i]; am[i] = m[i]; av[i] = v[i]; }
}

}  // namespace tsg

using namespace tsg;

extern "C" int tsg_knn_predict(const float* train, const int64_t* train_labels, const float* query,
                               int64_t num_train, int64_t num_query, int64_t dim, int k, int num_classes,
                               int64_t* pred, void* stream) {
  TSG_REQUIRE(num_train > 0 && num_query >= 0 && dim > 0, "knn_predict: bad shape");
  TSG_REQUIRE(k >= 1 && k <= KNN_MAXK, "knn_predict: k must be in [1, %d]", KNN_MAXK);
  TSG_REQUIRE(num_classes >= 1 && num_classes <= 16, "knn_predict: at most 16 classes");
  TSG_REQUIRE(num_train < (int64_t)0x7fffffff && num_query < (int64_t)0x7fffffff, "knn_predict: too large");
  if (num_query == 0) return TSG_OK;
  TSG_REQUIRE(train && train_labels && query && pred, "knn_predict: null pointer");
  const int wpb = 4;
  const size_t smem = (size_t)wpb * ((size_t)dim + 64 * KNN_MAXK) * sizeof(float);
  TSG_REQUIRE(smem <= 200 * 1024, "knn_predict: embedding width %lld too large", (long long)dim);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_knn_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_knn_predict<<<grid_for(num_query, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(train, train_labels, query, (int)num_train,
                                                                                   (int)num_query, (int)dim, k, num_classes, pred);
  return check_launch("knn_predict");
}

extern "C" int tsg_mlp1_train(const float* emb, const int64_t* labels, int64_t num_samples, int64_t dim,
                              int64_t hidden1, int64_t hidden2, int64_t num_classes,
                              float* params, float* adam_m, float* adam_v, int64_t step0,
                              float lr, float beta1, float beta2, float eps, float slope,
                              float* losses, void* stream) {
  TSG_REQUIRE(num_samples >= 0 && dim > 0 && hidden1 > 0 && hidden2 > 0 && num_classes > 0, "mlp1_train: bad shape");
  TSG_REQUIRE(hidden1 <= 256 && hidden2 <= 256 && num_classes <= 32, "mlp1_train: layer too wide for one CTA");
  if (num_samples == 0) return TSG_OK;
  TSG_REQUIRE(emb && labels && params && adam_m && adam_v, "mlp1_train: null pointer");
  const size_t P = (size_t)(hidden1 * dim + hidden1 + hidden2 * hidden1 + hidden2 + num_classes * hidden2 + num_classes);
  const size_t smem = (3 * P + (size_t)dim + 3 * hidden1 + 3 * hidden2 + num_classes + 8) * sizeof(float);
  TSG_REQUIRE(smem <= 220 * 1024, "mlp1_train: %zu parameters do not fit shared memory", P);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_mlp1_train, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  AdamCfg cfg{lr, beta1, beta2, eps};
  k_mlp1_train<<<1, 256, smem, (cudaStream_t)stream>>>(emb, labels, (int)num_samples, (int)dim, (int)hidden1, (int)hidden2,
                                                      (int)num_classes, params, adam_m, adam_v, (long long)step0, cfg, slope, losses);
  return check_launch("mlp1_train");
}
