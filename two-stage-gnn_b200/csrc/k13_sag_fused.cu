// K13 -- graph-resident SAGPool encoder FORWARD (Code/sag/network.py:33-46 + layers.py:14-26 over a packed batch).
//
// The kernel-per-operator executor (k10_sag_exec.cu) runs 36 kernels per forward and every level's xw, h, xg and CSR make
// a round trip through HBM between them.  A DD graph is 269 nodes x 32 floats = 34 KB plus a 3 KB CSR: it fits an SM's
// shared memory many times over.  Here ONE CTA owns ONE graph at a time and carries it through all three levels:
//
//   edge list -> CSR of A_hat (the list is coalesced: a row's edges are a run, so row starts fall out of the run
//   boundaries -- no counting pass, no atomics, no scan; symmetric, so the same CSR serves both orientations) ->
//   conv_l = ReLU(A_hat (x_l W_l) + b_l) (level 0: W_1[label] gathered from the L1-resident table) ->
//   score_l = A_hat (h_l ws_l) + bs_l -> top-k (every warp sorts a run of keys in registers, ranks by binary search
//   across runs: two barriers instead of the 45 of a 512-key bitonic network) -> gate -> [max || mean] readout ->
//   next level's CSR (filter + renormalise) -> next level's x W in place, all in shared memory.
//
// HBM sees labels + edges in (4n + 8E bytes) and h_l / score_l / perm_l / argmax_l + z out.  Launches are per SIZE CLASS
// (fz_classes): the shared-memory layout of a launch is sized for its class, so the 73 % of DD graphs below ~330 nodes run
// four 256-thread CTAs per SM instead of paying for the 1,000-node graph in the batch.
//
// Arithmetic is the executor's, term by term and in the same order: sequential per-row FMA accumulation in CSR order,
// bias added after, the h . ws dot as the per-lane FMA chain + xor butterfly of k_spmm_g, the score SpMM with separately
// rounded products, x W as the k-ascending FMA chain of k_linear_fwd_dense, the readout mean as four strided partial
// sums combined pairwise.  Outputs are bit-identical to the executor's (tests/test_sag_fused_gpu.py).
//
// Forward only (tsg_sag_encoder_embed_compact: what torch.no_grad() forwards take -- the evaluation loops of the
// reference embed every graph without a backward).  Measured on B200 (profiles/r02_fused_forward.md): 760 us for the
// 3,504-graph bench batch against ~860 us of operator kernels, 3 launches instead of 36; the kernel is bound by issue
// slots (46 % issue utilisation, 54 k warp instructions per 269-node graph: index arithmetic of 6-neighbour rows, row
// prologues / epilogues), not by memory -- which is why the training step keeps the operator kernels, whose saved
// tensors the backward needs anyway.
//
// Requirements (checked on the host: fused_supported; otherwise the executor runs): hidden in {32, 64, 128}, the
// per-graph working set fits 227 KB of shared memory, < 65,536 CSR entries and <= 32,767 nodes per graph.  Per graph,
// in the kernel: endpoints in range, no self loops, list sorted by (row, col) and symmetric -- what TUDataset / the TU
// loader / tsg.synth produce; a violation sets a status bit (the host raises) and the graph is skipped.
#include "sag_fused.cuh"
#include <float.h>

namespace tsg {

typedef unsigned short u16;

struct FzLayout {
  unsigned rpA, rpB, colA, colB, disA, disB, label, w1, w23, vecs, sw, score, keys, perml, inv, ibuf, feat, outs, misc, total;
};

#define FZ_TAKE(field, bytes) do { o.field = (unsigned)off; off += (((size_t)(bytes)) + 15) & ~(size_t)15; } while (0)
__host__ __device__ inline FzLayout fz_layout(int H, int L, int nmax0, int nmax1, int emax) {
  FzLayout o;
  size_t off = 0;
  const int nnz = emax + nmax0;
  int npad = 32;
  while (npad < nmax0) npad <<= 1;
  FZ_TAKE(rpA, (nmax0 + 2) * 2); FZ_TAKE(rpB, (nmax0 + 2) * 2);
  FZ_TAKE(colA, nnz * 2); FZ_TAKE(colB, nnz * 2);
  FZ_TAKE(disA, nmax0 * 4); FZ_TAKE(disB, nmax0 * 4);
  FZ_TAKE(label, nmax0 * 4);
  o.w1 = 0;                                   // conv1's [L, H] table is gathered through L1 (__ldg), not staged
  FZ_TAKE(w23, (size_t)2 * H * H * 4); FZ_TAKE(vecs, (6 * H + 4) * 4);
  FZ_TAKE(sw, nmax0 * 4); FZ_TAKE(score, nmax0 * 4);
  FZ_TAKE(keys, (size_t)(npad > 512 ? npad : 512) * 8); FZ_TAKE(perml, nmax0 * 2); FZ_TAKE(inv, nmax0 * 2);
  FZ_TAKE(ibuf, (nmax0 + 2) * 4);
  FZ_TAKE(feat, (size_t)nmax1 * H * 4);
  FZ_TAKE(outs, 3 * 2 * H * 4); FZ_TAKE(misc, 256);
  o.total = (unsigned)off;
  return o;
}

// order-preserving map (same as k5_topk.cu): larger key <=> earlier in a descending sort. NaN -> max, -0 -> +0.
__device__ __forceinline__ uint32_t fz_score_key(float s) {
  if (s != s) return 0xFFFFFFFFu;
  if (s == 0.f) s = 0.f;
  uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// in-place exclusive scan of a[0..n) (ints in shared memory), total -> a[n].  All FZ_T threads call it.
template <int FZ_T>
__device__ void fz_scan(int* a, int n, int* wsum) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int per = (n + FZ_T - 1) / FZ_T;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int s = 0;
  for (int i = beg; i < end; ++i) s += a[i];
  int x = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    int v = lane < FZ_T / 32 ? wsum[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += y; }
    wsum[lane] = v;
  }
  __syncthreads();
  int base = (w > 0 ? wsum[w - 1] : 0) + x - s;
  for (int i = beg; i < end; ++i) { const int t = a[i]; a[i] = base; base += t; }
  if (tid == 0) a[n] = wsum[FZ_T / 32 - 1];
  __syncthreads();
}

// descending bitonic sort of npad (power of two >= 32) 64-bit keys in shared memory: one barrier per step.  Only for
// graphs whose keys exceed what the register sort below holds (npad > 4 * FZ_T).
template <int FZ_T>
__device__ void fz_sort_desc(unsigned long long* sk, int npad) {
  for (int kk = 2; kk <= npad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += FZ_T) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long x = sk[i], y = sk[l];
        const bool desc = (i & kk) == 0;
        if ((x < y) == desc) { sk[i] = y; sk[l] = x; }
      }
      __syncthreads();
    }
  }
}

// Top-k selection without a sorting network across the CTA (the 45 barriers of a 512-key bitonic sort were 26 % of the
// first version's stall samples): every warp sorts its own run of 32 * KPT keys in REGISTERS (element e = slot * 32 +
// lane; strides < 32 are shuffles, strides >= 32 register swaps: no barrier), runs go to shared memory, and every key's
// rank in the union is its position in its own run plus, for every other run, the number of keys greater than it
// (binary search; keys are distinct).  Two barriers in total.  Writes perml[rank] / inv for rank < k.
template <int FZ_T, int KPT>
__device__ void fz_topk_runs(const float* score, int n, int k, unsigned long long* runs, u16* perml, short* inv) {
  constexpr int NW = FZ_T / 32, RL = 32 * KPT;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long key[KPT];
#pragma unroll
  for (int s = 0; s < KPT; ++s) {
    const int i = w * RL + s * 32 + lane;                   // node id
    key[s] = i < n ? (((unsigned long long)fz_score_key(score[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i)) : 0ull;
  }
#pragma unroll
  for (int kk = 2; kk <= RL; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int js = j >> 5;
#pragma unroll
        for (int s = 0; s < KPT; ++s) {
          if ((s & js) == 0) {
            const int e = s * 32 + lane;
            const bool desc = (e & kk) == 0;
            const unsigned long long x = key[s], y = key[s | js];
            if ((x < y) == desc) { key[s] = y; key[s | js] = x; }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < KPT; ++s) {
          const int e = s * 32 + lane;
          const unsigned long long x = key[s];
          const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
          const bool lower = (lane & j) == 0;               // this lane holds the lower index of the pair
          const bool desc = (e & kk) == 0;
          // descending pair: lower index keeps the larger key
          const bool keep_max = lower == desc;
          key[s] = keep_max ? (x > y ? x : y) : (x < y ? x : y);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < KPT; ++s) runs[w * RL + s * 32 + lane] = key[s];
  __syncthreads();
#pragma unroll
  for (int s = 0; s < KPT; ++s) {
    const unsigned long long x = key[s];
    if (x == 0ull) continue;                                // padding
    int rank = s * 32 + lane;
    for (int r = 0; r < NW; ++r) {
      if (r == w) continue;
      const unsigned long long* run = runs + r * RL;
      if (r * RL >= n) break;                               // runs past the graph hold only padding
      int lo = 0, hi = RL;                                  // first position whose key is < x (descending run)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (run[mid] > x) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      const int v = (int)(0xFFFFFFFFu - (unsigned)(x & 0xFFFFFFFFull));
      perml[rank] = (u16)v; inv[v] = (short)rank;
    }
  }
  __syncthreads();
}

struct FzCsr { u16* rp; u16* col; float* dis; };

// Level-0 CSR of A_hat from the graph's coalesced edge list: row r = [cols of the edges (r, *) in list order | r]
// (self loop last, PyG add_remaining_self_loops order).  Symmetric list => this IS the dst-major CSR and its
// transpose.  Returns the OR of the violated TSG_FUSED_* bits (uniform across the CTA).
template <int FZ_T>
__device__ int fz_build_csr0(const int32_t* __restrict__ lrow, const int32_t* __restrict__ lcol, int64_t eb, int e, int n,
                             FzCsr c, int* ibuf, int* wsum, int* s_bad) {
  const int tid = threadIdx.x;
  if (tid == 0) *s_bad = 0;
  __syncthreads();
  // the list is sorted by row: row r's edges are the run of lrow == r, so ibuf[r] = first edge of row r falls out of
  // the run boundaries (rows without edges inherit the next boundary) -- no counting pass, no atomics, no scan
  int bad = 0;
  for (int p = tid; p <= e; p += FZ_T) {
    const int r = p < e ? lrow[eb + p] : n;
    const int r0 = p > 0 ? lrow[eb + p - 1] : -1;
    if (p < e) {
      const int q = lcol[eb + p];
      if ((unsigned)r >= (unsigned)n || (unsigned)q >= (unsigned)n || r == q) { bad |= TSG_FUSED_BAD_EDGE; continue; }
      if (p > 0 && !(r0 < r || (r0 == r && lcol[eb + p - 1] < q))) { bad |= TSG_FUSED_UNSORTED; continue; }
    }
    if ((unsigned)r0 >= (unsigned)n && r0 != -1) continue;           // previous edge was bad: flagged by its own thread
    for (int rr = r0 + 1; rr <= r; ++rr) ibuf[rr] = p;
  }
  if (bad) atomicOr(s_bad, bad);
  __syncthreads();
  if (*s_bad) return *s_bad;
  for (int r = tid; r <= n; r += FZ_T) c.rp[r] = (u16)(ibuf[r] + r);
  for (int r = tid; r < n; r += FZ_T) {
    const int deg = ibuf[r + 1] - ibuf[r];
    c.dis[r] = __fdiv_rn(1.0f, __fsqrt_rn((float)(deg + 1)));           // K1b: deg counts the self loop
    c.col[ibuf[r + 1] + r] = (u16)r;                                    // self loop closes the row
  }
  for (int p = tid; p < e; p += FZ_T) c.col[p + lrow[eb + p]] = (u16)lcol[eb + p];
  __syncthreads();
  // symmetry: (r, q) listed => (q, r) listed (binary search in row q, self loop excluded)
  bad = 0;
  for (int p = tid; p < e; p += FZ_T) {
    const int r = lrow[eb + p], q = lcol[eb + p];
    int lo = c.rp[q], hi = c.rp[q + 1] - 1;
    bool found = false;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = c.col[mid];
      if (v == r) { found = true; break; }
      if (v < r) lo = mid + 1; else hi = mid;
    }
    if (!found) bad |= TSG_FUSED_ASYMMETRIC;
  }
  if (bad) atomicOr(s_bad, bad);
  __syncthreads();
  return *s_bad;
}

// K1c in shared memory: pooled level's CSR = rows perm[i] of the current one restricted to surviving columns,
// relabelled by inv, same order, renormalised from the new degrees (self loop included).
template <int FZ_T>
__device__ void fz_filter_csr(FzCsr cur, FzCsr nxt, const u16* perml, const short* inv, int k, int* ibuf, int* wsum,
                              int first_thread) {
  const int T = FZ_T - first_thread;
  const int tid = (int)threadIdx.x - first_thread;
  if (tid >= 0)
    for (int i = tid; i < k; i += T) {
      const int v = perml[i];
      int cnt = 0;
      for (int p = cur.rp[v]; p < cur.rp[v + 1]; ++p) cnt += inv[cur.col[p]] >= 0;
      ibuf[i] = cnt;
    }
  __syncthreads();
  fz_scan<FZ_T>(ibuf, k, wsum);
  for (int i = threadIdx.x; i < k; i += FZ_T) {
    const int v = perml[i];
    int o = ibuf[i];
    nxt.rp[i] = (u16)o;
    nxt.dis[i] = __fdiv_rn(1.0f, __fsqrt_rn((float)(ibuf[i + 1] - o)));
    for (int p = cur.rp[v]; p < cur.rp[v + 1]; ++p) {
      const int j = inv[cur.col[p]];
      if (j >= 0) nxt.col[o++] = (u16)j;
    }
  }
  if (threadIdx.x == 0) nxt.rp[k] = (u16)ibuf[k];
  __syncthreads();
}

// Work distribution: a launch serves ONE size class (graphs with cls_nlo < n0 or cls_elo < e, and n0 <= nmax[0],
// e <= emax); its shared-memory layout is sized for that class, so small graphs run 3-4 CTAs per SM instead of
// paying for the 1,000-node graph in the batch.  Thread 0 draws tickets until it finds a graph of the class.
template <int FZ_T>
__device__ __forceinline__ int fz_next_graph(const FusedArgs& a, int* s_g) {
  if (threadIdx.x == 0) {
    int g;
    for (;;) {
      g = (int)atomicAdd(a.sched + a.cls, 1u);
      if (g >= a.G) { g = -1; break; }
      const int n = (int)(a.level_ptr[g + 1] - a.level_ptr[g]);
      const int e = (int)(a.edge_ptr[g + 1] - a.edge_ptr[g]);
      const bool smaller = n <= a.cls_nlo && e <= a.cls_elo;
      if (!smaller && n <= a.nmax[0] && e <= a.emax) break;
    }
    *s_g = g;
  }
  __syncthreads();
  return *s_g;
}

template <int H, int FZ_T, int MINB>
__global__ void __launch_bounds__(FZ_T, MINB)
k_sag_fused_fwd(const FusedArgs a) {
  constexpr int F4 = H / 4, LPR = F4, RPP = FZ_T / LPR, RPW = 32 / LPR;
  static_assert(LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "hidden / 4 must be a power of two <= 32");
  extern __shared__ __align__(16) unsigned char smraw[];
  const FzLayout lo = fz_layout(H, a.L, a.nmax[0], a.nmax[1], a.emax);
  FzCsr csr[2] = {{(u16*)(smraw + lo.rpA), (u16*)(smraw + lo.colA), (float*)(smraw + lo.disA)},
                  {(u16*)(smraw + lo.rpB), (u16*)(smraw + lo.colB), (float*)(smraw + lo.disB)}};
  int* lab = (int*)(smraw + lo.label);
  float* W23 = (float*)(smraw + lo.w23);
  float* vecs = (float*)(smraw + lo.vecs);            // [b1 b2 b3 | ws1 ws2 ws3 | bs1 bs2 bs3]
  float* sw = (float*)(smraw + lo.sw);
  float* score = (float*)(smraw + lo.score);
  unsigned long long* keys = (unsigned long long*)(smraw + lo.keys);
  u16* perml = (u16*)(smraw + lo.perml);
  short* inv = (short*)(smraw + lo.inv);
  int* ibuf = (int*)(smraw + lo.ibuf);
  float4* feat = (float4*)(smraw + lo.feat);
  float* outs = (float*)(smraw + lo.outs);
  int* misc = (int*)(smraw + lo.misc);                // [0..31] scan warp sums, [32] graph ticket, [33] bad flags
  int* wsum = misc; int* s_g = misc + 32; int* s_bad = misc + 33;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = tid % LPR;
  const int64_t G1 = (int64_t)a.G + 1;
  const float4* __restrict__ W1g = reinterpret_cast<const float4*>(a.params[0]);   // [L, H] table: L1-resident gathers

  // parameters -> shared memory, once per CTA
  for (int i = tid; i < H * H; i += FZ_T) { W23[i] = a.params[4][i]; W23[H * H + i] = a.params[8][i]; }
  for (int i = tid; i < H; i += FZ_T) {
    vecs[i] = a.params[1][i]; vecs[H + i] = a.params[5][i]; vecs[2 * H + i] = a.params[9][i];
    vecs[3 * H + i] = a.params[2][i]; vecs[4 * H + i] = a.params[6][i]; vecs[5 * H + i] = a.params[10][i];
  }
  if (tid < 3) vecs[6 * H + tid] = a.params[4 * tid + 3][0];

  for (;;) {
    const int g = fz_next_graph<FZ_T>(a, s_g);
    if (g < 0) break;
    const int64_t eb = a.edge_ptr[g];
    const int e = (int)(a.edge_ptr[g + 1] - eb);
    int64_t nb = a.level_ptr[g];
    int n = (int)(a.level_ptr[g + 1] - nb);
    for (int i = tid; i < n; i += FZ_T) lab[i] = a.label[nb + i];
    const int bad = fz_build_csr0<FZ_T>(a.lrow, a.lcol, eb, e, n, csr[0], ibuf, wsum, s_bad);
    if (bad) {
      if (tid == 0) atomicOr(a.status, bad);
      for (int i = tid; i < 2 * H; i += FZ_T) a.z[(int64_t)g * 2 * H + i] = 0.f;
      continue;
    }
    for (int lvl = 0; lvl < 3; ++lvl) {
      const FzCsr cur = csr[lvl & 1], nxt = csr[(lvl + 1) & 1];
      const int64_t kb = a.level_ptr[(lvl + 1) * G1 + g];
      const int k = (int)(a.level_ptr[(lvl + 1) * G1 + g + 1] - kb);
      float4* hg = reinterpret_cast<float4*>(a.h[lvl]) + nb * F4;
      // ---- conv: h = ReLU(A_hat (x W) + b), sw = h . ws        (network.py:34,38,42; layers.py:18's product)
      {
        const float4 b4 = reinterpret_cast<const float4*>(vecs + lvl * H)[l];
        const float4 w4 = reinterpret_cast<const float4*>(vecs + (3 + lvl) * H)[l];
        const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
        for (int r = tid / LPR; r < n; r += RPP) {
          const int s = cur.rp[r], t = cur.rp[r + 1];
          const float dr = cur.dis[r];
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int p = s; p < t; ++p) {
            const int c = cur.col[p];
            float v = __fmul_rn(__fmul_rn(cur.dis[c], 1.0f), dr);
            float4 hv;
            if (lvl == 0) {
              int lb = lab[c];
              if ((unsigned)lb >= (unsigned)a.L) { lb = 0; v = 0.f; }       // all-zero one-hot row
              hv = __ldg(W1g + lb * F4 + l);
            } else {
              hv = feat[c * F4 + l];
            }
            acc.x = fmaf(v, hv.x, acc.x); acc.y = fmaf(v, hv.y, acc.y);
            acc.z = fmaf(v, hv.z, acc.z); acc.w = fmaf(v, hv.w, acc.w);
          }
          acc.x = __fadd_rn(acc.x, b4.x); acc.y = __fadd_rn(acc.y, b4.y);
          acc.z = __fadd_rn(acc.z, b4.z); acc.w = __fadd_rn(acc.w, b4.w);
          acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
          hg[(int64_t)r * F4 + l] = acc;
          float dot = 0.f;
          dot = fmaf(acc.x, w4.x, dot); dot = fmaf(acc.y, w4.y, dot);
          dot = fmaf(acc.z, w4.z, dot); dot = fmaf(acc.w, w4.w, dot);
#pragma unroll
          for (int d = LPR / 2; d > 0; d >>= 1) dot += __shfl_xor_sync(gmask, dot, d, LPR);
          if (l == 0) sw[r] = dot;
        }
      }
      __syncthreads();
      // ---- score = A_hat sw + bs (separately rounded products: k_spmm_scalar)
      {
        const float bs = vecs[6 * H + lvl];
        for (int r = tid; r < n; r += FZ_T) {
          const float dr = cur.dis[r];
          float acc = 0.f;
          for (int p = cur.rp[r]; p < cur.rp[r + 1]; ++p) {
            const int c = cur.col[p];
            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__fmul_rn(cur.dis[c], 1.0f), dr), sw[c]));
          }
          acc = __fadd_rn(acc, bs);
          score[r] = acc;
          a.score[lvl][nb + r] = acc;
          inv[r] = -1;
        }
      }
      __syncthreads();
      // ---- top-k (layers.py:20): register-sorted runs + rank merge; huge graphs: shared-memory bitonic network
      if (n <= FZ_T) fz_topk_runs<FZ_T, 1>(score, n, k, keys, perml, inv);
      else if (n <= 2 * FZ_T) fz_topk_runs<FZ_T, 2>(score, n, k, keys, perml, inv);
      else if (n <= 4 * FZ_T) fz_topk_runs<FZ_T, 4>(score, n, k, keys, perml, inv);
      else {
        int npad = 32;
        while (npad < n) npad <<= 1;
        for (int r = tid; r < npad; r += FZ_T)
          keys[r] = r < n ? (((unsigned long long)fz_score_key(score[r]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r)) : 0ull;
        __syncthreads();
        fz_sort_desc<FZ_T>(keys, npad);
        for (int i = tid; i < k; i += FZ_T) {
          const int v = (int)(0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull));
          perml[i] = (u16)v; inv[v] = (short)i;
        }
        __syncthreads();
      }
      // ---- gate: x' = h[perm] * tanh(score[perm])  (layers.py:21) into shared memory (the x W buffer is dead now)
      for (int i = tid / LPR; i < k; i += RPP) {
        const int v = perml[i];
        if (l == 0) a.perm[lvl][kb + i] = nb + v;
        const float th = tanhf(score[v]);
        float4 x = hg[(int64_t)v * F4 + l];
        x.x *= th; x.y *= th; x.z *= th; x.w *= th;
        feat[i * F4 + l] = x;
      }
      __syncthreads();
      // ---- readout [gmp || gap] by warp 0 (k_gate_readout_fwd's summation order), pooled CSR by everyone else
      if (warp == 0) {
        const int sub = lane / LPR;
        float mx[4], sm[4]; int am[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) { mx[v] = -FLT_MAX; sm[v] = 0.f; am[v] = -1; }
        for (int i = sub; i < k; i += RPW) {
          const float4 w = feat[i * F4 + l];
          const float vals[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            sm[v] += vals[v];
            if (vals[v] > mx[v] || am[v] < 0) { mx[v] = vals[v]; am[v] = i; }
          }
        }
#pragma unroll
        for (int d = LPR; d < 32; d <<= 1) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float omx = __shfl_xor_sync(0xffffffffu, mx[v], d);
            const int oam = __shfl_xor_sync(0xffffffffu, am[v], d);
            const float osm = __shfl_xor_sync(0xffffffffu, sm[v], d);
            sm[v] = (sub & (d / LPR)) ? osm + sm[v] : sm[v] + osm;
            const bool take = oam >= 0 && (am[v] < 0 || omx > mx[v] || (omx == mx[v] && oam < am[v]));
            if (take) { mx[v] = omx; am[v] = oam; }
          }
        }
        if (sub == 0) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            outs[lvl * 2 * H + l * 4 + v] = am[v] >= 0 ? mx[v] : 0.f;
            outs[lvl * 2 * H + H + l * 4 + v] = sm[v] / (float)(k > 0 ? k : 1);
            a.argmax[lvl][(int64_t)g * H + l * 4 + v] = am[v] >= 0 ? (int)(kb + am[v]) : -1;
          }
        }
      }
      if (lvl < 2) {
        fz_filter_csr<FZ_T>(cur, nxt, perml, inv, k, ibuf, wsum, 32);          // layers.py:23 (filter_adj) as K1c
        // ---- next level's x W in place (row-local product, k-ascending FMA chain: k_linear_fwd_dense)
        const float4* Wn = reinterpret_cast<const float4*>(W23 + lvl * H * H);
        const int sub = lane / LPR;
        for (int ib = warp * RPW; ib < k; ib += (FZ_T / 32) * RPW) {
          const int i = ib + sub;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < k) {
            const float4* xr = feat + i * F4;
#pragma unroll 2
            for (int k4 = 0; k4 < F4; ++k4) {
              const float4 x = xr[k4];
              const float4 w0 = Wn[(k4 * 4 + 0) * F4 + l], w1 = Wn[(k4 * 4 + 1) * F4 + l];
              const float4 w2 = Wn[(k4 * 4 + 2) * F4 + l], w3 = Wn[(k4 * 4 + 3) * F4 + l];
              acc.x = fmaf(x.x, w0.x, acc.x); acc.y = fmaf(x.x, w0.y, acc.y); acc.z = fmaf(x.x, w0.z, acc.z); acc.w = fmaf(x.x, w0.w, acc.w);
              acc.x = fmaf(x.y, w1.x, acc.x); acc.y = fmaf(x.y, w1.y, acc.y); acc.z = fmaf(x.y, w1.z, acc.z); acc.w = fmaf(x.y, w1.w, acc.w);
              acc.x = fmaf(x.z, w2.x, acc.x); acc.y = fmaf(x.z, w2.y, acc.y); acc.z = fmaf(x.z, w2.z, acc.z); acc.w = fmaf(x.z, w2.w, acc.w);
              acc.x = fmaf(x.w, w3.x, acc.x); acc.y = fmaf(x.w, w3.y, acc.y); acc.z = fmaf(x.w, w3.z, acc.z); acc.w = fmaf(x.w, w3.w, acc.w);
            }
          }
          __syncwarp();
          if (i < k) feat[i * F4 + l] = acc;
        }
      }
      __syncthreads();
      nb = kb; n = k;
    }
    // z = (out_0 + out_1) + out_2                                             (network.py:46)
    for (int i = tid; i < 2 * H; i += FZ_T)
      a.z[(int64_t)g * 2 * H + i] = __fadd_rn(__fadd_rn(outs[i], outs[2 * H + i]), outs[4 * H + i]);
  }
}

}  // namespace tsg

using namespace tsg;

namespace tsg {

static int fz_max_smem() {
  static int v = -1;
  if (v < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) v = 0;
  }
  return v;
}

int fused_num_ctas() { return TSG_NUM_SMS; }

bool fused_supported(int H, int L, int nmax0, int nmax1, int emax) {
  if (H != 32 && H != 64 && H != 128) return false;
  if (L <= 0 || nmax0 <= 0 || nmax0 > 32767 || nmax1 > nmax0 || emax < 0 || emax + nmax0 >= 65536) return false;
  const FzLayout lo = fz_layout(H, L, nmax0, nmax1, emax);
  return (int)lo.total <= fz_max_smem();
}

// Size classes of one batch (host side, from the batch maxima and its average degree): class caps are chosen so that the
// class's shared-memory layout allows `ctas` CTAs per SM; the last class is the batch maximum itself.
struct FzClass { int ncap, kcap, ecap, ctas, threads; };

static int fz_classes(const FusedArgs& a, FzClass* out) {
  const double davg = a.avg_degree > 1.0 ? a.avg_degree : 1.0;
  auto kof = [&](int n) { int k = (int)ceilf((float)a.ratio * (float)n); if (k > a.nmax[1]) k = a.nmax[1]; if (k < 1) k = 1; return k; };
  const int budget[2] = {fz_max_smem() / 4 - 1024, fz_max_smem() / 2 - 1024};
  const int start[2] = {384, 768};
  int nc = 0, prev = 0;
  for (int c = 0; c < 2; ++c) {
    int ncap = start[c];
    while (ncap > 32) {
      const int ecap = (int)(ncap * davg * 1.5) + 16;
      if ((int)fz_layout(a.H, a.L, ncap, kof(ncap), ecap).total <= budget[c]) break;
      ncap = ncap * 7 / 8;
    }
    if (ncap <= prev + 32) continue;                   // class would be (almost) empty
    if (ncap >= a.nmax[0]) break;                      // the batch maximum already fits this budget: final class below
    out[nc++] = {ncap, kof(ncap), (int)(ncap * davg * 1.5) + 16, c == 0 ? 4 : 2, 256};
    prev = ncap;
  }
  const int tot = (int)fz_layout(a.H, a.L, a.nmax[0], a.nmax[1], a.emax).total;
  const int ctas = tot <= budget[0] ? 4 : (tot <= budget[1] ? 2 : 1);
  out[nc++] = {a.nmax[0], a.nmax[1], a.emax, ctas, ctas == 1 ? 512 : 256};
  return nc;
}

template <int H, int T, int MINB>
static int launch_fwd_cls(FusedArgs a, const FzClass& c, int cls, int nlo, int elo, cudaStream_t st) {
  a.nmax[0] = c.ncap; a.nmax[1] = c.kcap; a.emax = c.ecap; a.cls = cls; a.cls_nlo = nlo; a.cls_elo = elo;
  const FzLayout lo = fz_layout(H, a.L, c.ncap, c.kcap, c.ecap);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_sag_fused_fwd<H, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, fz_max_smem()) != cudaSuccess)
      return check_launch("sag_fused_fwd(attr)");
    configured = true;
  }
  int grid = fused_num_ctas() * c.ctas;
  if (grid > a.G) grid = a.G;
  k_sag_fused_fwd<H, T, MINB><<<grid, T, lo.total, st>>>(a);
  return check_launch("sag_fused_fwd");
}

template <int H>
static int launch_fwd_t(const FusedArgs& a, cudaStream_t st) {
  FzClass cls[3];
  const int nc = fz_classes(a, cls);
  cudaMemsetAsync(a.sched, 0, 8 * sizeof(unsigned), st);
  int nlo = 0, elo = -1;
  for (int c = 0; c < nc; ++c) {
    int rc;
    if (cls[c].threads == 512) rc = launch_fwd_cls<H, 512, 1>(a, cls[c], c, nlo, elo, st);
    else if (cls[c].ctas >= 4) rc = launch_fwd_cls<H, 256, 4>(a, cls[c], c, nlo, elo, st);
    else rc = launch_fwd_cls<H, 256, 2>(a, cls[c], c, nlo, elo, st);
    if (rc != TSG_OK) return rc;
    nlo = cls[c].ncap; elo = cls[c].ecap;
  }
  return TSG_OK;
}

int launch_sag_fused_fwd(const FusedArgs& a, cudaStream_t st) {
  switch (a.H) {
    case 32: return launch_fwd_t<32>(a, st);
    case 64: return launch_fwd_t<64>(a, st);
    case 128: return launch_fwd_t<128>(a, st);
  }
  set_error("sag_fused_fwd: unsupported hidden width %d", a.H);
  return TSG_EINVAL;
}

}  // namespace tsg
