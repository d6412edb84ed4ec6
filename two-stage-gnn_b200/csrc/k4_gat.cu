// K4 -- dense-GAT head on the packed CSR: edge score, segmented edge-softmax, weighted aggregation.
//
// Replaces DGATHead.forward (Code/sage+gat+diffpool/encoders_GAT.py:29-49), which materialises an
// [N*N, 2F] tensor (244 MiB per head at N = 1000) to evaluate
//     e_ij = LeakyReLU(a[:F].h_i + a[F:].h_j)   masked to adj[i,j] > 0,
//     att  = softmax(e, dim=1)   -- over the ROW index i for every column j (adj [1,N,N] broadcasts),
//     h'_i = sum_j att_ij h_j.
// On the CSR only the edges exist: per column j the softmax runs over its in-list (src-major CSR),
// per row i the aggregation runs over its out-list (dst-major CSR).  Masked entries contribute exactly
// 0 upstream (exp(-9e15 - max) == 0), so the sparse form is exact for every column that has an edge;
// columns without any edge (isolated / padded nodes: uniform 1/N upstream) are handled by the host
// layer as a per-graph correction vector (tsg/gat.py).
//
// Forward : k_gat_colstats (per column: max and sum of exp)  +  k_gat_aggregate (per row gather).
// Backward: k_gat_bwd_col (per column: d_alpha, softmax backward, dh_j, ds2_j, dz per edge in COO
//           order)  +  k_gat_bwd_row (per row: ds1_i).  Sequential per segment: deterministic, no atomics.
// HBM-bound gathers; scores are 4-byte scattered reads that stay in L2 (the whole graph block is ~100 KB).
#include "common.cuh"
#include <stdlib.h>
#include <float.h>

namespace tsg {

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : slope * x; }

// one thread per (column j, head): online max / sum over the column's entries (src-major CSR)
__global__ void __launch_bounds__(256)
k_gat_colstats(const int* __restrict__ t_rowptr, const int* __restrict__ t_colidx,
               const float* __restrict__ s1, const float* __restrict__ s2, int n, int heads, float slope,
               float* __restrict__ mx, float* __restrict__ zs) {
  const int64_t total = (int64_t)n * heads;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx / heads), hd = (int)(idx - (int64_t)j * heads);
    const int p0 = t_rowptr[j], p1 = t_rowptr[j + 1];
    const float sj = s2[idx];
    float m = -FLT_MAX;
    for (int p = p0; p < p1; ++p) m = fmaxf(m, lrelu(s1[(int64_t)t_colidx[p] * heads + hd] + sj, slope));
    float z = 0.f;
    for (int p = p0; p < p1; ++p) z += expf(lrelu(s1[(int64_t)t_colidx[p] * heads + hd] + sj, slope) - m);
    mx[idx] = m; zs[idx] = z;
  }
}

// LPR lanes per row i, each lane owns columns f = l, l+LPR, ... of the heads*F wide feature row.
template <int LPR>
__global__ void __launch_bounds__(256)
k_gat_aggregate(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ h,
                const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ mx,
                const float* __restrict__ zs, int n, int heads, int F, float slope, float* __restrict__ hp) {
  const int l = threadIdx.x % LPR;
  const int W = heads * F;
  const int rpc = (n + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * rpc, r_end = min(n, r_begin + rpc);
  for (int i = r_begin + threadIdx.x / LPR; i < r_end; i += 256 / LPR) {
    const int p0 = rowptr[i], p1 = rowptr[i + 1];
    for (int f = l; f < W; f += LPR) {
      const int hd = f / F;
      const float si = s1[(int64_t)i * heads + hd];
      float acc = 0.f;
      for (int p = p0; p < p1; ++p) {
        const int j = colidx[p];
        const int64_t jh = (int64_t)j * heads + hd;
        const float a = expf(lrelu(si + s2[jh], slope) - mx[jh]) / zs[jh];
        acc = fmaf(a, h[(int64_t)j * W + f], acc);
      }
      hp[(int64_t)i * W + f] = acc;
    }
  }
}

// float4 variant of the aggregation for F % 4 == 0: LPR = heads*F/4 lanes per row (rounded up to a power of two),
// one 128-bit gather per neighbour per lane, 32 / LPR rows per warp (the scalar kernel above used 32 lanes per
// 64-wide row and recomputed alpha in every lane for every column).
template <int LPR>
__global__ void __launch_bounds__(256)
k_gat_aggregate_v4(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float4* __restrict__ h,
                   const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ mx,
                   const float* __restrict__ zs, int n, int heads, int F4, float slope, float4* __restrict__ hp) {
  const int l = threadIdx.x % LPR;
  const int W4 = heads * F4;
  const int rpc = (n + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * rpc, r_end = min(n, r_begin + rpc);
  for (int i = r_begin + threadIdx.x / LPR; i < r_end; i += 256 / LPR) {
    const int p0 = rowptr[i], p1 = rowptr[i + 1];
    for (int f = l; f < W4; f += LPR) {
      const int hd = f / F4;
      const float si = s1[(int64_t)i * heads + hd];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int p = p0;
      for (; p + 2 <= p1; p += 2) {
        const int j0 = colidx[p], j1 = colidx[p + 1];
        const int64_t a0 = (int64_t)j0 * heads + hd, a1 = (int64_t)j1 * heads + hd;
        const float e0 = lrelu(si + s2[a0], slope) - mx[a0], e1 = lrelu(si + s2[a1], slope) - mx[a1];
        const float4 h0 = h[(int64_t)j0 * W4 + f], h1 = h[(int64_t)j1 * W4 + f];
        const float w0 = expf(e0) / zs[a0], w1 = expf(e1) / zs[a1];
        acc.x = fmaf(w0, h0.x, acc.x); acc.y = fmaf(w0, h0.y, acc.y); acc.z = fmaf(w0, h0.z, acc.z); acc.w = fmaf(w0, h0.w, acc.w);
        acc.x = fmaf(w1, h1.x, acc.x); acc.y = fmaf(w1, h1.y, acc.y); acc.z = fmaf(w1, h1.z, acc.z); acc.w = fmaf(w1, h1.w, acc.w);
      }
      if (p < p1) {
        const int j0 = colidx[p];
        const int64_t a0 = (int64_t)j0 * heads + hd;
        const float w0 = expf(lrelu(si + s2[a0], slope) - mx[a0]) / zs[a0];
        const float4 h0 = h[(int64_t)j0 * W4 + f];
        acc.x = fmaf(w0, h0.x, acc.x); acc.y = fmaf(w0, h0.y, acc.y); acc.z = fmaf(w0, h0.z, acc.z); acc.w = fmaf(w0, h0.w, acc.w);
      }
      hp[(int64_t)i * W4 + f] = acc;
    }
  }
}

// warp per column j.  Two sweeps over the column's entries; d_alpha kept per (slot, head) in `dal`.
__global__ void __launch_bounds__(256)
k_gat_bwd_col(const int* __restrict__ t_rowptr, const int* __restrict__ t_colidx, const int* __restrict__ t_eid,
              const float* __restrict__ h, const float* __restrict__ s1, const float* __restrict__ s2,
              const float* __restrict__ mx, const float* __restrict__ zs, const float* __restrict__ dhp,
              int n, int heads, int F, float slope, float* __restrict__ dal, float* __restrict__ dz_coo,
              float* __restrict__ dh, float* __restrict__ ds2) {
  const int lane = threadIdx.x & 31;
  const int W = heads * F;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < n; j += nwarps) {
    const int p0 = t_rowptr[j], p1 = t_rowptr[j + 1];
    for (int hd = 0; hd < heads; ++hd) {
      const int64_t jh = j * heads + hd;
      const float sj = s2[jh], m = mx[jh], z = zs[jh];
      // sweep 1: d_alpha_ij = dhp_i[hd] . h_j[hd] ;  t = sum_i alpha_ij d_alpha_ij
      float t = 0.f;
      for (int p = p0; p < p1; ++p) {
        const int i = t_colidx[p];
        float part = 0.f;
        for (int f = lane; f < F; f += 32) part += dhp[(int64_t)i * W + hd * F + f] * h[j * W + hd * F + f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const float a = expf(lrelu(s1[(int64_t)i * heads + hd] + sj, slope) - m) / z;
        t += a * part;
        if (lane == 0) dal[(int64_t)p * heads + hd] = part;
      }
      __syncwarp();
      // sweep 2: dz_ij, ds2_j, and the transposed aggregation dh_j += alpha_ij dhp_i
      float dsum = 0.f;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};          // F <= 128: up to 4 columns per lane
      for (int p = p0; p < p1; ++p) {
        const int i = t_colidx[p];
        const float pre = s1[(int64_t)i * heads + hd] + sj;
        const float a = expf(lrelu(pre, slope) - m) / z;
        const float de = a * (dal[(int64_t)p * heads + hd] - t);
        const float dzv = de * (pre > 0.f ? 1.f : slope);
        dsum += dzv;
        if (lane == 0) dz_coo[(int64_t)t_eid[p] * heads + hd] = dzv;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int f = lane + 32 * q;
          if (f < F) acc[q] = fmaf(a, dhp[(int64_t)i * W + hd * F + f], acc[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = lane + 32 * q;
        if (f < F) dh[j * W + hd * F + f] = acc[q];
      }
      if (lane == 0) ds2[jh] = dsum;
    }
  }
}

// v2 of the column kernel for F % 4 == 0: warp per (column j, head); the 32 lanes are 4 entry slots x 8 lanes,
// each 8-lane group owns one column entry per step and reads its dhp row as float4 (128-bit gathers, 4
// independent entries in flight instead of one entry at a time with a 5-step shuffle reduction per entry).
// Two sweeps like v1 (the second re-reads the just-gathered rows from L1/L2); d_alpha is recomputed instead of
// round-tripping through a global scratch array.  Fixed reduction order: deterministic.
template <int GB_MAXF4>                             // float4 per lane: F <= 32 * GB_MAXF4
__global__ void __launch_bounds__(256)
k_gat_bwd_col_v4(const int* __restrict__ t_rowptr, const int* __restrict__ t_colidx, const int* __restrict__ t_eid,
                 const float* __restrict__ h, const float* __restrict__ s1, const float* __restrict__ s2,
                 const float* __restrict__ mx, const float* __restrict__ zs, const float* __restrict__ dhp,
                 int n, int heads, int F, float slope, float* __restrict__ dz_coo,
                 float* __restrict__ dh, float* __restrict__ ds2) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, l = lane & 7;
  const int W4 = heads * F / 4, F4 = F / 4;
  const int64_t items = (int64_t)n * heads;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4* h4 = reinterpret_cast<const float4*>(h);
  const float4* d4p = reinterpret_cast<const float4*>(dhp);
  for (int64_t it = warp; it < items; it += nwarps) {
    const int j = (int)(it / heads), hd = (int)(it - (int64_t)j * heads);
    const int p0 = t_rowptr[j], p1 = t_rowptr[j + 1];
    const float sj = s2[it], m = mx[it], z = zs[it];
    float4 hj[GB_MAXF4];
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q)
      hj[q] = (l + 8 * q < F4) ? h4[(int64_t)j * W4 + hd * F4 + l + 8 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
    // sweep 1: t = sum_i alpha_ij (dhp_i . h_j)
    float t = 0.f;
    for (int pb = p0; pb < p1; pb += 4) {
      const int p = pb + grp;
      const bool valid = p < p1;
      const int i = valid ? t_colidx[p] : j;
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < GB_MAXF4; ++q) {
        if (l + 8 * q < F4) {
          const float4 d = d4p[(int64_t)i * W4 + hd * F4 + l + 8 * q];
          part += d.x * hj[q].x + d.y * hj[q].y + d.z * hj[q].z + d.w * hj[q].w;
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      const float a = expf(lrelu(s1[(int64_t)i * heads + hd] + sj, slope) - m) / z;
      if (valid) t += a * part;
    }
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    // sweep 2: dz_ij, ds2_j and dh_j = sum_i alpha_ij dhp_i
    float dsum = 0.f;
    float4 acc[GB_MAXF4];
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pb = p0; pb < p1; pb += 4) {
      const int p = pb + grp;
      const bool valid = p < p1;
      const int i = valid ? t_colidx[p] : j;
      float4 d[GB_MAXF4];
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < GB_MAXF4; ++q) {
        if (l + 8 * q < F4) {
          d[q] = d4p[(int64_t)i * W4 + hd * F4 + l + 8 * q];
          part += d[q].x * hj[q].x + d[q].y * hj[q].y + d[q].z * hj[q].z + d[q].w * hj[q].w;
        } else {
          d[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      const float pre = s1[(int64_t)i * heads + hd] + sj;
      const float a = expf(lrelu(pre, slope) - m) / z;
      if (valid) {
        const float dzv = a * (part - t) * (pre > 0.f ? 1.f : slope);
        dsum += dzv;
        if (l == 0) dz_coo[(int64_t)t_eid[p] * heads + hd] = dzv;
#pragma unroll
        for (int q = 0; q < GB_MAXF4; ++q) {
          acc[q].x = fmaf(a, d[q].x, acc[q].x); acc[q].y = fmaf(a, d[q].y, acc[q].y);
          acc[q].z = fmaf(a, d[q].z, acc[q].z); acc[q].w = fmaf(a, d[q].w, acc[q].w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
      }
      if (grp == 0 && l + 8 * q < F4)
        reinterpret_cast<float4*>(dh)[(int64_t)j * W4 + hd * F4 + l + 8 * q] = acc[q];
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 8);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 16);
    if (lane == 0) ds2[it] = dsum;
  }
}


// v5 (round 2): ONE gathering sweep instead of two.  ncu on the config-3 shape (profiles/r02_kernels_ncu.md): v4 is
// issue bound (76 % issue active, 7 % of the HBM peak, 1.14 ms = 39 % of the GAT step) and half of its instructions are
// the second sweep re-gathering the dhp rows.  Only dz_ij = alpha_ij (dhp_i . h_j - t) lrelu'(pre_ij) needs the column
// total t; dh_j = sum_i alpha_ij dhp_i does not.  So the sweep accumulates dh_j and t together and parks (alpha_ij *
// lrelu', dhp_i . h_j) per entry in a per-warp shared-memory strip; a second, gather-free pass (one LANE per entry)
// finishes dz and ds2.  Columns longer than the strip (GB5_CAP entries) take v4's two-sweep body.
constexpr int GB5_CAP = 192;
template <int GB_MAXF4>
__global__ void __launch_bounds__(256)
k_gat_bwd_col_v5(const int* __restrict__ t_rowptr, const int* __restrict__ t_colidx, const int* __restrict__ t_eid,
                 const float* __restrict__ h, const float* __restrict__ s1, const float* __restrict__ s2,
                 const float* __restrict__ mx, const float* __restrict__ zs, const float* __restrict__ dhp,
                 int n, int heads, int F, float slope, float* __restrict__ dz_coo,
                 float* __restrict__ dh, float* __restrict__ ds2) {
  __shared__ float s_asf[8][GB5_CAP], s_part[8][GB5_CAP];
  const int lane = threadIdx.x & 31, grp = lane >> 3, l = lane & 7, wib = threadIdx.x >> 5;
  const int W4 = heads * F / 4, F4 = F / 4;
  const int64_t items = (int64_t)n * heads;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4* h4 = reinterpret_cast<const float4*>(h);
  const float4* d4p = reinterpret_cast<const float4*>(dhp);
  for (int64_t it = warp; it < items; it += nwarps) {
    const int j = (int)(it / heads), hd = (int)(it - (int64_t)j * heads);
    const int p0 = t_rowptr[j], p1 = t_rowptr[j + 1];
    const int deg = p1 - p0;
    const float sj = s2[it], m = mx[it], z = zs[it];
    const bool strip = deg <= GB5_CAP;
    float4 hj[GB_MAXF4];
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q)
      hj[q] = (l + 8 * q < F4) ? h4[(int64_t)j * W4 + hd * F4 + l + 8 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
    float t = 0.f;
    float4 acc[GB_MAXF4];
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the gathering sweep: t = sum_i alpha_ij (dhp_i . h_j), dh_j = sum_i alpha_ij dhp_i
    for (int pb = p0; pb < p1; pb += 4) {
      const int p = pb + grp;
      const bool valid = p < p1;
      const int i = valid ? t_colidx[p] : j;
      float4 d[GB_MAXF4];
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < GB_MAXF4; ++q) {
        if (l + 8 * q < F4) {
          d[q] = d4p[(int64_t)i * W4 + hd * F4 + l + 8 * q];
          part += d[q].x * hj[q].x + d[q].y * hj[q].y + d[q].z * hj[q].z + d[q].w * hj[q].w;
        } else {
          d[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      const float pre = s1[(int64_t)i * heads + hd] + sj;
      const float a = expf(lrelu(pre, slope) - m) / z;
      if (valid) {
        t += a * part;
        if (strip && l == 0) { s_asf[wib][p - p0] = a * (pre > 0.f ? 1.f : slope); s_part[wib][p - p0] = part; }
#pragma unroll
        for (int q = 0; q < GB_MAXF4; ++q) {
          acc[q].x = fmaf(a, d[q].x, acc[q].x); acc[q].y = fmaf(a, d[q].y, acc[q].y);
          acc[q].z = fmaf(a, d[q].z, acc[q].z); acc[q].w = fmaf(a, d[q].w, acc[q].w);
        }
      }
    }
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
#pragma unroll
    for (int q = 0; q < GB_MAXF4; ++q) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
      }
      if (grp == 0 && l + 8 * q < F4)
        reinterpret_cast<float4*>(dh)[(int64_t)j * W4 + hd * F4 + l + 8 * q] = acc[q];
    }
    float dsum = 0.f;
    if (strip) {                                     // gather-free finish: one lane per entry
      __syncwarp();
      for (int e = lane; e < deg; e += 32) {
        const float dzv = s_asf[wib][e] * (s_part[wib][e] - t);
        dsum += dzv;
        dz_coo[(int64_t)t_eid[p0 + e] * heads + hd] = dzv;
      }
      __syncwarp();                                  // the strip is reused by this warp's next item
#pragma unroll
      for (int o = 1; o <= 16; o <<= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    } else {                                         // long column: second gathering sweep (v4)
      for (int pb = p0; pb < p1; pb += 4) {
        const int p = pb + grp;
        const bool valid = p < p1;
        const int i = valid ? t_colidx[p] : j;
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < GB_MAXF4; ++q) {
          if (l + 8 * q < F4) {
            const float4 d = d4p[(int64_t)i * W4 + hd * F4 + l + 8 * q];
            part += d.x * hj[q].x + d.y * hj[q].y + d.z * hj[q].z + d.w * hj[q].w;
          }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        const float pre = s1[(int64_t)i * heads + hd] + sj;
        const float a = expf(lrelu(pre, slope) - m) / z;
        if (valid) {
          const float dzv = a * (part - t) * (pre > 0.f ? 1.f : slope);
          dsum += dzv;
          if (l == 0) dz_coo[(int64_t)t_eid[p] * heads + hd] = dzv;
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 8);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 16);
    }
    if (lane == 0) ds2[it] = dsum;
  }
}

__global__ void __launch_bounds__(256)
k_gat_bwd_row(const int* __restrict__ rowptr, const int* __restrict__ eid, const float* __restrict__ dz_coo,
              int n, int heads, float* __restrict__ ds1) {
  const int64_t total = (int64_t)n * heads;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / heads), hd = (int)(idx - (int64_t)i * heads);
    float s = 0.f;
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) s += dz_coo[(int64_t)eid[p] * heads + hd];
    ds1[idx] = s;
  }
}

}  // namespace tsg

using namespace tsg;

extern "C" int tsg_gat_fwd(const int32_t* rowptr, const int32_t* colidx, const int32_t* t_rowptr,
                           const int32_t* t_colidx, const float* h, const float* s1, const float* s2,
                           int64_t n, int64_t heads, int64_t F, float slope,
                           float* mx, float* zs, float* hp, void* stream) {
  TSG_REQUIRE(n >= 0 && heads > 0 && F > 0, "gat_fwd: bad shape");
  TSG_REQUIRE(n * heads * F < (int64_t)0x7fffffff, "gat_fwd: too large");
  if (n == 0) return TSG_OK;
  TSG_REQUIRE(rowptr && colidx && t_rowptr && t_colidx && h && s1 && s2 && mx && zs && hp, "gat_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  k_gat_colstats<<<grid_for(n * heads, 256), 256, 0, st>>>(t_rowptr, t_colidx, s1, s2, (int)n, (int)heads, slope, mx, zs);
  int W = (int)(heads * F);
  if (F % 4 == 0 && ((((uintptr_t)h) | ((uintptr_t)hp)) & 15) == 0) {
    int W4 = W / 4;
    int l4 = 1; while (l4 < W4 && l4 < 32) l4 <<= 1;
    int grid4 = grid_for(n, 256 / l4, 32);
#define TSG_GO4(L) k_gat_aggregate_v4<L><<<grid4, 256, 0, st>>>(rowptr, colidx, (const float4*)h, s1, s2, mx, zs, (int)n, (int)heads, (int)(F / 4), slope, (float4*)hp)
    switch (l4) {
      case 1: TSG_GO4(1); break; case 2: TSG_GO4(2); break; case 4: TSG_GO4(4); break;
      case 8: TSG_GO4(8); break; case 16: TSG_GO4(16); break; default: TSG_GO4(32); break;
    }
#undef TSG_GO4
    return check_launch("gat_fwd");
  }
  int lpr = 1; while (lpr < W && lpr < 32) lpr <<= 1;
  int grid = grid_for(n, 256 / lpr, 32);
#define TSG_GO(L) k_gat_aggregate<L><<<grid, 256, 0, st>>>(rowptr, colidx, h, s1, s2, mx, zs, (int)n, (int)heads, (int)F, slope, hp)
  switch (lpr) {
    case 1: TSG_GO(1); break; case 2: TSG_GO(2); break; case 4: TSG_GO(4); break;
    case 8: TSG_GO(8); break; case 16: TSG_GO(16); break; default: TSG_GO(32); break;
  }
#undef TSG_GO
  return check_launch("gat_fwd");
}

extern "C" size_t tsg_gat_bwd_workspace_bytes(int64_t nnz, int64_t heads) {
  return 2 * ws_bytes((size_t)(nnz + 1) * (size_t)heads, 4) + 512;
}

extern "C" int tsg_gat_bwd(const int32_t* rowptr, const int32_t* eid, const int32_t* t_rowptr,
                           const int32_t* t_colidx, const int32_t* t_eid, const float* h, const float* s1,
                           const float* s2, const float* mx, const float* zs, const float* dhp,
                           int64_t n, int64_t nnz, int64_t heads, int64_t F, float slope,
                           float* dh, float* ds1, float* ds2,
                           void* workspace, size_t workspace_bytes, void* stream) {
  TSG_REQUIRE(n >= 0 && heads > 0 && F > 0 && nnz >= 0, "gat_bwd: bad shape");
  TSG_REQUIRE(F <= 128, "gat_bwd: per-head width %lld > 128 not supported", (long long)F);
  if (n == 0) return TSG_OK;
  TSG_REQUIRE(rowptr && eid && t_rowptr && t_colidx && t_eid && h && s1 && s2 && mx && zs && dhp && dh && ds1 && ds2,
              "gat_bwd: null pointer");
  if (workspace_bytes < tsg_gat_bwd_workspace_bytes(nnz, heads)) { set_error("gat_bwd: workspace too small"); return TSG_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  float* dal = ws.take<float>((nnz + 1) * heads);
  float* dz = ws.take<float>((nnz + 1) * heads);
  if (F % 4 == 0 && ((((uintptr_t)h) | ((uintptr_t)dhp) | ((uintptr_t)dh)) & 15) == 0)
  {
    static const bool v4 = getenv("TSG_GAT_BWD_V4") != nullptr;        // A/B: the two-sweep kernel of round 1
#define TSG_GB(Q) do { if (v4) k_gat_bwd_col_v4<Q><<<grid_for(n * heads, 8), 256, 0, st>>>(t_rowptr, t_colidx, t_eid, h, s1, s2, mx, zs, dhp, (int)n, (int)heads, (int)F, slope, dz, dh, ds2); \
                      else k_gat_bwd_col_v5<Q><<<grid_for(n * heads, 8), 256, 0, st>>>(t_rowptr, t_colidx, t_eid, h, s1, s2, mx, zs, dhp, (int)n, (int)heads, (int)F, slope, dz, dh, ds2); } while (0)
    if (F <= 32) TSG_GB(1); else if (F <= 64) TSG_GB(2); else TSG_GB(4);
#undef TSG_GB
  }
  else
    k_gat_bwd_col<<<grid_for(n, 8), 256, 0, st>>>(t_rowptr, t_colidx, t_eid, h, s1, s2, mx, zs, dhp, (int)n, (int)heads,
                                                  (int)F, slope, dal, dz, dh, ds2);
  k_gat_bwd_row<<<grid_for(n * heads, 256), 256, 0, st>>>(rowptr, eid, dz, (int)n, (int)heads, ds1);
  return check_launch("gat_bwd");
}
