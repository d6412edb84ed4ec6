// k7_tc.cuh -- tcgen05 implementation of the DiffPool segment contraction
//     C_g [Kx <= 128, Ky <= 256] = sum_{r in graph g} X[r,:]^T Y[r,:]
// One CTA (4 warps) per graph; the contraction index (graph rows) is the MMA K dimension.
//   * operands: staged by the CTA into shared memory in the K-major, no-swizzle canonical layout
//     (8-row x 16-byte core matrices; LBO = K-direction core stride, SBO = M/N-direction 8-row group
//     stride), TRANSPOSED on the fly (X and Y are row-major with the contraction index as the row)
//     and split hi/lo for the error-compensated 3xTF32 product  hi*hi + hi*lo + lo*hi  that keeps
//     the 1e-5 fp32 parity bar (plain TF32 is ~1e-3);
//   * MMA: tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = Ky rounded up to 16, K = 8 per
//     instruction, issued by one thread, fp32 accumulators in TMEM (N columns x 128 lanes);
//   * completion: tcgen05.commit -> mbarrier; epilogue tcgen05.ld 32x32b (thread = accumulator row).
// SASS evidence: UTCHMMA / UTCCP-free path with LDTM in the epilogue (see profiles/).
#pragma once
#include "common.cuh"
#include <stdlib.h>

namespace tsg {

constexpr int TC_KC = 32;                 // contraction rows per staged chunk (4 MMA k-steps of 8)
constexpr int TC_M = 128;
constexpr int TC_A_BYTES = TC_M * TC_KC * 4;          // 16 KB per (hi | lo)
constexpr int TC_SPIN_LIMIT = 1 << 22;   // hang guard: a completed MMA batch arrives within microseconds

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// The hi / lo split of the 3xTF32 product in four instructions per element instead of nine (cvt.rna.tf32.f32 compiles to
// add + inf/NaN test + select + mask, twice, plus the subtraction -- and this split is what bounds the tcgen05 kernels here):
//   hi = round-to-nearest-away TF32 of x   = (bits(x) + 0x1000) & ~0x1fff      (same value as cvt.rna for finite x)
//   lo = TF32 TRUNCATION of (x - hi)       = bits(x - hi) & ~0x1fff            (x - hi is exact)
// Used by the row-local products (sums of <= 256 terms; measured 2-4e-6 against float64), not by the contraction.
// Error of hi*hi' + hi*lo' + lo*hi' against x*x': the dropped lo*lo' (<= 2^-22 |x x'|) plus the truncation of the two lo
// terms (<= 2^-10 |lo| <= 2^-21 |x|, one-sided): ~1.2e-6 relative per product against ~0.7e-6 with a rounded lo.
// NaN: hi may lose an all-ones-mantissa NaN (the add carries out), lo = NaN - hi keeps it, so the product is still NaN.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
// K-major SWIZZLE_NONE descriptor: addr(row, k) = (row/8)*SBO + (k/4)*LBO + (row%8)*16 + (k%4)*4
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  return d;                               // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// kind::tf32, fp32 accumulate, A and B K-major, dense
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                           // c_format = F32
  d |= 2u << 7;                           // a_format = TF32
  d |= 2u << 10;                          // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;          // n_dim
  d |= (uint32_t)(M >> 4) << 24;          // m_dim
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < TC_SPIN_LIMIT; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  // hang guard tripped: carrying on would read TMEM / overwrite shared memory under an unfinished MMA and return
  // silently corrupt results, so the kernel dies loudly instead (the launch error reaches the caller at its next
  // CUDA call; `err` is still set by kernels that get that far)
  __trap();
  return false;
}

// stage one operand chunk: rows = `width` feature columns (<= NR), k = TC_KC contraction rows.
// src row-major [rows_total, width]; element (n, k) = src[(r0 + k) * width + n].
template <int NR>
__device__ __forceinline__ void stage_operand(const float* __restrict__ src, int64_t r0, int rows, int width,
                                              char* hi, char* lo) {
  constexpr uint32_t SBO = (TC_KC / 4) * 128;
  for (int n = threadIdx.x; n < NR; n += 128) {
#pragma unroll
    for (int kq = 0; kq < TC_KC / 4; ++kq) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = kq * 4 + i;
        v[i] = (n < width && k < rows) ? __ldg(src + (r0 + k) * width + n) : 0.f;
      }
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h[i] = to_tf32(v[i]);
        l[i] = to_tf32(v[i] - __uint_as_float(h[i]));
      }
      uint32_t off = (uint32_t)(n >> 3) * SBO + (uint32_t)kq * 128 + (uint32_t)(n & 7) * 16;
      *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

template <int NPAD>     // NPAD: 64, 128 or 256 = TMEM columns and staged B rows
__global__ void __launch_bounds__(128)
k_seg_contract_tc(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ gptr,
                  int Kx, int Ky, int Nmma, float* __restrict__ C, int* __restrict__ err) {
  extern __shared__ __align__(128) char tc_smem[];
  char* a_hi = tc_smem;
  char* a_lo = a_hi + TC_A_BYTES;
  char* b_hi = a_lo + TC_A_BYTES;
  char* b_lo = b_hi + NPAD * TC_KC * 4;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int64_t lo_r = gptr[g], hi_r = gptr[g + 1];

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)NPAD) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc(TC_M, Nmma);
  constexpr uint32_t SBO = (TC_KC / 4) * 128, LBO = 128;
  uint32_t phase = 0;
  bool any = false, ok = true;

  for (int64_t r0 = lo_r; r0 < hi_r; r0 += TC_KC) {
    const int rows = (int)min((int64_t)TC_KC, hi_r - r0);
    stage_operand<TC_M>(X, r0, rows, Kx, a_hi, a_lo);
    stage_operand<NPAD>(Y, r0, rows, Ky, b_hi, b_lo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < TC_KC / 8; ++ks) {
        const uint32_t koff = ks * 2 * 128;                            // two 16-byte K cores per MMA
        uint64_t ah = make_desc(smem_u32(a_hi) + koff, LBO, SBO), al = make_desc(smem_u32(a_lo) + koff, LBO, SBO);
        uint64_t bh = make_desc(smem_u32(b_hi) + koff, LBO, SBO), bl = make_desc(smem_u32(b_lo) + koff, LBO, SBO);
        mma_tf32(tmem, al, bh, idesc, (any || ks > 0) ? 1u : 0u);      // small terms first
        mma_tf32(tmem, ah, bl, idesc, 1u);
        mma_tf32(tmem, ah, bh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    any = true;
    ok = mbar_wait(smem_u32(&mbar), phase) && ok;                       // smem may be overwritten after this
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && threadIdx.x == 0) atomicExch(err, 1);

  // epilogue: thread = accumulator row m (cluster index), 32 columns per tcgen05.ld
  const int m = warp * 32 + lane;
  float* Cg = C + (int64_t)g * Kx * Ky;
  for (int c0 = 0; c0 < Nmma; c0 += 32) {
    uint32_t v[32];
    if (any) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    if (m < Kx) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < Ky) Cg[(int64_t)m * Ky + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)NPAD) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// v2: TMA-fed, pipelined.  v1 above serialises  stage (scalar global loads) -> MMA -> wait  per 32-row
// chunk (ncu: tensor pipe 12.7 %, issue 19.5 %: latency bound).  Here
//   * the raw fp32 rows of a chunk (KC2 contraction rows of X and of Y; contiguous in global memory
//     because both operands are row-major with the contraction index as the row) are fetched by the TMA
//     engine: one thread arms an mbarrier with the byte count and issues two `cp.async.bulk` copies;
//     three raw stages are in flight, so HBM latency never sits on the critical path;
//   * the 128 threads transpose + hi/lo-split a landed raw stage into the K-major MMA tiles (the 3xTF32
//     split has to pass through registers, so this is the only SIMT work left) -- the MMA tiles are
//     double buffered and guarded by their own mbarriers (tcgen05.commit), so the tensor pipe works on
//     chunk i while the threads transform chunk i+1 and the TMA loads chunks i+2..i+4.
// Needs Kx % 4 == 0 and Ky % 4 == 0 (16-byte TMA granularity); other widths use v1.
// ------------------------------------------------------------------------------------------
constexpr int KC2 = 16;                   // contraction rows per chunk (2 MMA k-steps of 8)
constexpr int RAW_STAGES = 3;     // 3 raw stages keep the NPAD <= 128 kernel under 104 KB: two CTAs per SM
constexpr int TC2_SPIN_LIMIT = 1 << 22;   // hang guard (a wait normally returns within microseconds)

__device__ __forceinline__ bool mbar_wait2(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < TC2_SPIN_LIMIT; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  // hang guard tripped: carrying on would read TMEM / overwrite shared memory under an unfinished MMA and return
  // silently corrupt results, so the kernel dies loudly instead (the launch error reaches the caller at its next
  // CUDA call; `err` is still set by kernels that get that far)
  __trap();
  return false;
}

__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// raw [k][width] fp32 (as landed by the TMA) -> K-major no-swizzle hi / lo tiles; k >= rows zero-filled.
// 512 threads = 4 groups of contraction rows (kq = tid / 128) x 128 feature columns (n = tid % 128, + 128
// for the wide operand): no index division, consecutive shared-memory words per warp, and a branch-free
// fast path for full chunks.  Columns beyond the operand width never reach a stored output.
constexpr int TC2_THREADS = 512;
static_assert(KC2 == 16, "transform_operand maps tid / 128 to the 4 row groups of a 16-row chunk");

template <bool FULL>
__device__ __forceinline__ void transform_item(const float* __restrict__ raw, int rows, int width, int n, int kq,
                                               char* hi, char* lo) {
  constexpr uint32_t SBO = (KC2 / 4) * 128;
  const float* src = raw + (kq * 4) * width + n;
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v = (FULL || kq * 4 + i < rows) ? src[i * width] : 0.f;
    h[i] = to_tf32(v);                       // both parts rounded here: the contraction sums up to 1,000 rows and already sits at
    l[i] = to_tf32(v - __uint_as_float(h[i]));   // 8e-6 of the 1e-5 gate there (TMEM accumulation): no room for split_tf32's one-sided lo
  }
  const uint32_t off = (uint32_t)(n >> 3) * SBO + (uint32_t)kq * 128 + (uint32_t)(n & 7) * 16;
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void transform_operand(const float* __restrict__ raw, int rows, int width,
                                                  char* hi, char* lo) {
  const int kq = threadIdx.x >> 7, n0 = threadIdx.x & 127;
  if (rows == KC2) {
    for (int n = n0; n < width; n += 128) transform_item<true>(raw, rows, width, n, kq, hi, lo);
  } else {
    for (int n = n0; n < width; n += 128) transform_item<false>(raw, rows, width, n, kq, hi, lo);
  }
}

template <int NPAD>
__global__ void __launch_bounds__(TC2_THREADS + 32)
k_seg_contract_tc2(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ gptr,
                   int Kx, int Ky, int Nmma, float* __restrict__ C, int* __restrict__ err) {
  extern __shared__ __align__(128) char tc_smem[];
  constexpr int A_BYTES = TC_M * KC2 * 4, B_BYTES = NPAD * KC2 * 4;
  constexpr int MMA_SET = 2 * A_BYTES + 2 * B_BYTES;                 // a_hi a_lo b_hi b_lo
  constexpr int TWARPS = TC2_THREADS / 32;                           // transform warps
  char* mma_buf = tc_smem;                                           // 2 sets
  char* raw_buf = tc_smem + 2 * MMA_SET;                             // RAW_STAGES x raw_stride
  const int raw_stride = (KC2 * (Kx + Ky) * 4 + 127) & ~127;
  __shared__ __align__(8) uint64_t bar_raw[RAW_STAGES];              // TMA landed (tx bytes)
  __shared__ __align__(8) uint64_t bar_full[2];                      // MMA tile set written (TWARPS arrivals)
  __shared__ __align__(8) uint64_t bar_mma[2];                       // MMAs that read the set are done (commit)
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int64_t lo_r = gptr[g], hi_r = gptr[g + 1];
  const int nch = (int)((hi_r - lo_r + KC2 - 1) / KC2);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)NPAD) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < RAW_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar_raw[i])), "r"(1u) : "memory");
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar_full[i])), "r"((uint32_t)TWARPS) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar_mma[i])), "r"(1u) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;

  if (warp == TWARPS) {
    // ---------------- issuer warp: one lane feeds the TMA engine and the tensor pipe
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TC_M, Nmma);
      constexpr uint32_t SBO = (KC2 / 4) * 128, LBO = 128;
      auto issue_tma = [&](int i) {
        const int64_t r0 = lo_r + (int64_t)i * KC2;
        const int rows = (int)min((int64_t)KC2, hi_r - r0);
        const int s = i % RAW_STAGES;
        const uint32_t bar = smem_u32(&bar_raw[s]);
        const uint32_t bx = (uint32_t)rows * Kx * 4, by = (uint32_t)rows * Ky * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bx + by) : "memory");
        const uint32_t dst = smem_u32(raw_buf + (size_t)s * raw_stride);
        tma_load_1d(dst, X + r0 * Kx, bx, bar);
        tma_load_1d(dst + (uint32_t)KC2 * Kx * 4, Y + r0 * Ky, by, bar);
      };
      for (int i = 0; i < RAW_STAGES && i < nch; ++i) issue_tma(i);
      for (int i = 0; i < nch && ok; ++i) {
        const int t = i & 1;
        ok = mbar_wait2(smem_u32(&bar_full[t]), (uint32_t)((i >> 1) & 1));       // set t written, raw stage consumed
        if (i + RAW_STAGES < nch) issue_tma(i + RAW_STAGES);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(mma_buf + (size_t)t * MMA_SET), a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < KC2 / 8; ++ks) {
          const uint32_t koff = ks * 2 * 128;
          const uint64_t ah = make_desc(a_hi + koff, LBO, SBO), al = make_desc(a_lo + koff, LBO, SBO);
          const uint64_t bh = make_desc(b_hi + koff, LBO, SBO), bl = make_desc(b_lo + koff, LBO, SBO);
          mma_tf32(tmem, al, bh, idesc, (i > 0 || ks > 0) ? 1u : 0u);  // small terms first
          mma_tf32(tmem, ah, bl, idesc, 1u);
          mma_tf32(tmem, ah, bh, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma[t])) : "memory");
      }
    }
  } else {
    // ---------------- transform warps
    for (int i = 0; i < nch && ok; ++i) {
      const int s = i % RAW_STAGES, t = i & 1;
      const int rows = (int)min((int64_t)KC2, hi_r - (lo_r + (int64_t)i * KC2));
      ok = mbar_wait2(smem_u32(&bar_raw[s]), (uint32_t)((i / RAW_STAGES) & 1));              // raw chunk landed
      if (i >= 2) ok = mbar_wait2(smem_u32(&bar_mma[t]), (uint32_t)(((i >> 1) - 1) & 1)) && ok;  // MMA(i-2) left set t
      char* set = mma_buf + (size_t)t * MMA_SET;
      const float* raw = reinterpret_cast<const float*>(raw_buf + (size_t)s * raw_stride);
      transform_operand(raw, rows, Kx, set, set + A_BYTES);
      transform_operand(raw + KC2 * Kx, rows, Ky, set + 2 * A_BYTES, set + 2 * A_BYTES + B_BYTES);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> async proxy
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar_full[t])) : "memory");
    }
  }
  if (nch > 0 && ok) {                                                 // the last commit covers every earlier MMA
    const int last = nch - 1;
    ok = mbar_wait2(smem_u32(&bar_mma[last & 1]), (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) atomicExch(err, 1);                          // any role's hang guard tripped

  // epilogue: 16 warps; warp w reads TMEM lanes 32*(w%4).. (its hardware lane quarter) and the column slabs
  // c0 = 32*(w/4), +128, ...; thread = accumulator row m
  const int m = (warp & 3) * 32 + lane;
  const bool epi = warp < TC2_THREADS / 32;              // the issuer warp has no TMEM lane quarter of its own here
  float* Cg = C + (int64_t)g * Kx * Ky;
  for (int c0 = (warp >> 2) * 32; c0 < Nmma; c0 += 32 * (TC2_THREADS / 128)) {
    uint32_t v[32];
    if (!epi) break;
    if (nch > 0) {
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    if (m < Kx) {
      if (c0 + 32 <= Ky) {                                   // full 32-column slab: 128-bit stores (Ky % 4 == 0)
        float4* dst = reinterpret_cast<float4*>(Cg + (int64_t)m * Ky + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < Ky) Cg[(int64_t)m * Ky + c0 + i] = __uint_as_float(v[i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)NPAD) : "memory");
  }
}

static int launch_seg_contract_tc(const float* X, const float* Y, const int64_t* gptr, int G, int Kx, int Ky,
                                  float* C, int* err, cudaStream_t st) {
  if (Kx > 128 || Ky > 256) { set_error("seg_contract(tcgen05): needs Kx <= 128 and Ky <= 256 (got %d, %d)", Kx, Ky); return TSG_EINVAL; }
  int Nmma = (Ky + 15) / 16 * 16;
  if (Nmma < 16) Nmma = 16;
  int npad = Nmma <= 64 ? 64 : (Nmma <= 128 ? 128 : 256);
  static const bool force_v1 = getenv("TSG_K7_V1") != nullptr;
  if (!force_v1 && Kx % 4 == 0 && Ky % 4 == 0 && (((uintptr_t)X | (uintptr_t)Y | (uintptr_t)C) & 15) == 0) {
    const size_t raw_stride = ((size_t)KC2 * (Kx + Ky) * 4 + 127) & ~(size_t)127;
    const size_t smem2 = 2 * (2 * (size_t)TC_M * KC2 * 4 + 2 * (size_t)npad * KC2 * 4) + RAW_STAGES * raw_stride + 128;
#define TSG_GO2(NP)                                                                                        \
    cudaFuncSetAttribute(k_seg_contract_tc2<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);   \
    k_seg_contract_tc2<NP><<<G, TC2_THREADS + 32, smem2, st>>>(X, Y, gptr, Kx, Ky, Nmma, C, err)
    if (npad == 64) { TSG_GO2(64); } else if (npad == 128) { TSG_GO2(128); } else { TSG_GO2(256); }
#undef TSG_GO2
    return check_launch("seg_contract(tcgen05 v2)");
  }
  size_t smem = 2 * TC_A_BYTES + 2 * (size_t)npad * TC_KC * 4 + 128;
#define TSG_GO(NP)                                                                                        \
  cudaFuncSetAttribute(k_seg_contract_tc<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
  k_seg_contract_tc<NP><<<G, 128, smem, st>>>(X, Y, gptr, Kx, Ky, Nmma, C, err)
  if (npad == 64) { TSG_GO(64); } else if (npad == 128) { TSG_GO(128); } else { TSG_GO(256); }
#undef TSG_GO
  return check_launch("seg_contract(tcgen05)");
}

// ------------------------------------------------------------------------------------------
// Row-local product on tcgen05 (round 2):  Y[r, :] = X[r, :] . W_g   (W_g [Kin, M])   or   X[r, :] . W_g^T   (W_g [M, Kin])
// for r in graph g -- the backward of the contraction above (dS = [Z | AS] dC^T, d[Z | AS] = S dC: 970 k rows x
// 100 / 196 columns at config-4 size, 1.6 ms each on the SIMT kernel = 21 % of the DiffPool step).
//   * MMA: D [128 rows, N = M <= 256] += A [128 rows, K] . B [N, K]^T, the contraction index is the FEATURE index Kin,
//     walked in chunks of SL_KC = 16 (2 k-steps of 8); a CTA takes a graph and walks its rows 128 at a time; two CTAs per SM
//     (<= 96 KB of shared memory, <= 256 TMEM columns and 56 registers each) so one CTA's epilogue and barrier waits hide
//     behind the other's transform.
//   * operands go global -> registers -> hi/lo TF32 tiles in the canonical K-major no-swizzle layout (the 3xTF32 split has
//     to pass through registers anyway); A rows are K-major in memory (no transpose), B is K-major when W is stored
//     transposed and MN-major otherwise (transposed on the fly, as in k_seg_contract_tc2).  Lane mapping: 8 rows x 4
//     k-quads per warp, so a warp's 16-byte shared-memory stores cover all 32 banks per quarter and its global reads are
//     64-byte row pieces.
//   * pipeline: two MMA tile sets; the 16 transform warps fill set (i & 1) while the tensor pipe works on the other;
//     the issuer lane commits each chunk to an mbarrier; the same 16 warps run the epilogue (tcgen05.ld, thread =
//     row) at the end of every 128-row tile.
// Needs Kin % 4 == 0, M % 4 == 0, M <= 256, 16-byte aligned operands.
// ------------------------------------------------------------------------------------------
constexpr int SL_KC = 16;
constexpr int SL_THREADS = 512;

// One chunk of one operand held in registers between its global loads and its shared-memory stores, so that ALL of a
// chunk's loads are in flight together and the next chunk's loads overlap the barrier wait of this one (the first cut
// loaded, converted and stored item by item: four exposed load latencies per chunk, ~4,400 cycles).
//   K-major source (A rows; W stored transposed): element (row, k) = src[row * ld + k];
//   MN-major source (W as [Kin, M]):              element (n, k)   = src[k * ld + n].
// Item j of a tile with `trows` rows (a multiple of 8): K-major -> lane = 8 rows x 4 k-quads (64-byte row pieces from
// global, all 32 banks per quarter-warp on the 16-byte stores); MN-major -> consecutive n per lane.
template <int ITEMS>
struct SlChunk {
  float4 v[ITEMS];
};
// (row, k-quad) of every item of a thread, packed row | kq << 16, -1 = no item: chunk- and tile-invariant, so the
// divisions by the runtime tile height happen once per thread, not once per item per chunk (ncu, first cut: 25 % of
// the kernel's instructions were this index arithmetic)
template <int ITEMS>
struct SlMap {
  int rk[ITEMS];
};

template <int ITEMS, bool KMAJOR>
__device__ __forceinline__ void sl_map_init(SlMap<ITEMS>& m, int trows) {
  const int ngroups = trows >> 3;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int j = threadIdx.x + it * SL_THREADS;
    int row, kq;
    if (KMAJOR) { const int u = j >> 5; row = (u % ngroups) * 8 + (j & 7); kq = (u / ngroups) * 4 + ((j >> 3) & 3); }
    else { row = j % trows; kq = j / trows; }
    m.rk[it] = j < trows * (SL_KC / 4) ? (row | (kq << 16)) : -1;
  }
}

template <int ITEMS>
__device__ __forceinline__ void sl_load_kmajor(SlChunk<ITEMS>& c, const SlMap<ITEMS>& m, const float* __restrict__ src,
                                               int64_t ld, int nrows, int klen) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int row = m.rk[it] & 0xffff, kq = m.rk[it] >> 16;
    c.v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m.rk[it] >= 0 && row < nrows && kq * 4 < klen)
      c.v[it] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)row * ld + kq * 4));
  }
}

template <int ITEMS>
__device__ __forceinline__ void sl_load_mnmajor(SlChunk<ITEMS>& c, const SlMap<ITEMS>& m, const float* __restrict__ src,
                                                int64_t ld, int ncols, int klen) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int n = m.rk[it] & 0xffff, kq = m.rk[it] >> 16;
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = kq * 4 + i;
      f[i] = (m.rk[it] >= 0 && n < ncols && k < klen) ? __ldg(src + (int64_t)k * ld + n) : 0.f;
    }
    c.v[it] = make_float4(f[0], f[1], f[2], f[3]);
  }
}

template <int ITEMS>
__device__ __forceinline__ void sl_store(const SlChunk<ITEMS>& c, const SlMap<ITEMS>& m, char* hi, char* lo) {
  constexpr uint32_t SBO = (SL_KC / 4) * 128;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    if (m.rk[it] < 0) continue;
    const int row = m.rk[it] & 0xffff, kq = m.rk[it] >> 16;
    const float f[4] = {c.v[it].x, c.v[it].y, c.v[it].z, c.v[it].w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(f[i], h[i], l[i]);
    const uint32_t off = (uint32_t)(row >> 3) * SBO + (uint32_t)kq * 128 + (uint32_t)(row & 7) * 16;
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <int NPAD>
__global__ void __launch_bounds__(SL_THREADS + 32, 2)
k_seg_linear_tc(const float* __restrict__ X, const float* __restrict__ W, const int64_t* __restrict__ gptr, int64_t flat_rows,
                int Kin, int M, int Nmma, int w_transposed, float* __restrict__ Y, int* __restrict__ err) {
  extern __shared__ __align__(128) char sltc_smem[];
  constexpr int A_BYTES = TC_M * SL_KC * 4, B_BYTES = NPAD * SL_KC * 4;
  constexpr int MMA_SET = 2 * A_BYTES + 2 * B_BYTES;                 // a_hi a_lo b_hi b_lo
  constexpr int TWARPS = SL_THREADS / 32;
  __shared__ __align__(8) uint64_t bar_full[2];                      // tile set written (TWARPS arrivals)
  __shared__ __align__(8) uint64_t bar_mma[2];                       // MMAs that read the set are done (commit)
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // gptr == nullptr: ONE shared weight, CTA b takes rows [128 b, 128 b + 128) of flat_rows (tsg_linear_tc)
  const int g = blockIdx.x;
  const int64_t lo_r = gptr ? gptr[g] : (int64_t)g * TC_M;
  const int64_t hi_r = gptr ? gptr[g + 1] : min(flat_rows, lo_r + TC_M);
  const int nk = (Kin + SL_KC - 1) / SL_KC;
  const float* Wg = gptr ? W + (int64_t)g * Kin * M : W;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)NPAD) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar_full[i])), "r"((uint32_t)TWARPS) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar_mma[i])), "r"(1u) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;
  int ci = 0;                                                        // chunks issued so far (every role counts alike)
  constexpr int A_ITEMS_ = TC_M * (SL_KC / 4) / SL_THREADS, B_ITEMS_ = (NPAD * (SL_KC / 4) + SL_THREADS - 1) / SL_THREADS;
  SlMap<A_ITEMS_> map_a; SlMap<B_ITEMS_> map_b;
  if (warp < TWARPS) {
    sl_map_init<A_ITEMS_, true>(map_a, TC_M);
    if (w_transposed) sl_map_init<B_ITEMS_, true>(map_b, Nmma); else sl_map_init<B_ITEMS_, false>(map_b, Nmma);
  }

  SlChunk<A_ITEMS_> ca0, ca1; SlChunk<B_ITEMS_> cb0, cb1;
  auto load_chunk = [&](int64_t r0, int rows, int c, SlChunk<A_ITEMS_>& ca, SlChunk<B_ITEMS_>& cb) {
    const int k0 = c * SL_KC, klen = min(SL_KC, Kin - k0);
    sl_load_kmajor<A_ITEMS_>(ca, map_a, X + r0 * Kin + k0, Kin, rows, klen);
    if (w_transposed) sl_load_kmajor<B_ITEMS_>(cb, map_b, Wg + k0, Kin, M, klen);
    else sl_load_mnmajor<B_ITEMS_>(cb, map_b, Wg + (int64_t)k0 * M, M, M, klen);
  };

  for (int64_t r0 = lo_r; r0 < hi_r; r0 += TC_M) {
    const int rows = (int)min((int64_t)TC_M, hi_r - r0);
    if (warp == TWARPS) {
      // ---------------- issuer lane
      if (lane == 0) {
        const uint32_t idesc = make_idesc(TC_M, Nmma);
        constexpr uint32_t SBO = (SL_KC / 4) * 128, LBO = 128;
        for (int c = 0; c < nk && ok; ++c) {
          const int i = ci + c, t = i & 1;
          ok = mbar_wait2(smem_u32(&bar_full[t]), (uint32_t)((i >> 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(sltc_smem + (size_t)t * MMA_SET), a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int ks = 0; ks < SL_KC / 8; ++ks) {
            const uint32_t koff = ks * 2 * 128;
            const uint64_t ah = make_desc(a_hi + koff, LBO, SBO), al = make_desc(a_lo + koff, LBO, SBO);
            const uint64_t bh = make_desc(b_hi + koff, LBO, SBO), bl = make_desc(b_lo + koff, LBO, SBO);
            mma_tf32(tmem, al, bh, idesc, (c > 0 || ks > 0) ? 1u : 0u);  // small terms first
            mma_tf32(tmem, ah, bl, idesc, 1u);
            mma_tf32(tmem, ah, bh, idesc, 1u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma[t])) : "memory");
        }
      }
    } else {
      // ---------------- transform warps: chunks c + 1 and c + 2 are in flight (registers) while chunk c's tile set is
      // awaited and written
      auto put_chunk = [&](int c, SlChunk<A_ITEMS_>& ca, SlChunk<B_ITEMS_>& cb) {
        const int i = ci + c, t = i & 1;
        if (i >= 2) ok = mbar_wait2(smem_u32(&bar_mma[t]), (uint32_t)(((i >> 1) - 1) & 1));     // MMA(i-2) left set t
        char* set = sltc_smem + (size_t)t * MMA_SET;
        sl_store<A_ITEMS_>(ca, map_a, set, set + A_BYTES);
        sl_store<B_ITEMS_>(cb, map_b, set + 2 * A_BYTES, set + 2 * A_BYTES + B_BYTES);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> async proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar_full[t])) : "memory");
        if (c + 2 < nk) load_chunk(r0, rows, c + 2, ca, cb);
      };
      load_chunk(r0, rows, 0, ca0, cb0);
      if (nk > 1) load_chunk(r0, rows, 1, ca1, cb1);
      for (int c = 0; c < nk && ok; c += 2) {
        put_chunk(c, ca0, cb0);
        if (c + 1 < nk && ok) put_chunk(c + 1, ca1, cb1);
      }
    }
    ci += nk;
    if (ok) {                                                          // the last commit covers every earlier MMA of the tile
      const int last = ci - 1;
      ok = mbar_wait2(smem_u32(&bar_mma[last & 1]), (uint32_t)((last >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads its TMEM lane quarter 32 * (w % 4) and the column slabs 32 * (w / 4), + 128; thread = row
    if (warp < TWARPS) {
      const int m = (warp & 3) * 32 + lane;
      for (int c0 = (warp >> 2) * 32; c0 < Nmma; c0 += 32 * (SL_THREADS / 128)) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < rows) {
          float* dst = Y + (r0 + m) * M + c0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (c0 + 4 * i < M)                                          // M % 4 == 0: whole float4s
              reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                     // TMEM is overwritten by the next tile's first MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (!ok) atomicExch(err, 1);
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)NPAD) : "memory");
  }
}

// gptr == nullptr: flat mode over flat_rows rows with one shared W (G is ignored)
static int launch_seg_linear_tc(const float* X, const float* W, const int64_t* gptr, int G, int64_t flat_rows, int Kin, int M,
                                int w_transposed, float* Y, int* err, cudaStream_t st) {
  if (M > 256 || M % 4 || Kin % 4 || ((((uintptr_t)X) | ((uintptr_t)W) | ((uintptr_t)Y)) & 15)) {
    set_error("seg_linear(tcgen05): needs M <= 256, M %% 4 == 0, Kin %% 4 == 0 and 16-byte aligned operands (got Kin %d, M %d)", Kin, M);
    return TSG_EINVAL;
  }
  int Nmma = (M + 15) / 16 * 16;
  const int npad = Nmma <= 64 ? 64 : (Nmma <= 128 ? 128 : 256);
  const size_t smem = 2 * (2 * (size_t)TC_M * SL_KC * 4 + 2 * (size_t)npad * SL_KC * 4) + 128;
  const int grid = gptr ? G : (int)((flat_rows + TC_M - 1) / TC_M);
#define TSG_GOSL(NP)                                                                                     \
  cudaFuncSetAttribute(k_seg_linear_tc<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
  k_seg_linear_tc<NP><<<grid, SL_THREADS + 32, smem, st>>>(X, W, gptr, flat_rows, Kin, M, Nmma, w_transposed, Y, err)
  if (npad == 64) { TSG_GOSL(64); } else if (npad == 128) { TSG_GOSL(128); } else { TSG_GOSL(256); }
#undef TSG_GOSL
  return check_launch("seg_linear(tcgen05)");
}

}  // namespace tsg
