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

namespace tsg {

constexpr int TC_KC = 32;                 // contraction rows per staged chunk (4 MMA k-steps of 8)
constexpr int TC_M = 128;
constexpr int TC_A_BYTES = TC_M * TC_KC * 4;          // 16 KB per (hi | lo)
constexpr int TC_SPIN_LIMIT = 1 << 14;   // hang guard: a completed MMA batch arrives within microseconds

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
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

static int launch_seg_contract_tc(const float* X, const float* Y, const int64_t* gptr, int G, int Kx, int Ky,
                                  float* C, int* err, cudaStream_t st) {
  if (Kx > 128 || Ky > 256) { set_error("seg_contract(tcgen05): needs Kx <= 128 and Ky <= 256 (got %d, %d)", Kx, Ky); return TSG_EINVAL; }
  int Nmma = (Ky + 15) / 16 * 16;
  if (Nmma < 16) Nmma = 16;
  int npad = Nmma <= 64 ? 64 : (Nmma <= 128 ? 128 : 256);
  size_t smem = 2 * TC_A_BYTES + 2 * (size_t)npad * TC_KC * 4 + 128;
#define TSG_GO(NP)                                                                                        \
  cudaFuncSetAttribute(k_seg_contract_tc<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
  k_seg_contract_tc<NP><<<G, 128, smem, st>>>(X, Y, gptr, Kx, Ky, Nmma, C, err)
  if (npad == 64) { TSG_GO(64); } else if (npad == 128) { TSG_GO(128); } else { TSG_GO(256); }
#undef TSG_GO
  return check_launch("seg_contract(tcgen05)");
}

}  // namespace tsg
