// Tensor-core GEMM for the chained products of K3 (EFB projection) and K5 (matrix-normal draw):
//   C(m,n) = alpha * sum_k A(m,k) * B(k,n) [+ beta * C]     fp32 operands read as TF32, fp32 accumulation in TMEM
// with the same fused epilogues as the CUDA-core kernel (store / accumulate the square / add the posterior mean
// and split into weight and bias).  Operands are row-major fp32 matrices addressed through element strides, so each
// of A and B is either contraction-contiguous ("K-major": TMA boxes of [128 rows][32 k], SWIZZLE_128B) or
// row-contiguous ("MN-major": four boxes of [32 k][32 rows], SWIZZLE_128B_ATOM_32B -- the only MN-major layout
// tcgen05 accepts for 32-bit operands, see syrk_tc.cu); the instruction descriptor's major bits select the form per
// operand, no transpose is ever materialised.
//   CTA = one 128 x 128 tile of C; 6 warps: 0 = TMA, 1 = MMA issuer + TMEM owner, 2-5 = TMA + epilogue;
//   ring of 6 stages x 32 KB (A 16 KB | B 16 KB, 32 contraction indices), 4 x tcgen05.mma (M=128, N=128, K=8) per stage.
// Operand precision: the tensor core truncates fp32 words to TF32.  Callers that need round-to-nearest behaviour
// pass operands already rounded (crv_round_tf32; the EFB eigenbases and the inverse factors are rounded once), and
// `round_out` rounds an intermediate product in the epilogue so that the second GEMM of a chain sees rounded input.
#include "common.cuh"
#include "../../include/curvature_b200.h"
#include <cuda.h>

namespace crv {
namespace {

constexpr int GM = 128, GN = 128, GK = 32;
constexpr int G_STAGE = 32 * 1024, G_NSTAGE = 6, G_THREADS = 6 * 32, G_NPROD = 5;
constexpr int G_SMEM = G_NSTAGE * G_STAGE + 1024 + 1024;
constexpr uint32_t G_SPIN = 1u << 22;

struct GtParams {
  int m, n, k, ldc;
  float alpha, beta;
  int epi, round_out, a_mn, b_mn;
  float* C;
  SampleEpilogue se;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > G_SPIN) asm volatile("trap;");
  } while (!ok);
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// K-major SWIZZLE_128B: 128-byte rows (32 fp32 of the contraction axis), 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SW128_32B: 128-byte rows (32 fp32 along M/N), 4-row atoms, 512 B between the two atoms of a K = 8 step,
// `lbo` bytes between 32-row chunks along M/N
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ float rna_tf32(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const GtParams p, const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (s32(raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + G_NSTAGE * G_STAGE;          // full[6] | empty[6] | tmem_full
  const uint32_t bar_done = bars + 8 * (2 * G_NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw + (sbase - s32(raw)) + G_NSTAGE * G_STAGE + 8 * (2 * G_NSTAGE + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int nk = (p.k + GK - 1) / GK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_NSTAGE; ++s) {
      bar_init(bars + 8 * s, G_NPROD);
      bar_init(bars + 8 * (G_NSTAGE + s), 1);
    }
    bar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ta)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tb)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp != 1) {
    if (lane == 0) {
      const int me = warp == 0 ? 0 : warp - 1;
      const int na = p.a_mn ? 4 : 1, nb = p.b_mn ? 4 : 1, total = na + nb;
      uint32_t mine = 0;
      for (int e = me; e < total; e += G_NPROD) mine += (e < na ? (p.a_mn ? 4096u : 16384u) : (p.b_mn ? 4096u : 16384u));
      for (int it = 0; it < nk; ++it) {
        const int s = it % G_NSTAGE;
        const uint32_t ph = (uint32_t)(it / G_NSTAGE) & 1u;
        bar_wait(bars + 8 * (G_NSTAGE + s), ph ^ 1u);
        if (mine) bar_expect(bars + 8 * s, mine); else bar_arrive(bars + 8 * s);
        const uint32_t st = sbase + (uint32_t)s * G_STAGE;
        const int k0 = it * GK;
        for (int e = me; e < total; e += G_NPROD) {
          if (e < na) {
            if (p.a_mn) tma_2d(st + (uint32_t)e * 4096u, &ta, m0 + 32 * e, k0, bars + 8 * s);
            else tma_2d(st, &ta, k0, m0, bars + 8 * s);
          } else {
            const int q = e - na;
            if (p.b_mn) tma_2d(st + 16384u + (uint32_t)q * 4096u, &tb, n0 + 32 * q, k0, bars + 8 * s);
            else tma_2d(st + 16384u, &tb, k0, n0, bars + 8 * s);
          }
        }
      }
    }
    __syncwarp();
    if (warp >= 2) {
      // ---- epilogue: TMEM lane quadrant = warp & 3 ----
      const int quad = warp & 3;
      bar_wait(bar_done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int gm = m0 + quad * 32 + lane;
      for (int cc = 0; cc < GN; cc += 16) {
        if (n0 + cc >= p.n) break;                              // warp-uniform
        uint32_t a[16];
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)cc;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                       "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (gm >= p.m) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int gn = n0 + cc + j;
          if (gn >= p.n) break;
          const float v = __uint_as_float(a[j]);
          if (p.epi == EPI_STORE) {
            float* c = p.C + (size_t)gm * p.ldc + gn;
            float o = (p.beta == 0.f) ? p.alpha * v : p.alpha * v + p.beta * *c;
            *c = p.round_out ? rna_tf32(o) : o;
          } else if (p.epi == EPI_SQUARE_ACCUM) {
            float* c = p.C + (size_t)gm * p.ldc + gn;
            *c += v * v;
          } else {
            const float sv = p.alpha * v;
            if (p.se.s_out) p.se.s_out[(size_t)gm * p.n + gn] = sv;
            if (gn < p.se.K0) {
              if (p.se.w_out) p.se.w_out[(size_t)gm * p.se.K0 + gn] = p.se.mu_w[(size_t)gm * p.se.K0 + gn] + sv;
            } else {
              if (p.se.b_out) p.se.b_out[gm] = p.se.mu_b[gm] + sv;
            }
          }
        }
      }
    }
  } else {
    // MMA issuer: the whole warp runs the loop in uniform control flow, only the tcgen05 instructions are predicated on
    // one elected lane, so that ptxas keeps descriptors and addresses in uniform registers (see syrk_tc.cu: an
    // `if (lane == 0)` loop costs ~224 cycles per instruction against 64 for the N = 128 MMA itself).
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                           ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
    const uint32_t u_tmem = __reduce_max_sync(0xffffffffu, tmem);
    // descriptor = (low word: start address + leading byte offset, constant high word); one k-step = +1024 B (MN) / +32 B (K)
    const uint64_t da = p.a_mn ? desc_mn(0u, 4096u) : desc_k(0u), db = p.b_mn ? desc_mn(0u, 4096u) : desc_k(0u);
    const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
    const uint32_t a_step = p.a_mn ? (1024u >> 4) : (32u >> 4), b_step = p.b_mn ? (1024u >> 4) : (32u >> 4);
    uint32_t acc = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nk; ++it) {
      bar_wait(bars + 8 * s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = sbase + (uint32_t)s * G_STAGE, sb = sa + 16384u;
      const uint32_t a0 = a_lo | ((sa >> 4) & 0x3FFFu), b0 = b_lo | ((sb >> 4) & 0x3FFFu);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                       "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, q;\n\t}"
                       ::"r"(u_tmem), "r"(a0 + (uint32_t)ks * a_step), "r"(a_hi), "r"(b0 + (uint32_t)ks * b_step), "r"(b_hi),
                         "r"(idesc), "r"(ks == 0 ? acc : 1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8 * (G_NSTAGE + s)) : "memory");
      }
      acc = 1;
      if (++s == G_NSTAGE) { s = 0; ph ^= 1u; }
    }
    if (leader)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

__global__ void __launch_bounds__(256) round_tf32_inplace_kernel(const float* in, float* out, size_t n) /* in == out allowed: no __restrict__ */ {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = rna_tf32(in[i]);
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncFn encoder() {
  static EncFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncFn)ptr;
    else
      cudaGetLastError();
  }
  return fn;
}

// rows x depth operand with element strides (s_row, s_k): which form, and its tensor map
bool operand_map(const float* base, int rows, int depth, long long s_row, long long s_k, int& mn, CUtensorMap& map) {
  if (((uintptr_t)base & 15) != 0) return false;
  cuuint64_t gd[2], gs[1];
  cuuint32_t box[2], es[2] = {1, 1};
  CUtensorMapSwizzle sw;
  if (s_k == 1 && (s_row % 4) == 0 && s_row >= depth) {          // contraction-contiguous
    mn = 0;
    gd[0] = (cuuint64_t)depth; gd[1] = (cuuint64_t)rows; gs[0] = (cuuint64_t)s_row * 4;
    box[0] = GK; box[1] = 128;
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  } else if (s_row == 1 && (s_k % 4) == 0 && s_k >= rows) {      // row-contiguous
    mn = 1;
    gd[0] = (cuuint64_t)rows; gd[1] = (cuuint64_t)depth; gs[0] = (cuuint64_t)s_k * 4;
    box[0] = 32; box[1] = GK;
    sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  } else {
    return false;
  }
  return encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns 0 on success, -1 if the operands cannot be fed by TMA (alignment / leading dimension not a multiple of
// 4 floats / no driver entry point) -- the caller then uses the CUDA-core kernel --, > 0 on error.
int gemm_tc_launch(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k, long long sb_n,
                   float* C, int ldc, int m, int n, int k, float alpha, float beta, int epilogue,
                   const SampleEpilogue* sample, int round_out, cudaStream_t s) {
  CRV_CHECK(A && B, "null GEMM operand");
  CRV_CHECK(m > 0 && n > 0 && k > 0, "bad GEMM shape %d x %d x %d", m, n, k);
  if (!encoder()) return -1;
  GtParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap ta, tb;
  if (!operand_map(A, m, k, sa_m, sa_k, p.a_mn, ta)) return -1;
  if (!operand_map(B, n, k, sb_n, sb_k, p.b_mn, tb)) return -1;
  p.m = m; p.n = n; p.k = k; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.epi = epilogue; p.round_out = round_out;
  p.C = C;
  if (sample) p.se = *sample;
  if (epilogue == 2) CRV_CHECK(sample != nullptr, "sample epilogue needs its descriptor");
  else CRV_CHECK(C != nullptr, "null GEMM output");
  dim3 grid((n + GN - 1) / GN, (m + GM - 1) / GM, 1);
  CRV_CHECK(grid.y < 65536, "GEMM m too large");
  CRV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
  gemm_tc_kernel<<<grid, G_THREADS, G_SMEM, s>>>(p, ta, tb);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

int round_tf32_launch(const float* in, float* out, size_t n, cudaStream_t s) {
  CRV_CHECK(in && out, "null pointer");
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  round_tf32_inplace_kernel<<<blocks, 256, 0, s>>>(in, out, n);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace crv
