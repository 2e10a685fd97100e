// K1 (tensor-core tiers): fused implicit-im2col + SYRK on the 5th-generation tensor cores.
//
//   F[k1,k2] += alpha * sum_r X[k1,r] * X[k2,r]       X = im2col(x) (+ ones row), never in HBM
//
// Two kernel families live in this file.
//
// (1) Channels-last operands -- the fast path (second half of the file: syrk_nhwc_kernel, its stream-K partition,
//     reductions, pre-passes, launch logic).  See the comment block "channels-last (NHWC) operands" below and DESIGN.md
//     section 4 / K1.
//
// (2) NCHW-dense operands and bias rows (first half: syrk_tc_kernel / syrk_tc_tma_kernel) -- one work item per CTA, 17
//     warps, 1 CTA / SM:
//   * the factor is cut into 256-row blocks; only block pairs (I >= J) are computed, and the
//     contraction axis R is split S ways so that pairs*S items fill the 148 SMs;
//   * 16 PRODUCER warps gather activations straight from the NCHW tensor (lane <-> contraction
//     position r, so global reads are coalesced along ow), round them to TF32 with cvt.rna and
//     store them into shared memory in the canonical K-major SWIZZLE_128B layout that tcgen05
//     expects (row = 128 bytes = 32 positions; 16-byte chunk index XOR (row & 7));
//     rows of a block are enumerated TAP-MAJOR (k' = (i*kw+j)*C + c) so that the 32 consecutive
//     rows a warp owns share one filter tap: one bounds test / offset per tap instead of per row;
//   * ONE thread of the MMA warp issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N<=256, K=8):
//     per stage 2 row halves x 4 k-steps, accumulating a 256x256 fp32 tile in TMEM (2 x 256 columns
//     = all 512 columns); tcgen05.commit releases the smem stage / signals the epilogue;
//     (this family is only used for the few operands family (1) cannot take -- on ResNet-50 the fc layer's A factor --
//     and still issues from one divergent thread; the channels-last kernel issues in uniform control flow)
//   * mbarrier full/empty ring of 3 x 64 KB stages between producers and the MMA thread;
//   * EPILOGUE (producer warps 0-3): tcgen05.ld the accumulator (32 lanes x 16 columns per
//     instruction) and store the partial tile to the workspace; a second, small kernel sums the S
//     partial tiles in a fixed order (deterministic), scales by alpha, undoes the tap-major
//     permutation and adds the tile and its mirror image into the factor arena.
#include "common.cuh"
#include "../../include/curvature_b200.h"
#include <math.h>
#include <stdlib.h>
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <vector>
#include <algorithm>
#include <cmath>
#include <utility>

namespace crv {
namespace {

constexpr int TB = 256;                       // rows of X per block (tile edge)
constexpr int STAGE_ROWS = 512;               // (row, 32-position segment) slots per pipeline stage
constexpr int STAGE_BYTES = STAGE_ROWS * 128; // 64 KB
constexpr int NSTAGE = 3;
constexpr int NPROD = 16;                     // producer warps
constexpr int NTHREADS = (NPROD + 1) * 32;
constexpr int TILE_ELEMS = TB * TB;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*barriers*/ + 1024 /*alignment slack*/;
constexpr uint32_t SPIN_LIMIT = 1u << 21;     // watchdog: trap instead of hanging the GPU

__device__ __align__(16) float g_zero_page[64];   // zeros: where padded / out-of-range positions read from

struct FastDiv {
  uint32_t mul, shift, one;  // q = one ? n : umulhi(n, mul) >> shift   (exact for n < 2^31)
};

struct TcParams {
  ConvGeom g;
  FastDiv divL, divOW, divC, divKW;
  int T, pairs, splits, chunks, cps;  // blocks, block pairs, R-splits, 32-chunks, chunks per split
  int HW, CHW, KK;
  int vec_ok;                         // operand rows are 16-byte aligned runs of 4 positions (LDG.128 path)
  float* ws;
};

struct TmaGeom {          // position patches and channel segments of the TMA-fed variant
  FastDiv divPPI, divPCW; // patches per image, patches per patch-row
  int bw, bh;             // patch = bh x bw output positions (bw * bh = 32)
  int chbox;              // channels per TMA box = rows per segment (min(C, 256))
  int flat;               // 1: (L, 1, C, N) view of a 1x1 operand; 0: (W, H, C, N)
  int pcw, ppi;           // patches per patch-row, patches per image
  int nimg;               // N
};

__host__ FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  if (d <= 1) { f.mul = 0; f.shift = 0; f.one = 1; return f; }
  uint32_t s = 0;
  while ((1u << (s + 1)) <= d) ++s;          // s = floor(log2 d)
  if ((d & (d - 1)) == 0) {                  // power of two: mul = 2^(32-s) would overflow for s = 0 only
    f.mul = (uint32_t)(1ull << (32 - s));
    f.shift = 0;
    f.one = 0;
    return f;
  }
  f.mul = (uint32_t)(((1ull << (32 + s)) / d) + 1);
  f.shift = s;
  f.one = 0;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.one ? n : (__umulhi(n, f.mul) >> f.shift);
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > SPIN_LIMIT) asm volatile("trap;");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Warp-uniform value -> uniform register (REDUX writes a UR): lets ptxas keep everything derived from it uniform.
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
// tcgen05.mma with the two shared-memory descriptors given as (low word, common high word): the low word holds the
// start-address and leading-byte-offset fields, so stepping through a stage is one 32-bit add per operand.
__device__ __forceinline__ void tc_mma_lohi(bool bf16, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                            uint32_t idesc, uint32_t accumulate) {
  if (bf16)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// launch-level trace: first-start / last-end of a kernel's CTAs into two words of the launch's slot
__device__ __forceinline__ void trace_begin(unsigned long long* tr, int w) {
  if (tr && threadIdx.x == 0) atomicMin(tr + w, gtimer());
}
__device__ __forceinline__ void trace_end(unsigned long long* tr, int w) {
  if (tr && threadIdx.x == 0) atomicMax(tr + w + 1, gtimer());
}
struct TraceScope {
  unsigned long long* tr; int w;
  __device__ __forceinline__ TraceScope(unsigned long long* t, int word) : tr(t), w(word) { trace_begin(tr, w); }
  __device__ __forceinline__ ~TraceScope() { trace_end(tr, w); }
};
__device__ __forceinline__ long long nv_kg(int NB, int left, int kpb) { return (long long)min(NB, left) * kpb; }
// 1 in exactly one (elected) lane of a fully converged warp
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t cvt_tf32(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return u;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in bits
// [0,14), leading byte offset (unused for one swizzle atom along K) in [16,30), stride byte offset
// = 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), swizzle mode 2 (128 B)
// in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = fp32 (bits 4-5 = 1), A and B = TF32 (format 2 at bits 7-9 / 10-12),
// both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t umma_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void decode_pair(int pair, int T, int& I, int& J) {
  const int noff = T * (T - 1) / 2;  // off-diagonal pairs first (they are the heavy items)
  if (pair < noff) {
    int i = (int)((1.f + sqrtf(1.f + 8.f * (float)pair)) * 0.5f);
    while (i * (i - 1) / 2 > pair) --i;
    while ((i + 1) * i / 2 <= pair) ++i;
    I = i;
    J = pair - i * (i - 1) / 2;
  } else {
    I = J = pair - noff;
  }
}

// ---- pieces shared by the thread-staged and the TMA-fed kernels ------------------------------------
struct ItemShape {
  int I, J, rowsA, mh, ncols, RP, SC, cb, ce, nstage_it;
  int ncols0;      // columns of the first 128-row half that hold results (channels-last kernel: 128 for a two-half diagonal tile)
  bool diag;
};
__device__ __forceinline__ ItemShape item_shape(const TcParams& p, int item) {
  ItemShape t;
  const int pair = item / p.splits, split = item - pair * p.splits;
  decode_pair(pair, p.T, t.I, t.J);
  t.diag = (t.I == t.J);
  t.rowsA = min(TB, p.g.D - t.I * TB);
  t.mh = (t.rowsA + 127) >> 7;                                   // 128-row halves of the A block
  t.ncols = t.diag ? ((t.rowsA + 15) & ~15) : TB;                // UMMA N
  t.RP = t.diag ? (t.rowsA <= 128 ? 128 : 256) : 512;            // rows per sub-chunk in a stage
  t.SC = STAGE_ROWS / t.RP;                                      // 32-position sub-chunks per stage
  t.cb = split * p.cps;
  t.ce = min(p.chunks, t.cb + p.cps);
  t.nstage_it = (t.ce - t.cb + t.SC - 1) / t.SC;
  return t;
}

// One thread: for every stage wait for the operands, issue 4 k-steps x mh row halves of
// tcgen05.mma.kind::tf32 (M=128, N=ncols, K=8) per sub-chunk, then commit to free the stage.
__device__ __forceinline__ void mma_issue_all(const ItemShape& t, uint32_t sbase, uint32_t bars, uint32_t bar_tmem_full,
                                              uint32_t tmem) {
  const uint32_t idesc = umma_idesc(128, (uint32_t)t.ncols);
  uint32_t acc = 0;
  for (int it = 0; it < t.nstage_it; ++it) {
    const int s = it % NSTAGE;
    const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
    mbar_wait(bars + 8 * s, ph);
    tc_fence_after();
    const uint32_t st = sbase + (uint32_t)s * STAGE_BYTES;
    for (int q = 0; q < t.SC; ++q) {
      const uint32_t sub = st + (uint32_t)(q * t.RP) * 128u;
      const uint32_t bsub = t.diag ? sub : sub + TB * 128u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bdesc = umma_desc(bsub + ks * 32);
        for (int h = 0; h < t.mh; ++h) {
          const uint64_t adesc = umma_desc(sub + (uint32_t)h * (128u * 128u) + ks * 32);
          tc_mma_tf32(tmem + (uint32_t)h * 256u, adesc, bdesc, idesc, acc);
        }
        acc = 1;
      }
    }
    tc_commit(bars + 8 * (NSTAGE + s));                // frees the stage when these MMAs retire
  }
  tc_commit(bar_tmem_full);                            // accumulator complete -> epilogue
}

// One warp (TMEM lane quadrant `quad`): accumulator -> registers -> partial tile in the workspace.
__device__ __forceinline__ void epilogue_store(const ItemShape& t, uint32_t bar_tmem_full, uint32_t parity, uint32_t tmem,
                                               float* __restrict__ wsp, int quad, int lane) {
  mbar_wait(bar_tmem_full, parity);
  tc_fence_after();
  for (int h = 0; h < t.mh; ++h) {
    const int row = h * 128 + quad * 32 + lane;
    for (int cc = 0; cc < t.ncols; cc += 16) {
      uint32_t a[16];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256 + cc);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
            "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < t.rowsA) {
        uint4* dst = reinterpret_cast<uint4*>(wsp + (size_t)row * TB + cc);
        dst[0] = make_uint4(a[0], a[1], a[2], a[3]);
        dst[1] = make_uint4(a[4], a[5], a[6], a[7]);
        dst[2] = make_uint4(a[8], a[9], a[10], a[11]);
        dst[3] = make_uint4(a[12], a[13], a[14], a[15]);
      }
    }
  }
}

// Same, with coalesced stores: a lane holds 32 columns of ITS row, so a direct store instruction touches 32 rows
// (32 sectors for 512 bytes).  Each warp instead passes its 32 x 32 block through a private 4 KB shared-memory tile
// (16-byte chunks XOR-swizzled with the row, conflict-free both ways) and writes four whole 128-byte row segments per
// instruction: 16x fewer memory transactions, the accumulator drains in ~1/3 of the time.
__device__ __forceinline__ void epilogue_store_coalesced(const ItemShape& t, uint32_t bar_tmem_full, uint32_t parity, uint32_t tmem,
                                                         float* __restrict__ wsp, int quad, int lane, uint32_t stg) {
  mbar_wait(bar_tmem_full, parity);
  tc_fence_after();
  const int sub = lane >> 3, ch = lane & 7;             // read-back role: row within a group of 4, 16-byte chunk
  for (int h = 0; h < t.mh; ++h) {
    const int row0 = h * 128 + quad * 32;
    const int nc = h == 0 ? t.ncols0 : t.ncols;
    for (int cc = 0; cc < nc; cc += 32) {
      uint32_t a[32];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256 + cc);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
            "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
            "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),
            "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();                                     // the previous block has been read back
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((k ^ (lane & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[4 * k]), "r"(a[4 * k + 1]),
                     "r"(a[4 * k + 2]), "r"(a[4 * k + 3]) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int r4 = 0; r4 < 8; ++r4) {
        const int r = r4 * 4 + sub;
        uint32_t v0, v1, v2, v3;
        const uint32_t addr = stg + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) * 16);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr) : "memory");
        if (row0 + r < t.rowsA && cc + ch * 4 < t.ncols)
          *reinterpret_cast<uint4*>(wsp + (size_t)(row0 + r) * TB + cc + ch * 4) = make_uint4(v0, v1, v2, v3);
      }
    }
  }
}

// The same drain with the global side handed to the TMA unit (experiment, off by default: CURVATURE_B200_DBG=8): a
// 32 x 32 block goes TMEM -> registers -> the SWIZZLE_128B staging tile and one lane issues a tensor store of the 4 KB
// tile into the partial-tile buffer (`part`: see GroupMaps; rows of the slot start at `slot_row0`).  In the chain kernel
// this halved the drain (profiles/r2_chain_ablation.txt); here it does not pay (profiles/r2_syrk_tma_flush_ab.txt:
// 24.36 k img/s against 24.53 k with the warps' own stores) -- all 148 CTAs flush their 256 KB accumulators at the same
// moment at the end of a launch, and 38 MB in ~11 us is what L2 takes, whoever issues the stores.
__device__ __forceinline__ void epilogue_store_tma(const ItemShape& t, uint32_t bar_tmem_full, uint32_t parity, uint32_t tmem,
                                                   const CUtensorMap* part, int slot_row0, int quad, int lane, uint32_t stg) {
  mbar_wait(bar_tmem_full, parity);
  tc_fence_after();
  for (int h = 0; h < t.mh; ++h) {
    const int row0 = h * 128 + quad * 32;
    if (row0 >= t.rowsA) continue;                        // (warp-uniform)
    const int nc = h == 0 ? t.ncols0 : t.ncols;
    for (int cc = 0; cc < nc; cc += 32) {
      uint32_t a[32];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256 + cc);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
            "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
            "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),
            "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the previous block's store has read the tile
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((k ^ (lane & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[4 * k]), "r"(a[4 * k + 1]),
                     "r"(a[4 * k + 2]), "r"(a[4 * k + 3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(reinterpret_cast<uint64_t>(part)), "r"(cc), "r"(slot_row0 + row0), "r"(stg) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
}

// ---- TMA-fed kernel ---------------------------------------------------------------------------------
// Same tiles, same MMA loop, same epilogue; the operands are fetched by the TMA unit instead of by threads:
// one cp.async.bulk.tensor.4d per (filter tap, channel segment) box of [channels][bh x bw positions] lands
// directly in the K-major SWIZZLE_128B layout, with the hardware's out-of-bounds zero fill providing the
// convolution padding (box coordinates are shifted by the tap offset and may be negative).  No LSU traffic, no
// per-element instructions.  The tensor core reads the fp32 words as TF32 (low 13 mantissa bits ignored), so this
// variant has TF32-truncation accuracy (stated 1e-3 tier) instead of the round-to-nearest of the staged path.
// Warps: 0 = TMA producer (one thread), 1 = MMA issuer (one thread) + TMEM owner, 2..5 = epilogue.
constexpr int TMA_THREADS = 6 * 32;
__global__ void __launch_bounds__(TMA_THREADS, 1)
syrk_tc_tma_kernel(const TcParams p, const TmaGeom tg, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const uint32_t bars = sbase + NSTAGE * STAGE_BYTES;
  const uint32_t bar_tmem_full = bars + 8 * (2 * NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw) + NSTAGE * STAGE_BYTES + 8 * (2 * NSTAGE + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  const ItemShape t = item_shape(p, blockIdx.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (NSTAGE + s), 1);
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // channel segments of the two blocks: rows [sg*chbox, +chbox) <-> tap-major index k' = blk*256 + sg*chbox
      const int segs = TB / tg.chbox;
      const int nsegA = min(segs, (g.K0 - t.I * TB + tg.chbox - 1) / tg.chbox);
      const int nsegB = t.diag ? 0 : segs;                      // off-diagonal: block J is always complete
      const uint32_t box_bytes = (uint32_t)tg.chbox * 128u;
      const uint32_t stage_tx = (uint32_t)(t.SC * (nsegA + nsegB)) * box_bytes;
      for (int it = 0; it < t.nstage_it; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(bars + 8 * (NSTAGE + s), ph ^ 1u);
        mbar_arrive_expect_tx(bars + 8 * s, stage_tx);
        const uint32_t st = sbase + (uint32_t)s * STAGE_BYTES;
        for (int q = 0; q < t.SC; ++q) {
          // chunk -> (image, patch row, patch column); chunks past the end map to image index >= N: all zero fill
          const uint32_t chunk = (uint32_t)(t.cb + it * t.SC + q);
          const uint32_t n = fdiv(chunk, tg.divPPI);
          const uint32_t rem = chunk - n * (uint32_t)tg.ppi;
          const uint32_t pr = fdiv(rem, tg.divPCW);
          const uint32_t pc = rem - pr * (uint32_t)tg.pcw;
          const int w0 = (int)pc * tg.bw, h0 = (int)pr * tg.bh;
          const uint32_t sub = st + (uint32_t)(q * t.RP) * 128u;
          for (int sg = 0; sg < nsegA + nsegB; ++sg) {
            const bool isB = sg >= nsegA;
            const int kp = (isB ? t.J * TB + (sg - nsegA) * tg.chbox : t.I * TB + sg * tg.chbox);
            const int tap = (int)fdiv((uint32_t)kp, p.divC);
            const int c0 = kp - tap * g.C;
            const int ti = (int)fdiv((uint32_t)tap, p.divKW);
            const int tj = tap - ti * g.kw;
            const uint32_t dst = sub + (uint32_t)((isB ? TB + (sg - nsegA) * tg.chbox : sg * tg.chbox)) * 128u;
            tma_load_4d(dst, &tmap, w0 + tj - g.pw, h0 + ti - g.ph, c0, (int)n, bars + 8 * s);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) mma_issue_all(t, sbase, bars, bar_tmem_full, tmem);
    __syncwarp();
  } else {
    epilogue_store(t, bar_tmem_full, 0u, tmem, p.ws + (size_t)blockIdx.x * TILE_ELEMS, warp & 3, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- main kernel ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) syrk_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;     // swizzle-128B atoms need 1024-byte alignment
  const uint32_t bars = sbase + NSTAGE * STAGE_BYTES;
  // full[s] = bars + 8 s ; empty[s] = bars + 8 (NSTAGE + s) ; tmem_full = bars + 8 * 2 NSTAGE
  const uint32_t bar_tmem_full = bars + 8 * (2 * NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw) + NSTAGE * STAGE_BYTES + 8 * (2 * NSTAGE + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;

  const int item = blockIdx.x;
  const int pair = item / p.splits, split = item - pair * p.splits;
  int I, J;
  decode_pair(pair, p.T, I, J);
  const bool diag = (I == J);
  const int rowsA = min(TB, g.D - I * TB);
  const int mh = (rowsA + 127) >> 7;                               // 128-row halves of the A block
  const int ncols = diag ? ((rowsA + 15) & ~15) : TB;              // UMMA N
  const int RP = diag ? (rowsA <= 128 ? 128 : 256) : 512;          // rows per sub-chunk in a stage
  const int SC = STAGE_ROWS / RP;                                  // 32-position sub-chunks per stage
  const int cb = split * p.cps;
  const int ce = min(p.chunks, cb + p.cps);
  const int nstage_it = (ce - cb + SC - 1) / SC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bars + 8 * s, NPROD);
      mbar_init(bars + 8 * (NSTAGE + s), 1);
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPROD) {  // TMEM: all 512 columns (two 128 x 256 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < NPROD) {
    // ================= producers =================
    const int q0 = warp * 32;
    const int sc = q0 / RP;
    const int row0 = q0 - sc * RP;                         // first of this warp's 32 rows in the sub-chunk
    int kp0;                                               // tap-major row index of that row
    if (diag) kp0 = I * TB + row0;
    else kp0 = (row0 < TB) ? (I * TB + row0) : (J * TB + row0 - TB);
    const bool active = kp0 < g.D && (diag || row0 >= TB || row0 < mh * 128);
    const int t0 = (int)fdiv((uint32_t)min(kp0, g.K0), p.divC);
    const int c0 = min(kp0, g.K0) - t0 * g.C;
    uint32_t swz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) swz[j] = ((((uint32_t)lane >> 2) ^ (uint32_t)j) << 4) | (((uint32_t)lane & 3u) << 2);
    const uint32_t sub_off = (uint32_t)(sc * RP + row0) * 128u;
    const uint32_t r_end = (uint32_t)min((long long)ce * 32, g.R);

    // Stage hand-over shared by both producer paths: wait until the MMAs that read this slot have retired,
    // store the 32 staged rows (TF32-rounded) into the swizzled layout, publish to the async proxy, arrive.
    auto commit_stage = [&](int it, const float (&v)[32], bool has_data) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      mbar_wait(bars + 8 * (NSTAGE + s), ph ^ 1u);
      if (has_data) {
        const uint32_t dst = sbase + (uint32_t)s * STAGE_BYTES + sub_off;
#pragma unroll
        for (int j = 0; j < 32; ++j) sts32(dst + (uint32_t)j * 128u + swz[j & 7], cvt_tf32(v[j]));
      }
      fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * s);
    };

    // FAST PATH: the warp's 32 rows are 32 consecutive channels of ONE filter tap (true for every row group
    // of a layer whose channel count is a multiple of 32, i.e. all ResNet 1x1 / 3x3 convolutions and all
    // output-gradient operands).  Per stage each lane derives one pointer (or the zero page when its position
    // is padding / beyond the item) and walks the channels with a constant stride: 2 instructions per row,
    // no predicates.  Loads of stage it+1 are issued before stage it is stored (register double buffer), so
    // global/L2 latency overlaps the previous stage's stores and the tensor core never waits on a cold load.
    const bool fast = active && (kp0 + 32 <= g.K0) && (c0 + 32 <= g.C);
    if (fast && p.vec_ok) {
      // VECTOR PATH (1x1 / stride 1 / no padding, L % 4 == 0: every output-gradient operand and every 1x1
      // convolution input on 56^2, 28^2, 14^2 maps): a lane owns 4 consecutive positions = one 16-byte swizzle
      // chunk, a warp instruction covers 4 rows x 32 positions.  LDG.128 + 4 cvt.rna + STS.128 per 4 elements:
      // a quarter of the LSU requests and a third of the instructions of the scalar path.
      const int rsub = lane >> 3, chk = lane & 7;
      uint32_t vo[2];
#pragma unroll
      for (int par = 0; par < 2; ++par)
        vo[par] = (uint32_t)(rsub * 128) + (((uint32_t)chk ^ (uint32_t)(par * 4 + rsub)) << 4);
      const int cbase = (c0 + rsub) * p.HW;
      auto issue_vec = [&](int it, float4 (&v)[8]) {
        const uint32_t chunk = (uint32_t)(cb + it * SC + sc);
        uint32_t r = chunk * 32u + (uint32_t)chk * 4u;
        const bool vr = r < r_end;
        if (!vr) r = 0;
        const uint32_t n = fdiv(r, p.divL);
        const uint32_t l = r - n * (uint32_t)g.L;
        const float* __restrict__ ptr = vr ? (g.x + (size_t)n * (size_t)p.CHW + (size_t)(cbase + (int)l)) : g_zero_page;
        const size_t stride = vr ? (size_t)(4 * p.HW) : 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q] = __ldg(reinterpret_cast<const float4*>(ptr));
          ptr += stride;
        }
      };
      auto commit_vec = [&](int it, const float4 (&v)[8]) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(bars + 8 * (NSTAGE + s), ph ^ 1u);
        const uint32_t dst = sbase + (uint32_t)s * STAGE_BYTES + sub_off;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)q * 512u + vo[q & 1]),
                       "r"(cvt_tf32(v[q].x)), "r"(cvt_tf32(v[q].y)), "r"(cvt_tf32(v[q].z)), "r"(cvt_tf32(v[q].w))
                       : "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * s);
      };
      float4 va[8], vb[8];
      issue_vec(0, va);
      for (int it = 0; it < nstage_it; it += 2) {
        if (it + 1 < nstage_it) issue_vec(it + 1, vb);
        commit_vec(it, va);
        if (it + 1 < nstage_it) {
          if (it + 2 < nstage_it) issue_vec(it + 2, va);
          commit_vec(it + 1, vb);
        }
      }
    } else if (fast) {
      const int ti = (int)fdiv((uint32_t)t0, p.divKW);
      const int tj = t0 - ti * g.kw;
      const int dih = ti - g.ph, diw = tj - g.pw;
      const int cbase = c0 * p.HW;
      auto issue_stage = [&](int it, float (&v)[32]) {
        const uint32_t chunk = (uint32_t)(cb + it * SC + sc);
        uint32_t r = chunk * 32u + (uint32_t)lane;
        const bool vr = r < r_end;
        if (!vr) r = 0;
        const uint32_t n = fdiv(r, p.divL);
        const uint32_t l = r - n * (uint32_t)g.L;
        const uint32_t oh = fdiv(l, p.divOW);
        const uint32_t ow = l - oh * (uint32_t)g.OW;
        const int ih = (int)oh * g.sh + dih, iw = (int)ow * g.sw + diw;
        const bool ok = vr && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W;
        const float* __restrict__ ptr = ok ? (g.x + (size_t)n * (size_t)p.CHW + (size_t)(cbase + ih * g.W + iw))
                                           : g_zero_page;
        const size_t stride = ok ? (size_t)p.HW : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __ldg(ptr);
          ptr += stride;
        }
      };
      float va[32], vb[32];
      issue_stage(0, va);
      for (int it = 0; it < nstage_it; it += 2) {
        if (it + 1 < nstage_it) issue_stage(it + 1, vb);
        commit_stage(it, va, true);
        if (it + 1 < nstage_it) {
          if (it + 2 < nstage_it) issue_stage(it + 2, va);
          commit_stage(it + 1, vb, true);
        }
      }
    } else {
    // GENERIC PATH: any geometry (channel counts that are not multiples of 32, the bias row, ragged ends).
    for (int it = 0; it < nstage_it; ++it) {
      float v[32];
      if (active) {
        const uint32_t chunk = (uint32_t)(cb + it * SC + sc);
        uint32_t r = chunk * 32u + (uint32_t)lane;
        const bool vr = r < r_end;
        if (!vr) r = 0;
        const uint32_t n = fdiv(r, p.divL);
        const uint32_t l = r - n * (uint32_t)g.L;
        const uint32_t oh = fdiv(l, p.divOW);
        const uint32_t ow = l - oh * (uint32_t)g.OW;
        const int ihb = (int)oh * g.sh - g.ph, iwb = (int)ow * g.sw - g.pw;
        const float* __restrict__ xn = g.x + (size_t)n * (size_t)p.CHW;
        int t = t0, c = c0, coff = c0 * p.HW, kp = kp0, toff = 0;
        bool ok = false;
        auto settap = [&]() {
          ok = false;
          if (t < p.KK) {
            const int i = (int)fdiv((uint32_t)t, p.divKW);
            const int jj = t - i * g.kw;
            const int ih = ihb + i, iw = iwb + jj;
            ok = vr && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W;
            toff = ih * g.W + iw;
          }
        };
        settap();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float val = 0.f;
          if (t < p.KK) {
            if (ok) val = __ldg(xn + (toff + coff));
          } else if (kp == g.K0 && g.has_bias && vr) {
            val = 1.f;                                     // the ones row of the bias
          }
          v[j] = val;
          ++kp; ++c; coff += p.HW;
          if (c == g.C) { c = 0; coff = 0; ++t; settap(); }
        }
      }
      commit_stage(it, v, active);
    }
    }

    // ================= epilogue (warps 0-3: TMEM lane quadrant = warp index) =================
    if (warp < 4) {
      mbar_wait(bar_tmem_full, 0);
      tc_fence_after();
      float* __restrict__ wsp = p.ws + (size_t)item * TILE_ELEMS;
      for (int h = 0; h < mh; ++h) {
        const int row = h * 128 + warp * 32 + lane;
        for (int cc = 0; cc < ncols; cc += 16) {
          uint32_t a[16];
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * 256 + cc);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (row < rowsA) {
            uint4* dst = reinterpret_cast<uint4*>(wsp + (size_t)row * TB + cc);
            dst[0] = make_uint4(a[0], a[1], a[2], a[3]);
            dst[1] = make_uint4(a[4], a[5], a[6], a[7]);
            dst[2] = make_uint4(a[8], a[9], a[10], a[11]);
            dst[3] = make_uint4(a[12], a[13], a[14], a[15]);
          }
        }
      }
    }
  } else {
    // ================= MMA issuer: one thread =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, (uint32_t)ncols);
      uint32_t acc = 0;
      for (int it = 0; it < nstage_it; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(bars + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = sbase + (uint32_t)s * STAGE_BYTES;
        for (int q = 0; q < SC; ++q) {
          const uint32_t sub = st + (uint32_t)(q * RP) * 128u;
          const uint32_t bsub = diag ? sub : sub + TB * 128u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bdesc = umma_desc(bsub + ks * 32);
            for (int h = 0; h < mh; ++h) {
              const uint64_t adesc = umma_desc(sub + (uint32_t)h * (128u * 128u) + ks * 32);
              tc_mma_tf32(tmem + (uint32_t)h * 256u, adesc, bdesc, idesc, acc);
            }
            acc = 1;
          }
        }
        tc_commit(bars + 8 * (NSTAGE + s));                // frees the stage when these MMAs retire
      }
      tc_commit(bar_tmem_full);                            // accumulator complete -> epilogue
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NPROD) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- fixed-order reduction of the S partial tiles into the factor ------------------------------
// One CTA of 1024 threads per 32x32 sub-tile of a block pair, one element per thread.  Every thread sums its S
// partials in split order (deterministic) with 16 independent loads in flight -- the partials were just written
// and sit in L2, so the kernel is a latency chain of S/16 round trips, not a bandwidth problem.  The direct block
// is added row-wise and the mirror image through a shared-memory transpose, so both read-modify-writes of the
// factor are coalesced along the tap-major index.
__global__ void __launch_bounds__(1024) syrk_tc_reduce_kernel(const TcParams p, const float alpha, float* __restrict__ F) {
  __shared__ float tile[32][33];
  const int pair = blockIdx.x >> 6, sub = blockIdx.x & 63;
  const int br = sub >> 3, bc = sub & 7;
  int I, J;
  decode_pair(pair, p.T, I, J);
  const bool diag = (I == J);
  if (diag && bc > br) return;            // diagonal blocks: lower triangle only, mirrored below (exact symmetry)
  const ConvGeom& g = p.g;
  const int rowsA = min(TB, g.D - I * TB);
  const int colsB = diag ? rowsA : TB;
  if (br * 32 >= rowsA || bc * 32 >= colsB) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  auto perm = [&](int kp) -> int {        // tap-major k' -> the reference's row index c*kh*kw + tap
    if (kp >= g.K0) return kp;
    const int t = (int)fdiv((uint32_t)kp, p.divC);
    const int c = kp - t * g.C;
    return c * p.KK + t;
  };
  const float* __restrict__ base = p.ws + (size_t)pair * p.splits * TILE_ELEMS;
  {
    const int row = br * 32 + w, col = bc * 32 + lane;
    const bool valid = row < rowsA && col < colsB;
    float v = 0.f;
    if (valid) {
      const float* __restrict__ b = base + row * TB + col;
      float sum = 0.f;
      int s = 0;
      for (; s + 16 <= p.splits; s += 16) {
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = __ldcg(b + (size_t)(s + u) * TILE_ELEMS);
#pragma unroll
        for (int u = 0; u < 16; ++u) sum += t[u];
      }
      for (; s < p.splits; ++s) sum += __ldcg(b + (size_t)s * TILE_ELEMS);
      v = alpha * sum;
      if (!(diag && col > row)) F[(size_t)perm(I * TB + row) * g.D + perm(J * TB + col)] += v;
    }
    tile[w][lane] = v;
  }
  __syncthreads();
  {                                       // mirror image: lanes run along the original rows
    const int col = bc * 32 + w, row = br * 32 + lane;
    if (row < rowsA && col < colsB && !(diag && col >= row))
      F[(size_t)perm(J * TB + col) * g.D + perm(I * TB + row)] += tile[lane][w];
  }
}

struct Plan {
  int T, pairs, splits, chunks, cps;
  size_t ws_bytes;
};

Plan make_plan(const ConvGeom& g, int sms, int chunks) {
  Plan pl;
  pl.T = (g.D + TB - 1) / TB;
  pl.pairs = pl.T * (pl.T + 1) / 2;
  pl.chunks = chunks;
  if (sms <= 0) sms = 148;
  // choose the number of R-splits: estimated makespan (in chunk units) = waves * (chunks/S + fixed overhead)
  // + cost of reducing S partial tiles per pair
  const int maxS = pl.chunks / 16 > 0 ? pl.chunks / 16 : 1;
  double best = 1e300;
  int bestS = 1;
  for (int S = 1; S <= maxS && S <= 8 * sms; ++S) {
    const long long items = (long long)pl.pairs * S;
    const long long waves = (items + sms - 1) / sms;
    const double est = (double)waves * ((double)pl.chunks / S + 10.0) + 0.14 * (double)items;
    if (est < best * 0.999) { best = est; bestS = S; }
  }
  pl.cps = (pl.chunks + bestS - 1) / bestS;
  pl.cps = (pl.cps + 3) & ~3;                         // whole stages for every sub-chunk count (1, 2, 4)
  pl.splits = (pl.chunks + pl.cps - 1) / pl.cps;
  pl.ws_bytes = (size_t)pl.pairs * pl.splits * TILE_ELEMS * sizeof(float);
  return pl;
}

// Can this operand be fetched by TMA?  Needs a 1x1 / unit-stride / unpadded operand (its flattened (L, C, N) view
// lets the out-of-bounds fill zero the ragged last chunk of every image), no bias row, a power-of-two channel
// count (boxes of min(C,256) channels) and 16-byte aligned rows (L % 4 == 0).
bool tma_geometry(const ConvGeom& g, TmaGeom& tg, int& chunks) {
  if (g.has_bias || g.sh != 1 || g.sw != 1) return false;
  if (((uintptr_t)g.x & 15) != 0) return false;
  if (g.C < 16 || (g.C & (g.C - 1)) != 0) return false;
  tg.chbox = g.C < TB ? g.C : TB;
  tg.nimg = g.N;
  const int KK = g.kh * g.kw;
  tg.flat = (KK == 1 && g.ph == 0 && g.pw == 0) ? 1 : 0;
  int pch;
  if (tg.flat) {
    if (g.L % 4) return false;
    tg.bw = 32; tg.bh = 1;
    tg.pcw = (g.L + 31) / 32;
    pch = 1;
  } else {
    // Measured on B200 (scripts/experiments/tma_align_probe.cu): cp.async.bulk.tensor raises an illegal-instruction
    // fault when the innermost start coordinate is not a multiple of 16 bytes, so the +-1 column shifts of a k x k
    // filter cannot be expressed as TMA box coordinates on fp32 NCHW data.  k x k operands stay thread-staged.
    return false;
  }
  tg.ppi = tg.pcw * pch;
  tg.divPPI = make_fastdiv((uint32_t)tg.ppi);
  tg.divPCW = make_fastdiv((uint32_t)tg.pcw);
  const long long c = (long long)g.N * tg.ppi;
  if (c >= (1LL << 26)) return false;
  chunks = (int)c;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
    else
      cudaGetLastError();
  }
  return fn;
}


// ---- channels-last (NHWC) operands: MN-major, TMA-fed kernel -------------------------------------------
// When the recorded tensor is channels-last -- [N][H][W][C] for activations / output gradients, [N][F] for
// Linear operands -- the channel axis is contiguous, so for a fixed position the 32 channels of a chunk are one
// 128-byte run: exactly one row of the MN-major operand layout of tcgen05 (rows = contraction
// index = position, 128 bytes = 32 consecutive operand rows).  A filter tap is then a shift of the box
// COORDINATES in the W / H dimensions (never of the innermost dimension), which TMA handles at any alignment,
// with out-of-bounds zero fill supplying the convolution padding and elementStrides supplying the stride.
// Every operand of every layer (k x k A, 1 x 1 A, G, Linear) becomes the same thing: per stage, per 32-channel
// chunk, NB boxes of [bh x bw positions] x [32 channels] fetched by one cp.async.bulk.tensor each.  No thread
// touches the data on its way from L2 to the tensor core.
//
// Positions of a box that do not exist are never loaded: box extents divide the output grid (bw | OW, bh | OH),
// and where bw*bh is not a multiple of 8 (the contraction depth of one tf32 MMA) the remaining rows of the
// box's shared-memory slot are zeroed once at kernel start -- TMA never writes them.
constexpr int NH_THREADS = 13 * 32;            // warp 0: TMA; 1: MMA + TMEM owner; 2-5: epilogue; 6-12: TMA.  Registers are
                                               // allocated in groups of 4 warps: the kernel is capped at 96 registers (16 warps x
                                               // 96 x 32 = 48 K) so that a reduction / pre-pass CTA fits beside it on the SM
constexpr int NH_MAXSTAGE = 8;
constexpr int NH_NPROD = 8;                   // TMA-issuing warps (one elected lane each)
constexpr int NH_DATA_BYTES = 192 * 1024;     // operand stage ring (3 x 64 KB / 6 x 32 KB); the CTA leaves ~16 KB of the SM's
                                              // shared memory free so that reduction / pre-pass CTAs can be co-resident
constexpr int NH_EPI_BYTES = 16 * 1024;       // epilogue staging: 4 warps x (32 rows x 128 B)
constexpr int NH_SMEM_BYTES = NH_DATA_BYTES + NH_EPI_BYTES + 1024 /*barriers, chunk table*/ + 1024 /*alignment slack*/;
static const int NH_STAGE_TARGET = getenv("CURVATURE_B200_STAGE_KB") ? atoi(getenv("CURVATURE_B200_STAGE_KB")) * 1024 : 64 * 1024;

struct NhParams {
  int D, C, KK, kw;
  FastDiv divC, divKW;
  int T, splits;
  int nbox, bps;             // boxes in total / per split
  int PB, PBv;               // rows per box slot (multiple of 8) / rows a box really holds
  int NBoff, NBdiag;         // boxes per pipeline stage for off-diagonal / diagonal items
  FastDiv divPPI, divPCW;
  int ppi, pcw, bw, bh, bn;  // boxes per image group, boxes per box-row, box extent in output positions / images
  int sh, sw, ph, pw, flat;
  int K0;                    // rows that are tap-major permuted (the reduction undoes it)
  int ldF;                   // order of the factor F (= D except on the packed path, where D counts padding rows)
  int pk_kh, pk_kw, pk_c;    // packed small-C path: the original filter and channel count (pk_c = 0: not packed)
  int pack2;                 // D <= 64 single-tile TF32 factor: two position sets stacked along M (see seg_geom)
  int x3;                    // bf16x3 tier: the operand copy has two bf16 planes, x = hi + lo; three MMAs per k-group
  float alpha;
  float* F;                  // the factor this item accumulates into
};

// One launch works on a GROUP of factors: the work list is the concatenation of their pair lists (global pair index
// q; factor i owns q in [qbeg[i], qbeg[i+1])).  Read-once, HBM-bound factors of many layers share one launch (their
// fixed costs -- launch, pipeline ramp, accumulator flush, reduction -- are paid once per CTA instead of once per
// factor per CTA); a factor whose operand is re-read many times from L2 is a group of one, so that all SMs walk the
// same tensor at the same time.
constexpr int GRP_MAXF = 48;
struct GroupParams {
  int nf;
  int dbg;                   // ablation switches for profiling (bit 0: no TMA loads, bit 1: no MMAs); 0 in production
  long long* tl;             // per-CTA timeline (profiling aid, see crv_debug_timeline); null in production
  unsigned long long* trace; // launch-level trace slot (crv_debug_trace); null in production
  float* ws;                 // partial tiles, slot = CTA + q
  int qbeg[GRP_MAXF + 1];
  NhParams f[GRP_MAXF];
};
struct alignas(64) GroupMaps {
  CUtensorMap m[GRP_MAXF];
  CUtensorMap part;          // the launch's partial-tile buffer as a [(slots x 256) rows][256 columns] fp32 tensor, boxes of
                             // 32 x 32 in SWIZZLE_128B form: the accumulator flush goes out through the TMA unit
};

// MN-major descriptor for 32-bit operands.  tcgen05 accepts exactly one shared-memory layout for MN-major TF32
// (measured with scripts/experiments/mn_major_probe.cu; CUTLASS calls it SW128_32B): 128-byte rows (32 fp32 along
// M/N) whose 32-byte quarters are XOR-swizzled with (row & 3) -- what TMA writes in CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// mode -- in atoms of 4 consecutive rows (contraction index).  Layout type 1 in bits [61,64); stride byte offset =
// 512 B between the two 4-row atoms a K = 8 instruction reads; leading byte offset = distance between 32-row
// chunks along M/N.  (The plain SWIZZLE_128B layout, type 2, silently yields an all-zero product.)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ uint32_t umma_idesc_mn(uint32_t M, uint32_t N) {
  return umma_idesc(M, N) | (1u << 15) | (1u << 16);   // A and B both MN-major
}
// 16-bit operands (bf16): the ordinary SWIZZLE_128B MN-major layout -- 128-byte rows (64 bf16 along M/N), 16-byte
// chunks XOR-swizzled with (row & 7), atoms of 8 rows; a K = 16 instruction reads two atoms, 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn16(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D = fp32, A = B = BF16 (format 1), both MN-major
__device__ __forceinline__ uint32_t umma_idesc_mn16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- stream-K work partition -----------------------------------------------------------------------------
// The work of one factor is the list of (block pair q, box b) units in pair-major order.  Pairs differ in cost
// (a diagonal pair loads one block, the last row block may be short: one MMA per k-group instead of two), and
// pairs * splits rarely equals the SM count, so instead of "every pair is cut into the same number of equal splits"
// the host cuts the COST axis into G equal ranges, one per CTA (boundaries rounded to whole pipeline stages).  A CTA
// walks its range pair by pair; every (CTA, pair) segment accumulates in TMEM and is flushed to its own partial tile
// (slot = CTA + pair, unique), which the fixed-order reduction sums per pair in CTA order: still bit-reproducible.
constexpr int SK_MAXG = 160;
constexpr int SK_MAXP = 1024;
struct SkTable {
  int G;                      // CTAs
  int have_slots;             // plo / pn are filled (the launch has <= SK_MAXP pairs); else the reduction scans q / b
  uint16_t q[SK_MAXG + 1];    // boundary c = (pair q[c], box b[c]); strictly increasing; q[G] = pairs, b[G] = 0
  uint32_t b[SK_MAXG + 1];
  uint16_t plo[SK_MAXP];      // per pair: first CTA whose range intersects it ...
  uint16_t pn[SK_MAXP];       // ... and how many do (its partial tiles are slots plo + q .. plo + q + pn - 1)
};

struct SegGeom {
  int I, J, rowsA, mh, ncols, nchA, nchB, nslots, NB, nstage;
  int pk, poff, nbh;           // pack2: on; row / column offset of the second position set's block; boxes per set
  bool diag;
  uint32_t chunk_bytes, stage_bytes, plane_bytes;
};
template <int CH>
__device__ __forceinline__ SegGeom seg_geom(const NhParams& p, int q) {
  SegGeom t;
  decode_pair(q, p.T, t.I, t.J);
  t.diag = (t.I == t.J);
  t.rowsA = min(TB, p.D - t.I * TB);
  t.mh = (t.rowsA + 127) >> 7;
  t.ncols = t.diag ? ((t.rowsA + 15) & ~15) : TB;
  t.nchA = (t.rowsA + CH - 1) / CH;
  t.nchB = t.diag ? 0 : TB / CH;
  // A stage holds only the chunks that are really loaded: A chunks first, then B chunks.  An M = 128 descriptor
  // always spans 128 / CH chunks, so for a short A block it reads on into the B chunks (or, for a diagonal item,
  // past the stage into the next one / the ring's tail pad): those are accumulator rows >= rowsA, never stored.
  const int slotsA = t.mh * (128 / CH);
  t.nslots = t.nchA + t.nchB;
  t.NB = t.diag ? p.NBdiag : p.NBoff;
  t.pk = 0; t.poff = 0; t.nbh = t.NB;
  if (p.pack2) {
    // A factor of order <= 64 would use half (or a quarter) of the M = 128 rows of every MMA and is issue-bound at one
    // 2 KB k-group per ~70 ns.  Instead the boxes of a stage are split into two position sets whose chunks are stacked
    // along M: A = B = [set 1 channels | set 2 channels], so that ONE instruction contracts 2 x KPOS positions; the two
    // diagonal blocks of the accumulator are the two sets' partial sums (the reduction adds them), the cross blocks
    // are never read.
    t.pk = 1;
    t.poff = t.nchA * CH;
    t.nbh = t.NB / 2;
    t.nslots = 2 * t.nchA;
    t.ncols = 2 * t.poff;
    t.chunk_bytes = (uint32_t)(t.nbh * p.PB) * 128u;
    t.stage_bytes = (uint32_t)t.nslots * t.chunk_bytes;
    t.plane_bytes = t.stage_bytes;
    const uint32_t tail_pad = (uint32_t)(slotsA - t.nslots) * t.chunk_bytes;
    t.nstage = min(NH_MAXSTAGE, (int)((NH_DATA_BYTES - tail_pad) / t.stage_bytes));
    return t;
  }
  t.chunk_bytes = (uint32_t)(t.NB * p.PB) * 128u;
  // bf16x3: a stage holds the chunks of the hi plane, then the same chunks of the lo plane
  t.plane_bytes = (uint32_t)t.nslots * t.chunk_bytes;
  t.stage_bytes = (p.x3 ? 2u : 1u) * t.plane_bytes;
  const uint32_t tail_pad = t.diag ? (uint32_t)(slotsA - t.nchA) * t.chunk_bytes : 0u;
  t.nstage = min(NH_MAXSTAGE, (int)((NH_DATA_BYTES - tail_pad) / t.stage_bytes));
  return t;
}

// BF16 = false: fp32 words read as TF32, 32 channels per 128-byte row, 8 positions per MMA.
// BF16 = true : bf16 copy of the operand (made by the cast pre-pass), 64 channels per row, 16 positions per MMA:
//               half the bytes per operand element through L2 -> SM, twice the MMA rate.
// Warps: 0 and 6..12 issue TMA (8 issuers), 1 issues the MMAs and owns TMEM, 2..5 drain the accumulator.
template <bool BF16>
__global__ void __maxnreg__(96)
syrk_nhwc_kernel(const __grid_constant__ GroupParams gp, const __grid_constant__ SkTable sk,
                 const __grid_constant__ GroupMaps maps) {
  constexpr int CH = BF16 ? 64 : 32;        // operand rows (channels) per chunk = per 128-byte smem row
  constexpr int KPOS = BF16 ? 16 : 8;       // contraction positions per MMA instruction
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const uint32_t epi = sbase + NH_DATA_BYTES;                  // epilogue staging
  const uint32_t bars = epi + NH_EPI_BYTES;                    // full[8] | empty[8] | tmem_full | tmem_empty
  const uint32_t bar_tmem_full = bars + 8 * (2 * NH_MAXSTAGE);
  const uint32_t bar_tmem_empty = bar_tmem_full + 8;
  uint8_t* aux = smem_raw + (sbase - raw) + NH_DATA_BYTES + NH_EPI_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 8 * (2 * NH_MAXSTAGE + 2));
  int4* tab = reinterpret_cast<int4*>(aux + 256);              // per loaded chunk: {c0, dx, dy, slot}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  long long* tl = gp.tl ? gp.tl + 8 * (size_t)blockIdx.x : nullptr;
  if (tl && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[0] = (long long)gtimer(); tl[7] = smid;
  }
  trace_begin(gp.trace, 0);
  if (gp.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    gp.trace[6] = (unsigned long long)gp.f[0].D | ((unsigned long long)gp.nf << 20) | ((unsigned long long)(BF16 ? 1 : 0) << 30);
    gp.trace[7] = (unsigned long long)gp.qbeg[gp.nf];
  }
  // this CTA's range of the work list
  const int q0 = (int)sk.q[blockIdx.x], q1 = (int)sk.q[blockIdx.x + 1];
  const int bq0 = (int)sk.b[blockIdx.x], bq1 = (int)sk.b[blockIdx.x + 1];

  if (threadIdx.x == 0) {
    for (int s = 0; s < NH_MAXSTAGE; ++s) {
      mbar_init(bars + 8 * s, NH_NPROD);
      mbar_init(bars + 8 * (NH_MAXSTAGE + s), 1);
    }
    mbar_init(bar_tmem_full, 1);
    mbar_init(bar_tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tl && threadIdx.x == 0) tl[1] = (long long)gtimer();
  // factor of a global pair index (q0 <= q <= q1; factors are few: linear scan)
  auto factor_of = [&](int q) -> int {
    int f = 0;
    while (f + 1 < gp.nf && q >= gp.qbeg[f + 1]) ++f;
    return f;
  };

  if (warp == 0 || warp >= 6) {
    // ---- TMA producers.  Issue is spread over NH_NPROD warps (scripts/experiments/tma_rate_probe.cu: one issuer
    // sustains one cp.async.bulk.tensor per ~150 cycles whatever the box size; 4-8 issuers reach the TMA unit's
    // ~66 B/cycle/SM).  Each issuer arms the stage barrier with its own bytes.  The loop runs warp-uniformly, only
    // the mbarrier / TMA instructions are predicated on one elected lane.
    const uint32_t leader = elect_one();
    const int me = (int)uni((uint32_t)(warp == 0 ? 0 : warp - 5));     // 0 .. NH_NPROD-1
    const int ptid = me * 32 + lane;
    uint32_t pph = 0;                                                   // per-stage parity of the empty barriers
    int nseg = 0;
    for (int q = q0; q <= q1 && q < gp.qbeg[gp.nf]; ++q) {
      const int fi = factor_of(q);
      const NhParams& p = gp.f[fi];
      const CUtensorMap* tmap = &maps.m[fi];
      const CUtensorMap* tmap_lo = &maps.m[fi + 1];          // bf16x3 (always a group of one): the lo plane's map
      const int b_begin = q == q0 ? bq0 : 0, b_end = q == q1 ? bq1 : p.nbox;
      if (b_begin >= b_end) continue;
      const SegGeom t = seg_geom<CH>(p, q - gp.qbeg[fi]);
      const uint32_t box_bytes = (uint32_t)p.PBv * 128u;
      const int planes = (int)uni(p.x3 ? 2u : 1u);
      if (leader) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
      if (leader && planes == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap_lo)) : "memory");
      // the previous segment's MMAs have all retired (its accumulator is complete): every stage is free, whatever
      // the stage geometry of this segment is
      if (nseg > 0) mbar_wait(bar_tmem_full, (uint32_t)(nseg - 1) & 1u);
      asm volatile("bar.sync 1, %0;" ::"r"(NH_NPROD * 32) : "memory");  // nobody still reads the old chunk table
      const int loaded = t.pk ? t.nchA : t.nslots;           // chunks fetched per box
      if (me == 0 && lane < loaded) {
        const bool isB = lane >= t.nchA;
        const int kp = isB ? t.J * TB + (lane - t.nchA) * CH : t.I * TB + lane * CH;
        const int tap = (int)fdiv((uint32_t)kp, p.divC);
        const int ti = (int)fdiv((uint32_t)tap, p.divKW);
        tab[lane] = make_int4(kp - tap * p.C, p.flat ? 0 : (tap - ti * p.kw) - p.pw, p.flat ? 0 : ti - p.ph, lane);
      }
      if (p.PB > p.PBv) {   // zero the rows of every box slot that no box ever writes (TMA never touches them)
        const int pad16 = (p.PB - p.PBv) * 8;                       // 16-byte words per box slot
        const int total = t.nstage * t.nslots * t.NB * pad16 * planes;
        for (int e = ptid; e < total; e += NH_NPROD * 32) {
          const int slot = e / pad16, w = e - slot * pad16;
          const uint32_t a = sbase + (uint32_t)slot * (uint32_t)p.PB * 128u + (uint32_t)p.PBv * 128u + (uint32_t)w * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0) : "memory");
        }
        fence_proxy_async();
      }
      asm volatile("bar.sync 1, %0;" ::"r"(NH_NPROD * 32) : "memory");
      const int NB = (int)uni((uint32_t)t.NB), nstage = (int)uni((uint32_t)t.nstage), nld = (int)uni((uint32_t)loaded);
      const uint32_t stage_bytes = uni(t.stage_bytes), chunk_bytes = uni(t.chunk_bytes), plane_bytes = uni(t.plane_bytes);
      const int pk = (int)uni((uint32_t)t.pk), nbh = (int)uni((uint32_t)t.nbh), pk_nch = (int)uni((uint32_t)t.nchA);
      const int ub = (int)uni((uint32_t)b_begin), ue = (int)uni((uint32_t)b_end);
      const int nit = (ue - ub + NB - 1) / NB;
      int s = 0;
      for (int it = 0; it < nit; ++it) {
        const int b0 = ub + it * NB;
        const int nv = min(NB, ue - b0);
        const int per_plane = nv * nld;
        const int total = per_plane * planes;                   // TMA instructions of this stage
        const int mine = total > me ? (total - me + NH_NPROD - 1) / NH_NPROD : 0;
        mbar_wait(bars + 8 * (NH_MAXSTAGE + s), ((pph >> s) & 1u) ^ 1u);
        pph ^= 1u << s;
        const uint32_t st = sbase + (uint32_t)s * stage_bytes;
        if (pk && nv < NB) {
          // ragged last stage of a pack2 factor: the MMA still contracts both sets over all their box slots, so the slots
          // no box is loaded into must hold zeros (stale rows of an earlier stage would be summed)
          const int words = p.PB * 8;                            // 16-byte words per box slot
          const int tot = (NB - nv) * nld * words;
          for (int e = ptid; e < tot; e += NH_NPROD * 32) {
            const int bx = e / words, wd = e - bx * words;
            const int j = nv + bx / nld, qq = bx - (bx / nld) * nld;
            const int set = j >= nbh ? 1 : 0, jj = j - set * nbh;
            const uint32_t a = st + (uint32_t)(set * pk_nch + qq) * chunk_bytes + (uint32_t)(jj * p.PB) * 128u + (uint32_t)wd * 16u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0) : "memory");
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, %0;" ::"r"(NH_NPROD * 32) : "memory");
        }
        if (leader) {
          if (mine && !(gp.dbg & 1)) mbar_arrive_expect_tx(bars + 8 * s, (uint32_t)mine * box_bytes);
          else mbar_arrive(bars + 8 * s);
        }
        if (!(gp.dbg & 1)) {
          for (int e2 = me; e2 < total; e2 += NH_NPROD) {
            const int plane = e2 >= per_plane ? 1 : 0;
            const int e = e2 - plane * per_plane;
            const int j = e / nld, qq = e - j * nld;
            const uint32_t b = (uint32_t)(b0 + j);
            int X0, Y0, Nn;
            if (p.flat) {
              X0 = (int)b * p.PB; Y0 = 0; Nn = 0;
            } else {
              const uint32_t n = fdiv(b, p.divPPI);
              const uint32_t rem = b - n * (uint32_t)p.ppi;
              const uint32_t pr = fdiv(rem, p.divPCW);
              const uint32_t pc = rem - pr * (uint32_t)p.pcw;
              X0 = (int)pc * p.bw * p.sw; Y0 = (int)pr * p.bh * p.sh; Nn = (int)n * p.bn;
            }
            const int4 t4 = tab[qq];
            const int set = (pk && j >= nbh) ? 1 : 0;            // pack2: second half of the stage's boxes = second set
            const int jj = j - set * nbh, slot = t4.w + set * pk_nch;
            if (leader)
              tma_load_4d(st + (uint32_t)plane * plane_bytes + (uint32_t)(jj * p.PB) * 128u + (uint32_t)slot * chunk_bytes,
                          plane ? tmap_lo : tmap, t4.x, X0 + t4.y, Y0 + t4.z, Nn, bars + 8 * s);
          }
        }
        if (++s == nstage) s = 0;
      }
      ++nseg;
    }
  } else if (warp == 1) {
    // ---- MMA issuer.  The WHOLE warp runs this loop in uniform control flow and only the tcgen05 instructions
    // themselves are predicated on one elected lane: ptxas then keeps descriptors, addresses and the predicate in
    // uniform registers and the loop body is UIADD3 + UTCHMMA.  With `if (lane == 0) { loop }` it emitted an R2UR /
    // ELECT waterfall per instruction that took ~224 cycles per MMA against the tensor pipe's 128
    // (scripts/experiments/mma_rate_probe.cu).  Per-segment quantities come out of decode_pair / runtime divisions
    // in vector registers; a warp reduction (REDUX writes a uniform register) moves them to the uniform datapath.
    const uint32_t leader = elect_one();
    const uint32_t u_tmem = uni(tmem);
    constexpr uint32_t KSTEP = (uint32_t)((KPOS * 128) >> 4);             // one MMA's positions, in 16-byte units
    const bool run = !(gp.dbg & 2);
    uint32_t cph = 0;                                                     // per-stage parity of the full barriers
    int nseg = 0;
    for (int q = q0; q <= q1 && q < gp.qbeg[gp.nf]; ++q) {
      const int fi = factor_of(q);
      const NhParams& p = gp.f[fi];
      const int b_begin = q == q0 ? bq0 : 0, b_end = q == q1 ? bq1 : p.nbox;
      if (b_begin >= b_end) continue;
      const SegGeom t = seg_geom<CH>(p, q - gp.qbeg[fi]);
      const int kpb = (int)uni((uint32_t)(p.PB / KPOS));
      const uint32_t u_chunk = uni(t.chunk_bytes), u_stage = uni(t.stage_bytes);
      const int u_nstage = (int)uni((uint32_t)t.nstage), u_NB = (int)uni((uint32_t)t.NB);
      const int ub = (int)uni((uint32_t)b_begin), ue = (int)uni((uint32_t)b_end);
      const int u_nit = (ue - ub + u_NB - 1) / u_NB;
      const uint32_t idesc = uni(BF16 ? umma_idesc_mn16(128, (uint32_t)t.ncols) : umma_idesc_mn(128, (uint32_t)t.ncols));
      // diagonal tile of two row halves: only the lower triangle is ever read, so rows 0-127 need columns 0-127 only --
      // the first half's instruction runs with N = 128 (half the tensor-pipe time; a quarter of the tile's work saved)
      const uint32_t nc0 = uni((t.diag && t.mh == 2 && !(gp.dbg & 4)) ? 128u : (uint32_t)t.ncols);
      const uint32_t idesc0 = uni(BF16 ? umma_idesc_mn16(128, nc0) : umma_idesc_mn(128, nc0));
      const uint64_t dfull = (BF16 ? umma_desc_mn16(0u, u_chunk) : umma_desc_mn(0u, u_chunk));
      const uint32_t dlo = (uint32_t)dfull, dhi32 = (uint32_t)(dfull >> 32);
      const uint32_t hstep = (uint32_t)(128 / CH) * u_chunk;              // second 128-row half of the A block
      const uint32_t boff = uni(t.diag ? 0u : (uint32_t)t.nchA * t.chunk_bytes);
      const bool two = uni((uint32_t)t.mh) == 2u;
      const bool x3 = uni((uint32_t)p.x3) != 0u;
      const uint32_t lo16 = uni(t.plane_bytes) >> 4;                       // hi plane -> lo plane of a stage, in 16-byte units
      const int u_pk = (int)uni((uint32_t)t.pk), u_nbh = (int)uni((uint32_t)t.nbh);
      if (nseg > 0) {                                                     // accumulator drained by the epilogue warps
        mbar_wait(bar_tmem_empty, (uint32_t)(nseg - 1) & 1u);
        tc_fence_after();
      }
      uint32_t acc = 0;
      int s = 0;
      for (int it = 0; it < u_nit; ++it) {
        mbar_wait(bars + 8 * s, (cph >> s) & 1u);
        cph ^= 1u << s;
        tc_fence_after();
        if (tl && nseg == 0 && it == 0 && lane == 0) tl[2] = (long long)gtimer();
        if (tl && lane == 0) tl[6] += nv_kg(u_NB, ue - (ub + it * u_NB), kpb);
        const uint32_t st = sbase + (uint32_t)s * u_stage;
        const int nv = min(u_NB, ue - (ub + it * u_NB));
        const int nkg = run ? (u_pk ? u_nbh : nv) * kpb : 0;             // pack2: both sets in one instruction (ragged slots are zero)
        const uint32_t a0 = dlo | ((st >> 4) & 0x3FFFu);
        const uint32_t a1 = dlo | (((st + hstep) >> 4) & 0x3FFFu);
        const uint32_t b0 = dlo | (((st + boff) >> 4) & 0x3FFFu);
        if (leader) {
          if (BF16 && x3) {
            // x = hi + lo (two bf16 planes): X X^T ~= hi hi^T + hi lo^T + lo hi^T, three instructions per row half
            for (int kg = 0; kg < nkg; ++kg) {
              const uint32_t ko = (uint32_t)kg * KSTEP;
              tc_mma_lohi(true, u_tmem, a0 + ko, b0 + ko, dhi32, idesc0, acc);
              tc_mma_lohi(true, u_tmem, a0 + ko, b0 + lo16 + ko, dhi32, idesc0, 1u);
              tc_mma_lohi(true, u_tmem, a0 + lo16 + ko, b0 + ko, dhi32, idesc0, 1u);
              if (two) {
                tc_mma_lohi(true, u_tmem + 256u, a1 + ko, b0 + ko, dhi32, idesc, acc);
                tc_mma_lohi(true, u_tmem + 256u, a1 + ko, b0 + lo16 + ko, dhi32, idesc, 1u);
                tc_mma_lohi(true, u_tmem + 256u, a1 + lo16 + ko, b0 + ko, dhi32, idesc, 1u);
              }
              acc = 1;
            }
          } else {
            for (int kg = 0; kg < nkg; ++kg) {
              const uint32_t ko = (uint32_t)kg * KSTEP;
              tc_mma_lohi(BF16, u_tmem, a0 + ko, b0 + ko, dhi32, idesc0, acc);
              if (two) tc_mma_lohi(BF16, u_tmem + 256u, a1 + ko, b0 + ko, dhi32, idesc, acc);
              acc = 1;
            }
          }
        }
        if (nkg > 0) acc = 1;
        if (leader) tc_commit(bars + 8 * (NH_MAXSTAGE + s));
        if (++s == u_nstage) s = 0;
      }
      if (leader) tc_commit(bar_tmem_full);
      ++nseg;
      if (tl && lane == 0) { tl[3] = (long long)gtimer(); tl[5] = nseg; }
    }
    __syncwarp();
  } else {
    // ---- epilogue warps 2..5 (TMEM lane quadrant = warp & 3): accumulator -> registers -> partial tile `CTA + pair`
    int nseg = 0;
    for (int q = q0; q <= q1 && q < gp.qbeg[gp.nf]; ++q) {
      const int fi = factor_of(q);
      const NhParams& p = gp.f[fi];
      const int b_begin = q == q0 ? bq0 : 0, b_end = q == q1 ? bq1 : p.nbox;
      if (b_begin >= b_end) continue;
      const SegGeom g = seg_geom<CH>(p, q - gp.qbeg[fi]);
      ItemShape t;
      t.mh = g.mh; t.ncols = g.ncols; t.rowsA = g.pk ? g.poff + g.rowsA : g.rowsA;
      t.ncols0 = (g.diag && g.mh == 2 && !(gp.dbg & 4)) ? 128 : g.ncols;
      if (gp.dbg & 8)          // (A/B switch, CURVATURE_B200_DBG=8: flush through the TMA unit -- measured no faster here)
        epilogue_store_tma(t, bar_tmem_full, (uint32_t)nseg & 1u, tmem, &maps.part, (int)(blockIdx.x + q) * TB, warp & 3, lane,
                           epi + (uint32_t)(warp & 3) * 4096u);
      else
        epilogue_store_coalesced(t, bar_tmem_full, (uint32_t)nseg & 1u, tmem, gp.ws + (size_t)(blockIdx.x + q) * TILE_ELEMS,
                                 warp & 3, lane, epi + (uint32_t)(warp & 3) * 4096u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tmem_empty);
      ++nseg;
      if (tl && warp == 2 && lane == 0) tl[4] = (long long)gtimer();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // this warp's tensor stores have landed
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  trace_end(gp.trace, 0);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Fixed-order reduction for the stream-K partition: the partial tiles of global pair q are slots (c + q) for the
// CTAs c_lo..c_hi whose ranges intersect the pair, found from the boundary table; summed in CTA order, scaled,
// un-permuted and added (tile + mirror image) into the pair's factor.  One launch reduces every factor of a group.
__global__ void __launch_bounds__(256, 4) syrk_sk_reduce_kernel(const __grid_constant__ GroupParams gp, const __grid_constant__ SkTable sk) {
  // 256 threads / 4 KB of shared memory per CTA: small enough to be co-resident with a SYRK CTA of the next launch
  // (416 threads, ~211 KB), so that on the side stream the reduction really overlaps it.
  __shared__ float tile[32][33];
  __shared__ int s_lo, s_hi;
  TraceScope trace_scope(gp.trace, 2);
  const int q = blockIdx.x >> 6, sub = blockIdx.x & 63;
  int fi = 0;
  while (fi + 1 < gp.nf && q >= gp.qbeg[fi + 1]) ++fi;
  const NhParams& p = gp.f[fi];
  const int pair = q - gp.qbeg[fi];
  const int br = sub >> 3, bc = sub & 7;
  int I, J;
  decode_pair(pair, p.T, I, J);
  const bool diag = (I == J);
  if (diag && bc > br) return;            // diagonal blocks: lower triangle only, mirrored below (exact symmetry)
  const int rowsA = min(TB, p.D - I * TB);
  const int colsB = diag ? rowsA : TB;
  if (br * 32 >= rowsA || bc * 32 >= colsB) return;
  int c_lo, nsl;
  if (sk.have_slots) {
    c_lo = (int)sk.plo[q]; nsl = (int)sk.pn[q];
  } else {
    if (threadIdx.x == 0) { s_lo = 0; s_hi = 0; }
    __syncthreads();
    if ((int)threadIdx.x < sk.G) {
      const int c = threadIdx.x;
      const int cq = (int)sk.q[c];
      const uint32_t cb = sk.b[c];
      if (cq < q || (cq == q && cb == 0)) atomicMax(&s_lo, c);            // boundary(c) <= (q, 0)
      if (cq <= q) atomicMax(&s_hi, c);                                   // boundary(c) <  (q + 1, 0)
    }
    __syncthreads();
    c_lo = s_lo; nsl = s_hi - s_lo + 1;
  }
  const int lane = threadIdx.x & 31, w0 = threadIdx.x >> 5;
  const int D = p.ldF, K0 = p.K0, C = p.C, KK = p.KK;
  float* __restrict__ F = p.F;
  const float alpha = p.alpha;
  const int pk_c = p.pk_c, pk_kh = p.pk_kh, pk_kw = p.pk_kw;
  auto perm = [&](int kp) -> int {        // tap-major k' -> the reference's row index c*kh*kw + tap (-1: padding row)
    if (pk_c) {                           // packed path: k' = ip*64 + i2*32 + j*4 + c, filter row i = 2*ip + i2
      const int i = 2 * (kp >> 6) + ((kp >> 5) & 1), j = (kp >> 2) & 7, c = kp & 3;
      return (i < pk_kh && j < pk_kw && c < pk_c) ? (c * pk_kh + i) * pk_kw + j : -1;
    }
    if (kp >= K0) return kp;
    const int t = (int)fdiv((uint32_t)kp, p.divC);
    const int c = kp - t * C;
    return c * KK + t;
  };
  const float* __restrict__ base = gp.ws + (size_t)(c_lo + q) * TILE_ELEMS;
  const int col = bc * 32 + lane;
  const int pcol = col < colsB ? perm(J * TB + col) : -1;
  // pack2 factors: the partial tile holds two diagonal blocks (one per position set), `poff` rows / columns apart
  const size_t second = p.pack2 ? (size_t)((p.D + 31) / 32 * 32) * (TB + 1) : 0;
  // the thread's four rows are summed together: 4 x 8 (pack2: 4 x 4 x 2) independent loads in flight per round, each
  // row's partial tiles still added in slot order (the reduction is a latency chain of nsl / 8 round trips)
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  const float* bp[4];
  bool valid[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = br * 32 + w0 + 8 * k;
    valid[k] = row < rowsA && col < colsB;
    bp[k] = base + (valid[k] ? row * TB + col : 0);
  }
  if (second) {
    int s = 0;
    for (; s + 4 <= nsl; s += 4) {
      float t[4][4], t2[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          t[k][u] = valid[k] ? __ldcg(bp[k] + (size_t)(s + u) * TILE_ELEMS) : 0.f;
          t2[k][u] = valid[k] ? __ldcg(bp[k] + (size_t)(s + u) * TILE_ELEMS + second) : 0.f;
        }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < 4; ++u) sum[k] += t[k][u] + t2[k][u];
    }
    for (; s < nsl; ++s)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (valid[k]) sum[k] += __ldcg(bp[k] + (size_t)s * TILE_ELEMS) + __ldcg(bp[k] + (size_t)s * TILE_ELEMS + second);
  } else {
    int s = 0;
    for (; s + 8 <= nsl; s += 8) {
      float t[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < 8; ++u) t[k][u] = valid[k] ? __ldcg(bp[k] + (size_t)(s + u) * TILE_ELEMS) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < 8; ++u) sum[k] += t[k][u];
    }
    for (; s < nsl; ++s)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (valid[k]) sum[k] += __ldcg(bp[k] + (size_t)s * TILE_ELEMS);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int w = w0 + 8 * k;
    const int row = br * 32 + w;
    float v = 0.f;
    if (valid[k]) {
      v = alpha * sum[k];
      const int pr = perm(I * TB + row);
      if (!(diag && col > row) && pr >= 0 && pcol >= 0) F[(size_t)pr * D + pcol] += v;
    }
    tile[w][lane] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {           // mirror image: lanes run along the original rows
    const int w = w0 + 8 * k;
    const int colm = bc * 32 + w, rowm = br * 32 + lane;
    if (rowm < rowsA && colm < colsB && !(diag && colm >= rowm)) {
      const int pr = perm(J * TB + colm), pc = perm(I * TB + rowm);
      if (pr >= 0 && pc >= 0) F[(size_t)pr * D + pc] += tile[lane][w];
    }
  }
}

// Reduction for k x k convolution factors (KK = kh*kw taps, 2 <= KK <= 9, one factor per launch).  The generic kernel
// above scatters every element of a tile to F[c1*KK + t1][c2*KK + t2] -- KK floats apart in both directions, one 32-byte
// sector per 4-byte element, which made the 4608^2 factor's reduction cost as much as its contraction.  Here one CTA
// owns the F rows (32 channels c1, one tap t1) x the 32*KK CONTIGUOUS columns of 32 channels c2 and all taps t2: it
// gathers the KK sub-tiles (one per t2, each from its own block pair, transposed where only the mirror pair was
// computed), sums their partial tiles in the same fixed CTA order, interleaves them in shared memory and adds whole
// 32*KK-float row segments to F.  Same sums, same order, same (exactly symmetric) result.
__global__ void __launch_bounds__(256, 4) syrk_sk_reduce_taps_kernel(const __grid_constant__ GroupParams gp, const __grid_constant__ SkTable sk) {
  extern __shared__ float outbuf[];                  // [8][KK*32 + 1]
  __shared__ int s_lo[9], s_hi[9];
  TraceScope trace_scope(gp.trace, 2);
  const NhParams& p = gp.f[0];
  const int C = p.C, KK = p.KK, T = p.T, D = p.ldF;
  const int ncb = C >> 5;
  const int cb2 = blockIdx.x % ncb;                  // 32 channels c2 (columns)
  const int t1 = (blockIdx.x / ncb) % KK;
  const int rb = blockIdx.x / (ncb * KK);            // 8 channels c1 (rows)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int pitch = KK * 32 + 1;
  const int noff = T * (T - 1) / 2;
  const int k1 = t1 * C + rb * 8;
  const int I1 = k1 >> 8, r1 = k1 & 255;
  auto pair_of = [&](int t2, int& I2, int& r2) -> int {
    const int k2 = t2 * C + cb2 * 32;
    I2 = k2 >> 8; r2 = k2 & 255;
    const int hi = max(I1, I2), lo = min(I1, I2);
    return hi == lo ? noff + hi : hi * (hi - 1) / 2 + lo;
  };
  if (!sk.have_slots) {
    if (tid < 9) { s_lo[tid] = 0; s_hi[tid] = 0; }
    __syncthreads();
    for (int t2 = w; t2 < KK; t2 += 8) {             // one warp per tap: slots of its pair from the boundary table
      int I2, r2;
      const int q = pair_of(t2, I2, r2);
      int lo = 0, hi = 0;
      for (int c = lane; c < sk.G; c += 32) {
        const int cq = (int)sk.q[c];
        if (cq < q || (cq == q && sk.b[c] == 0)) lo = c;
        if (cq <= q) hi = c;
      }
      lo = __reduce_max_sync(0xffffffffu, lo);
      hi = __reduce_max_sync(0xffffffffu, hi);
      if (lane == 0) { s_lo[t2] = lo; s_hi[t2] = hi; }
    }
    __syncthreads();
  }
  // per tap: where this thread's element lives and how many partial tiles it has; then all taps' loads of one slot
  // index are issued together (KK independent loads in flight per round instead of a chain of KK * nsl round trips)
  const float* src[9];
  int nsl[9], oidx[9];
  int nmax = 0;
#pragma unroll
  for (int t2 = 0; t2 < 9; ++t2) {
    src[t2] = gp.ws; nsl[t2] = 0; oidx[t2] = 0;
    if (t2 < KK) {
      int I2, r2;
      const int q = pair_of(t2, I2, r2);
      // (a, b) = (c1, c2) offsets of this thread's element; the stored element is [max][min] of the two k' indices
      // (block-wise, then within the diagonal block), read along the stored rows so that the loads coalesce
      int a, b, sr, sc;
      if (I1 > I2 || (I1 == I2 && r1 >= r2 + 32)) { a = tid >> 5; b = tid & 31; sr = r1 + a; sc = r2 + b; }
      else if (I1 < I2 || r1 + 8 <= r2) { b = tid >> 3; a = tid & 7; sr = r2 + b; sc = r1 + a; }
      else {
        a = tid >> 5; b = tid & 31;
        const int ka = r1 + a, kb = r2 + b;
        sr = max(ka, kb); sc = min(ka, kb);
      }
      const int c_lo = sk.have_slots ? (int)sk.plo[q] : s_lo[t2];
      nsl[t2] = sk.have_slots ? (int)sk.pn[q] : s_hi[t2] - c_lo + 1;
      nmax = max(nmax, nsl[t2]);
      src[t2] = gp.ws + (size_t)(c_lo + q) * TILE_ELEMS + sr * TB + sc;
      oidx[t2] = a * pitch + b * KK + t2;
    }
  }
  float sum[9];
#pragma unroll
  for (int t2 = 0; t2 < 9; ++t2) sum[t2] = 0.f;
  for (int sl = 0; sl < nmax; ++sl) {
    float v[9];
#pragma unroll
    for (int t2 = 0; t2 < 9; ++t2) v[t2] = sl < nsl[t2] ? __ldcg(src[t2] + (size_t)sl * TILE_ELEMS) : 0.f;
#pragma unroll
    for (int t2 = 0; t2 < 9; ++t2) sum[t2] += v[t2];          // (+0 for taps that have fewer partial tiles: exact)
  }
#pragma unroll
  for (int t2 = 0; t2 < 9; ++t2)
    if (t2 < KK) outbuf[oidx[t2]] = sum[t2];
  __syncthreads();
  const float alpha = p.alpha;
  float* __restrict__ Frow = p.F + (size_t)((rb * 8 + w) * KK + t1) * D + (size_t)cb2 * 32 * KK;
  for (int k = lane; k < KK * 32; k += 32) Frow[k] += alpha * outbuf[w * pitch + k];
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi), both round-to-nearest-even: 16 significand bits of the fp32 value
// survive (|x - hi - lo| <= 2^-16 |x|), which is what the bf16x3 tier contracts with three bf16 MMAs per k-group.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));          // low half = first element
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// Pack pre-pass of the small-C path: Q[n][r][ow][i2*32 + j*4 + c] = bf16(x[n][c][2r + i2 - ph][ow*sw + j - pw]) (0 outside
// the image, for j >= kw and for c >= C).  One thread writes the 64 bytes of one (n, r, ow, i2); x is addressed through
// element strides, so NCHW-dense and channels-last inputs both work.
__global__ void __launch_bounds__(256) pack_smallc_kernel(const float* __restrict__ x, uint4* __restrict__ Q, uint4* __restrict__ Qlo,
                                                          int N, int C, int H, int W,
                                                          long long sN, long long sC, long long sH, long long sW, int Hq, int OW,
                                                          int kw, int sw, int ph, int pw, unsigned long long* tr) {
  TraceScope trace_scope(tr, 4);
  const long long total = (long long)N * Hq * OW * 2;
  {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;      // one 64-byte group per thread: short-lived blocks
    if (t >= total) return;
    const int i2 = (int)(t & 1);
    long long u = t >> 1;
    const int ow = (int)(u % OW); u /= OW;
    const int r = (int)(u % Hq);
    const int n = (int)(u / Hq);
    const int h = 2 * r + i2 - ph;
    uint32_t o[16], ol[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { o[k] = 0u; ol[k] = 0u; }
    if (h >= 0 && h < H) {
      const float* __restrict__ row = x + n * sN + h * sH;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int w = ow * sw + j - pw;
        if (j < kw && w >= 0 && w < W) {
          float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < C) v[c] = __ldg(row + c * sC + w * sW);
          split_bf16x2(v[0], v[1], o[2 * j], ol[2 * j]);
          split_bf16x2(v[2], v[3], o[2 * j + 1], ol[2 * j + 1]);
        }
      }
    }
    uint4* dst = Q + t * 4;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
    dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
    if (Qlo) {                                  // bf16x3 tier: the second plane
      uint4* dl = Qlo + t * 4;
      dl[0] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      dl[1] = make_uint4(ol[4], ol[5], ol[6], ol[7]);
      dl[2] = make_uint4(ol[8], ol[9], ol[10], ol[11]);
      dl[3] = make_uint4(ol[12], ol[13], ol[14], ol[15]);
    }
  }
}

// Pre-pass kernels are SHORT-LIVED blocks (256 threads x 4 float4 each, no grid-stride loop): they run on a low-priority
// stream beside the contraction kernels, and a contraction CTA (416 threads, ~211 KB of shared memory) can only be placed
// on an SM once enough resident pre-pass CTAs have exited.  Long-lived grid-stride blocks (8 x 256 threads per SM until the
// whole copy is done) kept every SM full and delayed the next contraction kernel by the rest of the copy (15-50 us gaps in
// the launch trace); with 16 KB blocks an SM frees up within a microsecond or two.
constexpr int PRE_V4 = 4;                     // float4 per thread

// out[i] = round-to-nearest TF32 of in[i] (layout preserved): the rounding pre-pass of the `tf32` tier.
__global__ void __launch_bounds__(256) round_tf32_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4, unsigned long long* tr) {
  TraceScope trace_scope(tr, 4);
  const size_t i0 = (size_t)blockIdx.x * (256 * PRE_V4) + threadIdx.x;
  float4 v[PRE_V4];
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4) v[u] = __ldg(in + i);
  }
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4) {
      float4 o;
      o.x = __uint_as_float(cvt_tf32(v[u].x)); o.y = __uint_as_float(cvt_tf32(v[u].y));
      o.z = __uint_as_float(cvt_tf32(v[u].z)); o.w = __uint_as_float(cvt_tf32(v[u].w));
      out[i] = o;
    }
  }
}

// out[i] = bf16(in[i]) (round to nearest even, layout preserved): the pre-pass of the `bf16` tier.
// The fp32 source is read exactly once (L2 evict-first), the bf16 copy is about to be read several times by the
// contraction kernel (L2 evict-last): without the hints the streaming reads push most of the freshly written copy out
// to DRAM before the contraction starts (ncu: 76 of 103 MB written back during the cast, then read again from DRAM).
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4, int hints, unsigned long long* tr) {
  TraceScope trace_scope(tr, 4);
  uint64_t pol_first = 0, pol_last = 0;
  if (hints) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  }
  const size_t i0 = (size_t)blockIdx.x * (256 * PRE_V4) + threadIdx.x;
  float4 v[PRE_V4];
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4) {
      if (hints)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(in + i), "l"(pol_first));
      else
        v[u] = __ldg(in + i);
    }
  }
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4) {
      uint2 o;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(v[u].y), "f"(v[u].x));   // low half = first element
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(v[u].w), "f"(v[u].z));
      if (hints)
        asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(out + i), "r"(o.x), "r"(o.y), "l"(pol_last) : "memory");
      else
        out[i] = o;
    }
  }
}

// Split pre-pass of the bf16x3 tier for a channels-last (or 2-D) operand: same layout, two bf16 planes.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                         size_t n4, unsigned long long* tr) {
  TraceScope trace_scope(tr, 4);
  const size_t i0 = (size_t)blockIdx.x * (256 * PRE_V4) + threadIdx.x;
  float4 v[PRE_V4];
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(in + i));
  }
#pragma unroll
  for (int u = 0; u < PRE_V4; ++u) {
    const size_t i = i0 + (size_t)u * 256;
    if (i < n4) {
      uint2 h, l;
      split_bf16x2(v[u].x, v[u].y, h.x, l.x);
      split_bf16x2(v[u].z, v[u].w, h.y, l.y);
      hi[i] = h;
      lo[i] = l;
    }
  }
}

// Layout-normalising pre-pass: an NCHW-dense fp32 operand [N][C][HW] becomes the channels-last bf16 copy [N][HW][C] the
// TMA-fed kernel reads (plus the lo plane on the bf16x3 tier).  One 64-channel x 64-position tile per CTA through shared
// memory: reads are 256-byte runs along HW, writes 128-byte runs along C.
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, uint32_t* __restrict__ hi,
                                                                uint32_t* __restrict__ lo, int C, int HW, int ctiles, int ptiles,
                                                                unsigned long long* tr) {
  __shared__ float t[64][65];
  TraceScope trace_scope(tr, 4);
  const int pt = blockIdx.x % ptiles;
  const int ct = (blockIdx.x / ptiles) % ctiles;
  const int n = blockIdx.x / (ptiles * ctiles);
  const int c0 = ct * 64, p0 = pt * 64;
  const float* __restrict__ src = x + ((size_t)n * C + c0) * (size_t)HW + p0;
  {
    const int pp = threadIdx.x & 63, cc0 = threadIdx.x >> 6;
    const bool pv = p0 + pp < HW;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int cc = cc0 + 4 * i;
      float v = 0.f;
      if (pv && c0 + cc < C) v = __ldg(src + (size_t)cc * HW + pp);
      t[cc][pp] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ch = 2 * lane;                                  // this thread's channel pair (C % 8 == 0: pairs never straddle C)
  if (c0 + ch < C) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pp = w + 8 * i;
      if (p0 + pp < HW) {
        const size_t o = (((size_t)n * HW + p0 + pp) * (size_t)C + c0 + ch) >> 1;
        uint32_t h, l;
        split_bf16x2(t[ch][pp], t[ch + 1][pp], h, l);
        hi[o] = h;
        if (lo) lo[o] = l;
      }
    }
  }
}

struct NhPlan {
  NhParams p;
  int pairs;
  int bf16;                      // operands go through the bf16 copy
  int pack;                      // packed small-C path: the operand the kernel sees is Q (geometry gq), made by the pack pre-pass
  int transpose;                 // the source is NCHW-dense: the pre-pass also transposes it to channels-last
  ConvGeom gq;
  size_t partial_bytes, copy_bytes;
  size_t plane_bytes;            // bytes of one bf16 plane of the copy (copy_bytes = planes * plane_bytes)
};

bool nhwc_plan(const ConvGeom& g, int precision, int sms, NhPlan& pl, bool src_is_bf16 = false);

// Small-C convolutions (the ResNet stem): see the header, "packed small-C path".
bool packable(const ConvGeom& g, int precision) {
  return (precision == CRV_PREC_BF16 || precision == CRV_PREC_BF16X3) && !g.has_bias && g.C <= 4 && g.kw <= 8 && g.sh == 2 &&
         g.kh >= 2 && g.kh <= 16 &&
         g.sw <= 8 && g.R < (1LL << 31) - 512;
}
bool pack_plan(const ConvGeom& g, int precision, int sms, NhPlan& pl) {
  const int nip = (g.kh + 1) / 2;
  ConvGeom gq;
  static const float dummy = 0.f;
  if (make_geom(gq, &dummy, g.N, 64, g.OH + nip - 1, g.OW, nip, 1, 1, 1, 0, 0, 0)) return false;
  gq.x = nullptr;
  if (!nhwc_plan(gq, precision, sms, pl, true) || !pl.bf16) return false;
  pl.pack = 1;
  pl.gq = gq;
  pl.p.ldF = g.D;
  pl.p.pk_kh = g.kh; pl.p.pk_kw = g.kw; pl.p.pk_c = g.C;
  pl.plane_bytes = ((size_t)gq.N * gq.H * gq.W * 64 * 2 + 1023) & ~(size_t)1023;
  pl.copy_bytes = (pl.p.x3 ? 2 : 1) * pl.plane_bytes;
  return true;
}

// operands the bf16 kernel instance can take: whole 16-byte channel groups, whole 64-channel chunks per filter tap
static bool bf16_geom_ok(const ConvGeom& g) {
  return g.C >= 64 && (g.C & 7) == 0 && (g.kh * g.kw == 1 || (g.C & 63) == 0) && !g.has_bias;
}

bool nhwc_plan(const ConvGeom& g, int precision, int sms, NhPlan& pl, bool src_is_bf16) {
  const int KK = g.kh * g.kw;
  pl.pack = 0;
  pl.transpose = 0;
  if (!src_is_bf16 && packable(g, precision)) return pack_plan(g, precision, sms, pl);
  const bool bf16_geom = bf16_geom_ok(g);
  // two bf16 planes: the bf16x3 tier, and -- inside the bf16 tier -- zero-mean operands with few contraction rows per
  // factor row (crv_syrk_item::zero_mean), whose single-plane error 2^-9 sqrt(2 D / R) would exceed the stated 1e-3
  const bool x3 = bf16_geom_ok(g) &&
                  (precision == CRV_PREC_BF16X3 || (!src_is_bf16 && precision == CRV_PREC_BF16 && g.zero_mean && g.R < 5LL * g.D));
  if (precision == CRV_PREC_BF16X3 && !x3) return false;
  // an NCHW-dense source is taken when a bf16 copy is made anyway: the pre-pass then transposes it to channels-last
  if (g.x_nchw && !(bf16_geom && (precision == CRV_PREC_BF16 || x3))) return false;
  if (g.has_bias || g.C < 32 || (g.C & 3) != 0) return false;
  if (KK > 1 && (g.C & 31) != 0) return false;
  if (((uintptr_t)g.x & 15) != 0) return false;
  if (g.sh > 8 || g.sw > 8) return false;
  if (g.R >= (1LL << 31) - 512) return false;
  NhParams& p = pl.p;
  p.D = g.D; p.C = g.C; p.KK = KK; p.kw = g.kw; p.K0 = g.K0; p.alpha = 0.f; p.F = nullptr;
  p.ldF = g.D; p.pk_kh = p.pk_kw = p.pk_c = 0; p.pack2 = 0; p.x3 = x3 ? 1 : 0;
  p.divC = make_fastdiv((uint32_t)g.C);
  p.divKW = make_fastdiv((uint32_t)g.kw);
  p.T = (g.D + TB - 1) / TB;
  pl.pairs = p.T * (p.T + 1) / 2;
  // The bf16 copy costs one extra pass over the tensor (read 4 B, write 2 B per element): it pays when every
  // element then travels L2 -> SM several times -- k x k convolutions (each element feeds kh*kw operand rows) and
  // factors of three or more row blocks.  Single-tile, read-once operands are HBM-bound and stay on the direct path.
  static const int bf16_min_t = getenv("CURVATURE_B200_BF16_MIN_T") ? atoi(getenv("CURVATURE_B200_BF16_MIN_T")) : 3;
  pl.bf16 = (src_is_bf16 || x3 || (precision == CRV_PREC_BF16 && (KK > 1 || p.T >= bf16_min_t || g.x_nchw) && bf16_geom)) ? 1 : 0;
  pl.transpose = (g.x_nchw && !src_is_bf16) ? 1 : 0;
  const int CH = pl.bf16 ? 64 : 32, gran = pl.bf16 ? 16 : 8;
  const int planes = x3 ? 2 : 1;
  p.sh = g.sh; p.sw = g.sw; p.ph = g.ph; p.pw = g.pw;
  p.flat = (KK == 1 && g.sh == 1 && g.sw == 1 && g.ph == 0 && g.pw == 0) ? 1 : 0;
  // chunk slots per stage of the two item kinds
  const int slots_diag = ((p.T > 1 ? TB : g.D) + CH - 1) / CH, slots_off = 2 * (TB / CH);   // chunks loaded per box
  const int slots_max = p.T > 1 ? slots_off : slots_diag;
  // positions per stage at the target stage size (single-tile factors stream from HBM: smaller stages, more of them)
  const int stage_target = p.T > 1 ? NH_STAGE_TARGET : NH_STAGE_TARGET / 2;   // single-tile factors stream from HBM
  const int pcap = stage_target / (slots_max * planes * 128);
  if (p.flat) {
    long long pb = (pcap < 256 ? pcap : 256) / gran * gran;     // whole MMA k-groups
    if (pb < gran) return false;
    static const bool pack2_on = !(getenv("CURVATURE_B200_PACK2") && atoi(getenv("CURVATURE_B200_PACK2")) == 0);
    if (pack2_on && !pl.bf16 && !src_is_bf16 && p.T == 1 && g.D <= 64 && pb >= 2 * gran && g.R >= 4 * pb) {
      p.pack2 = 1;                                               // two boxes per stage, one per position set
      pb = (pb / 2) / gran * gran;
    }
    const long long rr = (g.R + gran - 1) / gran * gran;
    if (pb > rr) pb = rr;
    p.PB = p.PBv = (int)pb;
    p.bw = (int)pb; p.bh = 1; p.bn = 1; p.pcw = 1; p.ppi = 1;
    p.nbox = (int)((g.R + pb - 1) / pb);
  } else {
    // box = bw x bh output positions of bn consecutive images, bw | OW, bh | OH (no box overhangs the output grid: an
    // overhanging position would read real pixels through a negative tap shift; an overhanging IMAGE is all zero
    // fill and harmless), padded to `gran` rows.  The image extent is what makes 14x14 / 7x7 / 28x28 maps fill
    // whole MMA k-groups (2x2x16, 1x1x64, 4x4x4 positions) with boxes of several KB instead of 87 % / 77 % fill or
    // 2 KB boxes.  Highest fill efficiency first; among equals the largest box that fits the target stage.
    const int pcap_hi = pcap;
    int best_bw = 0, best_bh = 0, best_bn = 1, best_pb = 0;
    double best_eff = -1.0;
    for (int bn = 1; bn <= 64; bn *= 2) {
      if (bn > 1 && bn > g.N) break;
      const double effn = (double)g.N / (double)((g.N + bn - 1) / bn * bn);
      for (int bw = 1; bw <= g.OW; ++bw) {
        if (g.OW % bw || bw * g.sw > 256) continue;
        for (int bh = 1; bh <= g.OH; ++bh) {
          if (g.OH % bh || bh * g.sh > 256) continue;
          const int pbv = bw * bh * bn, pb = (pbv + gran - 1) / gran * gran;
          if (pb > pcap_hi) continue;
          const double eff = (double)pbv / pb * effn;
          bool better;
          if (eff > best_eff + 1e-9) better = true;
          else if (eff < best_eff - 1e-9) better = false;
          else better = pb > best_pb || (pb == best_pb && (bn < best_bn || (bn == best_bn && bw > best_bw)));
          if (better) { best_eff = eff; best_bw = bw; best_bh = bh; best_bn = bn; best_pb = pb; }
        }
      }
    }
    if (best_bw == 0) return false;
    p.bw = best_bw; p.bh = best_bh; p.bn = best_bn; p.PBv = best_bw * best_bh * best_bn; p.PB = best_pb;
    p.pcw = g.OW / best_bw;
    p.ppi = p.pcw * (g.OH / best_bh);
    const long long nb = (long long)((g.N + best_bn - 1) / best_bn) * p.ppi;
    if (nb >= (1LL << 30)) return false;
    p.nbox = (int)nb;
  }
  p.divPPI = make_fastdiv((uint32_t)p.ppi);
  p.divPCW = make_fastdiv((uint32_t)p.pcw);
  auto boxes_per_stage = [&](int slots) {
    int nb = stage_target / (slots * planes * p.PB * 128);
    return nb < 1 ? 1 : nb;
  };
  p.NBoff = boxes_per_stage(slots_off);
  p.NBdiag = boxes_per_stage(slots_diag);
  if (p.pack2) {
    p.NBdiag = p.NBdiag / 2 * 2;
    if (p.NBdiag < 2) p.pack2 = 0;
  }
  if (p.T > 1) p.NBdiag = p.NBoff * (p.NBdiag / p.NBoff > 0 ? p.NBdiag / p.NBoff : 1);
  if (slots_max * planes * p.NBoff * p.PB * 128 * 2 > NH_DATA_BYTES && p.T > 1) return false;   // needs >= 2 stages
  if ((slots_diag * planes * 2 + 3) * p.NBdiag * p.PB * 128 > NH_DATA_BYTES) return false;      // (+ tail pad)
  p.bps = 0; p.splits = 0;                                   // (stream-K: the partition lives in the SkTable)
  pl.partial_bytes = (size_t)(SK_MAXG + pl.pairs) * TILE_ELEMS * sizeof(float);
  const size_t numel = (size_t)g.N * g.C * g.H * g.W;
  pl.plane_bytes = pl.bf16 ? ((numel * 2 + 1023) & ~(size_t)1023) : 0;
  pl.copy_bytes = pl.bf16 ? (size_t)planes * pl.plane_bytes
                          : (precision == CRV_PREC_TF32 ? ((numel * 4 + 1023) & ~(size_t)1023) : 0);
  return true;
}

void host_decode_pair(int pair, int T, int& I, int& J) {   // mirror of decode_pair
  const int noff = T * (T - 1) / 2;
  if (pair < noff) {
    int i = 1;
    while ((i + 1) * i / 2 <= pair) ++i;
    I = i; J = pair - i * (i - 1) / 2;
  } else {
    I = J = pair - noff;
  }
}

// Cut the cost axis of a group of factors into at most `sms` equal ranges (see SkTable).  Cost of one k-group of a
// pair, in ns = max(MMA issue/pipe time, operand bytes / beta, and -- for single-pair factors, which are read exactly
// once -- operand bytes / the HBM share of one SM); beta = L2 -> SM bytes per cycle per SM (the TMA side delivers
// ~14 TB/s = 50 B/cycle/SM with boxes of >= 8 KB, so with the default the MMA term decides for re-read operands).
void build_sk(const std::vector<const NhPlan*>& pls, int sms, GroupParams& gp, SkTable& sk) {
  static const double beta = getenv("CURVATURE_B200_SK_BETA") ? atof(getenv("CURVATURE_B200_SK_BETA")) : 64.0;
  static const double hbm = getenv("CURVATURE_B200_SK_HBM") ? atof(getenv("CURVATURE_B200_SK_HBM")) : 42.0;   // B/ns per SM
  std::vector<double> cbox;          // cost of one box of global pair q
  std::vector<long long> nbq, nboxq;
  long long iters = 0;
  gp.qbeg[0] = 0;
  for (size_t f = 0; f < pls.size(); ++f) {
    const NhPlan& pl = *pls[f];
    const NhParams& p = pl.p;
    const int CH = pl.bf16 ? 64 : 32, KPOS = pl.bf16 ? 16 : 8;
    for (int q = 0; q < pl.pairs; ++q) {
      int I, J;
      host_decode_pair(q, p.T, I, J);
      const bool diag = I == J;
      const int rowsA = std::min(TB, p.D - I * TB);
      const int mh = (rowsA + 127) >> 7;
      const int ncols = diag ? ((rowsA + 15) & ~15) : TB;
      const int nslots = (rowsA + CH - 1) / CH + (diag ? 0 : TB / CH);
      // measured with the per-CTA timeline (crv_debug_timeline), ns per k-group at ~1.85 GHz: two MMAs of N = 256:
      // 150; one MMA: 100 / 80 / 67 at N = 256 / 128 / 64 (one instruction per k-group is issue-bound, not pipe-bound)
      double mma = mh == 2 ? 150.0 * std::max(ncols / 256.0, 0.5) * (diag ? 0.75 : 1.0) : 56.0 + 0.17 * ncols;
      double bytes = (double)nslots * KPOS * 128;
      if (p.pack2) {      // one instruction per 2 x KPOS positions (N = twice the padded order): cost per KPOS positions
        const int n2 = 2 * ((rowsA + CH - 1) / CH) * CH;
        mma = (56.0 + 0.17 * n2) / 2.0;
      }
      if (p.x3) { mma *= 3.0; bytes *= 2.0; }
      double c = std::max(mma, bytes / beta / 1.85);
      if ((pls.size() > 1 && pl.copy_bytes == 0) || p.T == 1) {
        // read-once operand: HBM time.  An off-diagonal pair of a two-block factor finds about half of its second block
        // in L2 (the diagonal pairs of the same factor stream it at the same time on neighbouring CTAs): measured 248 ns
        // per 16 KB k-group against 196 ns per 8 KB one (per-CTA timeline of the ResNet-50 group launch).
        static const double offdiag_share = getenv("CURVATURE_B200_SK_OFFDIAG") ? atof(getenv("CURVATURE_B200_SK_OFFDIAG")) : 0.57;
        c = std::max(c, bytes * (diag ? 1.0 : offdiag_share) / hbm);
      }
      cbox.push_back(c * (p.PB / KPOS));
      nbq.push_back(diag ? p.NBdiag : p.NBoff);
      nboxq.push_back(p.nbox);
      iters += (p.nbox + nbq.back() - 1) / nbq.back();
    }
    gp.qbeg[f + 1] = gp.qbeg[f] + pl.pairs;
  }
  const int P = (int)cbox.size();
  std::vector<double> pre(P + 1, 0.0);
  for (int q = 0; q < P; ++q) pre[q + 1] = pre[q] + cbox[q] * (double)nboxq[q];
  int G = sms < SK_MAXG ? sms : SK_MAXG;
  if ((long long)G > iters) G = (int)iters;
  if (G < 1) G = 1;
  const double Wt = pre[P];
  std::vector<std::pair<int, long long>> bd;
  bd.push_back({0, 0});
  int q = 0;
  for (int c = 1; c < G; ++c) {
    const double x = Wt * c / G;
    while (q + 1 < P && pre[q + 1] <= x) ++q;
    int qq = q;
    const double box = (x - pre[qq]) / cbox[qq];
    const long long nb = nbq[qq], nbox = nboxq[qq];
    long long b = (long long)std::llround(box / nb) * nb;
    // a sliver at either end of a pair costs a pipeline ramp and an accumulator flush: snap it to the pair boundary
    const long long its = (nbox + nb - 1) / nb, snap = std::min<long long>(8, its / 4) * nb;
    if (b < snap) b = 0;
    if (b > nbox - snap || b >= nbox) { ++qq; b = 0; }
    if (qq > bd.back().first || (qq == bd.back().first && b > bd.back().second)) bd.push_back({qq, b});
  }
  if (bd.back().first >= P) bd.pop_back();
  sk.G = (int)bd.size();
  for (int c = 0; c < sk.G; ++c) { sk.q[c] = (uint16_t)bd[c].first; sk.b[c] = (uint32_t)bd[c].second; }
  sk.q[sk.G] = (uint16_t)P; sk.b[sk.G] = 0;
  static const bool no_slots = getenv("CURVATURE_B200_SK_NOSLOTS") && atoi(getenv("CURVATURE_B200_SK_NOSLOTS")) != 0;   // (test hook)
  sk.have_slots = (P <= SK_MAXP && !no_slots) ? 1 : 0;
  if (sk.have_slots) {
    int lo = 0, hi = 0;
    for (int qq = 0; qq < P; ++qq) {
      while (lo + 1 < sk.G && ((int)sk.q[lo + 1] < qq || ((int)sk.q[lo + 1] == qq && sk.b[lo + 1] == 0))) ++lo;   // last boundary <= (qq, 0)
      if (hi < lo) hi = lo;
      while (hi + 1 < sk.G && (int)sk.q[hi + 1] <= qq) ++hi;                                                     // last boundary < (qq + 1, 0)
      sk.plo[qq] = (uint16_t)lo;
      sk.pn[qq] = (uint16_t)(hi - lo + 1);
    }
  }
}

}  // namespace

size_t syrk_tc_workspace(const ConvGeom& g, int precision) {
  const int sms = device_sm_count();
  size_t b = make_plan(g, sms, (int)((g.R + 31) / 32)).ws_bytes;
  if (precision == CRV_PREC_TF32_TMA) {
    TmaGeom tg;
    int chunks = 0;
    if (tma_geometry(g, tg, chunks)) {
      const size_t b2 = make_plan(g, sms, chunks).ws_bytes;
      if (b2 > b) b = b2;
    }
  }
  return b;
}

int syrk_tc_launch(const ConvGeom& g, float alpha, float* F, int precision, void* ws, size_t ws_bytes,
                   cudaStream_t s) {
  CRV_CHECK(precision == CRV_PREC_TF32 || precision == CRV_PREC_TF32_TMA,
            "NCHW staged SYRK: tier %d is not built (available: tf32, tf32_tma)", precision);
  CRV_CHECK(F != nullptr, "null factor pointer");
  const int sms = device_sm_count();
  CRV_CHECK(sms > 0, "no CUDA device");
  CRV_CHECK(g.R < (1LL << 31) - 64, "contraction length too large");
  if (int rc = syrk_stream_join(s)) return rc;   // this path's workspace overlaps both halves of the channels-last one
  TmaGeom tg;
  int tma_chunks = 0;
  const bool use_tma = precision == CRV_PREC_TF32_TMA && tma_geometry(g, tg, tma_chunks) && tensor_map_encoder();
  const Plan pl = make_plan(g, sms, use_tma ? tma_chunks : (int)((g.R + 31) / 32));
  CRV_CHECK(ws != nullptr && ws_bytes >= pl.ws_bytes, "workspace too small: %zu < %zu", ws_bytes, pl.ws_bytes);
  CRV_CHECK(((uintptr_t)ws & 15) == 0, "workspace must be 16-byte aligned");
  TcParams p;
  p.g = g;
  p.divL = make_fastdiv((uint32_t)g.L);
  p.divOW = make_fastdiv((uint32_t)g.OW);
  p.divC = make_fastdiv((uint32_t)g.C);
  p.divKW = make_fastdiv((uint32_t)g.kw);
  p.T = pl.T; p.pairs = pl.pairs; p.splits = pl.splits; p.chunks = pl.chunks; p.cps = pl.cps;
  p.HW = g.H * g.W; p.CHW = g.C * g.H * g.W; p.KK = g.kh * g.kw;
  p.ws = (float*)ws;
  p.vec_ok = (p.KK == 1 && g.sh == 1 && g.sw == 1 && g.ph == 0 && g.pw == 0 && (g.L % 4) == 0 &&
              ((uintptr_t)g.x & 15) == 0) ? 1 : 0;
  if (use_tma) {
    CUtensorMap map;
    const cuuint64_t Wd = tg.flat ? (cuuint64_t)g.L : (cuuint64_t)g.W;
    const cuuint64_t Hd = tg.flat ? 1 : (cuuint64_t)g.H;
    const cuuint64_t gdim[4] = {Wd, Hd, (cuuint64_t)g.C, (cuuint64_t)g.N};
    const cuuint64_t gstr[3] = {Wd * 4, Hd * Wd * 4, (cuuint64_t)g.C * Hd * Wd * 4};
    const cuuint32_t box[4] = {(cuuint32_t)tg.bw, (cuuint32_t)tg.bh, (cuuint32_t)tg.chbox, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g.x, gdim, gstr, box,
                                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRV_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    CRV_CUDA(cudaFuncSetAttribute(syrk_tc_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    profile_begin(KC_SYRK_STAGED, (double)g.R * g.D * (g.D + 1), 4.0 * g.N * g.C * g.H * g.W, s);
    syrk_tc_tma_kernel<<<pl.pairs * pl.splits, TMA_THREADS, SMEM_BYTES, s>>>(p, tg, map);
    profile_end(s);
  } else {
    CRV_CUDA(cudaFuncSetAttribute(syrk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    profile_begin(KC_SYRK_STAGED, (double)g.R * g.D * (g.D + 1), 4.0 * g.N * g.C * g.H * g.W, s);
    syrk_tc_kernel<<<pl.pairs * pl.splits, NTHREADS, SMEM_BYTES, s>>>(p);
    profile_end(s);
  }
  CRV_CUDA(cudaGetLastError());
  profile_begin(KC_SYRK_REDUCE, 0.0, (double)pl.ws_bytes + 8.0 * g.D * g.D, s);
  syrk_tc_reduce_kernel<<<pl.pairs * 64, 1024, 0, s>>>(p, alpha, F);
  profile_end(s);
  CRV_CUDA(cudaGetLastError());
  return 0;
}


// ---- split reduction on a side stream --------------------------------------------------------------------
// The fixed-order reduction of a factor's partial tiles depends only on that factor's main kernel, and the next
// factor's main kernel depends on neither (different factor, and the workspace is double-buffered): so the reduction
// is enqueued on an internal side stream behind an event, where it overlaps the next main kernel (it needs 1024
// threads and 4 KB of shared memory per CTA and fits beside the resident SYRK CTA).  crv_stream_join() makes the
// caller's stream wait for the outstanding reductions; the host classes call it at the end of update().
namespace {
struct SideState {
  cudaStream_t side = nullptr;       // split reductions
  cudaStream_t cast = nullptr;       // cast / rounding pre-passes (only between crv_stream_fork and crv_stream_join)
  cudaStream_t hp = nullptr;         // contraction kernels between fork and join: highest priority, so that the block
                                     // scheduler places a SYRK CTA on every SM first and fills the rest of the SM with
                                     // reduction / pre-pass CTAs (queued earlier, they would otherwise crowd it out)
  cudaStream_t hp2 = nullptr;        // second contraction stream: consecutive contraction launches ALTERNATE between hp
                                     // and hp2, so launch k + 1 does not wait for the last CTA of launch k -- its CTAs
                                     // start on every SM a CTA of launch k has left (the factors are independent; the
                                     // workspace rings are ordered by events, not by stream order).  A stream-K launch
                                     // ends ragged (CTAs with two accumulator flushes finish 10-15 us after the median)
                                     // and a dependent launch costs another 5-15 us: ~40 such seams per ResNet-50 update.
  bool two_hp = true;
  int hp_toggle = 0;
  cudaEvent_t ev_hp = nullptr, ev_hp2 = nullptr;
  // Workspace rings.  Partial tiles: npart buffers -- contraction j writes buffer j % npart and first waits for reduction
  // j - npart.  Pre-pass copies: NCOPY slots of their own -- pre-pass c (the c-th launch that has one) writes slot
  // c % ncopy and waits only for the CONTRACTION that last read that slot (c - ncopy): pre-passes run up to ncopy - 1
  // launches ahead of the contractions, whatever the reductions are doing.  Two and two by default: a pre-pass one
  // launch ahead is enough to hide it, and its bf16 copy (51-411 MB) is then still largely in the 126 MB L2 when the
  // contraction reads it -- with deeper rings two more copies and their partial tiles pass through L2 in between.
  static constexpr int NCOPY_MAX = 6;
  static constexpr int NPART_MAX = 4;
  cudaEvent_t ev_main[NPART_MAX] = {}, ev_red[NPART_MAX] = {};
  cudaEvent_t ev_cast[NCOPY_MAX] = {}, ev_used[NCOPY_MAX] = {};
  cudaEvent_t ev_fork = nullptr;
  bool red_pending[NPART_MAX] = {}, main_pending[NPART_MAX] = {};
  bool used_pending[NCOPY_MAX] = {};
  bool forked = false;
  int toggle = 0, ctoggle = 0, ncopy = 2, npart = 2;   // (deeper rings measured 1.5-2.5 % slower: profiles/r2_ring_depth_ab.txt)
  size_t sig[4] = {0, 0, 0, 0};      // workspace layout of the last batch (base, bytes, partial size, copy size)
  bool enabled = true, init = false;
};
SideState g_side_state[16];

SideState* side_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return nullptr; }
  SideState& st = g_side_state[dev];
  if (!st.init) {
    st.init = true;
    const char* e = getenv("CURVATURE_B200_SIDE_STREAM");
    st.enabled = !(e && atoi(e) == 0);
    if (st.enabled) {
      int least = 0, greatest = 0;
      if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); least = greatest = 0; }
      bool ok = cudaStreamCreateWithPriority(&st.side, cudaStreamNonBlocking, least) == cudaSuccess &&
                cudaStreamCreateWithPriority(&st.cast, cudaStreamNonBlocking, least) == cudaSuccess &&
                cudaStreamCreateWithPriority(&st.hp, cudaStreamNonBlocking, greatest) == cudaSuccess &&
                cudaStreamCreateWithPriority(&st.hp2, cudaStreamNonBlocking, greatest) == cudaSuccess &&
                cudaEventCreateWithFlags(&st.ev_hp, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&st.ev_hp2, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&st.ev_fork, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < SideState::NPART_MAX && ok; ++i)
        ok = cudaEventCreateWithFlags(&st.ev_main[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st.ev_red[i], cudaEventDisableTiming) == cudaSuccess;
      const char* h2 = getenv("CURVATURE_B200_HP2");
      st.two_hp = !(h2 && atoi(h2) == 0);
      const char* np = getenv("CURVATURE_B200_NPART");
      if (np) st.npart = std::max(2, std::min((int)SideState::NPART_MAX, atoi(np)));
      for (int i = 0; i < SideState::NCOPY_MAX && ok; ++i)
        ok = cudaEventCreateWithFlags(&st.ev_cast[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st.ev_used[i], cudaEventDisableTiming) == cudaSuccess;
      const char* nc = getenv("CURVATURE_B200_NCOPY");
      if (nc) st.ncopy = std::max(1, std::min((int)SideState::NCOPY_MAX, atoi(nc)));
      if (!ok) { cudaGetLastError(); st.enabled = false; }
    }
  }
  return &st;
}
}  // namespace

// Make `s` wait for every reduction still outstanding on the side stream (no host synchronisation); ends a fork.
int syrk_stream_join(cudaStream_t s) {
  SideState* st = side_state();
  if (!st || !st->enabled) return 0;
  for (int b = 0; b < SideState::NPART_MAX; ++b)
    if (st->red_pending[b]) {
      CRV_CUDA(cudaStreamWaitEvent(s, st->ev_red[b], 0));
      st->red_pending[b] = false;
    }
  if (st->forked) {          // the contraction kernels of the fork ran on the internal high-priority stream
    CRV_CUDA(cudaEventRecord(st->ev_hp, st->hp));
    CRV_CUDA(cudaStreamWaitEvent(s, st->ev_hp, 0));
    CRV_CUDA(cudaEventRecord(st->ev_hp2, st->hp2));
    CRV_CUDA(cudaStreamWaitEvent(s, st->ev_hp2, 0));
  }
  st->forked = false;
  st->hp_toggle = 0;
  return 0;
}

// A launch of a fork .. join section failed: order the caller's stream behind everything the section enqueued so far and
// clear the section's bookkeeping, so that later (non-forked) calls do not run on the internal streams or skip a wait.
static void side_abort(cudaStream_t s) {
  SideState* st = side_state();
  if (!st || !st->enabled) return;
  for (int b = 0; b < SideState::NPART_MAX; ++b) {
    if (st->red_pending[b]) cudaStreamWaitEvent(s, st->ev_red[b], 0);
    if (st->main_pending[b]) cudaStreamWaitEvent(s, st->ev_main[b], 0);
    st->red_pending[b] = st->main_pending[b] = false;
  }
  for (int c = 0; c < SideState::NCOPY_MAX; ++c) {
    if (st->used_pending[c]) cudaStreamWaitEvent(s, st->ev_used[c], 0);
    st->used_pending[c] = false;
  }
  st->ctoggle = 0;
  if (st->forked) {
    if (cudaEventRecord(st->ev_hp, st->hp) == cudaSuccess) cudaStreamWaitEvent(s, st->ev_hp, 0);
    if (cudaEventRecord(st->ev_hp2, st->hp2) == cudaSuccess) cudaStreamWaitEvent(s, st->ev_hp2, 0);
    if (cudaEventRecord(st->ev_fork, st->cast) == cudaSuccess) cudaStreamWaitEvent(s, st->ev_fork, 0);
  }
  st->forked = false;
  st->toggle = 0;
  cudaGetLastError();
}

// Declare that every operand tensor of the SYRK calls that follow (until the next join) is complete on `s` NOW: the
// cast / rounding pre-pass of call i may then run on a side stream, concurrently with the main kernel of call i-1
// (an HBM-bound copy beside an L2/TMA-bound contraction), instead of in order behind it.
int syrk_stream_fork(cudaStream_t s) {
  SideState* st = side_state();
  if (!st || !st->enabled) return 0;
  CRV_CUDA(cudaEventRecord(st->ev_fork, s));
  CRV_CUDA(cudaStreamWaitEvent(st->cast, st->ev_fork, 0));
  CRV_CUDA(cudaStreamWaitEvent(st->hp, st->ev_fork, 0));
  CRV_CUDA(cudaStreamWaitEvent(st->hp2, st->ev_fork, 0));
  st->forked = true;
  st->hp_toggle = 0;
  return 0;
}

// ---- channels-last entry points ----------------------------------------------------------------------
bool syrk_nhwc_supported(const ConvGeom& g, int precision) {
  if (precision != CRV_PREC_TF32 && precision != CRV_PREC_TF32_TMA && precision != CRV_PREC_BF16 &&
      precision != CRV_PREC_BF16X3)
    return false;
  NhPlan pl;
  return nhwc_plan(g, precision, 148, pl);
}

namespace {
// which launch a factor goes into: re-read operands (bf16 copy, or several block pairs) get a launch of their own;
// single-pair, read-once operands share launches of up to GRP_MAXF factors
int group_max_t() {
  static const int t = getenv("CURVATURE_B200_GROUP_T") ? atoi(getenv("CURVATURE_B200_GROUP_T")) : 2;
  return t;
}
bool rides_in_group(const NhPlan& pl) { return !pl.bf16 && pl.p.T <= group_max_t() && pl.copy_bytes == 0; }

int make_tensor_map(const ConvGeom& g, const NhPlan& pl, const void* src, CUtensorMap* map) {
  const NhParams& p = pl.p;
  const cuuint64_t esz = pl.bf16 ? 2 : 4;
  const cuuint32_t chbox = pl.bf16 ? 64 : 32;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  if (p.flat) {
    gdim[0] = (cuuint64_t)g.C; gdim[1] = (cuuint64_t)g.R; gdim[2] = 1; gdim[3] = 1;
    gstr[0] = (cuuint64_t)g.C * esz; gstr[1] = (cuuint64_t)g.R * g.C * esz; gstr[2] = gstr[1];
    box[0] = chbox; box[1] = (cuuint32_t)p.PB; box[2] = 1; box[3] = 1;
    estr[0] = estr[1] = estr[2] = estr[3] = 1;
  } else {
    gdim[0] = (cuuint64_t)g.C; gdim[1] = (cuuint64_t)g.W; gdim[2] = (cuuint64_t)g.H; gdim[3] = (cuuint64_t)g.N;
    gstr[0] = (cuuint64_t)g.C * esz; gstr[1] = (cuuint64_t)g.W * g.C * esz; gstr[2] = (cuuint64_t)g.H * g.W * g.C * esz;
    box[0] = chbox; box[1] = (cuuint32_t)(p.bw * g.sw); box[2] = (cuuint32_t)(p.bh * g.sh); box[3] = (cuuint32_t)p.bn;
    estr[0] = 1; estr[1] = (cuuint32_t)g.sw; estr[2] = (cuuint32_t)g.sh; estr[3] = 1;
  }
  const CUresult rc = tensor_map_encoder()(map, pl.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                           4, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           pl.bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CRV_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
  return 0;
}

// One launch of the stream-K kernel (+ its reduction) over the factors idx[0..cnt) -- all of the same operand type.
int launch_group(const ConvGeom* gs, const float* alphas, float* const* Fs, const std::vector<NhPlan>& plans,
                 const int* idx, int cnt, void* ws, size_t ws_bytes, size_t copy_off, size_t max_copy, cudaStream_t caller,
                 SkTable* cached_sk = nullptr, std::vector<int>* cached_qbeg = nullptr, char* cached_valid = nullptr) {
  const int sms = device_sm_count();
  const bool bf16 = plans[idx[0]].bf16 != 0;
  size_t pairs = 0, copy_bytes = 0;
  for (int k = 0; k < cnt; ++k) { pairs += plans[idx[k]].pairs; copy_bytes += plans[idx[k]].copy_bytes; }
  for (int k = 0; cnt > 1 && k < cnt; ++k)
    CRV_CHECK(plans[idx[k]].copy_bytes == 0 || (plans[idx[k]].bf16 && !plans[idx[k]].p.x3 && !plans[idx[k]].pack),
              "internal: only single-plane bf16 copies share a launch");
  copy_bytes = 0;
  for (int k = 0; k < cnt; ++k) copy_bytes += (plans[idx[k]].copy_bytes + 1023) & ~(size_t)1023;
  const size_t partial_bytes = (size_t)(SK_MAXG + pairs) * TILE_ELEMS * sizeof(float);
  // Workspace layout of a batch: [partial tiles 0 | partial tiles 1 | copy slot 0 | ... | copy slot ncopy-1]; partial
  // buffers are `copy_off` bytes (the largest partial-tile region of any launch of the batch), copy slots `max_copy`.
  // The two rings are independent: see SideState.
  SideState* st = side_state();
  const int ncopy = st ? st->ncopy : 1, npart = st ? st->npart : 1;
  const size_t slot_bytes = (max_copy + 4095) & ~(size_t)1023;
  const size_t need = (size_t)npart * copy_off + (size_t)ncopy * slot_bytes + 2048;
  CRV_CHECK(partial_bytes <= copy_off && copy_bytes <= max_copy, "internal: launch exceeds the batch layout");
  CRV_CHECK(ws != nullptr && ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  CRV_CHECK(((uintptr_t)ws & 15) == 0, "workspace must be 16-byte aligned");
  // (while per-kernel event timing is on, everything runs in order on the caller's stream: an event bracket then
  // times the kernel alone, not the kernel plus whatever shares the SMs with it)
  const bool use_side = st && st->enabled && !profile_on();
  // between fork and join the contraction kernels run on the internal high-priority stream (see SideState::hp)
  cudaStream_t s = caller;
  if (use_side && st->forked) {
    s = (st->two_hp && st->hp_toggle) ? st->hp2 : st->hp;
    st->hp_toggle ^= 1;
  }
  int buf = 0;
  if (use_side) {
    buf = st->toggle;
    st->toggle = (st->toggle + 1) % npart;
    if (st->red_pending[buf]) CRV_CUDA(cudaStreamWaitEvent(s, st->ev_red[buf], 0));   // reduction j - npart read this buffer
  }
  char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  char* wsb = base + (size_t)buf * copy_off;
  static GroupParams gp;       // (8 KB: kept off the stack; the library is not re-entrant, see the header)
  static GroupMaps maps;
  static SkTable sk;
  gp.nf = cnt;
  gp.ws = (float*)wsb;
  gp.tl = debug_timeline_buffer();
  gp.trace = debug_trace_slot();
  {
    static const int tl_min_nf = getenv("CURVATURE_B200_TL_MIN_NF") ? atoi(getenv("CURVATURE_B200_TL_MIN_NF")) : 0;
    if (cnt < tl_min_nf) gp.tl = nullptr;      // profiling aid: record only the launches with at least that many factors
  }
  {
    const char* d = getenv("CURVATURE_B200_DBG");
    gp.dbg = d ? atoi(d) : 0;
  }
  double flops = 0.0, bytes = 0.0, fbytes = 0.0;
  int used_slot = -1;                  // copy slot this launch's contraction reads (one slot per launch: the copies of
  size_t slot_off = 0;                 // a launch's operands sit behind one another in it)
  cudaStream_t cast_stream = nullptr;
  std::vector<const NhPlan*> pls(cnt);
  for (int k = 0; k < cnt; ++k) {
    const ConvGeom& g = gs[idx[k]];
    const NhPlan& pl = plans[idx[k]];
    pls[k] = &pl;
    gp.f[k] = pl.p;
    gp.f[k].alpha = alphas[idx[k]];
    gp.f[k].F = Fs[idx[k]];
    const float* src = g.x;
    int cslot = -1;
    if (pl.copy_bytes) {   // pre-pass: bf16 copy (tiers bf16 / bf16x3) / TF32 round-to-nearest copy (tier tf32), or pack
      cudaStream_t cs = s;
      const bool side_cast = use_side && st->forked;
      if (side_cast) cs = st->cast;
      if (used_slot < 0) {
        cslot = use_side ? st->ctoggle : 0;
        if (use_side) st->ctoggle = (st->ctoggle + 1) % ncopy;
        // the contraction that last read this slot (ncopy launches with a copy ago) must have finished
        if (use_side && st->used_pending[cslot]) {
          CRV_CUDA(cudaStreamWaitEvent(cs, st->ev_used[cslot], 0));
          st->used_pending[cslot] = false;
        }
      } else {
        cslot = used_slot;
      }
      float* copy = (float*)(base + (size_t)npart * copy_off + (size_t)cslot * slot_bytes + slot_off);
      slot_off += (pl.copy_bytes + 1023) & ~(size_t)1023;
      const size_t n4 = (size_t)g.N * g.C * g.H * g.W / 4;
      if (pl.pack) {
        const ConvGeom& q = pl.gq;
        const long long nt = (long long)q.N * q.H * q.W * 2;
        const unsigned blocks = (unsigned)((nt + 255) / 256);
        long long sN, sC, sH, sW;
        if (g.x_nchw) { sW = 1; sH = g.W; sC = (long long)g.H * g.W; sN = sC * g.C; }
        else { sC = 1; sW = g.C; sH = (long long)g.W * g.C; sN = sH * g.H; }
        profile_begin(KC_PREPASS, 0.0, 4.0 * g.N * g.C * g.H * g.W + (double)nt * 64.0, cs);
        pack_smallc_kernel<<<blocks, 256, 0, cs>>>(g.x, (uint4*)copy, pl.p.x3 ? (uint4*)((char*)copy + pl.plane_bytes) : nullptr,
                                                   g.N, g.C, g.H, g.W, sN, sC, sH, sW, q.H, q.W, g.kw, g.sw, g.ph, g.pw, gp.trace);
      } else if (pl.transpose) {
        // NCHW-dense source: transposing cast into the channels-last bf16 copy (+ lo plane on the bf16x3 tier)
        const int HW = g.H * g.W;
        const int ctiles = (g.C + 63) / 64, ptiles = (HW + 63) / 64;
        const size_t plane = pl.plane_bytes;
        profile_begin(KC_PREPASS, 0.0, (pl.p.x3 ? 8.0 : 6.0) * (double)n4 * 4.0, cs);
        nchw_to_nhwc_bf16_kernel<<<(unsigned)((size_t)g.N * ctiles * ptiles), 256, 0, cs>>>(
            g.x, (uint32_t*)copy, pl.p.x3 ? (uint32_t*)((char*)copy + plane) : nullptr, g.C, HW, ctiles, ptiles, gp.trace);
      } else if (pl.p.x3) {
        const unsigned blocks = (unsigned)((n4 + 256 * PRE_V4 - 1) / (256 * PRE_V4));
        const size_t plane = pl.plane_bytes;
        profile_begin(KC_PREPASS, 0.0, 8.0 * (double)n4 * 4.0, cs);
        split_bf16_kernel<<<blocks, 256, 0, cs>>>((const float4*)g.x, (uint2*)copy, (uint2*)((char*)copy + plane), n4, gp.trace);
      } else {
        const unsigned blocks = (unsigned)((n4 + 256 * PRE_V4 - 1) / (256 * PRE_V4));
        profile_begin(KC_PREPASS, 0.0, (pl.bf16 ? 6.0 : 8.0) * (double)n4 * 4.0, cs);
        static const int cast_hints = getenv("CURVATURE_B200_CAST_HINT") ? atoi(getenv("CURVATURE_B200_CAST_HINT")) : 1;
        if (pl.bf16) cast_bf16_kernel<<<blocks, 256, 0, cs>>>((const float4*)g.x, (uint2*)copy, n4, cast_hints, gp.trace);
        else round_tf32_kernel<<<blocks, 256, 0, cs>>>((const float4*)g.x, (float4*)copy, n4, gp.trace);
      }
      profile_end(cs);
      CRV_CUDA(cudaGetLastError());
      if (side_cast) cast_stream = cs;
      src = copy;
      used_slot = cslot;
    }
    if (int rc = make_tensor_map(pl.pack ? pl.gq : g, pl, src, &maps.m[k])) return rc;
    if (pl.p.x3) {         // (a group of one: the lo plane's map sits in the next slot)
      if (int rc = make_tensor_map(pl.pack ? pl.gq : g, pl, (const char*)src + pl.plane_bytes, &maps.m[k + 1])) return rc;
    }
    flops += (double)g.R * g.D * (g.D + 1);
    bytes += pl.pack ? 2.0 * pl.gq.N * pl.gq.C * pl.gq.H * pl.gq.W : (pl.bf16 ? 2.0 : 4.0) * g.N * g.C * g.H * g.W;
    fbytes += 8.0 * g.D * g.D;
  }
  {
    const cuuint64_t gd[2] = {(cuuint64_t)TB, (cuuint64_t)(SK_MAXG + pairs) * TB}, gs[1] = {(cuuint64_t)TB * sizeof(float)};
    const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    const CUresult rc = tensor_map_encoder()(&maps.part, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wsb, gd, gs, box, es,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CRV_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(partial tiles) failed: %d", (int)rc);
  }
  if (cast_stream) {                            // the contraction waits for the (last) pre-pass of its launch
    CRV_CUDA(cudaEventRecord(st->ev_cast[used_slot], cast_stream));
    CRV_CUDA(cudaStreamWaitEvent(s, st->ev_cast[used_slot], 0));
  }
  if (cached_valid && *cached_valid) {          // same batch geometry as last time: the partition is unchanged
    sk = *cached_sk;
    for (int k = 0; k <= cnt; ++k) gp.qbeg[k] = (*cached_qbeg)[k];
  } else {
    build_sk(pls, sms, gp, sk);
    if (cached_valid) {
      *cached_sk = sk;
      cached_qbeg->assign(gp.qbeg, gp.qbeg + cnt + 1);
      *cached_valid = 1;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    attr_set = true;
    CRV_CUDA(cudaFuncSetAttribute(syrk_nhwc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NH_SMEM_BYTES));
    CRV_CUDA(cudaFuncSetAttribute(syrk_nhwc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NH_SMEM_BYTES));
    // same shared-memory carve-out as the SYRK kernel, or the two can never be resident on one SM at the same time
    cudaFuncSetAttribute(syrk_sk_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(syrk_sk_reduce_taps_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(cast_bf16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
  }
  profile_begin(bf16 ? KC_SYRK_NHWC_BF16 : KC_SYRK_NHWC_TF32, flops, bytes, s);
  if (bf16) syrk_nhwc_kernel<true><<<sk.G, NH_THREADS, NH_SMEM_BYTES, s>>>(gp, sk, maps);
  else syrk_nhwc_kernel<false><<<sk.G, NH_THREADS, NH_SMEM_BYTES, s>>>(gp, sk, maps);
  profile_end(s);
  CRV_CUDA(cudaGetLastError());
  cudaStream_t rs = s;
  if (use_side) {
    CRV_CUDA(cudaEventRecord(st->ev_main[buf], s));
    st->main_pending[buf] = true;
    CRV_CUDA(cudaStreamWaitEvent(st->side, st->ev_main[buf], 0));
    rs = st->side;
    if (used_slot >= 0) {
      CRV_CUDA(cudaEventRecord(st->ev_used[used_slot], s));
      st->used_pending[used_slot] = true;
    }
  }
  profile_begin(KC_SYRK_REDUCE, 0.0, (double)(sk.G + pairs) * TILE_ELEMS * 4.0 + fbytes, rs);
  static const bool taps_reduce = !(getenv("CURVATURE_B200_TAPS_REDUCE") && atoi(getenv("CURVATURE_B200_TAPS_REDUCE")) == 0);
  // (worth it when a pair has few partial tiles: with few pairs the generic kernel's 16-deep load batches win)
  if (taps_reduce && cnt == 1 && !plans[idx[0]].pack && gp.f[0].KK >= 2 && gp.f[0].KK <= 9 && (gp.f[0].C & 31) == 0 &&
      gp.f[0].T >= 8) {
    const int KK = gp.f[0].KK, C = gp.f[0].C;
    const size_t sm = (size_t)8 * (KK * 32 + 1) * sizeof(float);
    syrk_sk_reduce_taps_kernel<<<(unsigned)((C / 8) * KK * (C / 32)), 256, sm, rs>>>(gp, sk);
  } else {
    syrk_sk_reduce_kernel<<<(unsigned)pairs * 64, 256, 0, rs>>>(gp, sk);
  }
  profile_end(rs);
  CRV_CUDA(cudaGetLastError());
  if (use_side) {
    CRV_CUDA(cudaEventRecord(st->ev_red[buf], st->side));
    st->red_pending[buf] = true;
  }
  return 0;
}

// Host-side plan cache.  Planning a ResNet-50 step (108 factors: box search, stream-K partition of 40 launches) costs about
// as much host time as enqueueing it, and an estimation pass repeats the SAME geometry every step -- only the tensor
// pointers change.  The plans, the launch order and the stream-K tables of the last batch are kept, keyed by everything
// they depend on (geometry, layout flags, pointer alignment, tier, SM count).
struct BatchCache {
  std::vector<ConvGeom> key;
  int precision = -1, sms = 0;
  std::vector<NhPlan> plans;
  std::vector<std::vector<int>> launches;
  std::vector<SkTable> sk;              // per launch
  std::vector<std::vector<int>> qbeg;   // per launch: GroupParams::qbeg
  std::vector<char> sk_valid;
};
BatchCache g_batch_cache;
static const bool g_plan_cache_on = !(getenv("CURVATURE_B200_PLAN_CACHE") && atoi(getenv("CURVATURE_B200_PLAN_CACHE")) == 0);

bool cache_matches(const ConvGeom* gs, int n, int precision, int sms) {
  const BatchCache& c = g_batch_cache;
  if (!g_plan_cache_on || c.precision != precision || c.sms != sms || (int)c.key.size() != n) return false;
  for (int i = 0; i < n; ++i) {
    const ConvGeom& a = gs[i];
    const ConvGeom& b = c.key[i];
    if (((uintptr_t)a.x & 15) != (uintptr_t)b.x || a.N != b.N || a.C != b.C || a.H != b.H || a.W != b.W || a.kh != b.kh ||
        a.kw != b.kw || a.sh != b.sh || a.sw != b.sw || a.ph != b.ph || a.pw != b.pw || a.has_bias != b.has_bias ||
        a.x_nchw != b.x_nchw || a.zero_mean != b.zero_mean)
      return false;
  }
  return true;
}

int plan_batch_uncached(const ConvGeom* gs, int n, int precision, std::vector<NhPlan>& plans, std::vector<std::vector<int>>& launches,
                        int sms_override);

// plans / launches of the batch, through the cache (sms_override > 0: the host-only debug view, never cached)
int plan_batch(const ConvGeom* gs, int n, int precision, std::vector<NhPlan>& plans, std::vector<std::vector<int>>& launches,
               int sms_override = 0) {
  const int sms = sms_override > 0 ? sms_override : device_sm_count();
  if (sms_override <= 0 && cache_matches(gs, n, precision, sms)) {
    plans = g_batch_cache.plans;
    launches = g_batch_cache.launches;
    return 0;
  }
  if (int rc = plan_batch_uncached(gs, n, precision, plans, launches, sms_override)) return rc;
  if (sms_override <= 0 && g_plan_cache_on) {
    BatchCache& c = g_batch_cache;
    c.key.resize(n);
    for (int i = 0; i < n; ++i) {
      c.key[i] = gs[i];
      c.key[i].x = (const float*)((uintptr_t)gs[i].x & 15);
    }
    c.precision = precision; c.sms = sms;
    c.plans = plans; c.launches = launches;
    c.sk.assign(launches.size(), SkTable());
    c.qbeg.assign(launches.size(), std::vector<int>());
    c.sk_valid.assign(launches.size(), 0);
  }
  return 0;
}

int plan_batch_uncached(const ConvGeom* gs, int n, int precision, std::vector<NhPlan>& plans, std::vector<std::vector<int>>& launches,
                        int sms_override) {
  CRV_CHECK(precision == CRV_PREC_TF32 || precision == CRV_PREC_TF32_TMA || precision == CRV_PREC_BF16 ||
                precision == CRV_PREC_BF16X3,
            "channels-last SYRK: tier %d is not built (available: tf32, tf32_tma, bf16, bf16x3)", precision);
  const int sms = sms_override > 0 ? sms_override : device_sm_count();
  CRV_CHECK(sms > 0, "no CUDA device");
  plans.resize(n);
  std::vector<int> grp;
  static const int grp_max = getenv("CURVATURE_B200_GROUP") ? std::max(1, std::min(GRP_MAXF, atoi(getenv("CURVATURE_B200_GROUP")))) : GRP_MAXF;
  // Light re-read factors (the 1x1 convolutions of the 14^2 and 7^2 stages: 53 GF each, 30 us of tensor time against
  // ~30 us of ramp, accumulator flush and tail) CAN share launches (CURVATURE_B200_BF16_GROUP = operands per launch), but
  // it does not pay and is off by default: with four 103 MB copies live at once the operands no longer stay in L2
  // between the pairs that re-read them, and the launch becomes DRAM-bound (4 x 70 us apart -> 270-307 us together,
  // profiles/r2_bf16_grouping_ab.txt).
  static const int lgrp_max = getenv("CURVATURE_B200_BF16_GROUP") ? std::max(1, std::min(8, atoi(getenv("CURVATURE_B200_BF16_GROUP")))) : 1;
  static const double light_flops = getenv("CURVATURE_B200_BF16_GROUP_GF") ? atof(getenv("CURVATURE_B200_BF16_GROUP_GF")) * 1e9 : 120e9;
  size_t max_single_copy = 0;
  for (int i = 0; i < n; ++i) {
    CRV_CHECK(nhwc_plan(gs[i], precision, sms, plans[i]),
              "channels-last SYRK: unsupported geometry (needs no bias row, C %% 4 == 0, C >= 32, C %% 32 == 0 for k x k; "
              "bf16x3 tier and NCHW-dense sources: C >= 64, C %% 8 == 0, C %% 64 == 0 for k x k)");
    max_single_copy = std::max(max_single_copy, plans[i].copy_bytes);
  }
  int open_light = -1;                  // index into `launches` of the light-bf16 launch that still has room
  size_t open_copy = 0, open_pairs = 0;
  for (int i = 0; i < n; ++i) {
    const NhPlan& pl = plans[i];
    const bool light = lgrp_max > 1 && pl.bf16 && !pl.p.x3 && !pl.pack && pl.copy_bytes > 0 &&
                       (double)gs[i].R * gs[i].D * (gs[i].D + 1) <= light_flops;
    if (rides_in_group(pl)) {
      grp.push_back(i);
      if ((int)grp.size() == grp_max) { launches.push_back(grp); grp.clear(); }
    } else if (light) {
      const size_t cb = (pl.copy_bytes + 1023) & ~(size_t)1023;
      if (open_light >= 0 && (int)launches[open_light].size() < lgrp_max && open_copy + cb <= max_single_copy &&
          open_pairs + (size_t)pl.pairs <= 160) {
        launches[open_light].push_back(i);
        open_copy += cb; open_pairs += (size_t)pl.pairs;
      } else {
        launches.push_back(std::vector<int>(1, i));
        open_light = (int)launches.size() - 1;
        open_copy = cb; open_pairs = (size_t)pl.pairs;
      }
    } else {
      launches.push_back(std::vector<int>(1, i));
    }
  }
  if (!grp.empty()) launches.push_back(grp);
  // Order of the launches (the factors are independent: any order gives the same sums).  Group launches have no pre-pass
  // and are HBM-bound; launches of one re-read factor are tensor-bound and need their pre-pass copy first.  The lightest
  // group goes first (it covers the first pre-passes, which have nothing else to hide behind), the other groups go last
  // (they stream from HBM at full rate while only the last reductions are still running beside them).
  static const bool reorder = !(getenv("CURVATURE_B200_REORDER") && atoi(getenv("CURVATURE_B200_REORDER")) == 0);
  if (reorder && launches.size() > 2) {
    auto weight = [&](const std::vector<int>& l) {
      double w = 0.0;
      for (int i : l) w += (double)gs[i].N * gs[i].C * gs[i].H * gs[i].W;
      return w;
    };
    std::vector<std::vector<int>> groups, singles;
    for (auto& l : launches) {
      if (plans[l[0]].copy_bytes == 0 && rides_in_group(plans[l[0]])) groups.push_back(l);
      else singles.push_back(l);
    }
    if (!groups.empty() && !singles.empty()) {
      std::stable_sort(groups.begin(), groups.end(), [&](const std::vector<int>& a, const std::vector<int>& b) { return weight(a) < weight(b); });
      launches.clear();
      launches.push_back(groups[0]);
      // (the packed stem's pre-pass is the longest of all -- it writes a 422 MB tensor: not first among the singles, so
      // that it runs behind two contractions instead of in front of an idle GPU)
      std::vector<std::vector<int>> packed, rest;
      for (auto& l : singles) (plans[l[0]].pack ? packed : rest).push_back(l);
      singles.clear();
      for (size_t k = 0; k < rest.size(); ++k) {
        if (k == 2) for (auto& l : packed) singles.push_back(l);
        singles.push_back(rest[k]);
      }
      if (rest.size() <= 2) for (auto& l : packed) singles.push_back(l);
      for (auto& l : singles) launches.push_back(l);
      for (size_t k = 1; k < groups.size(); ++k) launches.push_back(groups[k]);
    }
  }
  return 0;
}
}  // namespace

namespace {
// largest partial-tile region and largest pre-pass copy over the launches of a batch
void batch_layout(const std::vector<NhPlan>& plans, const std::vector<std::vector<int>>& launches, size_t& copy_off, size_t& max_copy) {
  size_t max_partial = 0;
  max_copy = 0;
  for (const auto& l : launches) {
    size_t pairs = 0, copy = 0;
    for (int i : l) { pairs += plans[i].pairs; copy += (plans[i].copy_bytes + 1023) & ~(size_t)1023; }
    max_partial = std::max(max_partial, (size_t)(SK_MAXG + pairs) * TILE_ELEMS * sizeof(float));
    max_copy = std::max(max_copy, copy);
  }
  copy_off = (max_partial + 2047) & ~(size_t)1023;
}
}  // namespace

size_t syrk_nhwc_batch_workspace(const ConvGeom* gs, int n, int precision) {
  std::vector<NhPlan> plans;
  std::vector<std::vector<int>> launches;
  if (n <= 0 || plan_batch(gs, n, precision, plans, launches)) return 0;
  size_t copy_off, max_copy;
  batch_layout(plans, launches, copy_off, max_copy);
  SideState* st = side_state();
  const size_t slot_bytes = (max_copy + 4095) & ~(size_t)1023;
  return (size_t)(st ? st->npart : 1) * copy_off + (size_t)(st ? st->ncopy : 1) * slot_bytes + 2048;   // see launch_group
}

// F_i += alpha_i * X_i X_i^T for a batch of channels-last operands.
int syrk_nhwc_batch_launch(const ConvGeom* gs, const float* alphas, float* const* Fs, int n, int precision, void* ws,
                           size_t ws_bytes, cudaStream_t s) {
  CRV_CHECK(n > 0, "empty batch");
  for (int i = 0; i < n; ++i) CRV_CHECK(Fs[i] != nullptr, "null factor pointer");
  CRV_CHECK(tensor_map_encoder() != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
  std::vector<NhPlan> plans;
  std::vector<std::vector<int>> launches;
  if (int rc = plan_batch(gs, n, precision, plans, launches)) return rc;
  size_t copy_off, max_copy;
  batch_layout(plans, launches, copy_off, max_copy);
  if (SideState* st = side_state()) {
    // a batch whose workspace layout differs from the previous one's: nothing of the earlier calls may still be reading
    // or writing the buffer where this batch is about to put other things
    const size_t sig[4] = {(size_t)(uintptr_t)ws, ws_bytes, copy_off, max_copy};
    if (st->enabled && memcmp(sig, st->sig, sizeof(sig)) != 0) {
      cudaStream_t w = st->forked ? st->hp : s;
      for (int b = 0; b < SideState::NPART_MAX; ++b) {
        if (st->red_pending[b]) CRV_CUDA(cudaStreamWaitEvent(w, st->ev_red[b], 0));
        if (st->main_pending[b]) CRV_CUDA(cudaStreamWaitEvent(w, st->ev_main[b], 0));
      }
      if (st->forked) {        // pre-passes of this batch run on the cast stream, every other contraction on the second
                               // contraction stream: order them behind the same events
        for (int b = 0; b < SideState::NPART_MAX; ++b) {
          if (st->red_pending[b]) CRV_CUDA(cudaStreamWaitEvent(st->cast, st->ev_red[b], 0));
          if (st->main_pending[b]) CRV_CUDA(cudaStreamWaitEvent(st->cast, st->ev_main[b], 0));
          if (st->red_pending[b]) CRV_CUDA(cudaStreamWaitEvent(st->hp2, st->ev_red[b], 0));
          if (st->main_pending[b]) CRV_CUDA(cudaStreamWaitEvent(st->hp2, st->ev_main[b], 0));
        }
      }
      memcpy(st->sig, sig, sizeof(sig));
    }
  }
  const bool cached = cache_matches(gs, n, precision, device_sm_count()) && g_batch_cache.sk.size() == launches.size();
  for (size_t li = 0; li < launches.size(); ++li) {
    const auto& l = launches[li];
    BatchCache& c = g_batch_cache;
    if (int rc = launch_group(gs, alphas, Fs, plans, l.data(), (int)l.size(), ws, ws_bytes, copy_off, max_copy, s,
                              cached ? &c.sk[li] : nullptr, cached ? &c.qbeg[li] : nullptr, cached ? &c.sk_valid[li] : nullptr)) {
      side_abort(s);
      return rc;
    }
  }
  return 0;
}

// Host-only view of the scheduler (no device needed): how a batch is cut into launches and, for launch `which`, the
// stream-K boundary table and the per-pair cost / stage granularity it was built from.  Used by the CPU tests.
int syrk_nhwc_debug_partition(const ConvGeom* gs, int n, int precision, int sms, int which, int* launch_of_item, int* G,
                              int* q, unsigned* b, int cap, int* pairs, int* nbox_of_pair, int* nb_of_pair, int pair_cap) {
  std::vector<NhPlan> plans;
  std::vector<std::vector<int>> launches;
  if (int rc = plan_batch(gs, n, precision, plans, launches, sms)) return rc;
  for (size_t l = 0; l < launches.size(); ++l)
    for (int i : launches[l]) launch_of_item[i] = (int)l;
  CRV_CHECK(which >= 0 && which < (int)launches.size(), "launch %d of %d", which, (int)launches.size());
  std::vector<const NhPlan*> pls;
  for (int i : launches[which]) pls.push_back(&plans[i]);
  static GroupParams gp;
  static SkTable sk;
  gp.nf = (int)pls.size();
  gp.trace = nullptr;
  build_sk(pls, sms, gp, sk);
  CRV_CHECK(sk.G + 1 <= cap, "boundary table needs %d entries", sk.G + 1);
  *G = sk.G;
  for (int c = 0; c <= sk.G; ++c) { q[c] = sk.q[c]; b[c] = sk.b[c]; }
  int P = 0;
  for (const NhPlan* pl : pls)
    for (int k = 0; k < pl->pairs; ++k, ++P) {
      CRV_CHECK(P < pair_cap, "pair table too small");
      int I, J;
      host_decode_pair(k, pl->p.T, I, J);
      nbox_of_pair[P] = pl->p.nbox;
      nb_of_pair[P] = I == J ? pl->p.NBdiag : pl->p.NBoff;
    }
  *pairs = P;
  return (int)launches.size() << 16;      // (number of launches in the high half; 0 in the low half = ok)
}

size_t syrk_nhwc_workspace(const ConvGeom& g, int precision) { return syrk_nhwc_batch_workspace(&g, 1, precision); }

int syrk_nhwc_launch(const ConvGeom& g, float alpha, float* F, int precision, void* ws, size_t ws_bytes,
                     cudaStream_t s) {
  return syrk_nhwc_batch_launch(&g, &alpha, &F, 1, precision, ws, ws_bytes, s);
}

}  // namespace crv
