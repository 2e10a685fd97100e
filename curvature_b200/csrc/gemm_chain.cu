// K3 / K5 as ONE kernel launch per call: every two-GEMM chain of a model -- the EFB projection (QG^T g QA)^2
// (curvature/curvatures.py:424-433) or the matrix-normal draw mu + L_G z^T L_A^T (:387-392, :78-82, :119) of every layer,
// and with S stacked samples the posterior-sampling loop of scripts/evaluate.py:121-152 -- runs in a single persistent
// grid.  The work list holds 256 x 256 output tiles of ALL GEMMs of the call in a dependency-respecting order; a tile of a
// second GEMM waits (device-side counter, acquire) until the tiles of the first GEMM that produce its rows of the
// intermediate are complete.  The (M, K) intermediate lives in the caller's workspace but is produced and consumed
// inside one launch while it is still in L2 (<= 9.4 MB per layer against 126 MB of L2): no launch boundary, no host
// round trip, no HBM round trip between the two products.
//
//   tile      256 x 256 fp32 accumulator = all 512 TMEM columns (two M = 128 halves x N <= 256), like the SYRK kernel;
//   operands  fp32 words fed as TF32 by TMA, each of A and B either contraction-contiguous ("K-major", SWIZZLE_128B,
//             boxes of [32 k][128 rows]) or row-contiguous ("MN-major", SWIZZLE_128B_ATOM_32B, boxes of [32 rows][32 k]);
//             64 KB per 32-deep stage, 3 stages; per stage 4 k-steps x 2 halves of tcgen05.mma.kind::tf32 (M 128, N 256, K 8):
//             64 B per tensor-pipe clock per SM of L2 -> SM traffic, half of what the 128 x 128 tiles of gemm_tc.cu need
//             (those are ingest-bound at 30 % tensor pipe);
//   issue     the MMA warp runs its loop in uniform control flow, tcgen05 instructions predicated on one elected lane
//             (see syrk_tc.cu); TMA issue is spread over 4 warps;
//   epilogue  4 warps drain TMEM through a swizzled shared-memory tile and touch global memory in whole 128-byte row
//             segments; epilogues: store (optionally rounded to TF32 for the next GEMM), accumulate the square, add the
//             posterior mean and split weight / bias columns;
//   schedule  dynamic: the host sorts the tiles (first GEMMs before the GEMMs that consume them, longest first within each)
//             and the CTAs claim them in that order from an atomic cursor; a waiting tile only ever waits for tiles that were
//             claimed earlier -- no deadlock.
#include "common.cuh"
#include "../../include/curvature_b200.h"
#include <cuda.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace crv {
namespace {

constexpr int CT = 256;                       // tile edge
constexpr int CK = 32;                        // contraction depth per stage (one 128-byte row of fp32)
constexpr int C_STAGE = 64 * 1024, C_NSTAGE = 3;
constexpr int C_EPI = 32 * 1024;              // 8 epilogue warps x (32 rows x 128 B) staging
constexpr int C_NPROD = 3;                    // TMA-issuing warps: 0, 6, 7
constexpr int C_THREADS = 12 * 32;            // warp 1: MMA + TMEM owner; warps 2-5 and 8-11: epilogue; 0, 6, 7: TMA
                                              // (registers are allocated per 4 warps: 12 warps x 168 fit, 13 would cap at 128)
constexpr int C_NSCHED = 4;                    // entries of the CTA's tile queue
constexpr int C_SMEM = C_NSTAGE * C_STAGE + C_EPI + 1024 + 1024;
constexpr uint32_t C_SPIN = 1u << 26;

struct ChainItemDev {
  int m, n, k, ldc;
  int a_mn, b_mn, mapA, mapB;
  int epi, round_out, dep_counters, dep_need;   // dep_counters: first per-row-tile counter of the producer GEMM (-1: none)
  int counters, dep_div;                        // this GEMM's own per-row-tile counters (-1: nobody waits for it);
                                                // dep_div: rows of this GEMM per producer row (stacked samples)
  int dep_rows, mapC;                           // rows of the producer GEMM; tensor map of the output (-1: not TMA-storable)
  float alpha;
  float* C;
  SampleEpilogue se;
};
struct ChainTile { int item, tm, tn, pad; };

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
    if (!ok && ++spins > C_SPIN) asm volatile("trap;");
  } while (!ok);
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {          // K-major SWIZZLE_128B, 8-row atoms 1024 B apart
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {   // MN-major SW128_32B (see syrk_tc.cu)
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ float rna_tf32(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct TileShape { int m0, n0, rows, mh, ncols, nk; };
__device__ __forceinline__ TileShape tile_shape(const ChainItemDev& it, const ChainTile& t) {
  TileShape s;
  s.m0 = t.tm * CT; s.n0 = t.tn * CT;
  s.rows = min(CT, it.m - s.m0);
  s.mh = (s.rows + 127) >> 7;
  s.ncols = min(CT, (it.n - s.n0 + 15) & ~15);
  s.nk = (it.k + CK - 1) / CK;
  return s;
}

__global__ void __launch_bounds__(C_THREADS, 1)
gemm_chain_kernel(const ChainItemDev* __restrict__ items, const ChainTile* __restrict__ tiles, int ntiles, int* __restrict__ cursor,
                  const CUtensorMap* __restrict__ maps, int* __restrict__ counters, long long* __restrict__ dbg, int flags) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (s32(raw) + 1023u) & ~1023u;
  const uint32_t epi = sbase + C_NSTAGE * C_STAGE;
  const uint32_t bars = epi + C_EPI;                         // full[3] | empty[3] | tmem_full | tmem_empty
  const uint32_t bar_tfull = bars + 8 * (2 * C_NSTAGE), bar_tempty = bar_tfull + 8;
  // the CTA's tile queue: a ring of C_NSCHED entries (tile index, -1 = no more work) filled by the fetcher (one lane of
  // producer warp 0, atomicAdd on the global cursor) and read by all twelve warps
  const uint32_t sched_full = bar_tempty + 8, sched_empty = sched_full + 8 * C_NSCHED;
  const uint32_t sched_idx = sched_empty + 8 * C_NSCHED;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw + (sbase - s32(raw)) + C_NSTAGE * C_STAGE + C_EPI + 8 * (2 * C_NSTAGE + 2) +
                                                    8 * (2 * C_NSCHED) + 4 * C_NSCHED);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = dbg ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C_NSTAGE; ++s) {
      bar_init(bars + 8 * s, C_NPROD);
      bar_init(bars + 8 * (C_NSTAGE + s), 1);
    }
    bar_init(bar_tfull, 1);
    bar_init(bar_tempty, 8);
    for (int i = 0; i < C_NSCHED; ++i) {
      bar_init(sched_full + 8 * i, 1);
      bar_init(sched_empty + 8 * i, C_THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // Dynamic schedule: the tiles of the call are ONE list in a dependency-respecting order (first GEMMs, then the GEMMs that
  // consume them; the most expensive first within each) and every CTA takes the next unclaimed one.  Whatever a tile
  // really costs -- cold operands from HBM, a dependency that is late, a plain-store epilogue -- the CTAs finish within
  // one tile of each other (the static longest-first deal left them 416 k .. 522 k cycles apart), and a tile only ever
  // waits for tiles that were claimed before it: no deadlock.
  auto next_tile = [&](int it, bool fetcher) -> int {
    const uint32_t slot = (uint32_t)it % C_NSCHED, ph = ((uint32_t)it / C_NSCHED) & 1u;
    if (fetcher) {
      bar_wait(sched_empty + 8 * slot, ph ^ 1u);            // every warp has read the entry this one replaces
      if (lane == 0) {
        int idx = atomicAdd(cursor, 1);
        if (idx >= ntiles) idx = -1;
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(sched_idx + 4 * slot), "r"(idx) : "memory");
        bar_arrive(sched_full + 8 * slot);
      }
    }
    bar_wait(sched_full + 8 * slot, ph);
    int idx;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(idx) : "r"(sched_idx + 4 * slot) : "memory");
    __syncwarp();
    if (lane == 0) bar_arrive(sched_empty + 8 * slot);
    return idx;
  };

  if (warp == 0 || warp == 6 || warp == 7) {
    // ---- TMA producers ----
    const uint32_t leader = elect_one();
    const int me = (int)uni((uint32_t)(warp == 0 ? 0 : warp - 5));       // 0 .. 2
    uint32_t git = 0;                                                     // stage iterations since kernel start
    for (int it_ = 0;; ++it_) {
      const int ti = next_tile(it_, warp == 0);
      if (ti < 0) break;
      const ChainTile t = tiles[ti];
      const ChainItemDev& it = items[t.item];
      const TileShape sh = tile_shape(it, t);
      const CUtensorMap* ma = maps + it.mapA;
      const CUtensorMap* mb = maps + it.mapB;
      if (leader) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(ma)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(mb)) : "memory");
        if (it.mapC >= 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps + it.mapC)) : "memory");
      }
      long long t_dep0 = dbg ? clock64() : 0;
      if (it.dep_counters >= 0) {
        // rows [m0, m0 + 256) of the A operand are rows [m0 / div, (m0 + 255) / div] of the producer GEMM's output:
        // written by its row tiles lo .. hi (one tile when div = 1)
        const int lo = (sh.m0 / it.dep_div) / CT;
        const int hi = min(it.dep_rows - 1, (sh.m0 + sh.rows - 1) / it.dep_div) / CT;
        for (int rt = lo; rt <= hi; ++rt) {
          const int* c = counters + it.dep_counters + rt;
          uint32_t spins = 0;
          while (ld_acquire(c) < it.dep_need) {
            __nanosleep(64);
            if (++spins > (1u << 24)) asm volatile("trap;");
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy writes of the other CTAs -> this CTA's TMA reads
      }
      if (dbg && me == 0 && lane == 0) dbg[8 * blockIdx.x + 0] += clock64() - t_dep0;
      // Read-modify-write epilogues (`C += acc^2`, `mean + sample`) read a 256 x 256 region whose rows are `ld` floats
      // apart: one DRAM page per 1 KB row segment.  Left to the drain those reads stall the epilogue warps while the tensor
      // pipe idles (drain as long as the MMAs in the per-CTA counters); the last producer warp instead asks for the tile's
      // rows with L2 bulk prefetches NOW, a whole K loop before the drain needs them.
      if (me == C_NPROD - 1 && it.epi != EPI_STORE && it.mapC < 0) {
        const float* src = it.epi == EPI_SQUARE_ACCUM ? it.C : it.se.mu_w;
        const int ld = it.epi == EPI_SQUARE_ACCUM ? it.ldc : it.se.K0;
        const int ncol = min(sh.ncols, (it.epi == EPI_SQUARE_ACCUM ? it.n : it.se.K0) - sh.n0);
        if (src != nullptr && ncol >= 4 && (ld & 3) == 0 && ((uintptr_t)src & 15) == 0) {
          const uint32_t bytes = (uint32_t)(ncol & ~3) * 4u;
          for (int r = lane; r < sh.rows; r += 32) {
            const float* p = src + (size_t)(sh.m0 + r) * ld + sh.n0;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
          }
        }
      }
      // boxes of one stage: A then B; K-major 16 KB boxes of 128 rows, MN-major 4 KB boxes of 32 rows
      const int na = it.a_mn ? (sh.rows + 31) / 32 : sh.mh;
      const int nb = it.b_mn ? (sh.ncols + 31) / 32 : (sh.ncols + 127) / 128;
      const uint32_t ba = it.a_mn ? 4096u : 16384u, bb = it.b_mn ? 4096u : 16384u;
      const int total = na + nb;
      uint32_t mine = 0;
      for (int e = me; e < total; e += C_NPROD) mine += e < na ? ba : bb;
      const int u_na = (int)uni((uint32_t)na), u_total = (int)uni((uint32_t)total), u_nk = (int)uni((uint32_t)sh.nk);
      const int a_mn = (int)uni((uint32_t)it.a_mn), b_mn = (int)uni((uint32_t)it.b_mn);
      for (int kb = 0; kb < u_nk; ++kb, ++git) {
        const uint32_t s = git % C_NSTAGE, ph = (git / C_NSTAGE) & 1u;
        bar_wait(bars + 8 * (C_NSTAGE + s), ph ^ 1u);
        if (leader) {
          if (mine) bar_expect(bars + 8 * s, mine); else bar_arrive(bars + 8 * s);
        }
        const uint32_t sa = sbase + s * C_STAGE, sb = sa + 32768u;
        const int k0 = kb * CK;
        for (int e = me; e < u_total; e += C_NPROD) {
          if (e < u_na) {
            if (leader) {
              if (a_mn) tma_2d(sa + (uint32_t)e * 4096u, ma, sh.m0 + 32 * e, k0, bars + 8 * s);
              else tma_2d(sa + (uint32_t)e * 16384u, ma, k0, sh.m0 + 128 * e, bars + 8 * s);
            }
          } else {
            const int q = e - u_na;
            if (leader) {
              if (b_mn) tma_2d(sb + (uint32_t)q * 4096u, mb, sh.n0 + 32 * q, k0, bars + 8 * s);
              else tma_2d(sb + (uint32_t)q * 16384u, mb, k0, sh.n0 + 128 * q, bars + 8 * s);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (uniform control flow, one elected lane issues) ----
    const uint32_t leader = elect_one();
    const uint32_t u_tmem = uni(tmem);
    uint32_t git = 0;
    int ntile = 0;
    for (;; ++ntile) {
      const int ti = (int)uni((uint32_t)(next_tile(ntile, false) + 1)) - 1;
      if (ti < 0) break;
      const ChainTile t = tiles[ti];
      const ChainItemDev& it = items[t.item];
      const TileShape sh = tile_shape(it, t);
      const uint32_t a_mn = uni((uint32_t)it.a_mn), b_mn = uni((uint32_t)it.b_mn);
      const uint32_t idesc = uni((1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) |
                                 ((uint32_t)(sh.ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
      const uint64_t da = a_mn ? desc_mn(0u, 4096u) : desc_k(0u), db = b_mn ? desc_mn(0u, 4096u) : desc_k(0u);
      const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
      const uint32_t a_step = a_mn ? (1024u >> 4) : (32u >> 4), b_step = b_mn ? (1024u >> 4) : (32u >> 4);
      const bool two = uni((uint32_t)sh.mh) == 2u;
      const int u_nk = (int)uni((uint32_t)sh.nk);
      if (ntile > 0) {                                       // the previous tile's accumulator has been drained
        const long long t0 = dbg ? clock64() : 0;
        bar_wait(bar_tempty, (uint32_t)(ntile - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg && lane == 0) dbg[8 * blockIdx.x + 1] += clock64() - t0;
      }
      uint32_t acc = 0;
      for (int kb = 0; kb < u_nk; ++kb, ++git) {
        const uint32_t s = git % C_NSTAGE, ph = (git / C_NSTAGE) & 1u;
        const long long t0 = dbg ? clock64() : 0;
        bar_wait(bars + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg && lane == 0) dbg[8 * blockIdx.x + 2] += clock64() - t0;
        const uint32_t sa = sbase + s * C_STAGE, sb = sa + 32768u;
        const uint32_t a0 = a_lo | ((sa >> 4) & 0x3FFFu), a1 = a_lo | (((sa + 16384u) >> 4) & 0x3FFFu);
        const uint32_t b0 = b_lo | ((sb >> 4) & 0x3FFFu);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                         "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, q;\n\t}"
                         ::"r"(u_tmem), "r"(a0 + (uint32_t)ks * a_step), "r"(a_hi), "r"(b0 + (uint32_t)ks * b_step), "r"(b_hi),
                           "r"(idesc), "r"(ks == 0 ? acc : 1u) : "memory");
            if (two)
              asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                           "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                           "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, q;\n\t}"
                           ::"r"(u_tmem + 256u), "r"(a1 + (uint32_t)ks * a_step), "r"(a_hi), "r"(b0 + (uint32_t)ks * b_step),
                             "r"(b_hi), "r"(idesc), "r"(ks == 0 ? acc : 1u) : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                       ::"r"(bars + 8 * (C_NSTAGE + s)) : "memory");
        }
        acc = 1;
      }
      if (leader)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull) : "memory");
    }
    __syncwarp();
  } else {
    // ---- epilogue: EIGHT warps (2..5 and 8..11), two per TMEM lane quadrant (= warp & 3), each pair splitting the
    // tile's 32-column blocks between them.  The accumulator drain is not overlapped with the next tile's MMAs (one
    // accumulator fills TMEM), so it is on the critical path of every tile: the per-CTA cycle counters showed it as long
    // as the MMAs themselves with four warps and load-after-drain read-modify-writes.  Per 32 x 32 block a warp now
    // (1) issues the global loads of a read-modify-write epilogue FIRST (they do not depend on the accumulator),
    // (2) moves the block TMEM -> registers -> swizzled shared memory, (3) reads it back row-wise, combines and stores
    // with 16-byte accesses where rows are 16-byte aligned.
    const int quad = warp & 3;
    const int eset = warp >= 8 ? 1 : 0;
    const uint32_t stg = epi + (uint32_t)(eset * 4 + quad) * 4096u;
    const int sub = lane >> 3, ch = lane & 7;               // read-back role: row within a group of 4, 16-byte chunk
    int ntile = 0;
    for (;; ++ntile) {
      const int ti = next_tile(ntile, false);
      if (ti < 0) break;
      const ChainTile t = tiles[ti];
      const ChainItemDev& it = items[t.item];
      const TileShape sh = tile_shape(it, t);
      // the item's fields live in global memory: fetch what the epilogue needs once per tile
      const int e_m = it.m, e_n = it.n, e_ldc = it.ldc, e_kind = it.epi, e_round = it.round_out, e_K0 = it.se.K0;
      const float e_alpha = it.alpha;
      float* const e_C = it.C;
      float* const e_wout = it.se.w_out; float* const e_bout = it.se.b_out; float* const e_sout = it.se.s_out;
      const float* const e_muw = it.se.mu_w; const float* const e_mub = it.se.mu_b;
      const int e_mapC = it.mapC;
      const long long t_e0 = dbg ? clock64() : 0;
      bar_wait(bar_tfull, (uint32_t)ntile & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long t_e1 = dbg ? clock64() : 0;
      // this warp's 32 x 32 blocks of the tile: (row half h, column block cc), cc = eset * 32, + 64, ...
      const int per_h = sh.ncols > eset * 32 ? (sh.ncols - eset * 32 + 63) / 64 : 0;
      const int nblk = sh.mh * per_h;
      const bool tma_w = e_mapC >= 0 && e_kind == 2;      // sample epilogue: loads by the warps, the weight block by TMA
      if (e_mapC >= 0 && e_kind != 2) {
        // ---- TMA epilogue (outputs with 16-byte aligned rows).  The per-CTA phase counters put 57 % of the accumulator
        // drain on the warps' global STOREs and 21 % on the loads of `C += acc^2` (profiles/r2_chain_ablation.txt); the
        // TMEM reads and the staging tile cost 3 %.  So the warps no longer touch global memory: a block goes
        // TMEM -> registers -> (scale / round / square) -> the SWIZZLE_128B staging tile, and one lane hands the 4 KB
        // tile to the TMA unit -- a tensor store, or for the accumulating epilogue a tensor REDUCE-ADD that L2 applies,
        // so the old values never travel to the SM at all.  Rows and columns past the matrix edge are clipped by the map.
        const CUtensorMap* mc = maps + e_mapC;
        for (int idx = 0; idx < nblk; ++idx) {
          const int h = idx / per_h, cc = eset * 32 + (idx - h * per_h) * 64;
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
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(a[j]);
            float y;
            if (e_kind == EPI_SQUARE_ACCUM) y = x * x;
            else { y = e_alpha * x; if (e_round) y = rna_tf32(y); }
            a[j] = __float_as_uint(y);
          }
          // the previous block's tensor store has finished READING the staging tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((k ^ (lane & 7)) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[4 * k]), "r"(a[4 * k + 1]),
                         "r"(a[4 * k + 2]), "r"(a[4 * k + 3]) : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          const int c0 = sh.n0 + cc, r0 = sh.m0 + h * 128 + quad * 32;
          if (lane == 0 && c0 < e_n && r0 < e_m) {
            if (e_kind == EPI_SQUARE_ACCUM)
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(reinterpret_cast<uint64_t>(mc)), "r"(c0), "r"(r0), "r"(stg) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(reinterpret_cast<uint64_t>(mc)), "r"(c0), "r"(r0), "r"(stg) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        // all of this warp's tensor stores are COMPLETE (written, not merely read) before the tile is published and
        // before the staging tile is used by the generic path of a later tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
      } else {
      // where this lane's 4-column group of block (h, cc) lives in global memory
      struct Where { float* dst; const float* src; int ld, nv, row0, gn0; bool vec; };
      auto where = [&](int idx) {
        Where w;
        const int h = idx / per_h, cc = eset * 32 + (idx - h * per_h) * 64;
        w.row0 = sh.m0 + h * 128 + quad * 32;
        w.gn0 = sh.n0 + cc + ch * 4;
        w.dst = nullptr; w.src = nullptr; w.ld = 0; w.nv = 0;
        if (w.gn0 < e_n) {
          if (e_kind == 2) {
            if (w.gn0 < e_K0) { w.dst = e_wout; w.src = e_muw; w.ld = e_K0; w.nv = min(4, e_K0 - w.gn0); }
          } else {
            w.dst = e_C; w.src = e_kind == EPI_SQUARE_ACCUM ? e_C : nullptr; w.ld = e_ldc; w.nv = min(4, e_n - w.gn0);
          }
        }
        w.vec = w.nv == 4 && (w.ld & 3) == 0 && w.dst != nullptr && (((uintptr_t)w.dst) & 15) == 0;
        return w;
      };
      // (1) the old values of a read-modify-write epilogue do not depend on the accumulator: the loads of block i + 1 are
      // issued before block i is drained, so their latency hides behind the TMEM / shared-memory phase of block i
      auto load_old = [&](const Where& w, float (&old)[8][4]) {
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          old[r4][0] = old[r4][1] = old[r4][2] = old[r4][3] = 0.f;
          const int gm = w.row0 + r4 * 4 + sub;
          if (w.src != nullptr && gm < e_m) {
            const float* p = w.src + (size_t)gm * w.ld + w.gn0;
            if (w.vec) {
              const float4 t4 = *reinterpret_cast<const float4*>(p);
              old[r4][0] = t4.x; old[r4][1] = t4.y; old[r4][2] = t4.z; old[r4][3] = t4.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (j < w.nv) old[r4][j] = p[j];
            }
          }
        }
      };
      float old[8][4], nxt[8][4];
      if (nblk > 0) load_old(where(0), old);
      for (int idx = 0; idx < nblk; ++idx) {
        const Where w = where(idx);
        const int h = idx / per_h, cc = eset * 32 + (idx - h * per_h) * 64;
        if (idx + 1 < nblk && !(flags & 4)) load_old(where(idx + 1), nxt);
        // (2) TMEM -> registers -> swizzled staging tile
        if (!(flags & 2)) {
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
          if (tma_w && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tensor store
          __syncwarp();                                     // the previous block has been read back
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((k ^ (lane & 7)) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[4 * k]), "r"(a[4 * k + 1]),
                         "r"(a[4 * k + 2]), "r"(a[4 * k + 3]) : "memory");
          }
          __syncwarp();
        }
        // (3) read back row-wise, combine, store
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          const int r = r4 * 4 + sub;
          float v[4];
          const uint32_t addr = stg + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) * 16);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
          const int gm = w.row0 + r;
          const bool live = gm < e_m && w.gn0 < e_n && !(flags & 1);
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (e_kind == EPI_STORE) { const float y = e_alpha * v[j]; o[j] = e_round ? rna_tf32(y) : y; }
            else if (e_kind == EPI_SQUARE_ACCUM) o[j] = old[r4][j] + v[j] * v[j];
            else o[j] = old[r4][j] + e_alpha * v[j];
          }
          if (tma_w) {                           // mean + sample goes back into the staging tile (same swizzled slot)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
          } else if (live && w.dst != nullptr) {
            float* p = w.dst + (size_t)gm * w.ld + w.gn0;
            if (w.vec) *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (j < w.nv) p[j] = o[j];
            }
          }
          if (live && e_kind == 2) {
            if (e_sout) {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (w.gn0 + j < e_n) e_sout[(size_t)gm * e_n + w.gn0 + j] = e_alpha * v[j];
            }
            if (e_bout) {                      // the bias column K0 (at most one per row) may sit in this group
              const int jb = e_K0 - w.gn0;
              if (jb >= 0 && jb < 4 && e_K0 < e_n) {
                float xb = v[0];
                if (jb == 1) xb = v[1]; else if (jb == 2) xb = v[2]; else if (jb == 3) xb = v[3];
                e_bout[gm] = __ldg(e_mub + gm) + e_alpha * xb;
              }
            }
          }
        }
        if (tma_w) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          const int c0 = sh.n0 + cc, r0 = sh.m0 + h * 128 + quad * 32;
          if (lane == 0 && c0 < e_K0 && r0 < e_m) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(reinterpret_cast<uint64_t>(maps + e_mapC)), "r"(c0), "r"(r0), "r"(stg) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4)
#pragma unroll
          for (int j = 0; j < 4; ++j) old[r4][j] = nxt[r4][j];
      }
      if (tma_w) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(bar_tempty);
      if (dbg && warp == 2 && lane == 0) {
        dbg[8 * blockIdx.x + 3] += t_e1 - t_e0;            // epilogue waiting for the MMAs of the tile
        dbg[8 * blockIdx.x + 4] += clock64() - t_e1;       // accumulator drain
        dbg[8 * blockIdx.x + 5] += 1;                      // tiles
      }
      if (it.counters >= 0) {
        // publish: every epilogue thread's stores -> gpu scope, then ONE increment of the row tile's counter
        __threadfence();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (warp == 2 && lane == 0) atomicAdd(counters + it.counters + t.tm, 1);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[8 * blockIdx.x + 6] = clock64() - t_start;
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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

// rows x depth operand with element strides (s_row, s_k): which form it has and (map != null) its tensor map
bool operand_form(const float* base, int rows, int depth, long long s_row, long long s_k, int& mn, CUtensorMap* map) {
  if (((uintptr_t)base & 15) != 0) return false;
  cuuint64_t gd[2], gs[1];
  cuuint32_t box[2], es[2] = {1, 1};
  CUtensorMapSwizzle sw;
  if (s_k == 1 && (s_row % 4) == 0 && s_row >= depth) {          // contraction-contiguous
    mn = 0;
    gd[0] = (cuuint64_t)depth; gd[1] = (cuuint64_t)rows; gs[0] = (cuuint64_t)s_row * 4;
    box[0] = CK; box[1] = 128;
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  } else if (s_row == 1 && (s_k % 4) == 0 && s_k >= rows) {      // row-contiguous
    mn = 1;
    gd[0] = (cuuint64_t)rows; gd[1] = (cuuint64_t)depth; gs[0] = (cuuint64_t)s_k * 4;
    box[0] = 32; box[1] = CK;
    sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  } else {
    return false;
  }
  if (!map) return true;
  return encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// m x n row-major output with leading dimension ldc: boxes of [32 columns][32 rows] in the staging tile's SWIZZLE_128B form
bool output_map(float* C, int m, int n, int ldc, CUtensorMap* map) {
  if (C == nullptr || ((uintptr_t)C & 15) != 0 || (ldc & 3) != 0 || ldc < n) return false;
  cuuint64_t gd[2] = {(cuuint64_t)n, (cuuint64_t)m}, gs[1] = {(cuuint64_t)ldc * 4};
  cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  return encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)C, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

bool gemm_chain_supported(const ChainGemm& g) {
  int mn;
  return encoder() != nullptr && g.m > 0 && g.n > 0 && g.k > 0 &&
         operand_form(g.A, g.m, g.k, g.sa_m, g.sa_k, mn, nullptr) && operand_form(g.B, g.n, g.k, g.sb_n, g.sb_k, mn, nullptr);
}

size_t gemm_chain_workspace(const ChainGemm* gemms, int count) {
  size_t tiles = 0, rows = 0;
  for (int i = 0; i < count; ++i) {
    const size_t tm = (gemms[i].m + CT - 1) / CT, tn = (gemms[i].n + CT - 1) / CT;
    tiles += tm * tn;
    rows += tm;
  }
  return al(sizeof(ChainItemDev) * (size_t)count) + al(sizeof(ChainTile) * tiles) + al(sizeof(int) * (SK_MAX_CTAS + 1)) +
         al(sizeof(CUtensorMap) * 3 * (size_t)count) + al(sizeof(int) * rows) + 1024;
}

namespace {

bool same_gemm(const ChainGemm& a, const ChainGemm& b) {
  return a.A == b.A && a.sa_m == b.sa_m && a.sa_k == b.sa_k && a.B == b.B && a.sb_k == b.sb_k && a.sb_n == b.sb_n &&
         a.C == b.C && a.ldc == b.ldc && a.m == b.m && a.n == b.n && a.k == b.k && a.alpha == b.alpha && a.epi == b.epi &&
         a.round_out == b.round_out && a.se.mu_w == b.se.mu_w && a.se.mu_b == b.se.mu_b && a.se.w_out == b.se.w_out &&
         a.se.b_out == b.se.b_out && a.se.s_out == b.se.s_out && a.se.K0 == b.se.K0 && a.se.has_bias == b.se.has_bias &&
         a.dep == b.dep && a.dep_div == b.dep_div && a.dep_count == b.dep_count;
}

// The launch tables of one call (items, tile lists, tensor maps, zeroed counters) in the byte layout of the workspace
// prefix, kept in PINNED host memory: a call with the same GEMM list (the steady state of EFB.update and of the sampling
// loop: same factors, same buffers) is one asynchronous copy and one launch -- no tensor-map encoding, no scheduling.
struct TableCache {
  std::vector<ChainGemm> key;
  int sms = 0, grid = 0, dev = -1, ntiles = 0;
  char* pinned = nullptr;
  size_t bytes = 0, cap = 0, counters_off = 0;
  cudaEvent_t used = nullptr;      // the last copy out of `pinned` has executed
  unsigned long long stamp = 0;
};
constexpr int N_TABLE_CACHE = 6;
TableCache g_tables[N_TABLE_CACHE];
unsigned long long g_table_clock = 0;

}  // namespace

// One launch for `count` GEMMs; gemms[i].dep (if >= 0) names an EARLIER GEMM of the call whose output C is this one's A
// operand (same m, same row tiling).  Returns -1 if an operand cannot be fed by TMA.  (Callers hold the ApiGuard lock.)
int gemm_chain_launch(const ChainGemm* gemms, int count, void* ws, size_t ws_bytes, cudaStream_t s) {
  CRV_CHECK(gemms && count > 0, "empty GEMM chain");
  CRV_CHECK(ws && ws_bytes >= gemm_chain_workspace(gemms, count), "workspace too small: %zu < %zu", ws_bytes,
            gemm_chain_workspace(gemms, count));
  CRV_CHECK(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  if (!encoder()) return -1;
  const int sms = device_sm_count();
  CRV_CHECK(sms > 0, "no CUDA device");
  const int G = std::min(sms, (int)SK_MAX_CTAS);
  static const bool use_cache = !(getenv("CURVATURE_B200_PLAN_CACHE") && atoi(getenv("CURVATURE_B200_PLAN_CACHE")) == 0);
  static const int dbg_flags = getenv("CURVATURE_B200_CHAIN_DBG") ? atoi(getenv("CURVATURE_B200_CHAIN_DBG")) : 0;   // (profiling ablations)
  static bool attr = false;
  if (!attr) {
    attr = true;
    CRV_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
  }
  char* base = (char*)ws;
  auto launch = [&](const TableCache& tc) -> int {
    size_t off = 0;
    const ChainItemDev* d_items = (const ChainItemDev*)(base + off); off += al(sizeof(ChainItemDev) * (size_t)count);
    const ChainTile* d_tiles = (const ChainTile*)(base + off);
    int* d_cursor = (int*)(base + tc.counters_off - al(sizeof(CUtensorMap) * 3 * (size_t)count) - al(sizeof(int) * (SK_MAX_CTAS + 1)));
    const CUtensorMap* d_maps = (const CUtensorMap*)(base + tc.counters_off - al(sizeof(CUtensorMap) * 3 * (size_t)count));
    int* d_counters = (int*)(base + tc.counters_off);
    gemm_chain_kernel<<<tc.grid, C_THREADS, C_SMEM, s>>>(d_items, d_tiles, tc.ntiles, d_cursor, d_maps, d_counters,
                                                         debug_timeline_buffer(), dbg_flags);
    CRV_CUDA(cudaGetLastError());
    return 0;
  };
  TableCache* slot = nullptr;
  int dev = 0;
  CRV_CUDA(cudaGetDevice(&dev));
  if (use_cache) {
    for (auto& tc : g_tables) {
      if (tc.pinned == nullptr || tc.dev != dev || tc.sms != sms || (int)tc.key.size() != count) continue;
      bool same = true;
      for (int i = 0; i < count && same; ++i) same = same_gemm(tc.key[i], gemms[i]);
      if (!same) continue;
      tc.stamp = ++g_table_clock;
      CRV_CUDA(cudaMemcpyAsync(base, tc.pinned, tc.bytes, cudaMemcpyHostToDevice, s));
      CRV_CUDA(cudaEventRecord(tc.used, s));
      return launch(tc);
    }
    slot = &g_tables[0];
    for (auto& tc : g_tables) if (tc.stamp < slot->stamp) slot = &tc;
  }
  std::vector<ChainItemDev> items(count);
  std::vector<CUtensorMap> maps(3 * (size_t)count);
  static const bool tma_out = !(getenv("CURVATURE_B200_CHAIN_TMA_OUT") && atoi(getenv("CURVATURE_B200_CHAIN_TMA_OUT")) == 0);
  std::vector<int> tm_of(count), tn_of(count);
  int ncounters = 0;
  std::vector<int> counter_base(count, -1);
  for (int i = 0; i < count; ++i) {
    const ChainGemm& g = gemms[i];
    CRV_CHECK(g.A && g.B && g.m > 0 && g.n > 0 && g.k > 0, "bad GEMM %d: %d x %d x %d", i, g.m, g.n, g.k);
    CRV_CHECK(g.dep < i, "GEMM %d depends on a later one", i);
    ChainItemDev& d = items[i];
    memset(&d, 0, sizeof(d));
    if (!operand_form(g.A, g.m, g.k, g.sa_m, g.sa_k, d.a_mn, &maps[2 * i])) return -1;
    if (!operand_form(g.B, g.n, g.k, g.sb_n, g.sb_k, d.b_mn, &maps[2 * i + 1])) return -1;
    d.m = g.m; d.n = g.n; d.k = g.k; d.ldc = g.ldc; d.mapA = 2 * i; d.mapB = 2 * i + 1;
    d.mapC = -1;
    if (tma_out && g.epi != 2 && output_map(g.C, g.m, g.n, g.ldc, &maps[2 * (size_t)count + i])) d.mapC = 2 * count + i;
    if (tma_out && g.epi == 2 && g.se.w_out && g.se.mu_w && !g.se.s_out && ((uintptr_t)g.se.mu_w & 15) == 0 &&
        output_map(g.se.w_out, g.m, g.se.K0, g.se.K0, &maps[2 * (size_t)count + i]))
      d.mapC = 2 * count + i;
    d.epi = g.epi; d.round_out = g.round_out; d.alpha = g.alpha; d.C = g.C; d.se = g.se;
    if (g.epi == 2) CRV_CHECK(g.se.mu_w || g.se.s_out || !g.se.w_out, "sample epilogue needs its descriptor");
    else CRV_CHECK(g.C != nullptr, "null GEMM output");
    tm_of[i] = (g.m + CT - 1) / CT; tn_of[i] = (g.n + CT - 1) / CT;
    d.dep_counters = -1; d.counters = -1; d.dep_div = 1; d.dep_rows = g.m;
  }
  for (int i = 0; i < count; ++i)
    if (gemms[i].dep >= 0) {
      const int p = gemms[i].dep;
      const int div = gemms[i].dep_div > 1 ? gemms[i].dep_div : 1;
      const int np = gemms[i].dep_count > 1 ? gemms[i].dep_count : 1;
      CRV_CHECK(p + np <= i, "GEMM %d depends on a later one", i);
      CRV_CHECK(gemms[p].C == gemms[i].A, "GEMM %d: its producer %d does not write its A operand", i, p);
      int need = 0;
      for (int j = p; j < p + np; ++j) {       // the producers of one consumer share one set of per-row-tile counters
        CRV_CHECK(gemms[j].m * div == gemms[i].m, "GEMM %d: producer %d has %d rows, expected %d", i, j, gemms[j].m, gemms[i].m / div);
        CRV_CHECK(counter_base[j] < 0 || (j > p && counter_base[j] == counter_base[p]), "GEMM %d feeds two consumers", j);
        if (j == p) { if (counter_base[p] < 0) { counter_base[p] = ncounters; ncounters += tm_of[p]; } }
        else counter_base[j] = counter_base[p];
        items[j].counters = counter_base[p];
        need += tn_of[j];
      }
      items[i].dep_counters = counter_base[p];
      items[i].dep_need = need;
      items[i].dep_div = div;
      items[i].dep_rows = gemms[p].m;
    }
  // global order: producers (GEMMs somebody waits for, and independent ones) first, then consumers; within each phase the
  // most expensive tiles first.  The CTAs claim tiles from this list at run time (see next_tile in the kernel).
  struct T { int item, tm, tn, phase; double cost; };
  std::vector<T> all;
  std::vector<int> depth(count, 0);                // length of the dependency chain behind a GEMM (producers come earlier)
  for (int i = 0; i < count; ++i)
    if (gemms[i].dep >= 0) {
      const int np = gemms[i].dep_count > 1 ? gemms[i].dep_count : 1;
      for (int j = gemms[i].dep; j < gemms[i].dep + np; ++j) depth[i] = std::max(depth[i], depth[j] + 1);
    }
  for (int i = 0; i < count; ++i)
    for (int a = 0; a < tm_of[i]; ++a)
      for (int b = 0; b < tn_of[i]; ++b) {
        const int rows = std::min(CT, gemms[i].m - a * CT), cols = std::min(CT, gemms[i].n - b * CT);
        // cycles: MMA (4 k-steps x mh instructions of N/256 x 128 clocks per 32-deep stage) + accumulator drain (per warp
        // mh x N/32 blocks of 32 x 32; about half as long through the TMA unit) + fixed
        const int mh = (rows + 127) / 128, nc = (cols + 15) / 16 * 16;
        const double per_blk = items[i].mapC >= 0 ? 700.0 : 1500.0;
        const double c = (double)((gemms[i].k + CK - 1) / CK) * mh * 2.0 * std::max(nc, 64) + (double)mh * ((nc + 31) / 32) * per_blk + 3000.0;
        all.push_back({i, a, b, depth[i], c});
      }
  std::stable_sort(all.begin(), all.end(), [](const T& x, const T& y) {
    if (x.phase != y.phase) return x.phase < y.phase;
    return x.cost > y.cost;
  });
  std::vector<ChainTile> tiles;
  for (const T& t : all) tiles.push_back({t.item, t.tm, t.tn, 0});
  std::vector<int> begin(SK_MAX_CTAS + 1, 0);      // (first word: the tile cursor, zero at launch)
  // the workspace prefix: items | tiles | CTA ranges | tensor maps | counters (zero)
  const size_t o_items = 0, o_tiles = o_items + al(sizeof(ChainItemDev) * (size_t)count);
  const size_t o_begin = o_tiles + al(sizeof(ChainTile) * tiles.size());
  const size_t o_maps = o_begin + al(sizeof(int) * (SK_MAX_CTAS + 1));
  const size_t o_counters = o_maps + al(sizeof(CUtensorMap) * maps.size());
  const size_t total = o_counters + al(sizeof(int) * (size_t)std::max(ncounters, 1));
  TableCache local;
  TableCache& tc = slot ? *slot : local;
  std::vector<char> pageable;
  char* host;
  if (slot) {
    if (tc.used) {
      CRV_CUDA(cudaEventSynchronize(tc.used));                   // an earlier copy may still be reading the old tables
      if (tc.dev != dev) { cudaEventDestroy(tc.used); tc.used = nullptr; }      // (an event belongs to its device)
    }
    tc.key.clear();                                              // (not a valid entry until it is filled again below)
    if (!tc.used) CRV_CUDA(cudaEventCreateWithFlags(&tc.used, cudaEventDisableTiming));
    tc.dev = dev;
    if (tc.cap < total) {
      if (tc.pinned) CRV_CUDA(cudaFreeHost(tc.pinned));
      tc.pinned = nullptr; tc.cap = 0;
      CRV_CUDA(cudaMallocHost((void**)&tc.pinned, total));
      tc.cap = total;
    }
    host = tc.pinned;
  } else {
    pageable.resize(total);
    host = pageable.data();
  }
  memset(host, 0, total);
  memcpy(host + o_items, items.data(), sizeof(ChainItemDev) * (size_t)count);
  memcpy(host + o_tiles, tiles.data(), sizeof(ChainTile) * tiles.size());
  memcpy(host + o_begin, begin.data(), sizeof(int) * (SK_MAX_CTAS + 1));
  memcpy(host + o_maps, maps.data(), sizeof(CUtensorMap) * maps.size());
  tc.sms = sms; tc.grid = std::min(G, (int)tiles.size()); tc.bytes = total; tc.counters_off = o_counters; tc.ntiles = (int)tiles.size();
  if (slot) { tc.key.assign(gemms, gemms + count); tc.stamp = ++g_table_clock; }
  CRV_CUDA(cudaMemcpyAsync(base, host, total, cudaMemcpyHostToDevice, s));   // (pageable: staged before the call returns)
  if (slot) CRV_CUDA(cudaEventRecord(tc.used, s));
  return launch(tc);
}

}  // namespace crv
