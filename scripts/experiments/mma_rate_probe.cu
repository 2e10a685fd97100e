// Probe: issue rate of tcgen05.mma.cta_group::1 (M = 128) from shared-memory operands as a function of element
// type (bf16 / tf32), operand major-ness (K-major SWIZZLE_128B vs MN-major, the layouts syrk_tc.cu uses), N and the
// number of row halves that share one B operand.  One thread per CTA issues `iters` k-steps back to back over a
// ring of shared-memory stages (contents are arbitrary finite numbers), then commits and waits; the elapsed
// clock64() / #MMA is the sustained cost of one instruction.  No TMA traffic: this is the tensor pipe + its
// shared-memory operand fetch alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu && ./mma_rate_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0; int spins = 0;
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (!ok && ++spins > (1 << 26)) asm volatile("trap;");
  }
}
// K-major SWIZZLE_128B: SBO = 1024 B between 8-row groups
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major 32-bit (SW128 atom 32B, layout type 1): SBO 512 B, LBO = chunk stride
__device__ __forceinline__ uint64_t desc_mn32(uint32_t a, uint32_t lbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// MN-major 16-bit (SW128, type 2): SBO 1024 B, LBO = chunk stride
__device__ __forceinline__ uint64_t desc_mn16(uint32_t a, uint32_t lbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
struct P {
  int bf16, mn, N, halves, iters, kpos_per_stage, nstage;
  uint32_t stage_bytes, chunk_bytes;
  int same_addr;       // 1: every k-step reads the same shared-memory bytes
};
__global__ void __launch_bounds__(128) probe(const P p, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t sb = (s32(raw) + 1023u) & ~1023u;
  // finite contents
  for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 4; i += blockDim.x) {
    const uint32_t v = p.bf16 ? 0x3C003C00u + ((i * 2654435761u) >> 28) * 0x00010001u : 0x3C000000u + ((i * 2654435761u) >> 12 & 0xFF000u);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + i * 4), "r"(v) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    uint32_t idesc;
    if (p.bf16) idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)p.N >> 3) << 17) | ((128u >> 4) << 24);
    else idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)p.N >> 3) << 17) | ((128u >> 4) << 24);
    if (p.mn) idesc |= (1u << 15) | (1u << 16);
    const int CH = p.bf16 ? 64 : 32;
    const int KPOS = p.bf16 ? 16 : 8;
    const long long t0 = clock64();
    uint32_t acc = 0;
    int kg = 0, s = 0;
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t st = sb + (p.same_addr ? 0u : (uint32_t)s * p.stage_bytes);
      const int kk = p.same_addr ? 0 : kg;
      uint64_t bd;
      uint32_t a0;
      if (p.mn) {
        // stage = [A chunks (halves*128/CH)] [B chunks (N/CH)], chunk = kpos_per_stage rows x 128 B
        const uint32_t koff = (uint32_t)kk * (uint32_t)(KPOS * 128);
        const uint32_t bst = st + (uint32_t)(p.halves * 128 / CH) * p.chunk_bytes;
        bd = p.bf16 ? desc_mn16(bst + koff, p.chunk_bytes) : desc_mn32(bst + koff, p.chunk_bytes);
        a0 = st + koff;
      } else {
        // stage = A rows (halves*128 x 128 B) then B rows (N x 128 B); k-step = 32 B along the row; 4 k-steps / stage
        bd = desc_k(st + (uint32_t)(p.halves * 128) * 128u + (uint32_t)kk * 32u);
        a0 = st + (uint32_t)kk * 32u;
      }
      for (int h = 0; h < p.halves; ++h) {
        uint64_t ad;
        if (p.mn) ad = p.bf16 ? desc_mn16(a0 + (uint32_t)(h * 128 / CH) * p.chunk_bytes, p.chunk_bytes) : desc_mn32(a0 + (uint32_t)(h * 128 / CH) * p.chunk_bytes, p.chunk_bytes);
        else ad = desc_k(a0 + (uint32_t)h * 128u * 128u);
        if (p.bf16)
          asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;}" ::"r"(tmem + (uint32_t)h * 256u), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;}" ::"r"(tmem + (uint32_t)h * 256u), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      acc = 1;
      if (++kg == (p.mn ? p.kpos_per_stage / KPOS : 4)) { kg = 0; if (++s == p.nstage) s = 0; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    mwait(s32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- style 2: the whole warp runs the loop (uniform control flow), descriptors advance by adding a constant to the
// 14-bit start-address field, only the tcgen05.mma itself is predicated by elect.sync -- what lets ptxas keep the
// descriptors in uniform registers (no R2UR / ELECT waterfall per instruction).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{.reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P;}" : "=r"(pred));
  return pred;
}
template <bool BF16>
__global__ void __launch_bounds__(128) probe2(const P p, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t sb = (s32(raw) + 1023u) & ~1023u;
  for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 4; i += blockDim.x) {
    const uint32_t v = BF16 ? 0x3C003C00u + ((i * 2654435761u) >> 28) * 0x00010001u : 0x3C000000u + ((i * 2654435761u) >> 12 & 0xFF000u);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + i * 4), "r"(v) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x < 32) {
    constexpr int CH = BF16 ? 64 : 32;
    constexpr int KPOS = BF16 ? 16 : 8;
    uint32_t idesc = (1u << 4) | ((BF16 ? 1u : 2u) << 7) | ((BF16 ? 1u : 2u) << 10) | (((uint32_t)p.N >> 3) << 17) | ((128u >> 4) << 24) | (1u << 15) | (1u << 16);
    const uint32_t leader = elect_one();
    const uint64_t hi = ((uint64_t)((p.chunk_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((BF16 ? 1024 : 512) >> 4) << 32) | (1ull << 46) | ((BF16 ? 2ull : 1ull) << 61);
    const uint32_t a_h = (uint32_t)(128 / CH) * p.chunk_bytes;               // second row half
    const uint32_t b_off = (uint32_t)(p.halves * 128 / CH) * p.chunk_bytes;
    const int nkg = p.kpos_per_stage / KPOS;
    const long long t0 = clock64();
    uint32_t acc = 0;
    int s = 0;
    for (int it = 0; it < p.iters; it += nkg) {
      const uint32_t st = sb + (uint32_t)s * p.stage_bytes;
      uint64_t ad = hi | (uint64_t)((st >> 4) & 0x3FFF);
      uint64_t ad2 = hi | (uint64_t)(((st + a_h) >> 4) & 0x3FFF);
      uint64_t bd = hi | (uint64_t)(((st + b_off) >> 4) & 0x3FFF);
      for (int kg = 0; kg < nkg; ++kg) {
        if (leader) {
          if (BF16) {
            asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            if (p.halves == 2)
              asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;}" ::"r"(tmem + 256u), "l"(ad2), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          } else {
            asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            if (p.halves == 2)
              asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;}" ::"r"(tmem + 256u), "l"(ad2), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          }
        }
        acc = 1;
        ad += (KPOS * 128) >> 4; ad2 += (KPOS * 128) >> 4; bd += (KPOS * 128) >> 4;
      }
      if (++s == p.nstage) s = 0;
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    mwait(s32(&bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  cudaFuncSetAttribute(probe2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  cudaFuncSetAttribute(probe2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  printf("%-5s %-3s %4s %2s %5s %5s | %10s %10s\n", "type", "maj", "N", "h", "same", "grid", "cyc/MMA", "cyc/kstep");
  for (int grid : {148})
    for (int bf16 = 1; bf16 >= 0; --bf16)
      for (int mn = 0; mn <= 1; ++mn)
        for (int N : {256, 128, 64})
          for (int halves = 2; halves >= 1; --halves)
            for (int same = 0; same <= 1; ++same) {
              if (N != 256 && (halves == 1 || same)) continue;
              P p;
              p.bf16 = bf16; p.mn = mn; p.N = N; p.halves = halves; p.iters = 4096; p.same_addr = same;
              const int CH = bf16 ? 64 : 32;
              const int nch = halves * 128 / CH + N / CH;
              if (mn) {
                p.kpos_per_stage = 64 * 1024 / (nch * 128) / 16 * 16;
                if (p.kpos_per_stage > 256) p.kpos_per_stage = 256;
                p.chunk_bytes = (uint32_t)p.kpos_per_stage * 128u;
                p.stage_bytes = p.chunk_bytes * nch;
              } else {
                p.kpos_per_stage = 0;
                p.chunk_bytes = 0;
                p.stage_bytes = (uint32_t)(halves * 128 + N) * 128u;
              }
              p.nstage = (192 * 1024) / p.stage_bytes;
              if (p.nstage > 8) p.nstage = 8;
              probe<<<grid, 128, 202 * 1024>>>(p, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
              unsigned long long h[148];
              cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
              unsigned long long mx = 0;
              for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
              printf("%-5s %-3s %4d %2d %5d %5d | %10.1f %10.1f\n", bf16 ? "bf16" : "tf32", mn ? "MN" : "K", N, halves, same, grid,
                     (double)mx / (p.iters * halves), (double)mx / p.iters);
              if (mn && !same) {
                p.iters = 4096 / (p.kpos_per_stage / (bf16 ? 16 : 8)) * (p.kpos_per_stage / (bf16 ? 16 : 8));
                if (bf16) probe2<true><<<grid, 128, 202 * 1024>>>(p, d); else probe2<false><<<grid, 128, 202 * 1024>>>(p, d);
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error (style 2): %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                mx = 0;
                for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("%-5s %-3s %4d %2d %5s %5d | %10.1f %10.1f   <- uniform-warp issue loop\n", bf16 ? "bf16" : "tf32", "MN", N, halves, "u", grid,
                       (double)mx / (p.iters * halves), (double)mx / p.iters);
              }
            }
  return 0;
}
