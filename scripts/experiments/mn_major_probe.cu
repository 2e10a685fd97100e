// Probe: tcgen05.mma kind::tf32 with MN-major SWIZZLE_128B operands produced by TMA from a [R][C] fp32 matrix.
// D[m][n] = sum_r X[r][m] X[r][n].  Tries descriptor variants and prints which reproduces the host result.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>

constexpr int R = 64, C = 128;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0; int spins = 0;
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (!ok && ++spins > (1 << 22)) asm volatile("trap;");
  }
}
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, uint32_t lbo, uint32_t sbo, uint32_t idesc, uint32_t ltype,
                                             float* D, float* dump) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tslot;
  const uint32_t sb = (s32(raw) + 1023u) & ~1023u;
  const uint32_t b0 = s32(&bars[0]), b1 = s32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0), "r"(R * C * 4) : "memory");
    for (int q = 0; q < C / 32; ++q)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(sb + q * (R * 128)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(q * 32), "r"(0), "r"(b0) : "memory");
    wait(b0, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kg = 0; kg < R / 8; ++kg) {
      const uint32_t a = sb + kg * 1024;
      const uint64_t desc = (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                            ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)ltype << 61);
      const uint32_t acc = kg > 0;
      asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}"
                   ::"r"(tmem), "l"(desc), "l"(desc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b1) : "memory");
  }
  __syncthreads();
  {  // raw smem dump (after the TMA landed: thread 0 waited before the barrier above)
    const float* s = reinterpret_cast<const float*>(raw + (sb - s32(raw)));
    for (int i = threadIdx.x; i < R * C; i += 128) dump[i] = s[i];
  }
  wait(b1, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cc = 0; cc < C; cc += 16) {
    uint32_t v[16];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + cc;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * C + cc + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  std::vector<float> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (float)(((r * 7 + c * 3 + (r * c) % 5) % 7) - 3);
  std::vector<double> want(C * C, 0.0);
  for (int m = 0; m < C; ++m) for (int n = 0; n < C; ++n) { double s = 0; for (int r = 0; r < R; ++r) s += h[r * C + m] * h[r * C + n]; want[m * C + n] = s; }
  float *d, *D, *dump;
  cudaMalloc(&d, R * C * 4); cudaMemcpy(d, h.data(), R * C * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&D, C * C * 4); cudaMalloc(&dump, R * C * 4);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  CUtensorMap map;
  cuuint64_t gd[2] = {C, R}; cuuint64_t gs[1] = {C * 4}; cuuint32_t box[2] = {32, R}; cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, R * C * 4 + 2048);
  const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((C >> 3) << 17) | ((128 >> 4) << 24);
  struct V { const char* name; uint32_t lbo, sbo, idesc, ltype; } vs[] = {
    {"MN-major BASE32B lbo=chunk sbo=512", R * 128, 512, base | (1u << 15) | (1u << 16), 1},
    {"MN-major BASE32B lbo=512 sbo=chunk", 512, R * 128, base | (1u << 15) | (1u << 16), 1},
    {"MN-major BASE32B lbo=chunk sbo=1024", R * 128, 1024, base | (1u << 15) | (1u << 16), 1},
    {"K-major  (control, wrong layout)", 0, 1024, base, 2},
  };
  for (auto& v : vs) {
    cudaMemset(D, 0xff, C * C * 4);
    probe<<<1, 128, R * C * 4 + 2048>>>(map, v.lbo, v.sbo, v.idesc, v.ltype, D, dump);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s: %s", v.name, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 0; }
    std::vector<float> got(C * C), sm(R * C);
    cudaMemcpy(got.data(), D, C * C * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(sm.data(), dump, R * C * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0; double md = 0;
    for (int i = 0; i < C * C; ++i) { double dd = fabs(got[i] - want[i]); if (dd > 0) ++bad; if (got[i] == 0) ++zeros; if (dd > md) md = dd; }
    printf("  mismatches=%d zeros=%d maxdiff=%g  D[0][0..3]=%g %g %g %g want %g %g %g %g\n", bad, zeros, md, got[0], got[1], got[2], got[3],
           want[0], want[1], want[2], want[3]);
    if (&v == &vs[0]) {
      printf("  smem row0 (pos 0, chunk 0) first 8: "); for (int i = 0; i < 8; ++i) printf("%g ", sm[i]);
      printf("| X[0][0..7]: "); for (int i = 0; i < 8; ++i) printf("%g ", h[i]);
      printf("\n  smem row1 first 8: "); for (int i = 0; i < 8; ++i) printf("%g ", sm[32 + i]);
      printf("| X[1][0..7]: "); for (int i = 0; i < 8; ++i) printf("%g ", h[C + i]);
      printf("| X[1][4..11]: "); for (int i = 4; i < 12; ++i) printf("%g ", h[C + i]); printf("\n");
    }
  }
  return 0;
}
