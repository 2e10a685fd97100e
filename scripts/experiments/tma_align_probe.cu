// Probe: does cp.async.bulk.tensor (tile mode) accept box start coordinates whose innermost offset is not a
// multiple of 16 bytes?  (Decides whether +-1 column filter-tap shifts can be done by TMA on NCHW fp32 data.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, float* out, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  sb = (sb + 1023u) & ~1023u;
  uint32_t bb = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(n * 4));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(bb) : "memory");
    uint32_t ok = 0; int spins = 0;
    while (!ok && ++spins < (1 << 22))
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bb));
    out[n] = ok ? 1.f : -1.f;
  }
  __syncthreads();
  const float* s = reinterpret_cast<const float*>(smem + ((sb - (uint32_t)__cvta_generic_to_shared(smem))));
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}

int main() {
  const int W = 64, H = 16;
  std::vector<float> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (float)i;
  float* d; cudaMalloc(&d, W * H * 4); cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 4096);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  for (int sw = 0; sw < 2; ++sw) {
    CUtensorMap map;
    cuuint64_t gd[2] = {W, H}; cuuint64_t gs[1] = {W * 4}; cuuint32_t box[2] = {8, 4}; cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle=%d encode rc=%d\n", sw, (int)rc);
    for (int c0 : {0, 4, 1, 2, 3, -1, -4, 61}) {
      cudaMemset(out, 0, 4096);
      probe<<<1, 32, 4096>>>(map, c0, 2, out, 32);
      cudaError_t e = cudaDeviceSynchronize();
      float r[33]; 
      if (e == cudaSuccess) cudaMemcpy(r, out, 33 * 4, cudaMemcpyDeviceToHost);
      printf("  c0=%3d -> %s", c0, cudaGetErrorString(e));
      if (e == cudaSuccess) { printf("  done=%g  first row:", r[32]); for (int i = 0; i < 8; ++i) printf(" %g", r[i]); }
      printf("\n");
      if (e != cudaSuccess) { printf("context lost; stop\n"); return 0; }
    }
  }
  return 0;
}
