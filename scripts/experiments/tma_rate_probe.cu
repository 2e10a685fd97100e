// Probe: per-SM TMA load throughput (L2-resident source) as a function of box shape / rank / swizzle / element type.
// One producer thread per CTA issues boxes into an NST-stage ring; a consumer thread frees each stage as soon as it
// has landed.  Reports bytes/cycle/SM at grid = 148 (and 16, to separate a per-SM limit from an L2 limit).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0; int spins = 0;
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (!ok && ++spins > (1 << 24)) asm volatile("trap;");
  }
}
struct P {
  int nst, boxes_per_stage, iters, rank, ni;
  uint32_t box_bytes;
  int n0, n1, n2, n3;      // coordinate space: box index b -> (c0 = (b % n0)*s0, c1 = ((b/n0) % n1)*s1, ...)
  int s0, s1, s2, s3;
  int boxes_total;         // coordinate space size (wraps)
};
__global__ void __launch_bounds__(288) probe(const __grid_constant__ CUtensorMap map, const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t sb = (s32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.box_bytes * p.boxes_per_stage;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nst; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[i])), "r"(p.ni));
    for (int i = p.nst; i < 2 * p.nst; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const long long t0 = clock64();
  __shared__ int4 ctab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const uint32_t b = (uint32_t)((blockIdx.x * 977u + i * 13u) % (uint32_t)p.boxes_total);
    uint32_t r = b / p.n0;
    int4 c; c.x = (int)(b % p.n0) * p.s0; c.y = (int)(r % p.n1) * p.s1; r /= p.n1; c.z = (int)(r % p.n2) * p.s2; r /= p.n2; c.w = (int)(r % p.n3) * p.s3;
    ctab[i] = c;
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && wid < p.ni) {
    const int per = p.boxes_per_stage / p.ni;
    uint32_t b = wid * 64;
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.nst; const uint32_t ph = (it / p.nst) & 1;
      mwait(s32(&bars[p.nst + s]), ph ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[s])), "r"(p.box_bytes * per) : "memory");
      for (int j = wid * per; j < (wid + 1) * per; ++j) {
        const int4 cc = ctab[b & 255];
        const int c0 = cc.x, c1 = cc.y, c2 = cc.z, c3 = cc.w;
        const uint32_t dst = sb + s * stage_bytes + j * p.box_bytes;
        if (p.rank == 2)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(s32(&bars[s])) : "memory");
        else if (p.rank == 3)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(0), "r"(s32(&bars[s])) : "memory");
        else if (p.rank == 5)
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(0), "r"(s32(&bars[s])) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(s32(&bars[s])) : "memory");
        b += 1;
      }
    }
  } else if (threadIdx.x == 256) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.nst; const uint32_t ph = (it / p.nst) & 1;
      mwait(s32(&bars[s]), ph);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bars[p.nst + s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  Enc enc = (Enc)fnp;
  const size_t BUF = 64ull << 20;            // 64 MB source: L2-resident after the first sweep
  void* d; cudaMalloc(&d, BUF); cudaMemset(d, 1, BUF);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
  struct Case { std::string name; int rank; CUtensorMapDataType dt; int esz; cuuint64_t gd[5]; cuuint32_t box[5]; CUtensorMapSwizzle sw;
                int n[4]; int s[4]; int bps; int nst; cuuint64_t gs[4]; };
  std::vector<Case> cases;
  const cuuint64_t R64 = BUF / 128, R512 = BUF / 1024, R256 = BUF / 512;
  auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; auto F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  auto SW = CU_TENSOR_MAP_SWIZZLE_128B; auto SW32 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  cases.push_back({"bf16 2D C=64  box 64x64 (8 KB) x8/stage 3st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 64, 1, 1, 1}, SW, {1, (int)(R64 / 64), 1, 1}, {0, 64, 0, 0}, 8, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 2D C=64  box 64x32 (4 KB) x16/stage 3st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 32, 1, 1, 1}, SW, {1, (int)(R64 / 32), 1, 1}, {0, 32, 0, 0}, 16, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 2D C=64  box 64x16 (2 KB) x32/stage 3st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 16, 1, 1, 1}, SW, {1, (int)(R64 / 16), 1, 1}, {0, 16, 0, 0}, 32, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 2D C=64  box 64x128 (16 KB) x4/stage 3st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 128, 1, 1, 1}, SW, {1, (int)(R64 / 128), 1, 1}, {0, 128, 0, 0}, 4, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 2D C=64  box 64x256 (32 KB) x2/stage 3st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 256, 1, 1, 1}, SW, {1, (int)(R64 / 256), 1, 1}, {0, 256, 0, 0}, 2, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 2D C=64  box 64x256 (32 KB) x1/stage 6st", 2, BF, 2, {64, R64, 1, 1, 1}, {64, 256, 1, 1, 1}, SW, {1, (int)(R64 / 256), 1, 1}, {0, 256, 0, 0}, 1, 6, {0, 0, 0, 0}});
  // chunk dimension trick: [R][C=256] viewed as {64 ch, R, 4 chunks}, chunk stride 128 B (smaller than the row stride)
  cases.push_back({"bf16 3D C=256 box 64x64x4 chunks (32 KB) x2/stage 3st", 3, BF, 2, {64, R256, 4, 1, 1}, {64, 64, 4, 1, 1}, SW, {1, (int)(R256 / 64), 1, 1}, {0, 64, 0, 0}, 2, 3, {512, 128, 128 * 4, 128 * 4}});
  cases.push_back({"bf16 3D C=256 box 64x32x4 chunks (16 KB) x2/stage 6st", 3, BF, 2, {64, R256, 4, 1, 1}, {64, 32, 4, 1, 1}, SW, {1, (int)(R256 / 32), 1, 1}, {0, 32, 0, 0}, 2, 6, {512, 128, 128 * 4, 128 * 4}});
  // NHWC 5D with chunk dimension: C=256, 14x14: {64, W, H, N, chunk}, box 64 x 14 x 2 x 1 x 4 (14 KB)
  cases.push_back({"bf16 5D C=256 14x14 box 64x14x2x1x4 (14 KB) x4/stage 3st", 5, BF, 2, {64, 14, 14, BUF / (512 * 196), 4}, {64, 14, 2, 1, 4}, SW, {1, 1, 7, (int)(BUF / (512 * 196))}, {0, 0, 2, 1}, 4, 3, {512, 512 * 14, 512 * 196, 128}});
  cases.push_back({"bf16 4D C=256 14x14 box 64x14x2 (3.5 KB) x16/stage 3st", 4, BF, 2, {256, 14, 14, BUF / (512 * 196), 1}, {64, 14, 2, 1, 1}, SW, {4, 1, 7, (int)(BUF / (512 * 196))}, {64, 0, 2, 1}, 16, 3, {0, 0, 0, 0}});
  cases.push_back({"bf16 4D C=64 56x56 box 64x8x8 (8 KB) x8/stage 3st", 4, BF, 2, {64, 56, 56, BUF / (128 * 3136), 1}, {64, 8, 8, 1, 1}, SW, {1, 7, 7, (int)(BUF / (128 * 3136))}, {0, 8, 8, 1}, 8, 3, {0, 0, 0, 0}});
  cases.push_back({"fp32 2D C=64  box 32x128 (16 KB) x2/stage 3st", 2, F32, 4, {64, BUF / 256, 1, 1, 1}, {32, 128, 1, 1, 1}, SW32, {2, (int)(BUF / 256 / 128), 1, 1}, {32, 128, 0, 0}, 2, 3, {0, 0, 0, 0}});
  cases.push_back({"fp32 2D C=64  box 32x256 (32 KB) x2/stage 3st", 2, F32, 4, {64, BUF / 256, 1, 1, 1}, {32, 256, 1, 1, 1}, SW32, {2, (int)(BUF / 256 / 256), 1, 1}, {32, 256, 0, 0}, 2, 3, {0, 0, 0, 0}});
  cases.push_back({"fp32 3D C=64  box 32x128x2 chunks (32 KB) x1/stage 6st", 3, F32, 4, {32, BUF / 256, 2, 1, 1}, {32, 128, 2, 1, 1}, SW32, {1, (int)(BUF / 256 / 128), 1, 1}, {0, 128, 0, 0}, 1, 6, {256, 128, 256, 256}});
  for (auto& c : cases) {
    CUtensorMap map;
    cuuint64_t gs[4]; cuuint64_t acc = c.gd[0] * c.esz;
    for (int i = 0; i < 4; ++i) { gs[i] = acc; acc *= c.gd[i + 1 < 5 ? i + 1 : 4]; }
    if (c.gs[0]) for (int i = 0; i < 4; ++i) gs[i] = c.gs[i];
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult rc = enc(&map, c.dt, c.rank, d, c.gd, gs, c.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("%s: encode rc=%d\n", c.name.c_str(), (int)rc); continue; }
    int ni = 1; P p; p.nst = c.nst; p.boxes_per_stage = c.bps; p.iters = 400; p.rank = c.rank; p.ni = ni;
    p.box_bytes = c.box[0] * c.box[1] * c.box[2] * c.box[3] * c.box[4] * c.esz;
    p.n0 = c.n[0]; p.n1 = c.n[1]; p.n2 = c.n[2]; p.n3 = c.n[3]; p.s0 = c.s[0]; p.s1 = c.s[1]; p.s2 = c.s[2]; p.s3 = c.s[3];
    p.boxes_total = c.n[0] * c.n[1] * c.n[2] * c.n[3];
    const size_t smem = (size_t)p.box_bytes * c.bps * c.nst + 1024;
    for (int ni : {1, 2, 4, 8}) { if (c.bps % ni) continue; p.ni = ni;
    for (int grid : {148}) {
      for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms L2
        probe<<<grid, 288, smem>>>(map, p, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name.c_str(), cudaGetErrorString(e)); return 0; }
      }
      std::vector<unsigned long long> h(grid);
      cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      const double bytes = (double)p.box_bytes * c.bps * p.iters;
      printf("%-58s ni %d grid %3d: %6.1f B/cyc/SM  (%.0f cycles per stage of %u KB)\n", c.name.c_str(), ni, grid, bytes / mx, (double)mx / p.iters,
             p.box_bytes * c.bps / 1024);
    } }
  }
  return 0;
}
