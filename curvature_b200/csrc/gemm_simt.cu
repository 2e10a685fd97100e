// fp32 CUDA-core GEMM with the fused epilogues of K3 (EFB projection: accumulate the square)
// and K5 (matrix-normal draw: add the posterior mean and split into weight / bias).
//   C(m,n) = alpha * sum_k A(m,k) * B(k,n) [+ beta * C]
// Operands are addressed through explicit element strides, so transposes cost nothing:
//   A(m,k) = A[m*sa_m + k*sa_k],   B(k,n) = B[k*sb_k + n*sb_n].
#include "common.cuh"

namespace crv {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int GP = 68;  // smem pitch (floats)
constexpr int GT = 256;

template <int EPI>
__global__ void __launch_bounds__(GT, 2)
gemm_simt_kernel(const float* __restrict__ A, long long sa_m, long long sa_k,
                 const float* __restrict__ B, long long sb_k, long long sb_n,
                 float* __restrict__ C, int ldc, int m, int n, int k, float alpha, float beta,
                 SampleEpilogue se) {
  __shared__ __align__(16) float As[TK][GP];
  __shared__ __align__(16) float Bs[TK][GP];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;

  // loader coordinates: lanes run along whichever axis is contiguous in memory
  const bool a_kc = (sa_k == 1);
  const bool b_kc = (sb_k == 1) && (sb_n != 1);
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int mm, kk;
      if (a_kc) { kk = t & 15; mm = (t >> 4) + 16 * q; } else { mm = t & 63; kk = (t >> 6) + 4 * q; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[q] = (gm < m && gk < k) ? __ldg(A + gm * sa_m + gk * sa_k) : 0.f;
      int nn, kb;
      if (b_kc) { kb = t & 15; nn = (t >> 4) + 16 * q; } else { nn = t & 63; kb = (t >> 6) + 4 * q; }
      const int gn = n0 + nn, gkb = k0 + kb;
      rb[q] = (gn < n && gkb < k) ? __ldg(B + gkb * sb_k + gn * sb_n) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int mm, kk;
      if (a_kc) { kk = t & 15; mm = (t >> 4) + 16 * q; } else { mm = t & 63; kk = (t >> 6) + 4 * q; }
      As[kk][mm] = ra[q];
      int nn, kb;
      if (b_kc) { kb = t & 15; nn = (t >> 4) + 16 * q; } else { nn = t & 63; kb = (t >> 6) + 4 * q; }
      Bs[kb][nn] = rb[q];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  for (int k0 = 0; k0 < k; k0 += TK) {
    stage();
    __syncthreads();
    if (k0 + TK < k) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= n) continue;
      const float v = acc[i][j];
      if (EPI == 0) {
        float* c = C + (size_t)gm * ldc + gn;
        *c = (beta == 0.f) ? alpha * v : alpha * v + beta * *c;
      } else if (EPI == 1) {
        float* c = C + (size_t)gm * ldc + gn;
        *c += v * v;
      } else {
        const float sv = alpha * v;
        if (se.s_out) se.s_out[(size_t)gm * n + gn] = sv;
        if (gn < se.K0) {
          if (se.w_out) se.w_out[(size_t)gm * se.K0 + gn] = se.mu_w[(size_t)gm * se.K0 + gn] + sv;
        } else {
          if (se.b_out) se.b_out[gm] = se.mu_b[gm] + sv;
        }
      }
    }
  }
}

}  // namespace

int gemm_simt_launch(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                     long long sb_n, float* C, int ldc, int m, int n, int k, float alpha, float beta,
                     int epilogue, const SampleEpilogue* sample, cudaStream_t s) {
  CRV_CHECK(A && B, "null GEMM operand");
  CRV_CHECK(m > 0 && n > 0 && k > 0, "bad GEMM shape %d x %d x %d", m, n, k);
  dim3 grid((n + TN - 1) / TN, (m + TM - 1) / TM, 1);
  CRV_CHECK(grid.y < 65536, "GEMM m too large");
  SampleEpilogue se;
  memset(&se, 0, sizeof(se));
  if (sample) se = *sample;
  if (epilogue == EPI_STORE) {
    CRV_CHECK(C != nullptr, "null GEMM output");
    gemm_simt_kernel<0><<<grid, GT, 0, s>>>(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, se);
  } else if (epilogue == EPI_SQUARE_ACCUM) {
    CRV_CHECK(C != nullptr, "null GEMM output");
    gemm_simt_kernel<1><<<grid, GT, 0, s>>>(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, se);
  } else {
    CRV_CHECK(sample != nullptr, "sample epilogue needs its descriptor");
    gemm_simt_kernel<2><<<grid, GT, 0, s>>>(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, se);
  }
  CRV_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace crv
