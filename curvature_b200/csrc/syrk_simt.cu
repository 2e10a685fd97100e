// K1 (fp32 tier): fused implicit-im2col + SYRK + running-sum epilogue on the CUDA cores.
//
//   F[k1,k2] += alpha * sum_r X[k1,r] * X[k2,r]
//
// X is never materialised: every staged element is gathered straight from the NCHW
// activation tensor through the index map of ConvGeom (bit-exact restatement of the
// reference's F.unfold row/column order, curvature/curvatures.py:329-336).  Exact fp32
// products and fp32 FMA accumulation -> this is the 1e-5 parity tier and the checker for the
// tensor-core tiers.  Only lower-triangular 64x64 tile pairs are computed; off-diagonal tiles
// are mirrored in the epilogue.  The contraction axis R is split across CTAs so that small
// factors with huge R (ResNet stem: 147^2 over R = 3.2M) still fill 148 SMs; partial sums are
// merged with fp32 red.global.add straight into the factor arena (the `state +=` of
// curvatures.py:346-350 is the same RMW, so no separate accumulate pass exists).
#include "common.cuh"
#include <algorithm>

namespace crv {
namespace {

constexpr int BM = 64;     // tile edge (rows of X per operand tile)
constexpr int BK = 32;     // contraction chunk = one warp-wide coalesced read along r
constexpr int PITCH = 68;  // smem pitch in floats: 16B-aligned rows, conflict-free float4 access
constexpr int NT = 256;

// Row descriptor: channel base offset c*H*W (>= 0), or ROW_ONES / ROW_NONE.
constexpr int ROW_NONE = -1;
constexpr int ROW_ONES = -2;

__device__ __forceinline__ void decode_row(const ConvGeom& g, int k, int& base, int& dij) {
  if (k < g.K0) {
    const int khw = g.kh * g.kw;
    const int c = k / khw;
    const int t = k - c * khw;
    const int i = t / g.kw;
    const int j = t - i * g.kw;
    base = c * g.H * g.W;
    dij = ((i - g.ph) << 16) | ((j - g.pw) & 0xffff);
  } else {
    base = (k < g.D) ? ROW_ONES : ROW_NONE;
    dij = 0;
  }
}

// One CTA's share of one factor: tile pair `p`, contraction split `split`.  DEPTH = chunks of global loads in flight: 1 for
// the per-factor launch (FMA-bound large factors: two CTAs per SM at 128 registers), 2 for the small-model batch launch.
template <int DEPTH>
__device__ __forceinline__ void syrk_simt_body(const ConvGeom& g, const float alpha, float* __restrict__ F,
                                               const int chunks_per_split, const int p, const int split) {
  __shared__ __align__(16) float As[BK][PITCH];
  __shared__ __align__(16) float Bs[BK][PITCH];

  // lower-triangular tile pair (ti >= tj) from the linear block index
  int ti = (int)((sqrtf(8.f * (float)p + 1.f) - 1.f) * 0.5f);
  while (ti * (ti + 1) / 2 > p) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
  const int tj = p - ti * (ti + 1) / 2;
  const bool diag = (ti == tj);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // Each thread stages, per operand tile, two quads of 4 consecutive rows at one r:
  // rows 4*(warp + 8*q) .. +3, q = 0,1.  One float4 st.shared per quad (conflict-free).
  int baseA[8], dA[8], baseB[8], dB[8];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = 4 * (warp + 8 * q) + e;
      decode_row(g, ti * BM + row, baseA[q * 4 + e], dA[q * 4 + e]);
      decode_row(g, tj * BM + row, baseB[q * 4 + e], dB[q * 4 + e]);
    }

  const long long r_begin = (long long)split * chunks_per_split * BK;
  long long r_end = r_begin + (long long)chunks_per_split * BK;
  if (r_end > g.R) r_end = g.R;

  // two chunks of global loads in flight (a chunk is one round trip to L2 / HBM: with a single one the loop of a factor
  // with a long contraction axis and few rows -- LeNet's conv1, R = 78 400, D = 6 / 26 -- is pure load latency)
  float va0[8], vb0[8], va1[8], vb1[8];
  auto fetch = [&](long long r0, float (&va)[8], float (&vb)[8]) {
    const long long r = r0 + lane;
    const bool valid = r < r_end;
    int n = 0, oh = 0, ow = 0;
    if (valid) {
      n = (int)(r / g.L);
      const int l = (int)(r - (long long)n * g.L);
      oh = l / g.OW;
      ow = l - oh * g.OW;
    }
    const float* __restrict__ img = g.x + (size_t)n * g.C * g.H * g.W;
    const int ih0 = oh * g.sh, iw0 = ow * g.sw;
    auto load = [&](int base, int dij) -> float {
      if (!valid || base == ROW_NONE) return 0.f;
      if (base == ROW_ONES) return 1.f;
      const int ih = ih0 + (dij >> 16);
      const int iw = iw0 + (int)(short)(dij & 0xffff);
      if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) return __ldg(img + base + ih * g.W + iw);
      return 0.f;
    };
#pragma unroll
    for (int e = 0; e < 8; ++e) va[e] = load(baseA[e], dA[e]);
    if (!diag) {
#pragma unroll
      for (int e = 0; e < 8; ++e) vb[e] = load(baseB[e], dB[e]);
    }
  };

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  auto step = [&](long long r0, float (&va)[8], float (&vb)[8]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      *reinterpret_cast<float4*>(&As[lane][4 * (warp + 8 * q)]) =
          make_float4(va[q * 4 + 0], va[q * 4 + 1], va[q * 4 + 2], va[q * 4 + 3]);
      if (!diag)
        *reinterpret_cast<float4*>(&Bs[lane][4 * (warp + 8 * q)]) =
            make_float4(vb[q * 4 + 0], vb[q * 4 + 1], vb[q * 4 + 2], vb[q * 4 + 3]);
    }
    __syncthreads();
    if (r0 + DEPTH * BK < r_end) fetch(r0 + DEPTH * BK, va, vb);  // the next loads into this buffer fly during DEPTH rounds of FMAs
    const float(*Bp)[PITCH] = diag ? As : Bs;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bp[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  };
  if (r_begin < r_end) fetch(r_begin, va0, vb0);
  if (DEPTH == 2) {
    if (r_begin + BK < r_end) fetch(r_begin + BK, va1, vb1);
    for (long long r0 = r_begin; r0 < r_end; r0 += 2 * BK) {
      step(r0, va0, vb0);
      if (r0 + BK < r_end) step(r0 + BK, va1, vb1);
    }
  } else {
    for (long long r0 = r_begin; r0 < r_end; r0 += BK) step(r0, va0, vb0);
  }

  // epilogue: F += alpha * acc (and the mirrored element for off-diagonal tiles)
  const int D = g.D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = ti * BM + ty * 4 + i;
    if (row >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tj * BM + tx * 4 + j;
      if (col >= D) continue;
      const float v = alpha * acc[i][j];
      atomicAdd(&F[(size_t)row * D + col], v);
      if (!diag) atomicAdd(&F[(size_t)col * D + row], v);
    }
  }
}

__global__ void __launch_bounds__(NT, 2)
syrk_simt_kernel(const ConvGeom g, const float alpha, float* __restrict__ F, const int chunks_per_split) {
  syrk_simt_body<1>(g, alpha, F, chunks_per_split, (int)blockIdx.x, (int)blockIdx.y);
}

// Every factor of a SMALL model in one launch (LeNet-5: ten factors; launched one by one the update is launch-bound,
// 0.3 ms for 0.3 GFLOP): block b works on factor i with begin[i] <= b < begin[i + 1], on (pair, split) =
// ((b - begin[i]) % pairs[i], (b - begin[i]) / pairs[i]).
constexpr int SIMT_BATCH_MAX = 64;
struct SimtBatch {
  int n;
  int begin[SIMT_BATCH_MAX + 1];
  int pairs[SIMT_BATCH_MAX];
  int cps[SIMT_BATCH_MAX];
  float alpha[SIMT_BATCH_MAX];
  float* F[SIMT_BATCH_MAX];
  ConvGeom g[SIMT_BATCH_MAX];
};

__global__ void __launch_bounds__(NT, 1) syrk_simt_batch_kernel(const __grid_constant__ SimtBatch b) {
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.begin[i + 1]) ++i;
  const int local = (int)blockIdx.x - b.begin[i];
  const int split = local / b.pairs[i];
  syrk_simt_body<2>(b.g[i], b.alpha[i], b.F[i], b.cps[i], local - split * b.pairs[i], split);
}

// (pairs, splits, chunks per split) of one factor when it should occupy about `target` CTAs
void simt_partition(const ConvGeom& g, long long target, long long& pairs, long long& splits, long long& cps) {
  const int T = (g.D + BM - 1) / BM;
  pairs = (long long)T * (T + 1) / 2;
  const long long chunks = (g.R + BK - 1) / BK;
  splits = (target + pairs - 1) / pairs;
  // every split adds its partial sum into F with one fp32 atomic: S sequential roundings per element (same-sign ones where
  // all partials are alike, e.g. the bias row's count).  This is the 1e-5 checker tier: cap S so that they stay ~1e-6.
  if (splits > 32) splits = 32;
  if (splits > chunks / 4) splits = chunks / 4;     // keep >= 4 chunks per CTA
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  cps = (chunks + splits - 1) / splits;
  splits = (chunks + cps - 1) / cps;
}

}  // namespace

int syrk_simt_batch_launch(const ConvGeom* gs, const float* alphas, float* const* Fs, int n, cudaStream_t s) {
  CRV_CHECK(gs && alphas && Fs && n > 0, "empty batch");
  const int sms = device_sm_count();
  CRV_CHECK(sms > 0, "no CUDA device");
  static SimtBatch b;          // (7 KB: kept off the stack; callers hold the library lock)
  for (int i0 = 0; i0 < n; i0 += SIMT_BATCH_MAX) {
    const int cnt = std::min(SIMT_BATCH_MAX, n - i0);
    double total = 0.0;
    for (int k = 0; k < cnt; ++k) total += (double)gs[i0 + k].R * gs[i0 + k].D * (gs[i0 + k].D + 1);
    b.n = cnt;
    b.begin[0] = 0;
    double flops = 0.0, bytes = 0.0;
    for (int k = 0; k < cnt; ++k) {
      const ConvGeom& g = gs[i0 + k];
      CRV_CHECK(Fs[i0 + k] != nullptr, "null factor pointer");
      // ~3 waves of 2 resident CTAs per SM for the whole launch, shared out by work
      const double w = (double)g.R * g.D * (g.D + 1);
      long long pairs, splits, cps;
      // ... but never fewer CTAs than one per 16 chunks of the contraction axis (a chunk is a latency-bound round trip)
      const long long by_work = (long long)((double)sms * 6.0 * w / total) + 1;
      const long long T = (g.D + BM - 1) / BM, by_chunks = T * (T + 1) / 2 * (((g.R + BK - 1) / BK + 15) / 16);
      simt_partition(g, std::max(by_work, by_chunks), pairs, splits, cps);
      CRV_CHECK(pairs * splits < (1LL << 24) && b.begin[k] + pairs * splits < (1LL << 30), "batch too large for one launch");
      b.pairs[k] = (int)pairs; b.cps[k] = (int)cps; b.alpha[k] = alphas[i0 + k]; b.F[k] = Fs[i0 + k]; b.g[k] = g;
      b.begin[k + 1] = b.begin[k] + (int)(pairs * splits);
      flops += w; bytes += 4.0 * g.N * g.C * g.H * g.W;
    }
    profile_begin(KC_SYRK_SIMT, flops, bytes, s);
    syrk_simt_batch_kernel<<<(unsigned)b.begin[cnt], NT, 0, s>>>(b);
    profile_end(s);
    CRV_CUDA(cudaGetLastError());
  }
  return 0;
}

int syrk_simt_launch(const ConvGeom& g, float alpha, float* F, cudaStream_t s) {
  CRV_CHECK(F != nullptr, "null factor pointer");
  const int sms = device_sm_count();
  CRV_CHECK(sms > 0, "no CUDA device");
  long long pairs, splits, cps;
  simt_partition(g, (long long)sms * 2 * 3, pairs, splits, cps);   // ~3 waves of 2 resident CTAs per SM
  CRV_CHECK(pairs < (1LL << 31), "factor too large");
  dim3 grid((unsigned)pairs, (unsigned)splits, 1);
  profile_begin(KC_SYRK_SIMT, (double)g.R * g.D * (g.D + 1), 4.0 * g.N * g.C * g.H * g.W, s);
  syrk_simt_kernel<<<grid, NT, 0, s>>>(g, alpha, F, (int)cps);
  profile_end(s);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace crv
