// K4: batched damped "Cholesky of the inverse" (KFAC.invert, curvature/curvatures.py:368-379).
//
// The reference computes L = chol_lower(inverse(reg)) with reg = sym(sqrt(s) F + sqrt(n) I).
// That L is reproduced WITHOUT forming an inverse and without an LU:
//     P = J reg J  (J = exchange matrix),   P = C C^T  (lower Cholesky),
//     reg = (J C J)(J C J)^T = U U^T with U upper  =>  reg^{-1} = U^{-T} U^{-1},
// and U^{-T} is lower triangular with a positive diagonal, i.e. it is THE Cholesky factor of
// reg^{-1}:  L = U^{-T} = J C^{-T} J,  L[i][j] = Cinv[D-1-j][D-1-i].
// (C^{-T} alone -- the obvious "invert the Cholesky factor" -- is an upper factor of the same
// covariance and gives different samples for the same noise)
//
// Blocked right-looking factorisation, two levels: 32-wide inner panels (diagonal blocks factored and inverted in fp64
// in shared memory by one warp, panel solve in the same launch, rank-32 fp32 updates confined to the 128-wide outer panel)
// and one rank-128 fp32 update of the trailing matrix per outer panel; followed by a triangular inversion by recursive
// doubling (log-depth, all GEMM tiles).  All matrices of a call advance together:
// gridDim.z indexes the matrix, CTAs of finished / smaller matrices exit at once.
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace crv {
namespace {

constexpr int NB = 32;

struct MatDesc {
  const float* F;   // input factor (D x D)
  float* W;         // workspace: flipped damped matrix -> its lower Cholesky factor C
  float* X;         // workspace: inverse of C (lower part)
  float* Dinv;      // workspace: nb inverted diagonal blocks, 32 x 32 each
  float* L;         // output
  int D, nb;
  float sqrt_mul, sqrt_add;
};

// ---- prologue: W = J * sym(sqrt(s) F + sqrt(n) I) * J -----------------------------------------
__global__ void __launch_bounds__(256) chol_prologue_kernel(const MatDesc* __restrict__ descs) {
  const MatDesc d = descs[blockIdx.z];
  const size_t total = (size_t)d.D * d.D;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int i = (int)(e / d.D), j = (int)(e - (size_t)i * d.D);
    const int a = d.D - 1 - i, b = d.D - 1 - j;
    // same rounding sequence as the reference: (s*F + diag) then (reg + reg^T) / 2
    float rab = __fmul_rn(d.sqrt_mul, __ldg(d.F + (size_t)a * d.D + b));
    float rba = __fmul_rn(d.sqrt_mul, __ldg(d.F + (size_t)b * d.D + a));
    if (a == b) { rab = __fadd_rn(rab, d.sqrt_add); rba = __fadd_rn(rba, d.sqrt_add); }
    d.W[e] = __fadd_rn(rab, rba) / 2.0f;
  }
}

// ---- step 1 + 2: factor + invert the diagonal block of panel j, and solve the panel below it -------------------
// One launch per 32-wide panel: CTA 0 owns the diagonal block (writes its inverse to Dinv and the failure flag), CTA
// ib - j > 0 owns the 32-row block ib below it.  EVERY CTA factors the (same, read-only) diagonal block itself -- a
// single warp, one lane per row, fp64 in shared memory, ~10 us -- instead of waiting for a separate launch to publish
// it: the redundant work is nothing, the launch and the dependency it removes are a third of the critical path.  The
// factored diagonal block itself is never written back: nothing reads it again (the triangular inversion starts from
// Dinv, the updates read only rows below the block).
constexpr int PANEL_RB = 4;      // 32-row blocks solved per CTA of the panel kernel (amortises the redundant factorisation)
__global__ void __launch_bounds__(128) chol_panel_kernel(const MatDesc* __restrict__ descs, int j, int* info) {
  const MatDesc d = descs[blockIdx.z];
  if (j >= d.nb) return;
  const int ib0 = j + 1 + PANEL_RB * blockIdx.x;      // this CTA's row blocks: ib0 .. ib0 + PANEL_RB - 1 (CTA 0 also owns the diagonal)
  if (ib0 >= d.nb && blockIdx.x > 0) return;
  const int ib = blockIdx.x == 0 ? j : ib0;            // (`ib == j` marks the CTA that publishes Dinv and the failure flag)
  __shared__ double S[NB][NB + 1];
  __shared__ double Iv[NB][NB + 1];
  __shared__ float Ab[NB][NB + 1];
  const int o = j * NB;
  const int bs = min(NB, d.D - o);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int e = t; e < NB * NB; e += 128) {
    const int r = e / NB, c = e % NB;
    double v = 0.0;
    if (r < bs && c < bs && c <= r) v = (double)d.W[(size_t)(o + r) * d.D + o + c];
    if (r >= bs && r == c) v = 1.0;   // pad to a full 32x32 block with the identity
    S[r][c] = v;
  }
  __syncthreads();
  if (warp == 0) {
    // Right-looking Cholesky of the 32 x 32 block entirely in REGISTERS: lane = row, the row's 32 entries live in
    // s[0..31] (loops fully unrolled, compile-time indices), values of other rows arrive by warp shuffles.  No shared
    // memory round trips and no barriers on the critical path: ~5 us instead of ~45 us for the shared-memory version.
    constexpr unsigned FULL = 0xffffffffu;
    double s_[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) s_[k] = S[lane][k];
    int bad = 0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      const double p = __shfl_sync(FULL, s_[c], c);          // the pivot S[c][c]
      if (!(p > 0.0) && !bad) bad = o + c + 1;
      const double piv = sqrt(p);
      const double l = (lane == c) ? piv : s_[c] / piv;      // L[lane][c] (rows above c hold unused upper-triangle values)
      s_[c] = l;
#pragma unroll
      for (int cc = c + 1; cc < NB; ++cc) {
        const double lcc = __shfl_sync(FULL, l, cc);         // L[cc][c]
        s_[cc] -= l * lcc;
      }
    }
    // inverse of the lower-triangular block by forward substitution: lane = column, iv[r] = Iv[r][lane]
    double iv[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      double acc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; ++k) {
        const double srk = __shfl_sync(FULL, s_[k], r);      // L[r][k]
        acc -= srk * iv[k];                                  // (iv[k] = 0 for k < lane)
      }
      const double srr = __shfl_sync(FULL, s_[r], r);
      iv[r] = (r < lane) ? 0.0 : acc / srr;
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) Iv[r][lane] = iv[r];
    if (ib == j && lane == 0 && bad && info[blockIdx.z] == 0) info[blockIdx.z] = bad;
  }
  __syncthreads();
  if (ib == j)
    for (int e = t; e < NB * NB; e += 128) d.Dinv[(size_t)j * NB * NB + e] = (float)Iv[e / NB][e % NB];
  // W[rb][j] <- W[rb][j] * inv(C_jj)^T for this CTA's row blocks
  for (int rb = ib0; rb < ib0 + PANEL_RB && rb < d.nb; ++rb) {
    const int ro = rb * NB;
    __syncthreads();
    for (int e = t; e < NB * NB; e += 128) {
      const int r = e / NB, c = e % NB;
      Ab[r][c] = (ro + r < d.D && o + c < d.D) ? d.W[(size_t)(ro + r) * d.D + o + c] : 0.f;
    }
    __syncthreads();
    for (int e = t; e < NB * NB; e += 128) {
      const int r = e / NB, c = e % NB;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < NB; ++k) acc = fmaf(Ab[r][k], (float)Iv[c][k], acc);
      if (ro + r < d.D && o + c < d.D) d.W[(size_t)(ro + r) * d.D + o + c] = acc;
    }
  }
}

// ---- step 3: symmetric rank-k update  W[r][c] -= sum_k W[r][kb + k] W[c][kb + k]  for cb <= c < ce, r >= c ----------
// Two uses per 128-wide outer panel: after each 32-wide inner panel the columns still inside the outer panel get its
// rank-32 update at once (the next inner panel needs them); everything beyond the outer panel gets ONE rank-128 update
// when the outer panel is complete -- a quarter of the passes over the trailing matrix, four times the arithmetic per
// byte.  64 x 64 tiles, lower triangle only (grid: row tile x column tile, relative to cb).
__global__ void __launch_bounds__(256) chol_update_kernel(const MatDesc* __restrict__ descs, int kb, int klen, int cb, int ce) {
  const MatDesc d = descs[blockIdx.z];
  if (cb >= d.D) return;
  const int cend = min(ce, d.D);
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (ti < tj) return;
  const int r0 = cb + ti * 64, c0 = cb + tj * 64;
  if (r0 >= d.D || c0 >= cend) return;
  __shared__ __align__(16) float Ps[NB][68];   // Ps[k][row]  panel rows of tile ti
  __shared__ __align__(16) float Qs[NB][68];   // Qs[k][row]  panel rows of tile tj
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < klen; k0 += NB) {
    const int ko = kb + k0;
    for (int e = t; e < 64 * NB; e += 256) {
      const int r = e / NB, k = e % NB;   // lanes along k: contiguous 128-byte panel rows
      const bool kv = k0 + k < klen && ko + k < d.D;
      Ps[k][r] = (kv && r0 + r < d.D) ? d.W[(size_t)(r0 + r) * d.D + ko + k] : 0.f;
      Qs[k][r] = (kv && c0 + r < d.D) ? d.W[(size_t)(c0 + r) * d.D + ko + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ps[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Qs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i], b[jj], acc[i][jj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= d.D) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c = c0 + tx * 4 + jj;
      if (c < cend && c <= r) d.W[(size_t)r * d.D + c] -= acc[i][jj];
    }
  }
}

// ---- triangular inverse X = C^{-1} by recursive doubling ------------------------------------------
// For a lower-triangular C = [[C11, 0], [C21, C22]]:  C^{-1} = [[X11, 0], [-X22 C21 X11, X22]].  Level 0 are the 32 x 32
// diagonal blocks (inverted in fp64 by the factorisation, Dinv); level l merges neighbouring inverted blocks of size
// b = 32 * 2^(l-1) with two GEMMs per pair -- T = C21 X11, then X21 = -X22 T -- and ALL pairs of all matrices of the call
// are tiles of the same two launches: log2(D / 32) levels (8 for D = 4608) of fully parallel fp32 GEMM tiles instead of a
// sequential sweep over block rows per column strip (which took 60 % of the whole invert).  The zero upper triangles of
// X11 / X22 are skipped tile-wise (k range) and masked element-wise inside the diagonal tiles; T lives in the output
// matrix L, which is only written by the epilogue.
__global__ void __launch_bounds__(256) chol_trtri_init_kernel(const MatDesc* __restrict__ descs) {
  const MatDesc d = descs[blockIdx.z];
  const int i = blockIdx.x;
  if (i >= d.nb) return;
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    if (i * NB + r < d.D && i * NB + c < d.D) d.X[(size_t)(i * NB + r) * d.D + i * NB + c] = d.Dinv[(size_t)i * NB * NB + e];
  }
}

// phase 0: T[bot, top] = C[bot, top] * X[top, top];  phase 1: X[bot, top] = -X[bot, bot] * T[bot, top]
__global__ void __launch_bounds__(256) chol_trtri_merge_kernel(const MatDesc* __restrict__ descs, int b, int phase) {
  const MatDesc d = descs[blockIdx.z];
  const int tpp = max(1, b / 64);                      // 64-wide tiles per block edge
  const int pair = blockIdx.x / (tpp * tpp);
  const int tr = (blockIdx.x / tpp) % tpp, tc = blockIdx.x % tpp;
  const int r0 = pair * 2 * b;                         // top block = [r0, r0 + b), bottom = [r0 + b, r0 + b + b2)
  if (r0 + b >= d.D) return;
  const int b2 = min(b, d.D - r0 - b);
  const int rb = tr * 64, cb = tc * 64;                // tile origin inside the (b2 x b) output block
  if (rb >= b2 || cb >= b) return;
  __shared__ __align__(16) float As[16][68];           // As[k][row]
  __shared__ __align__(16) float Bs[16][68];           // Bs[k][col]
  const float* __restrict__ Am; const float* __restrict__ Bm; float* __restrict__ Om;
  int a_r0, a_c0, b_r0, b_c0, klo, khi;
  if (phase == 0) {        // A = C[bot][top], B = X[top][top] (lower triangular: rows k >= column)
    Am = d.W; a_r0 = r0 + b; a_c0 = r0; Bm = d.X; b_r0 = r0; b_c0 = r0; Om = d.L;
    klo = cb & ~15; khi = b;
  } else {                 // A = X[bot][bot] (lower triangular: columns k <= row), B = T[bot][top]
    Am = d.X; a_r0 = r0 + b; a_c0 = r0 + b; Bm = d.L; b_r0 = r0 + b; b_c0 = r0; Om = d.X;
    klo = 0; khi = min(b2, rb + 64);
  }
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (int k0 = klo; k0 < khi; k0 += 16) {
    for (int e = t; e < 64 * 16; e += 256) {
      const int kk = e & 15, rr = e >> 4;              // lanes along k: 64-byte runs of a row of A
      const int gr = rb + rr, gk = k0 + kk;
      float v = 0.f;
      if (gr < b2 && gk < khi && !(phase == 1 && gk > gr)) v = Am[(size_t)(a_r0 + gr) * d.D + a_c0 + gk];
      As[kk][rr] = v;
    }
    for (int e = t; e < 64 * 16; e += 256) {
      const int cc = e & 63, kk = e >> 6;              // lanes along the columns: contiguous rows of B
      const int gc = cb + cc, gk = k0 + kk;
      float v = 0.f;
      if (gc < b && gk < khi && !(phase == 0 && gk < gc)) v = Bm[(size_t)(b_r0 + gk) * d.D + b_c0 + gc];
      Bs[kk][cc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float sign = phase == 0 ? 1.f : -1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = rb + ty * 4 + i;
    if (gr >= b2) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = cb + tx * 4 + j;
      if (gc < b) Om[(size_t)(r0 + b + gr) * d.D + r0 + gc] = sign * acc[i][j];
    }
  }
}

// ---- epilogue: L[i][j] = X[D-1-j][D-1-i] for j <= i, 0 above the diagonal ----------------------
__global__ void __launch_bounds__(256) chol_epilogue_kernel(const MatDesc* __restrict__ descs) {
  const MatDesc d = descs[blockIdx.z];
  const size_t total = (size_t)d.D * d.D;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int i = (int)(e / d.D), j = (int)(e - (size_t)i * d.D);
    d.L[e] = (j <= i) ? d.X[(size_t)(d.D - 1 - j) * d.D + (d.D - 1 - i)] : 0.f;
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

size_t chol_workspace(const int* dims, int count) {
  size_t bytes = align_up(sizeof(MatDesc) * (size_t)count, 256);
  for (int i = 0; i < count; ++i) {
    const size_t D = (size_t)dims[i];
    const size_t nb = (D + NB - 1) / NB;
    bytes += align_up(D * D * 4, 256) * 2 + align_up(nb * NB * NB * 4, 256);
  }
  return bytes;
}

int chol_inv_batched_launch(const float* const* F, const int* dims, int count, const float* add,
                            const float* mul, float* const* L_out, int* info, void* ws, size_t ws_bytes,
                            cudaStream_t s) {
  CRV_CHECK(count > 0, "empty batch");
  CRV_CHECK(F && dims && add && mul && L_out && info && ws, "null pointer");
  CRV_CHECK(count <= 65535, "too many matrices in one call");
  CRV_CHECK(ws_bytes >= chol_workspace(dims, count), "workspace too small: %zu < %zu", ws_bytes,
            chol_workspace(dims, count));
  MatDesc* host = new MatDesc[count];
  char* base = (char*)ws;
  size_t off = align_up(sizeof(MatDesc) * (size_t)count, 256);
  int maxD = 0;
  for (int i = 0; i < count; ++i) {
    const size_t D = (size_t)dims[i];
    if (dims[i] <= 0 || !F[i] || !L_out[i] || !(mul[i] >= 0.f) || !(add[i] >= 0.f)) {
      delete[] host;
      set_error("bad matrix %d: D=%d add=%g mul=%g", i, dims[i], (double)add[i], (double)mul[i]);
      return 1;
    }
    const size_t nb = (D + NB - 1) / NB;
    host[i].F = F[i];
    host[i].L = L_out[i];
    host[i].D = dims[i];
    host[i].nb = (int)nb;
    host[i].sqrt_mul = (float)sqrt((double)mul[i]);
    host[i].sqrt_add = (float)sqrt((double)add[i]);
    host[i].W = (float*)(base + off); off += align_up(D * D * 4, 256);
    host[i].X = (float*)(base + off); off += align_up(D * D * 4, 256);
    host[i].Dinv = (float*)(base + off); off += align_up(nb * NB * NB * 4, 256);
    if (dims[i] > maxD) maxD = dims[i];
  }
  cudaError_t e = cudaMemcpyAsync(ws, host, sizeof(MatDesc) * (size_t)count, cudaMemcpyHostToDevice, s);
  // pageable source: the copy has been staged when the call returns, so `host` can go
  delete[] host;
  CRV_CUDA(e);
  CRV_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * (size_t)count, s));
  const MatDesc* descs = (const MatDesc*)ws;
  const int nbmax = (maxD + NB - 1) / NB;
  const unsigned ew = (unsigned)min((size_t)1024, ((size_t)maxD * maxD + 255) / 256);
  chol_prologue_kernel<<<dim3(ew, 1, count), 256, 0, s>>>(descs);
  constexpr int NBO = 128;                               // outer panel: 4 inner panels of 32
  for (int J = 0; J * NBO < maxD; ++J) {
    const int oend = (J + 1) * NBO;
    for (int j = J * (NBO / NB); j < (J + 1) * (NBO / NB) && j < nbmax; ++j) {
      chol_panel_kernel<<<dim3((nbmax - j - 1 + PANEL_RB - 1) / PANEL_RB > 0 ? (nbmax - j - 1 + PANEL_RB - 1) / PANEL_RB : 1, 1, count), 128, 0, s>>>(descs, j, info);
      const int cb = (j + 1) * NB;                       // columns of the outer panel that still need this inner panel's update
      if (cb < oend && cb < maxD) {
        const int ntr = (maxD - cb + 63) / 64, ntc = (std::min(oend, maxD) - cb + 63) / 64;
        chol_update_kernel<<<dim3(ntr, ntc, count), 256, 0, s>>>(descs, j * NB, NB, cb, oend);
      }
    }
    if (oend < maxD) {
      const int nt = (maxD - oend + 63) / 64;
      chol_update_kernel<<<dim3(nt, nt, count), 256, 0, s>>>(descs, J * NBO, NBO, oend, maxD);
    }
  }
  chol_trtri_init_kernel<<<dim3(nbmax, 1, count), 256, 0, s>>>(descs);
  for (int b = NB; b < maxD; b *= 2) {
    const int tpp = b / 64 > 1 ? b / 64 : 1;
    const int pairs = (maxD + 2 * b - 1) / (2 * b);
    const dim3 grid((unsigned)(pairs * tpp * tpp), 1, count);
    chol_trtri_merge_kernel<<<grid, 256, 0, s>>>(descs, b, 0);
    chol_trtri_merge_kernel<<<grid, 256, 0, s>>>(descs, b, 1);
  }
  chol_epilogue_kernel<<<dim3(ew, 1, count), 256, 0, s>>>(descs);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace crv
