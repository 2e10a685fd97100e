// K2 and the other streaming (HBM-bound) kernels of the path: squared-gradient accumulation,
// the elementwise inverts and the Diagonal sampler.  All are one read + one RMW per element;
// grids are sized as a multiple of the SM count and threads walk the flat index grid-stride
// with four consecutive elements per thread (16-byte accesses on the aligned state arrays).
#include "common.cuh"

namespace crv {
namespace {

constexpr int ET = 256;

inline int stream_grid(size_t work_items) {
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  size_t blocks = (work_items + ET - 1) / ET;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// value of the concatenated gradient matrix [wgrad | bgrad] at flat index e of an (M, K) matrix
__device__ __forceinline__ float grad_at(const float* __restrict__ w, const float* __restrict__ b,
                                         int K0, int K, size_t e) {
  if (K == K0) return __ldg(w + e);
  const size_t m = e / K;
  const int k = (int)(e - m * K);
  return (k < K0) ? __ldg(w + m * K0 + k) : __ldg(b + m);
}

__global__ void __launch_bounds__(ET)
diag_accum_kernel(const float* __restrict__ w, const float* __restrict__ b, int K0, int K, size_t total,
                  float scale, float* __restrict__ state, float* __restrict__ grads_out) {
  const size_t nvec = (total + 3) / 4;
  for (size_t v = (size_t)blockIdx.x * ET + threadIdx.x; v < nvec; v += (size_t)gridDim.x * ET) {
    const size_t e0 = v * 4;
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = (e0 + i < total) ? grad_at(w, b, K0, K, e0 + i) : 0.f;
    if (e0 + 3 < total) {
      if (state) {
        float4 s4 = *reinterpret_cast<float4*>(state + e0);
        s4.x += g[0] * g[0] * scale; s4.y += g[1] * g[1] * scale;
        s4.z += g[2] * g[2] * scale; s4.w += g[3] * g[3] * scale;
        *reinterpret_cast<float4*>(state + e0) = s4;
      }
      if (grads_out) *reinterpret_cast<float4*>(grads_out + e0) = make_float4(g[0], g[1], g[2], g[3]);
    } else {
      for (int i = 0; i < 4 && e0 + i < total; ++i) {
        if (state) state[e0 + i] += g[i] * g[i] * scale;
        if (grads_out) grads_out[e0 + i] = g[i];
      }
    }
  }
}

// Whole-model form of diag_accum_kernel: ONE launch for every parameter group of an estimation step.  The flat work list
// is cut into fixed chunks of DB_CHUNK elements; block b looks its chunk's item up in the prefix table (kernel
// parameter) and streams it exactly like the per-layer kernel.  Per-layer launches move one layer's few MB in ~8 us
// each (latency-bound); over the 161 parameter groups of ResNet-50 this is the difference between ~0.5 and ~5 TB/s.
constexpr int DB_MAX = 320;                 // items per launch
constexpr int DB_CHUNK = ET * 4 * 8;        // elements per block: 8 float4 per thread
struct DiagBatch {
  int n;
  float scale;
  int first_chunk[DB_MAX + 1];              // prefix sum of chunks per item
  const float* w[DB_MAX];
  const float* b[DB_MAX];
  float* state[DB_MAX];
  float* grads_out[DB_MAX];
  int M[DB_MAX], K0[DB_MAX];
};
__global__ void __launch_bounds__(ET) diag_accum_batch_kernel(const __grid_constant__ DiagBatch db) {
  int lo = 0, hi = db.n - 1;                // item of this block's chunk (binary search on the prefix table)
  const int c = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (db.first_chunk[mid] <= c) lo = mid; else hi = mid - 1;
  }
  const int it = lo;
  const float* __restrict__ w = db.w[it];
  const float* __restrict__ b = db.b[it];
  float* __restrict__ state = db.state[it];
  float* __restrict__ grads_out = db.grads_out[it];
  const int K0 = db.K0[it], K = K0 + (b ? 1 : 0);
  const size_t total = (size_t)db.M[it] * K;
  const size_t base = (size_t)(c - db.first_chunk[it]) * DB_CHUNK;
  const float scale = db.scale;
  const bool vec_w = (K == K0) && (((uintptr_t)w & 15) == 0);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const size_t e0 = base + ((size_t)u * ET + threadIdx.x) * 4;
    if (e0 >= total) break;
    float g[4];
    if (vec_w && e0 + 3 < total) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(w + e0));
      g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = (e0 + i < total) ? grad_at(w, b, K0, K, e0 + i) : 0.f;
    }
    if (e0 + 3 < total) {
      if (state) {
        float4 s4 = *reinterpret_cast<float4*>(state + e0);
        s4.x += g[0] * g[0] * scale; s4.y += g[1] * g[1] * scale;
        s4.z += g[2] * g[2] * scale; s4.w += g[3] * g[3] * scale;
        *reinterpret_cast<float4*>(state + e0) = s4;
      }
      if (grads_out) *reinterpret_cast<float4*>(grads_out + e0) = make_float4(g[0], g[1], g[2], g[3]);
    } else {
      for (int i = 0; i < 4 && e0 + i < total; ++i) {
        if (state) state[e0 + i] += g[i] * g[i] * scale;
        if (grads_out) grads_out[e0 + i] = g[i];
      }
    }
  }
}

__global__ void __launch_bounds__(ET)
inv_sqrt_kernel(const float* __restrict__ v, float add, float mul, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * ET + threadIdx.x; i < n; i += (size_t)gridDim.x * ET)
    out[i] = sqrtf(1.0f / (mul * __ldg(v + i) + add));
}

__global__ void __launch_bounds__(ET)
diag_sample_kernel(const float* __restrict__ z, const float* __restrict__ inv, int M, int K0, int K,
                   const float* __restrict__ mu_w, const float* __restrict__ mu_b,
                   float* __restrict__ w_out, float* __restrict__ b_out, float* __restrict__ s_out) {
  const size_t total = (size_t)M * K;
  for (size_t e = (size_t)blockIdx.x * ET + threadIdx.x; e < total; e += (size_t)gridDim.x * ET) {
    const float sv = __ldg(z + e) * __ldg(inv + e);
    if (s_out) s_out[e] = sv;
    const size_t m = e / K;
    const int k = (int)(e - m * K);
    if (k < K0) {
      if (w_out) w_out[m * K0 + k] = mu_w[m * K0 + k] + sv;
    } else if (b_out) {
      b_out[m] = mu_b[m] + sv;
    }
  }
}

// out[k, m] = z[k, m] * row_scale[m, k]     (EFB.sample: z *= lambdas.t())
__global__ void __launch_bounds__(ET)
scale_by_transposed_kernel(const float* __restrict__ z, const float* __restrict__ rs, int K, int M,
                           float* __restrict__ out) {
  const size_t total = (size_t)K * M;
  for (size_t e = (size_t)blockIdx.x * ET + threadIdx.x; e < total; e += (size_t)gridDim.x * ET) {
    const size_t k = e / M;
    const size_t m = e - k * M;
    out[e] = __ldg(z + e) * __ldg(rs + m * K + k);
  }
}

}  // namespace

int diag_accum_batch_launch(const float* const* wgrad, const float* const* bgrad, const int* M, const int* K0, float scale,
                            float* const* state, float* const* grads_out, int n, cudaStream_t s) {
  CRV_CHECK(n > 0, "empty batch");
  static DiagBatch db;
  for (int i0 = 0; i0 < n; i0 += DB_MAX) {
    const int cnt = n - i0 < DB_MAX ? n - i0 : DB_MAX;
    db.n = cnt;
    db.scale = scale;
    int chunks = 0;
    for (int k = 0; k < cnt; ++k) {
      const int i = i0 + k;
      CRV_CHECK(wgrad[i] != nullptr, "null weight gradient");
      CRV_CHECK(M[i] > 0 && K0[i] > 0, "bad gradient shape %d x %d", M[i], K0[i]);
      float* st = state ? state[i] : nullptr;
      float* go = grads_out ? grads_out[i] : nullptr;
      CRV_CHECK(st || go, "nothing to write");
      CRV_CHECK((st == nullptr || ((uintptr_t)st & 15) == 0) && (go == nullptr || ((uintptr_t)go & 15) == 0),
                "state / grads_out must be 16-byte aligned");
      const size_t total = (size_t)M[i] * (K0[i] + (bgrad && bgrad[i] ? 1 : 0));
      db.first_chunk[k] = chunks;
      chunks += (int)((total + DB_CHUNK - 1) / DB_CHUNK);
      db.w[k] = wgrad[i]; db.b[k] = bgrad ? bgrad[i] : nullptr; db.state[k] = st; db.grads_out[k] = go;
      db.M[k] = M[i]; db.K0[k] = K0[i];
    }
    db.first_chunk[cnt] = chunks;
    diag_accum_batch_kernel<<<chunks, ET, 0, s>>>(db);
    CRV_CUDA(cudaGetLastError());
  }
  return 0;
}

int diag_accum_launch(const float* wgrad, const float* bgrad, int M, int K0, float scale, float* state,
                      float* grads_out, cudaStream_t s) {
  CRV_CHECK(wgrad != nullptr, "null weight gradient");
  CRV_CHECK(M > 0 && K0 > 0, "bad gradient shape %d x %d", M, K0);
  CRV_CHECK(state || grads_out, "nothing to write");
  const int K = K0 + (bgrad ? 1 : 0);
  const size_t total = (size_t)M * K;
  CRV_CHECK((state == nullptr || ((uintptr_t)state & 15) == 0) &&
            (grads_out == nullptr || ((uintptr_t)grads_out & 15) == 0),
            "state / grads_out must be 16-byte aligned");
  diag_accum_kernel<<<stream_grid((total + 3) / 4), ET, 0, s>>>(wgrad, bgrad, K0, K, total, scale, state,
                                                                  grads_out);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

int inv_sqrt_launch(const float* v, float add, float mul, float* out, size_t n, cudaStream_t s) {
  CRV_CHECK(v && out, "null pointer");
  if (n == 0) return 0;
  inv_sqrt_kernel<<<stream_grid(n), ET, 0, s>>>(v, add, mul, out, n);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

int diag_sample_launch(const float* z, const float* inv, int M, int K0, int has_bias, const float* mu_w,
                       const float* mu_b, float* w_out, float* b_out, float* s_out, cudaStream_t s) {
  CRV_CHECK(z && inv, "null pointer");
  CRV_CHECK(M > 0 && K0 > 0, "bad shape");
  CRV_CHECK(!w_out || mu_w, "w_out needs mu_w");
  CRV_CHECK(!b_out || mu_b, "b_out needs mu_b");
  const int K = K0 + (has_bias ? 1 : 0);
  diag_sample_kernel<<<stream_grid((size_t)M * K), ET, 0, s>>>(z, inv, M, K0, K, mu_w, mu_b, w_out, b_out,
                                                                s_out);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

int scale_transpose_launch(const float* z, const float* row_scale, int K, int M, float* out, cudaStream_t s) {
  CRV_CHECK(z && row_scale && out, "null pointer");
  scale_by_transposed_kernel<<<stream_grid((size_t)K * M), ET, 0, s>>>(z, row_scale, K, M, out);
  CRV_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace crv
