// Shared declarations of the curvature_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace crv {

// thread-local last-error text surfaced through crv_last_error()
void set_error(const char* fmt, ...);
const char* last_error();

#define CRV_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      ::crv::set_error(__VA_ARGS__);         \
      return 1;                              \
    }                                        \
  } while (0)

#define CRV_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::crv::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                       __FILE__, __LINE__);                                         \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

// The (D x R) matrix whose Gram matrix a SYRK call accumulates, described implicitly:
// row k = c*kh*kw + i*kw + j (+ one trailing row of ones iff has_bias), column
// r = n*L + oh*OW + ow, element x[n, c, oh*sh - ph + i, ow*sw - pw + j] or 0 outside.
// An (N, M, L) "rows" operand is the special case C=M, H=1, W=L, 1x1 kernel.
struct ConvGeom {
  const float* x;
  int N, C, H, W;
  int kh, kw, sh, sw, ph, pw;
  int OH, OW;
  int L;         // OH*OW
  int K0;        // C*kh*kw
  int D;         // K0 + has_bias
  int has_bias;
  long long R;   // N*L
  int x_nchw;    // channels-last entry points only: 1 = x is NCHW-dense (packed small-C path, copy tiers)
  int zero_mean; // channels-last entry points only: crv_syrk_item::zero_mean
};

inline int make_geom(ConvGeom& g, const float* x, int N, int C, int H, int W, int kh, int kw, int sh,
                     int sw, int ph, int pw, int has_bias) {
  CRV_CHECK(x != nullptr, "null input pointer");
  CRV_CHECK(N > 0 && C > 0 && H > 0 && W > 0, "bad tensor shape N=%d C=%d H=%d W=%d", N, C, H, W);
  CRV_CHECK(kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0,
            "bad conv geometry k=(%d,%d) s=(%d,%d) p=(%d,%d)", kh, kw, sh, sw, ph, pw);
  CRV_CHECK(H + 2 * ph >= kh && W + 2 * pw >= kw, "kernel larger than padded input");
  g.x = x; g.N = N; g.C = C; g.H = H; g.W = W;
  g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw; g.ph = ph; g.pw = pw;
  g.OH = (H + 2 * ph - kh) / sh + 1;
  g.OW = (W + 2 * pw - kw) / sw + 1;
  g.L = g.OH * g.OW;
  g.K0 = C * kh * kw;
  g.has_bias = has_bias ? 1 : 0;
  g.D = g.K0 + g.has_bias;
  g.R = (long long)N * g.L;
  g.x_nchw = 0;
  g.zero_mean = 0;
  CRV_CHECK((long long)N * C * H * W < (1LL << 31), "input tensor too large for 32-bit indexing");
  return 0;
}

int device_sm_count();

// Every launching entry point of the C ABI holds one of these for the duration of the call: it serialises the host-side
// launch paths (kernel parameter blocks, stream-K tables, side-stream bookkeeping and the profiler records are
// process-wide state) and makes the device that owns `ptr` (else the device of `s`, else the current one) current, so
// that the per-device stream pools / SM counts the library looks up belong to the tensors it is handed.
class ApiGuard {
 public:
  explicit ApiGuard(const void* ptr, cudaStream_t s = nullptr);
  ~ApiGuard();
  ApiGuard(const ApiGuard&) = delete;
  ApiGuard& operator=(const ApiGuard&) = delete;
 private:
  int prev_ = -1;
  bool switched_ = false;
};

// ---- optional per-kernel CUDA-event timing (crv_profile_*): a scope records one event pair on the launching stream
// around one kernel launch when profiling is enabled, and costs one predictable branch when it is not.
enum KernelClass {
  KC_SYRK_NHWC_BF16 = 0,   // syrk_nhwc_kernel<true>   (TMA-fed MN-major, bf16 operands)
  KC_SYRK_NHWC_TF32 = 1,   // syrk_nhwc_kernel<false>  (TMA-fed MN-major, TF32 operands)
  KC_SYRK_STAGED    = 2,   // syrk_tc_kernel / syrk_tc_tma_kernel (NCHW operands)
  KC_SYRK_REDUCE    = 3,   // syrk_tc_reduce_kernel
  KC_PREPASS        = 4,   // cast_bf16_kernel / round_tf32_kernel
  KC_SYRK_SIMT      = 5,   // syrk_simt_kernel (fp32 tier)
  KC_COUNT          = 6
};
long long* debug_timeline_buffer();
unsigned long long* debug_trace_slot();
bool profile_on();
void profile_begin(int kclass, double flops, double bytes, cudaStream_t s);
void profile_end(cudaStream_t s);

// ---- kernel launchers (one per .cu file) -------------------------------------------------
int syrk_simt_launch(const ConvGeom& g, float alpha, float* F, cudaStream_t s);
int syrk_simt_batch_launch(const ConvGeom* gs, const float* alphas, float* const* Fs, int n, cudaStream_t s);
int syrk_tc_launch(const ConvGeom& g, float alpha, float* F, int precision, void* ws, size_t ws_bytes,
                   cudaStream_t s);
size_t syrk_tc_workspace(const ConvGeom& g, int precision);
// channels-last operands (g.x is [N][H][W][C]); tensor-core tiers only
int syrk_stream_join(cudaStream_t s);
int syrk_stream_fork(cudaStream_t s);
bool syrk_nhwc_supported(const ConvGeom& g, int precision);
size_t syrk_nhwc_workspace(const ConvGeom& g, int precision);
size_t syrk_nhwc_batch_workspace(const ConvGeom* gs, int n, int precision);
int syrk_nhwc_debug_partition(const ConvGeom* gs, int n, int precision, int sms, int which, int* launch_of_item, int* G,
                              int* q, unsigned* b, int cap, int* pairs, int* nbox_of_pair, int* nb_of_pair, int pair_cap);
int syrk_nhwc_batch_launch(const ConvGeom* gs, const float* alphas, float* const* Fs, int n, int precision, void* ws,
                           size_t ws_bytes, cudaStream_t s);
int syrk_nhwc_launch(const ConvGeom& g, float alpha, float* F, int precision, void* ws, size_t ws_bytes,
                     cudaStream_t s);

enum GemmEpilogue {
  EPI_STORE = 0,       // C = alpha*acc + beta*C
  EPI_SQUARE_ACCUM = 1 // C += acc*acc
};
struct SampleEpilogue {   // EPI for K5: split (M,K) result into weight/bias with the mean added
  const float* mu_w; const float* mu_b; float* w_out; float* b_out; float* s_out; int K0; int has_bias;
};
int gemm_simt_launch(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                     long long sb_n, float* C, int ldc, int m, int n, int k, float alpha, float beta,
                     int epilogue, const SampleEpilogue* sample, cudaStream_t s);

// K3 / K5 chains: many GEMMs (with dependencies between them) in ONE persistent launch, 256 x 256 tcgen05 TF32 tiles
// (gemm_chain.cu).  C (m x n) = alpha * A (m x k) * B (k x n) through element strides, epilogues as below.
constexpr int SK_MAX_CTAS = 160;
struct ChainGemm {
  const float* A; long long sa_m, sa_k;
  const float* B; long long sb_k, sb_n;
  float* C; int ldc;
  int m, n, k;
  float alpha;
  int epi;          // EPI_STORE / EPI_SQUARE_ACCUM / 2 = sample epilogue (se)
  int round_out;    // EPI_STORE: round the result to the nearest TF32 (it is the next GEMM's A operand)
  SampleEpilogue se;
  int dep;          // index of an EARLIER GEMM of the same call whose output C is this GEMM's A operand, or -1
  int dep_div;      // rows of this GEMM per row of the producer (S stacked samples: row r of A is row r / S of the
                    // producer's output); 0 or 1 = same rows
  int dep_count;    // number of consecutive producer GEMMs dep, dep + 1, ... that together write the A operand (each a
                    // column slab of it, same rows); 0 or 1 = one
};
bool gemm_chain_supported(const ChainGemm& g);
size_t gemm_chain_workspace(const ChainGemm* gemms, int count);
int gemm_chain_launch(const ChainGemm* gemms, int count, void* ws, size_t ws_bytes, cudaStream_t s);

// tensor-core (TF32) GEMM with the same epilogues; returns -1 when the operands cannot be fed by TMA
int gemm_tc_launch(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                   long long sb_n, float* C, int ldc, int m, int n, int k, float alpha, float beta,
                   int epilogue, const SampleEpilogue* sample, int round_out, cudaStream_t s);
int round_tf32_launch(const float* in, float* out, size_t n, cudaStream_t s);

int diag_accum_batch_launch(const float* const* wgrad, const float* const* bgrad, const int* M, const int* K0, float scale,
                            float* const* state, float* const* grads_out, int n, cudaStream_t s);
int diag_accum_launch(const float* wgrad, const float* bgrad, int M, int K0, float scale, float* state,
                      float* grads_out, cudaStream_t s);
int inv_sqrt_launch(const float* v, float add, float mul, float* out, size_t n, cudaStream_t s);
int diag_sample_launch(const float* z, const float* inv, int M, int K0, int has_bias, const float* mu_w,
                       const float* mu_b, float* w_out, float* b_out, float* s_out, cudaStream_t s);
int scale_transpose_launch(const float* z, const float* row_scale, int K, int M, float* out, cudaStream_t s);

size_t chol_workspace(const int* dims, int count);
int chol_inv_batched_launch(const float* const* F, const int* dims, int count, const float* add,
                            const float* mul, float* const* L_out, int* info, void* ws, size_t ws_bytes,
                            cudaStream_t s);

}  // namespace crv
