/*
 * curvature_b200 -- C ABI of the B200 (sm_100a) kernels behind the Fisher-estimation
 * hot path of DLR-RM/curvature.
 *
 * The reference has no FFI of its own: its boundary is the Python class API of
 * curvature/curvatures.py, and every "kernel" is an ATen call made from there.  Each entry
 * point below replaces the ATen call sequence cited beside it (file:line in the reference
 * tree); INTEGRATION.md shows the ctypes stub a maintainer would add to the reference.
 *
 * Conventions (all entry points):
 *   - every data pointer is a DEVICE pointer to fp32 memory owned by the caller (torch);
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); calls only
 *     enqueue work, never synchronise the device, never allocate persistent memory;
 *   - the channels-last SYRK calls and the batched K3 / K5 calls run part of their work on a few internal streams
 *     (created lazily, per device); everything is ordered against `stream` with events, see crv_stream_fork / _join;
 *   - scratch memory is passed in as (ws, ws_bytes); size it with crv_workspace_bytes() or the call's own
 *     *_workspace() function; the SAME buffer must be passed to consecutive calls of a fork .. join section (its two
 *     halves are used alternately by overlapping launches);
 *   - return value 0 = ok, non-zero = error; crv_last_error() gives the message (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute call returns an error.
 *   - matrices are dense row-major.
 */
#ifndef CURVATURE_B200_H
#define CURVATURE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRV_ABI_VERSION 5

typedef void* crv_stream_t; /* cudaStream_t */

/* Arithmetic tier of the dense contractions (SYRK / GEMM kernels). */
enum crv_precision {
  CRV_PREC_FP32   = 0, /* CUDA-core fp32 FMA (exact fp32 products); parity tier 1e-5            */
  CRV_PREC_TF32   = 1, /* tcgen05 kind::tf32, operands rounded to nearest (cvt.rna), fp32 accum  */
  CRV_PREC_BF16X3 = 2, /* tcgen05 kind::f16 on a TWO-TERM bf16 split of the fp32 operand, x = hi + lo (both round-to-nearest,
                          16 significand bits survive), made by a pre-pass: X X^T = hi hi^T + hi lo^T + lo hi^T, three MMAs
                          per k-group, fp32 accumulation; parity tier 1e-5 on the tensor cores.  Channels-last entry points
                          (K1c / K1d / K1e); operands need C >= 64, C % 8 == 0 (C % 64 == 0 for k x k filters)            */
  CRV_PREC_BF16   = 3, /* tcgen05 kind::f16 on a bf16 copy of the operand (cast pre-pass), fp32 accum; parity tier
                          1e-3.  Channels-last entry points only; read-once operands stay on the TF32 path    */
  CRV_PREC_TF32_TMA = 4 /* tcgen05 kind::tf32 fed by TMA where the geometry allows (hardware TF32
                           truncation of the fp32 operands; parity tier 1e-3), else as CRV_PREC_TF32 */
};

/* Operations for crv_workspace_bytes(). dims as documented per op. */
enum crv_op {
  CRV_OP_SYRK_CONV   = 0, /* dims = {N,C,H,W,kh,kw,sh,sw,ph,pw,has_bias,precision} */
  CRV_OP_SYRK_ROWS   = 1, /* dims = {N,M,L,has_bias,precision}                     */
  CRV_OP_EFB_PROJECT = 2, /* dims = {M,K}                                          */
  CRV_OP_CHOL_INV    = 3, /* dims = {count, d_0, ..., d_{count-1}}                 */
  CRV_OP_SAMPLE_MN   = 4, /* dims = {M,K}                                          */
  CRV_OP_SYRK_CONV_NHWC = 5, /* dims as CRV_OP_SYRK_CONV; 0 if the geometry is unsupported */
  CRV_OP_SYRK_ROWS_NHWC = 6  /* dims as CRV_OP_SYRK_ROWS; 0 if the geometry is unsupported */
};

int         crv_abi_version(void);
const char* crv_last_error(void);
/* Number of SMs of the current device (0 if there is none). */
int         crv_device_sm_count(void);
size_t      crv_workspace_bytes(int op, const int64_t* dims, int ndims);

/* Measurement aid (bench.py's roofline): when enabled, every SYRK-family kernel launch is bracketed by a CUDA event
 * pair recorded on the launching stream.  crv_profile_collect() waits for the recorded events, returns per kernel
 * class the summed device time (ms), the algorithmic flops and bytes of those launches and their number, and
 * clears the records.  Classes: 0 channels-last SYRK bf16, 1 channels-last SYRK tf32, 2 NCHW staged SYRK,
 * 3 split reduction, 4 cast / rounding pre-pass, 5 fp32 SIMT SYRK.  While it is enabled the library does not use its
 * internal side streams (reductions and pre-passes run in order on the caller's stream), so that every bracket times
 * its kernel alone.  Not thread-safe; off by default. */
#define CRV_KERNEL_CLASSES 6
int         crv_profile_enable(int on);
int         crv_profile_collect(double* ms, double* flops, double* bytes, long long* launches, int nclasses);
/* Profiling aid: when `buf` (device memory, >= 160 * 8 int64) is non-null, every CTA of the channels-last SYRK kernel
 * launched afterwards writes 8 words at buf[8 * cta]: globaltimer ns at entry / after the prologue / at the first
 * operand arrival / when the last MMA was issued / when the accumulator was drained, the number of (CTA, pair)
 * segments, MMA k-groups issued, SM id.  Pass null to switch it off (the default). */
int         crv_debug_timeline(long long* buf);
/* Profiling aid: launch-level trace of the channels-last SYRK path WITH its stream overlap left on.  `buf` is device
 * memory of `slots` x 8 uint64, words {0, 2, 4} of every slot preset to ~0 and the others to 0; launch i (in enqueue
 * order since this call) records globaltimer ns into slot i: [0] / [1] first CTA start / last CTA end of the contraction
 * kernel, [2] / [3] of its reduction, [4] / [5] of its pre-pass (cast / rounding / pack), [6] = order of the first factor
 * | factors << 20 | bf16 << 30, [7] = block pairs.  crv_debug_trace_count() = launches recorded.  Null switches it off. */
int         crv_debug_trace(unsigned long long* buf, int slots);
int         crv_debug_trace_count(void);

/* K1a -- first Kronecker factor of a Conv2d layer, fused implicit im2col + SYRK + running sum:
 *   A[k1,k2] += alpha * sum_r X[k1,r] X[k2,r],   X = unfold(x) in the reference's row order
 *   k = c*kh*kw + i*kw + j, r = n*OH*OW + oh*OW + ow, zero padding, dilation 1, groups 1,
 *   plus a trailing row of ones iff has_bias.  A is (K,K), K = C*kh*kw + has_bias.
 * Replaces F.unfold + permute/contiguous + ones/cat + torch.mm + div + add_
 * (curvature/curvatures.py:329-336, 346-350).  The patch matrix is never written to HBM. */
int crv_syrk_conv_accum(const float* x, int N, int C, int H, int W,
                        int kh, int kw, int sh, int sw, int ph, int pw,
                        int has_bias, float alpha, float* A,
                        void* ws, size_t ws_bytes, int precision, crv_stream_t stream);

/* K1b -- Gram matrix over the non-channel axes of an (N, M, L) tensor (L = 1 for Linear):
 *   F[m1,m2] += alpha * sum_{n,l} g[n,m1,l] g[n,m2,l]    (+ trailing ones row iff has_bias)
 * Second factor G (curvatures.py:339-343, 346-350) and first factor of Linear layers
 * (curvatures.py:332-336).  F is (D,D), D = M + has_bias. */
int crv_syrk_rows_accum(const float* g, int N, int M, int L, int has_bias, float alpha, float* F,
                        void* ws, size_t ws_bytes, int precision, crv_stream_t stream);

/* K1c / K1d -- the same two contractions for CHANNELS-LAST operands: x is the (N,C,H,W) activation stored as
 * [N][H][W][C] (torch.channels_last), g the (N,M,L) operand stored as [N][L][M] (for Linear layers, L = 1, both
 * layouts coincide).  The logical tensor, the row order of the factor and the result are exactly those of K1a /
 * K1b; only the memory layout of the input differs.  This is the TMA-fed path: every filter tap is a shifted box
 * of cp.async.bulk.tensor (zero fill = padding), operands reach tcgen05.mma in MN-major form, no thread touches
 * them.  Tensor-core tiers only (CRV_PREC_TF32: round-to-nearest TF32 copy made by a pre-pass into ws;
 * CRV_PREC_TF32_TMA: the fp32 words are fed as they are, i.e. TF32 truncation; CRV_PREC_BF16: bf16 copy made by
 * a pre-pass for operands that are re-read -- k x k convolutions and factors of >= 3 row blocks --, else as TF32_TMA).  Requirements: has_bias == 0,
 * C >= 32, C % 4 == 0, and C % 32 == 0 when kh*kw > 1; crv_workspace_bytes() returns 0 for an unsupported
 * geometry and the call itself returns an error (callers then use K1a / K1b on an NCHW copy). */
int crv_syrk_conv_accum_nhwc(const float* x, int N, int C, int H, int W,
                             int kh, int kw, int sh, int sw, int ph, int pw,
                             int has_bias, float alpha, float* A,
                             void* ws, size_t ws_bytes, int precision, crv_stream_t stream);
int crv_syrk_rows_accum_nhwc(const float* g, int N, int M, int L, int has_bias, float alpha, float* F,
                             void* ws, size_t ws_bytes, int precision, crv_stream_t stream);

/* K1e -- a whole estimation step's factors in one call: F_i += alpha_i * X_i X_i^T for i < n, every operand
 * channels-last as in K1c / K1d (a rows operand (N, M, L) is the item C = M, H = 1, W = L, 1x1 kernel).  This is what
 * KFAC.update (curvature/curvatures.py:312-350, the loop over layers) maps to: read-once, HBM-bound factors of many
 * layers share kernel launches (one stream-K work list over all of them), re-read operands get a launch each.  Same
 * requirements, tiers and results as K1c / K1d called item by item (the partition of the work differs, every sum is
 * still reduced in a fixed order).  crv_syrk_batch_nhwc_workspace() returns the bytes `ws` must have (0 if any item's
 * geometry is unsupported).
 * Packed small-C path (the ResNet stem, 3 -> 64 channels, 7x7, stride 2): a convolution with C <= 4, kw <= 8 and
 * vertical stride 2 is first packed by a pre-pass into a bf16 tensor Q[N][OH + ceil(kh/2) - 1][OW][64] whose 64
 * "channels" are (input row parity, 8 horizontal taps, 4 channels) -- a kw-fold, not a kh*kw-fold expansion, in ws --
 * which turns it into a ceil(kh/2) x 1 convolution over 64 channels that the TMA-fed kernel takes; the reduction drops
 * the padding rows and writes the factor in the reference's row order.  x may be NCHW-dense (nchw = 1) there. */
typedef struct {
  const float* x;
  int N, C, H, W;
  int kh, kw, sh, sw, ph, pw;
  float alpha;
  float* F;
  int nchw;          /* 1: x is NCHW-dense instead of channels-last: accepted for the packed small-C path below and on the
                        tiers whose pre-pass writes a bf16 copy anyway (CRV_PREC_BF16, CRV_PREC_BF16X3; the pre-pass then
                        also transposes to channels-last; needs C >= 64, C % 8 == 0, C % 64 == 0 for k x k filters)       */
  int zero_mean;     /* 1: the operand may be zero-mean (an output GRADIENT rather than a post-activation input).  Tier
                        CRV_PREC_BF16 only: a Gram matrix of zero-mean rows has no dominant mean component to hide operand
                        rounding behind, its relative error is ~ 2^-9 sqrt(2 D / R): such operands get the two-term split
                        of CRV_PREC_BF16X3 when R < 5 D (few contraction rows per factor row, e.g. Linear layers), so that
                        the tier's stated 1e-3 holds for them too                                                          */
} crv_syrk_item;
size_t crv_syrk_batch_nhwc_workspace(const crv_syrk_item* items, int n, int precision);
int crv_syrk_batch_nhwc(const crv_syrk_item* items, int n, void* ws, size_t ws_bytes, int precision,
                        crv_stream_t stream);
/* K1f -- the factors of a SMALL model in ONE launch: F_i += alpha_i * X_i X_i^T for NCHW-dense operands exactly as K1a
 * (a rows operand (N, M, L) of K1b is the item C = M, H = L, W = 1, 1x1 kernel), on the CUDA cores with exact fp32
 * products -- every tier's tolerance holds.  For models whose factors are too small or too oddly shaped for the
 * TMA-fed kernel (LeNet-5, BASELINE configs[0]: ten factors, C = 1 / 6, bias rows), where the loop over layers of
 * curvature/curvatures.py:312-350 is launch-bound when every factor is a launch of its own. */
typedef struct {
  const float* x;
  int N, C, H, W;
  int kh, kw, sh, sw, ph, pw;
  int has_bias;
  float alpha;
  float* F;
} crv_syrk_dense_item;
int crv_syrk_batch_dense(const crv_syrk_dense_item* items, int n, crv_stream_t stream);

/* Host-only view of the K1e scheduler for `sms` SMs (no device needed; x / F of the items are not dereferenced but x must
 * be non-null and 16-byte aligned): launch_of_item[i] = the launch item i rides in, *n_launches, and for launch `which`
 * the stream-K boundary table -- CTA c starts at (pair q[c], box b[c]), c <= *G, (q[*G], b[*G]) = (*pairs, 0) -- plus
 * per pair of that launch the number of boxes and the stage granularity.  Used by the CPU tests of the partition. */
int crv_debug_partition(const crv_syrk_item* items, int n, int precision, int sms, int which, int* launch_of_item,
                        int* n_launches, int* G, int* q, unsigned* b, int cap, int* pairs, int* nbox_of_pair,
                        int* nb_of_pair, int pair_cap);

/* The channels-last SYRK calls enqueue their split reduction (the kernel that adds the result into the factor) on an
 * internal side stream, so that it overlaps the next call's main kernel.  crv_stream_join() makes `stream` wait (on the
 * device, no host synchronisation) for every reduction still outstanding; call it after the last SYRK call of an
 * estimation step and before anything else reads or writes the factors.  CURVATURE_B200_SIDE_STREAM=0 disables the side
 * stream (everything then runs in order on the caller's stream and the join is a no-op). */
int crv_stream_join(crv_stream_t stream);
/* Optional: declares that every operand tensor of the SYRK calls that follow, up to the next crv_stream_join(), is
 * complete on `stream` at this point.  The cast / rounding pre-pass of call i may then run on a second side stream,
 * concurrently with the main kernel of call i-1.  Without a fork the pre-pass runs in order on the caller's stream. */
int crv_stream_fork(crv_stream_t stream);

/* K2 -- squared-gradient accumulation (Diagonal.update, curvatures.py:151-158; the `diags`
 * part of EFB.update, curvatures.py:431-434):
 *   state[m, k] += scale * G[m,k]^2,  G = [wgrad.view(M,K0) | bgrad]  (bias column last).
 * bgrad may be NULL (no bias -> K = K0).  If grads_out != NULL the concatenated G (M, K) is
 * also written there (input of crv_efb_project_accum).  state may be NULL. */
int crv_diag_accum(const float* wgrad, const float* bgrad, int M, int K0, float scale,
                   float* state, float* grads_out, crv_stream_t stream);

/* K2b -- the same for every parameter group of the model in ONE launch (what Diagonal.update, curvatures.py:141-158,
 * and the diags half of EFB.update, :431-434, loop over): per-layer launches are latency-bound at a few MB each. */
typedef struct {
  const float* wgrad;   /* (M, K0) */
  const float* bgrad;   /* (M) or null */
  int M, K0;
  float* state;         /* (M, K0 + has_bias) running sum, or null */
  float* grads_out;     /* optional copy of [wgrad | bgrad], or null */
} crv_diag_item;
int crv_diag_accum_batch(const crv_diag_item* items, int n, float scale, crv_stream_t stream);

/* K3 -- EFB eigenbasis projection (curvatures.py:427-433):
 *   lambdas[m,k] += ((QG^T * G * QA)[m,k])^2,   QG (M,M), G (M,K), QA (K,K).
 * ws holds the (M,K) intermediate.  precision: CRV_PREC_FP32 = CUDA-core fp32 GEMMs; any tensor-core tier = two
 * tcgen05 TF32 GEMM launches (gemm_tc.cu; the intermediate is rounded to nearest TF32 in the first epilogue, the square
 * is fused into the second), CUDA-core fallback only if K or M is not a multiple of 4. */
int crv_efb_project_accum(const float* QG, const float* QA, const float* G, int M, int K,
                          float* lambdas, void* ws, size_t ws_bytes, int precision,
                          crv_stream_t stream);

/* K4 -- batched damped Cholesky-of-inverse (KFAC.invert, curvatures.py:368-379):
 *   reg = sqrt(mul_i) * F_i + sqrt(add_i) * I;  reg = (reg + reg^T)/2;
 *   L_i = lower Cholesky factor of inv(reg)     (L_i L_i^T = reg^{-1}).
 * F / L_out are HOST arrays of `count` device pointers, dims[i] the matrix orders, add / mul
 * host arrays of per-matrix scalars.  info is a DEVICE int array (count): 0 = ok, j > 0 = the
 * j-th leading minor of the flipped matrix is not positive (reference: RuntimeError -> numpy
 * fallback at curvatures.py:380-383; here the caller raises -- no CPU fallback). */
int crv_chol_inv_batched(const float* const* F, const int* dims, int count,
                         const float* add, const float* mul, float* const* L_out, int* info,
                         void* ws, size_t ws_bytes, crv_stream_t stream);

/* K5 -- matrix-normal posterior draw fused with the parameter write-back
 * (KFAC.sample curvatures.py:391-392 + Curvature._replace :78-82 + the load_state_dict
 * reload of the mean at :119):
 *   S = LG * z^T * LA^T  (M,K);  w_out[m, 0:K0] = mu_w + S[:, 0:K0];  b_out[m] = mu_b + S[:, K0]
 * z is (K, M) exactly as the reference draws it, LA (K,K), LG (M,M), K = K0 + has_bias.
 * If row_scale != NULL (EFB.sample, curvatures.py:458-460) z is first multiplied elementwise
 * by row_scale^T where row_scale is (M,K).  s_out (M,K), if non-NULL, receives S itself
 * (KFAC.sample's return value); w_out / b_out / mu_* may be NULL when only S is wanted.
 * precision as for crv_efb_project_accum. */
int crv_sample_matrix_normal(const float* LG, const float* LA, const float* z, const float* row_scale,
                             int M, int K0, int has_bias,
                             const float* mu_w, const float* mu_b, float* w_out, float* b_out,
                             float* s_out, void* ws, size_t ws_bytes, int precision,
                             crv_stream_t stream);

/* K3b / K5b -- the per-layer loops of EFB.update (curvatures.py:424-433) and sample_and_replace (:117-129 with
 * :387-392 / :453-460) as ONE call each.  The layers' GEMM chains are independent and mostly small, so the library
 * spreads them over a pool of internal streams (balanced by flops) and joins them back into `stream`; every item
 * computes exactly what K3 / K5 compute.  `round_g` rounds the item's gradient copy to the nearest TF32 in place
 * first (tensor-core tiers).  The *_workspace functions return the bytes `ws` must have. */
typedef struct {
  const float* QG;      /* (M, M) */
  const float* QA;      /* (K, K) */
  const float* G;       /* (M, K) gradient copy [wgrad | bgrad] */
  int M, K;
  float* lambdas;       /* (M, K), accumulated into */
  int round_g;
} crv_efb_item;
size_t crv_efb_project_batch_workspace(const crv_efb_item* items, int n);
int crv_efb_project_batch(const crv_efb_item* items, int n, void* ws, size_t ws_bytes, int precision,
                          crv_stream_t stream);
typedef struct {
  const float* LG; const float* LA; const float* z; const float* row_scale;
  int M, K0, has_bias;
  const float* mu_w; const float* mu_b;
  float* w_out; float* b_out; float* s_out;
} crv_sample_item;
size_t crv_sample_matrix_normal_batch_workspace(const crv_sample_item* items, int n);
int crv_sample_matrix_normal_batch(const crv_sample_item* items, int n, void* ws, size_t ws_bytes, int precision,
                                   crv_stream_t stream);

/* K5c -- S posterior samples of every layer in one call (the sampling half of the BNN evaluation loop,
 * scripts/evaluate.py:121-152, which calls sample_and_replace() once per sample): for every item
 *   s_out[m][s][:] = (LG * z_s^T * LA^T)[m][:]   for s < S,   z = [z_0; ...; z_{S-1}] stacked, (S*K, M),
 * i.e. s_out is (M, S, K) row-major.  On the tensor-core tiers the S draws of a layer are ONE pair of GEMMs
 * ((M x M)(M x S K), then (M S x K)(K x K)) and all layers share one persistent launch of the chain kernel; layers TMA
 * cannot address (K or M not a multiple of 4) and the fp32 tier run sample by sample through K5.  The mean is not added:
 * the caller adds S_s to the posterior mean when it installs sample s.  *_workspace gives the bytes `ws` must have. */
typedef struct {
  const float* LG; const float* LA; const float* z;
  int M, K;          /* K includes the bias column */
  float* s_out;      /* (M, S, K) */
} crv_sample_multi_item;
size_t crv_sample_matrix_normal_multi_workspace(const crv_sample_multi_item* items, int n, int S);
int crv_sample_matrix_normal_multi(const crv_sample_multi_item* items, int n, int S, void* ws, size_t ws_bytes,
                                   int precision, crv_stream_t stream);

/* out[i] = in[i] rounded to the nearest TF32 value (in place if out == in).  The tensor-core GEMMs of K3 / K5 read fp32
 * words as TF32 (truncation); operands that stay constant over many calls (EFB eigenbases, inverse factors) are rounded
 * once with this call so that the products carry round-to-nearest instead of truncation error. */
int crv_round_tf32(const float* in, float* out, size_t n, crv_stream_t stream);

/* out[i] = sqrt(1 / (mul * v[i] + add))  (Diagonal.invert :188, EFB.invert :450, INF :526). */
int crv_elementwise_inv_sqrt(const float* v, float add, float mul, float* out, size_t n,
                             crv_stream_t stream);

/* w_out = mu_w + (z * inv)[:, 0:K0], b_out = mu_b + (z * inv)[:, K0]   (Diagonal.sample :193 +
 * _replace); z, inv are (M,K); s_out (nullable) receives z * inv. */
int crv_diag_sample(const float* z, const float* inv, int M, int K0, int has_bias,
                    const float* mu_w, const float* mu_b, float* w_out, float* b_out, float* s_out,
                    crv_stream_t stream);

/* Plain row-major GEMM used by the INF low-rank algebra (curvatures.py:538-600) and tests:
 *   C (m,n) = alpha * op(A) * op(B) + beta * C,  op = transpose iff trans* != 0. */
int crv_gemm(const float* A, int lda, int transA, const float* B, int ldb, int transB,
             float* C, int ldc, int m, int n, int k, float alpha, float beta, int precision,
             crv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CURVATURE_B200_H */
