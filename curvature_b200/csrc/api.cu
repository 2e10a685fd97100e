// extern "C" boundary of libcurvature_b200.so (see include/curvature_b200.h).
#include "../../include/curvature_b200.h"
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <utility>
#include <mutex>

namespace crv {

static std::recursive_mutex g_api_mutex;

ApiGuard::ApiGuard(const void* ptr, cudaStream_t s) {
  g_api_mutex.lock();
  int want = -1;
  if (ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) want = at.device;
    else cudaGetLastError();
  }
  if (want < 0 && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread) {
    int d = -1;
    if (cudaStreamGetDevice(s, &d) == cudaSuccess) want = d;
    else cudaGetLastError();
  }
  if (want >= 0) {
    if (cudaGetDevice(&prev_) != cudaSuccess) { cudaGetLastError(); prev_ = -1; }
    if (prev_ != want && cudaSetDevice(want) == cudaSuccess) switched_ = true;
  }
}
ApiGuard::~ApiGuard() {
  if (switched_ && prev_ >= 0) cudaSetDevice(prev_);
  g_api_mutex.unlock();
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int device_sm_count() {
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

// ---- per-kernel event timing --------------------------------------------------------------------
namespace {
struct ProfRec { cudaEvent_t e0, e1; int kclass; double flops, bytes; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
}  // namespace

bool profile_on() { return g_prof_on; }
static long long* g_timeline = nullptr;
long long* debug_timeline_buffer() { return g_timeline; }
// launch-level trace (crv_debug_trace): slot i = 8 words of launch i of the channels-last SYRK path
static unsigned long long* g_trace = nullptr;
static int g_trace_cap = 0, g_trace_next = 0;
unsigned long long* debug_trace_slot() {
  if (!g_trace || g_trace_next >= g_trace_cap) return nullptr;
  return g_trace + 8 * (size_t)(g_trace_next++);
}

void profile_begin(int kclass, double flops, double bytes, cudaStream_t s) {
  if (!g_prof_on) return;
  ProfRec r;
  if (!g_prof_pool.empty()) { r.e0 = g_prof_pool.back().first; r.e1 = g_prof_pool.back().second; g_prof_pool.pop_back(); }
  else { cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); }
  r.kclass = kclass; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
}
void profile_end(cudaStream_t s) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().e1, s);
}

// GEMM dispatch: tensor-core tiers use the tcgen05 kernel and fall back to the CUDA-core kernel only when TMA cannot
// address the operands (leading dimension not a multiple of 4 floats, e.g. the 2049-wide fc factor with its bias column)
static int gemm_dispatch(int precision, const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                         long long sb_n, float* C, int ldc, int m, int n, int k, float alpha, float beta, int epilogue,
                         const SampleEpilogue* sample, int round_out, cudaStream_t s) {
  if (precision != CRV_PREC_FP32) {
    const int rc = gemm_tc_launch(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, epilogue, sample, round_out, s);
    if (rc >= 0) return rc;
  }
  return gemm_simt_launch(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, epilogue, sample, s);
}

static int syrk_dispatch(const ConvGeom& g, float alpha, float* F, void* ws, size_t ws_bytes, int precision,
                         cudaStream_t s) {
  if (precision == CRV_PREC_FP32) return syrk_simt_launch(g, alpha, F, s);
  if (precision == CRV_PREC_BF16X3) return syrk_simt_launch(g, alpha, F, s);   // 1e-5 tier, operand not TMA-addressable
  if (precision == CRV_PREC_BF16) precision = CRV_PREC_TF32;                   // 1e-3 tier, likewise: thread-staged TF32
  if (precision == CRV_PREC_TF32 || precision == CRV_PREC_TF32_TMA)
    return syrk_tc_launch(g, alpha, F, precision, ws, ws_bytes, s);
  set_error("unknown precision tier %d", precision);
  return 1;
}

}  // namespace crv

using namespace crv;

extern "C" {

int crv_abi_version(void) { return CRV_ABI_VERSION; }
const char* crv_last_error(void) { return last_error(); }
int crv_device_sm_count(void) { return device_sm_count(); }

int crv_debug_timeline(long long* buf) { g_timeline = buf; return 0; }
int crv_debug_trace(unsigned long long* buf, int slots) {
  g_trace = buf; g_trace_cap = buf ? slots : 0; g_trace_next = 0;
  return 0;
}
int crv_debug_trace_count(void) { return g_trace_next; }

int crv_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}

int crv_profile_collect(double* ms, double* flops, double* bytes, long long* launches, int nclasses) {
  ApiGuard guard_(nullptr);
  CRV_CHECK(ms && flops && bytes && launches && nclasses >= KC_COUNT, "crv_profile_collect: need %d classes", (int)KC_COUNT);
  for (int i = 0; i < nclasses; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  for (auto& r : g_prof) {
    CRV_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    CRV_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.kclass] += t; flops[r.kclass] += r.flops; bytes[r.kclass] += r.bytes; launches[r.kclass] += 1;
    g_prof_pool.push_back(std::make_pair(r.e0, r.e1));
  }
  g_prof.clear();
  return 0;
}

size_t crv_workspace_bytes(int op, const int64_t* dims, int ndims) {
  switch (op) {
    case CRV_OP_SYRK_CONV: {
      if (ndims < 12) return 0;
      ConvGeom g;
      static const float dummy = 0.f;
      if (make_geom(g, &dummy, (int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], (int)dims[4],
                    (int)dims[5], (int)dims[6], (int)dims[7], (int)dims[8], (int)dims[9], (int)dims[10]))
        return 0;
      return dims[11] == CRV_PREC_FP32 ? 0 : syrk_tc_workspace(g, (int)dims[11]);
    }
    case CRV_OP_SYRK_ROWS: {
      if (ndims < 5) return 0;
      ConvGeom g;
      static const float dummy = 0.f;
      if (make_geom(g, &dummy, (int)dims[0], (int)dims[1], 1, (int)dims[2], 1, 1, 1, 1, 0, 0, (int)dims[3]))
        return 0;
      return dims[4] == CRV_PREC_FP32 ? 0 : syrk_tc_workspace(g, (int)dims[4]);
    }
    case CRV_OP_SYRK_CONV_NHWC: {
      if (ndims < 12) return 0;
      ConvGeom g;
      static const float dummy = 0.f;
      if (make_geom(g, &dummy, (int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], (int)dims[4],
                    (int)dims[5], (int)dims[6], (int)dims[7], (int)dims[8], (int)dims[9], (int)dims[10]))
        return 0;
      g.x = nullptr;   // alignment of the real pointer is checked at launch
      return syrk_nhwc_workspace(g, (int)dims[11]);
    }
    case CRV_OP_SYRK_ROWS_NHWC: {
      if (ndims < 5) return 0;
      ConvGeom g;
      static const float dummy = 0.f;
      if (make_geom(g, &dummy, (int)dims[0], (int)dims[1], 1, (int)dims[2], 1, 1, 1, 1, 0, 0, (int)dims[3]))
        return 0;
      g.x = nullptr;
      return syrk_nhwc_workspace(g, (int)dims[4]);
    }
    case CRV_OP_EFB_PROJECT:
    case CRV_OP_SAMPLE_MN:
      if (ndims < 2) return 0;
      return (size_t)dims[0] * (size_t)dims[1] * sizeof(float) * (op == CRV_OP_SAMPLE_MN ? 2 : 1);
    case CRV_OP_CHOL_INV: {
      if (ndims < 1 || ndims < 1 + dims[0]) return 0;
      const int count = (int)dims[0];
      int* d = new int[count];
      for (int i = 0; i < count; ++i) d[i] = (int)dims[1 + i];
      const size_t b = chol_workspace(d, count);
      delete[] d;
      return b;
    }
    default:
      return 0;
  }
}

int crv_syrk_conv_accum(const float* x, int N, int C, int H, int W, int kh, int kw, int sh, int sw, int ph,
                        int pw, int has_bias, float alpha, float* A, void* ws, size_t ws_bytes,
                        int precision, crv_stream_t stream) {
  ApiGuard guard_(x, (cudaStream_t)stream);
  ConvGeom g;
  if (int rc = make_geom(g, x, N, C, H, W, kh, kw, sh, sw, ph, pw, has_bias)) return rc;
  return syrk_dispatch(g, alpha, A, ws, ws_bytes, precision, (cudaStream_t)stream);
}

int crv_syrk_rows_accum(const float* gptr, int N, int M, int L, int has_bias, float alpha, float* F,
                        void* ws, size_t ws_bytes, int precision, crv_stream_t stream) {
  ApiGuard guard_(gptr, (cudaStream_t)stream);
  ConvGeom g;
  if (int rc = make_geom(g, gptr, N, M, 1, L, 1, 1, 1, 1, 0, 0, has_bias)) return rc;
  return syrk_dispatch(g, alpha, F, ws, ws_bytes, precision, (cudaStream_t)stream);
}

int crv_syrk_conv_accum_nhwc(const float* x, int N, int C, int H, int W, int kh, int kw, int sh, int sw, int ph,
                             int pw, int has_bias, float alpha, float* A, void* ws, size_t ws_bytes,
                             int precision, crv_stream_t stream) {
  ApiGuard guard_(x, (cudaStream_t)stream);
  ConvGeom g;
  if (int rc = make_geom(g, x, N, C, H, W, kh, kw, sh, sw, ph, pw, has_bias)) return rc;
  return syrk_nhwc_launch(g, alpha, A, precision, ws, ws_bytes, (cudaStream_t)stream);
}

int crv_syrk_rows_accum_nhwc(const float* gptr, int N, int M, int L, int has_bias, float alpha, float* F,
                             void* ws, size_t ws_bytes, int precision, crv_stream_t stream) {
  ApiGuard guard_(gptr, (cudaStream_t)stream);
  ConvGeom g;
  if (int rc = make_geom(g, gptr, N, M, 1, L, 1, 1, 1, 1, 0, 0, has_bias)) return rc;
  return syrk_nhwc_launch(g, alpha, F, precision, ws, ws_bytes, (cudaStream_t)stream);
}

static int batch_geoms(const crv_syrk_item* items, int n, std::vector<ConvGeom>& gs, std::vector<float>& alphas,
                       std::vector<float*>& Fs) {
  CRV_CHECK(items != nullptr && n > 0, "empty batch");
  gs.resize(n); alphas.resize(n); Fs.resize(n);
  for (int i = 0; i < n; ++i) {
    const crv_syrk_item& it = items[i];
    if (int rc = make_geom(gs[i], it.x, it.N, it.C, it.H, it.W, it.kh, it.kw, it.sh, it.sw, it.ph, it.pw, 0)) return rc;
    gs[i].x_nchw = it.nchw ? 1 : 0;
    gs[i].zero_mean = it.zero_mean ? 1 : 0;
    alphas[i] = it.alpha;
    Fs[i] = it.F;
  }
  return 0;
}

size_t crv_syrk_batch_nhwc_workspace(const crv_syrk_item* items, int n, int precision) {
  std::vector<ConvGeom> gs;
  std::vector<float> alphas;
  std::vector<float*> Fs;
  if (batch_geoms(items, n, gs, alphas, Fs)) return 0;
  return syrk_nhwc_batch_workspace(gs.data(), n, precision);
}

int crv_debug_partition(const crv_syrk_item* items, int n, int precision, int sms, int which, int* launch_of_item,
                        int* n_launches, int* G, int* q, unsigned* b, int cap, int* pairs, int* nbox_of_pair,
                        int* nb_of_pair, int pair_cap) {
  std::vector<ConvGeom> gs;
  std::vector<float> alphas;
  std::vector<float*> Fs;
  if (int rc = batch_geoms(items, n, gs, alphas, Fs)) return rc;
  const int rc = syrk_nhwc_debug_partition(gs.data(), n, precision, sms, which, launch_of_item, G, q, b, cap, pairs,
                                           nbox_of_pair, nb_of_pair, pair_cap);
  if (rc & 0xFFFF) return rc;
  *n_launches = rc >> 16;
  return 0;
}

int crv_syrk_batch_nhwc(const crv_syrk_item* items, int n, void* ws, size_t ws_bytes, int precision,
                        crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].x : nullptr, (cudaStream_t)stream);
  std::vector<ConvGeom> gs;
  std::vector<float> alphas;
  std::vector<float*> Fs;
  if (int rc = batch_geoms(items, n, gs, alphas, Fs)) return rc;
  return syrk_nhwc_batch_launch(gs.data(), alphas.data(), Fs.data(), n, precision, ws, ws_bytes, (cudaStream_t)stream);
}

int crv_syrk_batch_dense(const crv_syrk_dense_item* items, int n, crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].x : nullptr, (cudaStream_t)stream);
  CRV_CHECK(items != nullptr && n > 0, "empty batch");
  std::vector<ConvGeom> gs(n);
  std::vector<float> alphas(n);
  std::vector<float*> Fs(n);
  for (int i = 0; i < n; ++i) {
    const crv_syrk_dense_item& it = items[i];
    if (int rc = make_geom(gs[i], it.x, it.N, it.C, it.H, it.W, it.kh, it.kw, it.sh, it.sw, it.ph, it.pw, it.has_bias ? 1 : 0))
      return rc;
    alphas[i] = it.alpha;
    Fs[i] = it.F;
  }
  return syrk_simt_batch_launch(gs.data(), alphas.data(), Fs.data(), n, (cudaStream_t)stream);
}

int crv_stream_join(crv_stream_t stream) {
  ApiGuard guard_(nullptr, (cudaStream_t)stream); return syrk_stream_join((cudaStream_t)stream); }
int crv_stream_fork(crv_stream_t stream) {
  ApiGuard guard_(nullptr, (cudaStream_t)stream); return syrk_stream_fork((cudaStream_t)stream); }

int crv_diag_accum(const float* wgrad, const float* bgrad, int M, int K0, float scale, float* state,
                   float* grads_out, crv_stream_t stream) {
  ApiGuard guard_(wgrad, (cudaStream_t)stream);
  return diag_accum_launch(wgrad, bgrad, M, K0, scale, state, grads_out, (cudaStream_t)stream);
}

int crv_diag_accum_batch(const crv_diag_item* items, int n, float scale, crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].wgrad : nullptr, (cudaStream_t)stream);
  CRV_CHECK(items != nullptr && n > 0, "empty batch");
  std::vector<const float*> w(n), b(n);
  std::vector<float*> st(n), go(n);
  std::vector<int> M(n), K0(n);
  for (int i = 0; i < n; ++i) {
    w[i] = items[i].wgrad; b[i] = items[i].bgrad; st[i] = items[i].state; go[i] = items[i].grads_out;
    M[i] = items[i].M; K0[i] = items[i].K0;
  }
  return diag_accum_batch_launch(w.data(), b.data(), M.data(), K0.data(), scale, st.data(), go.data(), n, (cudaStream_t)stream);
}

int crv_gemm(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
             int m, int n, int k, float alpha, float beta, int precision, crv_stream_t stream) {
  ApiGuard guard_(A, (cudaStream_t)stream);
  // op(A)(i,kk): A[i*lda + kk] or, transposed, A[kk*lda + i]
  const long long sa_m = transA ? 1 : lda, sa_k = transA ? lda : 1;
  const long long sb_k = transB ? 1 : ldb, sb_n = transB ? ldb : 1;
  return gemm_dispatch(precision, A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, m, n, k, alpha, beta, EPI_STORE, nullptr, 0,
                       (cudaStream_t)stream);
}

// ---- stream pool for the batched K3 / K5 calls: the per-layer GEMM chains of a model are independent and mostly
// small (a 64 x 576 layer is a handful of CTAs), so they are spread over a few internal streams and overlap; the
// caller's stream forks into the pool and joins it again (device-side dependencies only).
namespace {
constexpr int POOL = 4;
struct StreamPool {
  cudaStream_t s[POOL] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done[POOL] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr;
  bool init = false, ok = false;
};
StreamPool g_pool[16];
StreamPool* stream_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return nullptr; }
  StreamPool& p = g_pool[dev];
  if (!p.init) {
    p.init = true;
    p.ok = cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < POOL && p.ok; ++i)
      p.ok = cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming) == cudaSuccess;
    if (!p.ok) cudaGetLastError();
  }
  return p.ok ? &p : nullptr;
}
// greedy longest-processing-time assignment of items (cost[i]) to POOL lanes; returns the lane of each item
std::vector<int> lanes_by_cost(const std::vector<double>& cost) {
  std::vector<int> order(cost.size()), lane(cost.size());
  for (size_t i = 0; i < cost.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  double load[POOL] = {0, 0, 0, 0};
  for (int i : order) {
    int best = 0;
    for (int l = 1; l < POOL; ++l) if (load[l] < load[best]) best = l;
    lane[i] = best;
    load[best] += cost[i];
  }
  return lane;
}
}  // namespace

static int efb_project_one(const float* QG, const float* QA, const float* G, int M, int K, float* lambdas, float* T,
                           int precision, cudaStream_t stream);
static int sample_mn_one(const float* LG, const float* LA, const float* z, const float* row_scale, int M, int K0,
                         int has_bias, const float* mu_w, const float* mu_b, float* w_out, float* b_out, float* s_out,
                         float* T, int precision, cudaStream_t s);

// ---- K3 / K5 batches: every layer whose operands TMA can address rides in ONE persistent launch of the chain kernel
// (gemm_chain.cu: both GEMMs of every layer, the second waiting on the first through device-side counters, the intermediate
// consumed out of L2); the others (leading dimension not a multiple of 4 floats: the 147-wide stem, the 2049-wide fc
// factor) keep the per-layer path on the stream pool.  Workspace: [chain tables | one (M, K) intermediate per chained
// layer (+ one scaled noise matrix where row_scale is given) | POOL intermediates for the per-layer path].
static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
static bool chain_enabled() {
  static const bool on = !(getenv("CURVATURE_B200_CHAIN") && atoi(getenv("CURVATURE_B200_CHAIN")) == 0);
  return on;
}

// Operands whose rows are not 16-byte aligned (K % 4 != 0: a bias column, e.g. the 2049-wide fc factor) are copied into
// row-padded buffers of the workspace first (a 2-D device copy of a few MB); TMA never reads the padding (the tensor map
// has the logical extents).  Without this such a layer falls to the CUDA-core GEMM and, alone, takes longer than all
// other layers of the model together.
struct PadCopy { float* dst; size_t ld_dst; const float* src; size_t ld_src; int rows, cols; };
struct Bump {
  char* base; size_t off;
  float* take(size_t bytes) { float* p = (float*)(base ? base + off : (char*)16); off += al256(bytes); return p; }
};
static int run_pad_copies(const std::vector<PadCopy>& cps, cudaStream_t s) {
  for (const PadCopy& c : cps)
    CRV_CUDA(cudaMemcpy2DAsync(c.dst, c.ld_dst * sizeof(float), c.src, c.ld_src * sizeof(float), (size_t)c.cols * sizeof(float),
                               (size_t)c.rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}

struct EfbSplit { std::vector<int> chained, single; std::vector<ChainGemm> gemms; std::vector<PadCopy> copies;
                  size_t t_bytes = 0, chain_bytes = 0, pool_bytes = 0; };
static void efb_split(const crv_efb_item* items, int n, int precision, char* tbase, EfbSplit& sp) {
  size_t mx = 0;
  Bump bump{tbase, 0};
  for (int i = 0; items && i < n; ++i) {
    const crv_efb_item& it = items[i];
    const int M = it.M, K = it.K, Kp = (K + 3) & ~3;
    const bool pad = Kp != K;
    const size_t mark = bump.off, ncopies = sp.copies.size();
    float* T = bump.take((size_t)M * Kp * sizeof(float));
    const float* Gop = it.G; const float* QAop = it.QA;
    if (pad) {
      float* Gp = bump.take((size_t)M * Kp * sizeof(float));
      float* QAp = bump.take((size_t)K * Kp * sizeof(float));
      sp.copies.push_back({Gp, (size_t)Kp, it.G, (size_t)K, M, K});
      sp.copies.push_back({QAp, (size_t)Kp, it.QA, (size_t)K, K, K});
      Gop = Gp; QAop = QAp;
    }
    ChainGemm g[2];
    memset(g, 0, sizeof(g));
    // T = QG^T * G: A(i, kk) = QG[kk * M + i]; B(kk, n) = G[kk * ld + n]
    g[0].A = it.QG; g[0].sa_m = 1; g[0].sa_k = M; g[0].B = Gop; g[0].sb_k = Kp; g[0].sb_n = 1;
    g[0].C = T; g[0].ldc = Kp; g[0].m = M; g[0].n = K; g[0].k = M; g[0].alpha = 1.f; g[0].epi = EPI_STORE;
    g[0].round_out = 1; g[0].dep = -1;
    // lambdas += (T * QA)^2: A = T; B(kk, n) = QA[kk * ld + n]
    g[1].A = T; g[1].sa_m = Kp; g[1].sa_k = 1; g[1].B = QAop; g[1].sb_k = Kp; g[1].sb_n = 1;
    g[1].C = it.lambdas; g[1].ldc = K; g[1].m = M; g[1].n = K; g[1].k = K; g[1].alpha = 1.f;
    g[1].epi = EPI_SQUARE_ACCUM; g[1].dep = (int)sp.gemms.size();
    const bool ok = precision != CRV_PREC_FP32 && chain_enabled() && it.QG && it.QA && it.G && it.lambdas &&
                    gemm_chain_supported(g[0]) && gemm_chain_supported(g[1]);
    if (ok) {
      sp.chained.push_back(i);
      sp.gemms.push_back(g[0]); sp.gemms.push_back(g[1]);
    } else {
      bump.off = mark;
      sp.copies.resize(ncopies);
      sp.single.push_back(i);
      mx = std::max(mx, (size_t)M * (size_t)K);
    }
  }
  sp.t_bytes = bump.off;
  sp.chain_bytes = sp.gemms.empty() ? 0 : al256(gemm_chain_workspace(sp.gemms.data(), (int)sp.gemms.size()));
  sp.pool_bytes = (size_t)POOL * al256(mx * sizeof(float));
}

size_t crv_efb_project_batch_workspace(const crv_efb_item* items, int n) {
  EfbSplit sp, fp;                                                   // (the tier is not known here: enough for either)
  efb_split(items, n, CRV_PREC_TF32, nullptr, sp);
  efb_split(items, n, CRV_PREC_FP32, nullptr, fp);
  return std::max(sp.chain_bytes + sp.t_bytes + sp.pool_bytes, fp.pool_bytes) + 512;
}

int crv_efb_project_batch(const crv_efb_item* items, int n, void* ws, size_t ws_bytes, int precision,
                          crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].G : nullptr, (cudaStream_t)stream);
  CRV_CHECK(items != nullptr && n > 0, "empty batch");
  const size_t need = crv_efb_project_batch_workspace(items, n);
  CRV_CHECK(ws && ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  cudaStream_t caller = (cudaStream_t)stream;
  char* base = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  EfbSplit sz;
  efb_split(items, n, precision, nullptr, sz);                       // sizes first, then the real pointers
  EfbSplit sp;
  efb_split(items, n, precision, base + sz.chain_bytes, sp);
  for (int i = 0; i < n; ++i) {
    const crv_efb_item& it = items[i];
    CRV_CHECK(it.QG && it.QA && it.G && it.lambdas, "null pointer in item %d", i);
  }
  if (!sp.chained.empty()) {
    for (int i : sp.chained)
      if (items[i].round_g)     // gradient copy rounded to the nearest TF32 in place (the tensor core would truncate it)
        if (int rc = round_tf32_launch(items[i].G, const_cast<float*>(items[i].G), (size_t)items[i].M * items[i].K, caller)) return rc;
    if (int rc = run_pad_copies(sp.copies, caller)) return rc;
    const int rc = gemm_chain_launch(sp.gemms.data(), (int)sp.gemms.size(), base, sp.chain_bytes, caller);
    CRV_CHECK(rc >= 0, "internal: chain kernel rejected operands it had accepted");
    if (rc) return rc;
  }
  if (sp.single.empty()) return 0;
  StreamPool* pool = stream_pool();
  const int ns = (int)sp.single.size();
  std::vector<double> cost(ns);
  for (int j = 0; j < ns; ++j) {
    const crv_efb_item& it = items[sp.single[j]];
    cost[j] = 2.0 * it.M * it.K * ((double)it.M + it.K) + 2e7;
  }
  const std::vector<int> lane = lanes_by_cost(cost);
  if (pool) {
    CRV_CUDA(cudaEventRecord(pool->fork, caller));
    for (int l = 0; l < POOL; ++l) CRV_CUDA(cudaStreamWaitEvent(pool->s[l], pool->fork, 0));
  }
  const size_t per = sp.pool_bytes / POOL;
  char* pbase = base + sp.chain_bytes + sp.t_bytes;
  for (int j = 0; j < ns; ++j) {
    const crv_efb_item& it = items[sp.single[j]];
    cudaStream_t s = pool ? pool->s[lane[j]] : caller;
    float* T = (float*)(pbase + (pool ? (size_t)lane[j] * per : 0));
    if (it.round_g) {
      if (int rc = round_tf32_launch(it.G, const_cast<float*>(it.G), (size_t)it.M * it.K, s)) return rc;
    }
    if (int rc = efb_project_one(it.QG, it.QA, it.G, it.M, it.K, it.lambdas, T, precision, s)) return rc;
  }
  if (pool)
    for (int l = 0; l < POOL; ++l) {
      CRV_CUDA(cudaEventRecord(pool->done[l], pool->s[l]));
      CRV_CUDA(cudaStreamWaitEvent(caller, pool->done[l], 0));
    }
  return 0;
}

struct SampleSplit { std::vector<int> chained, single; std::vector<ChainGemm> gemms; std::vector<PadCopy> copies;
                     size_t t_bytes = 0, chain_bytes = 0, pool_bytes = 0; };
static void sample_split(const crv_sample_item* items, int n, int precision, char* tbase, SampleSplit& sp) {
  size_t mx = 0;
  Bump bump{tbase, 0};
  for (int i = 0; items && i < n; ++i) {
    const crv_sample_item& it = items[i];
    const int K = it.K0 + (it.has_bias ? 1 : 0), M = it.M, Kp = (K + 3) & ~3;
    const bool pad = Kp != K;
    const size_t mark = bump.off, ncopies = sp.copies.size();
    float* T = bump.take((size_t)M * Kp * sizeof(float));
    const float* zz = it.z;
    if (it.row_scale) zz = bump.take((size_t)K * M * sizeof(float));      // scaled noise (EFB's draw)
    const float* LAop = it.LA;
    if (pad) {
      float* LAp = bump.take((size_t)K * Kp * sizeof(float));
      sp.copies.push_back({LAp, (size_t)Kp, it.LA, (size_t)K, K, K});
      LAop = LAp;
    }
    ChainGemm g[2];
    memset(g, 0, sizeof(g));
    // T = LG * z^T: A = LG (M x M); B(kk, n) = z[n * M + kk]
    g[0].A = it.LG; g[0].sa_m = M; g[0].sa_k = 1; g[0].B = zz; g[0].sb_k = 1; g[0].sb_n = M;
    g[0].C = T; g[0].ldc = Kp; g[0].m = M; g[0].n = K; g[0].k = M; g[0].alpha = 1.f; g[0].epi = EPI_STORE; g[0].round_out = 1;
    g[0].dep = -1;
    // S = T * LA^T: B(kk, n) = LA[n * ld + kk]; epilogue adds the mean and splits weight / bias columns
    g[1].A = T; g[1].sa_m = Kp; g[1].sa_k = 1; g[1].B = LAop; g[1].sb_k = 1; g[1].sb_n = Kp;
    g[1].C = nullptr; g[1].ldc = K; g[1].m = M; g[1].n = K; g[1].k = K; g[1].alpha = 1.f; g[1].epi = 2;
    g[1].se.mu_w = it.mu_w; g[1].se.mu_b = it.mu_b; g[1].se.w_out = it.w_out; g[1].se.b_out = it.b_out; g[1].se.s_out = it.s_out;
    g[1].se.K0 = it.K0; g[1].se.has_bias = it.has_bias ? 1 : 0;
    g[1].dep = (int)sp.gemms.size();
    const bool ok = precision != CRV_PREC_FP32 && chain_enabled() && it.LG && it.LA && it.z && M > 0 && it.K0 > 0 &&
                    gemm_chain_supported(g[0]) && gemm_chain_supported(g[1]);
    if (ok) {
      sp.chained.push_back(i);
      sp.gemms.push_back(g[0]); sp.gemms.push_back(g[1]);
    } else {
      bump.off = mark;
      sp.copies.resize(ncopies);
      sp.single.push_back(i);
      mx = std::max(mx, (size_t)M * K * (it.row_scale ? 2 : 1));
    }
  }
  sp.t_bytes = bump.off;
  sp.chain_bytes = sp.gemms.empty() ? 0 : al256(gemm_chain_workspace(sp.gemms.data(), (int)sp.gemms.size()));
  sp.pool_bytes = (size_t)POOL * al256(mx * sizeof(float));
}

size_t crv_sample_matrix_normal_batch_workspace(const crv_sample_item* items, int n) {
  SampleSplit sp, fp;
  sample_split(items, n, CRV_PREC_TF32, nullptr, sp);
  sample_split(items, n, CRV_PREC_FP32, nullptr, fp);
  return std::max(sp.chain_bytes + sp.t_bytes + sp.pool_bytes, fp.pool_bytes) + 512;
}

int crv_sample_matrix_normal_batch(const crv_sample_item* items, int n, void* ws, size_t ws_bytes, int precision,
                                   crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].LG : nullptr, (cudaStream_t)stream);
  CRV_CHECK(items != nullptr && n > 0, "empty batch");
  const size_t need = crv_sample_matrix_normal_batch_workspace(items, n);
  CRV_CHECK(ws && ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  cudaStream_t caller = (cudaStream_t)stream;
  char* base = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  SampleSplit sz;
  sample_split(items, n, precision, nullptr, sz);
  SampleSplit sp;
  sample_split(items, n, precision, base + sz.chain_bytes, sp);
  if (!sp.chained.empty()) {
    for (size_t j = 0; j < sp.chained.size(); ++j) {
      const crv_sample_item& it = items[sp.chained[j]];
      CRV_CHECK(!it.w_out || it.mu_w, "w_out needs mu_w");
      CRV_CHECK(!it.b_out || it.mu_b, "b_out needs mu_b");
      if (it.row_scale) {         // EFB's draw: the noise is scaled elementwise first (curvatures.py:458)
        const int K = it.K0 + (it.has_bias ? 1 : 0);
        if (int rc = scale_transpose_launch(it.z, it.row_scale, K, it.M, const_cast<float*>(sp.gemms[2 * j].B), caller)) return rc;
      }
    }
    if (int rc = run_pad_copies(sp.copies, caller)) return rc;
    const int rc = gemm_chain_launch(sp.gemms.data(), (int)sp.gemms.size(), base, sp.chain_bytes, caller);
    CRV_CHECK(rc >= 0, "internal: chain kernel rejected operands it had accepted");
    if (rc) return rc;
  }
  if (sp.single.empty()) return 0;
  StreamPool* pool = stream_pool();
  const int ns = (int)sp.single.size();
  std::vector<double> cost(ns);
  for (int j = 0; j < ns; ++j) {
    const crv_sample_item& it = items[sp.single[j]];
    const double M = it.M, K = it.K0 + (it.has_bias ? 1 : 0);
    cost[j] = 2.0 * M * K * (M + K) + 2e7;
  }
  const std::vector<int> lane = lanes_by_cost(cost);
  if (pool) {
    CRV_CUDA(cudaEventRecord(pool->fork, caller));
    for (int l = 0; l < POOL; ++l) CRV_CUDA(cudaStreamWaitEvent(pool->s[l], pool->fork, 0));
  }
  const size_t per = sp.pool_bytes / POOL;
  char* pbase = base + sp.chain_bytes + sp.t_bytes;
  for (int j = 0; j < ns; ++j) {
    const crv_sample_item& it = items[sp.single[j]];
    cudaStream_t s = pool ? pool->s[lane[j]] : caller;
    float* T = (float*)(pbase + (pool ? (size_t)lane[j] * per : 0));
    if (int rc = sample_mn_one(it.LG, it.LA, it.z, it.row_scale, it.M, it.K0, it.has_bias, it.mu_w, it.mu_b, it.w_out,
                               it.b_out, it.s_out, T, precision, s))
      return rc;
  }
  if (pool)
    for (int l = 0; l < POOL; ++l) {
      CRV_CUDA(cudaEventRecord(pool->done[l], pool->s[l]));
      CRV_CUDA(cudaStreamWaitEvent(caller, pool->done[l], 0));
    }
  return 0;
}

// ---- K5c: S stacked samples per layer ----
struct MultiSplit { std::vector<int> chained, single; std::vector<ChainGemm> gemms; std::vector<PadCopy> copies;
                    size_t t_bytes = 0, chain_bytes = 0, pool_bytes = 0; };
static void multi_split(const crv_sample_multi_item* items, int n, int S, int precision, char* tbase, MultiSplit& sp) {
  size_t mx = 0;
  Bump bump{tbase, 0};
  for (int i = 0; items && i < n; ++i) {
    const crv_sample_multi_item& it = items[i];
    const int M = it.M, K = it.K, Kp = (K + 3) & ~3;
    const bool pad = Kp != K;
    const size_t mark = bump.off, ncopies = sp.copies.size(), ngemms = sp.gemms.size();
    float* T = bump.take((size_t)M * S * Kp * sizeof(float));               // (M, S, Kp)
    const float* LAop = it.LA;
    if (pad) {
      float* LAp = bump.take((size_t)K * Kp * sizeof(float));
      sp.copies.push_back({LAp, (size_t)Kp, it.LA, (size_t)K, K, K});
      LAop = LAp;
    }
    bool ok = precision != CRV_PREC_FP32 && chain_enabled() && it.LG && it.LA && it.z && it.s_out && M > 0 && K > 0;
    ChainGemm g;
    const int first = (int)sp.gemms.size();
    // first products: T[:, s, :] = LG * z_s^T.  Unpadded rows: ONE GEMM over the stacked noise (n = S K columns);
    // padded rows: one GEMM per sample, each writing its column slab of T
    const int nfirst = pad ? S : 1;
    for (int sidx = 0; sidx < nfirst && ok; ++sidx) {
      memset(&g, 0, sizeof(g));
      g.A = it.LG; g.sa_m = M; g.sa_k = 1; g.B = it.z + (size_t)sidx * K * M; g.sb_k = 1; g.sb_n = M;
      g.C = T + (size_t)sidx * Kp; g.ldc = S * Kp; g.m = M; g.n = pad ? K : S * K; g.k = M; g.alpha = 1.f; g.epi = EPI_STORE;
      g.round_out = 1; g.dep = -1;
      if (sidx == 0) ok = ok && gemm_chain_supported(g);
      sp.gemms.push_back(g);
    }
    // s_out (M*S, K) = T viewed as (M*S, Kp) * LA^T
    memset(&g, 0, sizeof(g));
    g.A = T; g.sa_m = Kp; g.sa_k = 1; g.B = LAop; g.sb_k = 1; g.sb_n = Kp;
    g.C = it.s_out; g.ldc = K; g.m = M * S; g.n = K; g.k = K; g.alpha = 1.f; g.epi = EPI_STORE;
    g.dep = first; g.dep_div = S; g.dep_count = nfirst;
    ok = ok && gemm_chain_supported(g);
    sp.gemms.push_back(g);
    if (ok) {
      sp.chained.push_back(i);
    } else {
      bump.off = mark;
      sp.copies.resize(ncopies);
      sp.gemms.resize(ngemms);
      sp.single.push_back(i);
      mx = std::max(mx, (size_t)M * K);
    }
  }
  sp.t_bytes = bump.off;
  sp.chain_bytes = sp.gemms.empty() ? 0 : al256(gemm_chain_workspace(sp.gemms.data(), (int)sp.gemms.size()));
  sp.pool_bytes = 2 * al256(mx * sizeof(float));      // per-layer path: intermediate + one dense (M, K) sample
}

size_t crv_sample_matrix_normal_multi_workspace(const crv_sample_multi_item* items, int n, int S) {
  if (S <= 0) return 0;
  MultiSplit sp, fp;
  multi_split(items, n, S, CRV_PREC_TF32, nullptr, sp);
  multi_split(items, n, S, CRV_PREC_FP32, nullptr, fp);
  return std::max(sp.chain_bytes + sp.t_bytes + sp.pool_bytes, fp.pool_bytes) + 512;
}

int crv_sample_matrix_normal_multi(const crv_sample_multi_item* items, int n, int S, void* ws, size_t ws_bytes,
                                   int precision, crv_stream_t stream) {
  ApiGuard guard_(items && n > 0 ? items[0].LG : nullptr, (cudaStream_t)stream);
  CRV_CHECK(items != nullptr && n > 0 && S > 0, "empty batch");
  const size_t need = crv_sample_matrix_normal_multi_workspace(items, n, S);
  CRV_CHECK(ws && ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  cudaStream_t caller = (cudaStream_t)stream;
  char* base = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  MultiSplit sz;
  multi_split(items, n, S, precision, nullptr, sz);
  MultiSplit sp;
  multi_split(items, n, S, precision, base + sz.chain_bytes, sp);
  if (!sp.chained.empty()) {
    if (int rc = run_pad_copies(sp.copies, caller)) return rc;
    const int rc = gemm_chain_launch(sp.gemms.data(), (int)sp.gemms.size(), base, sp.chain_bytes, caller);
    CRV_CHECK(rc >= 0, "internal: chain kernel rejected operands it had accepted");
    if (rc) return rc;
  }
  // layers the chain kernel cannot take: sample by sample through K5 into a dense (M, K) buffer, scattered to (M, S, K)
  char* pbase = base + sp.chain_bytes + sp.t_bytes;
  for (int i : sp.single) {
    const crv_sample_multi_item& it = items[i];
    CRV_CHECK(it.LG && it.LA && it.z && it.s_out, "null pointer in item %d", i);
    float* T = (float*)pbase;
    float* dense = (float*)(pbase + sp.pool_bytes / 2);
    for (int sidx = 0; sidx < S; ++sidx) {
      if (int rc = sample_mn_one(it.LG, it.LA, it.z + (size_t)sidx * it.K * it.M, nullptr, it.M, it.K, 0, nullptr, nullptr, nullptr,
                                 nullptr, dense, T, precision, caller))
        return rc;
      CRV_CUDA(cudaMemcpy2DAsync(it.s_out + (size_t)sidx * it.K, (size_t)S * it.K * sizeof(float), dense, (size_t)it.K * sizeof(float),
                                 (size_t)it.K * sizeof(float), (size_t)it.M, cudaMemcpyDeviceToDevice, caller));
    }
  }
  return 0;
}

int crv_efb_project_accum(const float* QG, const float* QA, const float* G, int M, int K, float* lambdas,
                          void* ws, size_t ws_bytes, int precision, crv_stream_t stream) {
  ApiGuard guard_(G, (cudaStream_t)stream);
  CRV_CHECK(QG && QA && G && lambdas, "null pointer");
  CRV_CHECK(ws && ws_bytes >= (size_t)M * K * sizeof(float), "workspace too small");
  return efb_project_one(QG, QA, G, M, K, lambdas, (float*)ws, precision, (cudaStream_t)stream);
}

static int efb_project_one(const float* QG, const float* QA, const float* G, int M, int K, float* lambdas, float* T,
                           int precision, cudaStream_t stream) {
  // T = QG^T * G        (M x M)^T (M x K)
  if (int rc = gemm_dispatch(precision, QG, 1, M, G, K, 1, T, K, M, K, M, 1.f, 0.f, EPI_STORE, nullptr,
                             precision != CRV_PREC_FP32, (cudaStream_t)stream))
    return rc;
  // lambdas += (T * QA)^2   (M x K)(K x K)
  return gemm_dispatch(precision, T, K, 1, QA, K, 1, lambdas, K, M, K, K, 1.f, 0.f, EPI_SQUARE_ACCUM, nullptr, 0,
                       (cudaStream_t)stream);
}

int crv_chol_inv_batched(const float* const* F, const int* dims, int count, const float* add,
                         const float* mul, float* const* L_out, int* info, void* ws, size_t ws_bytes,
                         crv_stream_t stream) {
  ApiGuard guard_(info, (cudaStream_t)stream);
  return chol_inv_batched_launch(F, dims, count, add, mul, L_out, info, ws, ws_bytes, (cudaStream_t)stream);
}

int crv_sample_matrix_normal(const float* LG, const float* LA, const float* z, const float* row_scale, int M,
                             int K0, int has_bias, const float* mu_w, const float* mu_b, float* w_out,
                             float* b_out, float* s_out, void* ws, size_t ws_bytes, int precision,
                             crv_stream_t stream) {
  ApiGuard guard_(LG, (cudaStream_t)stream);
  const size_t mk0 = (size_t)M * (K0 + (has_bias ? 1 : 0));
  CRV_CHECK(ws && ws_bytes >= mk0 * sizeof(float) * (row_scale ? 2 : 1), "workspace too small");
  return sample_mn_one(LG, LA, z, row_scale, M, K0, has_bias, mu_w, mu_b, w_out, b_out, s_out, (float*)ws, precision,
                       (cudaStream_t)stream);
}

static int sample_mn_one(const float* LG, const float* LA, const float* z, const float* row_scale, int M, int K0,
                         int has_bias, const float* mu_w, const float* mu_b, float* w_out, float* b_out, float* s_out,
                         float* ws, int precision, cudaStream_t stream) {
  const size_t ws_bytes = (size_t)M * (K0 + (has_bias ? 1 : 0)) * sizeof(float) * (row_scale ? 2 : 1);
  CRV_CHECK(LG && LA && z, "null pointer");
  CRV_CHECK(M > 0 && K0 > 0, "bad shape");
  CRV_CHECK(!w_out || mu_w, "w_out needs mu_w");
  CRV_CHECK(!b_out || mu_b, "b_out needs mu_b");
  const int K = K0 + (has_bias ? 1 : 0);
  const size_t mk = (size_t)M * K;
  CRV_CHECK(ws && ws_bytes >= mk * sizeof(float) * (row_scale ? 2 : 1), "workspace too small");
  float* T = (float*)ws;
  const float* zz = z;
  cudaStream_t s = (cudaStream_t)stream;
  if (row_scale) {
    float* Z2 = T + mk;
    if (int rc = scale_transpose_launch(z, row_scale, K, M, Z2, s)) return rc;
    zz = Z2;
  }
  // T = LG * z^T : A = LG (M x M), B(kk, n) = z[n, kk]  -> sb_k = 1, sb_n = M
  if (int rc = gemm_dispatch(precision, LG, M, 1, zz, 1, M, T, K, M, K, M, 1.f, 0.f, EPI_STORE, nullptr,
                             precision != CRV_PREC_FP32, s))
    return rc;
  // S = T * LA^T : B(kk, n) = LA[n, kk] -> sb_k = 1, sb_n = K
  SampleEpilogue se;
  se.mu_w = mu_w; se.mu_b = mu_b; se.w_out = w_out; se.b_out = b_out; se.s_out = s_out;
  se.K0 = K0; se.has_bias = has_bias ? 1 : 0;
  return gemm_dispatch(precision, T, K, 1, LA, 1, K, nullptr, K, M, K, K, 1.f, 0.f, 2, &se, 0, s);
}

int crv_round_tf32(const float* in, float* out, size_t n, crv_stream_t stream) {
  ApiGuard guard_(in, (cudaStream_t)stream);
  return round_tf32_launch(in, out, n, (cudaStream_t)stream);
}

int crv_elementwise_inv_sqrt(const float* v, float add, float mul, float* out, size_t n, crv_stream_t stream) {
  ApiGuard guard_(v, (cudaStream_t)stream);
  return inv_sqrt_launch(v, add, mul, out, n, (cudaStream_t)stream);
}

int crv_diag_sample(const float* z, const float* inv, int M, int K0, int has_bias, const float* mu_w,
                    const float* mu_b, float* w_out, float* b_out, float* s_out, crv_stream_t stream) {
  ApiGuard guard_(z, (cudaStream_t)stream);
  return diag_sample_launch(z, inv, M, K0, has_bias, mu_w, mu_b, w_out, b_out, s_out, (cudaStream_t)stream);
}

}  // extern "C"
