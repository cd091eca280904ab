// libb200fwd.so -- C ABI (include/b200fwd.h) over the hand-written sm_100a kernels.
// Host side only orchestrates: device memory, TMA descriptors, launch order, CUDA graph for decode.
// There is no CPU compute path: without an sm_100 device every compute call fails with B2_ERR_NO_GPU.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types only: the entry points are resolved with dlopen, libb200fwd.so has no link-time NCCL dependency

#include "../../include/b200fwd.h"
#include "attention_tcgen05.cuh"
#include "elementwise.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm2_tcgen05.cuh"
#include "gemm_sk_tcgen05.cuh"

using namespace b2;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CKR(call)            \
  do {                       \
    int _r = (call);         \
    if (_r != B2_OK) return _r; \
  } while (0)

extern "C" const char* b2_last_error(void) { return g_err; }

extern "C" int b2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

static int require_gpu(int device) {
  // a device that passed once is not re-queried (cudaGetDeviceProperties is ~1 ms); atomic: engines on different
  // devices may be created from different host threads
  static std::atomic<uint64_t> ok{0};
  if (device >= 0 && device < 64 && ((ok.load(std::memory_order_acquire) >> device) & 1ull)) return B2_OK;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(B2_ERR_NO_GPU, "no CUDA device: libb200fwd has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(B2_ERR_ARG, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return fail(B2_ERR_NO_GPU, "device %d is sm_%d%d; this library contains sm_100a code only", device, p.major, p.minor);
  if (device < 64) ok.fetch_or(1ull << device, std::memory_order_release);
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// NCCL (optional tensor-parallel path): resolved at run time.  The process usually already has a libnccl.so.2 loaded
// (torch's); RTLD_NOLOAD picks that one up, otherwise the system library is opened.
// ------------------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static const NcclApi* nccl_api() {
  static NcclApi api;
  static std::atomic<int> state{0};   // 0 = not tried, 1 = ok, -1 = unavailable
  static std::atomic_flag busy = ATOMIC_FLAG_INIT;
  if (state.load(std::memory_order_acquire) == 1) return &api;
  while (busy.test_and_set(std::memory_order_acquire)) {}
  if (state.load() == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    bool ok = h != nullptr;
    if (ok) {
      api.h = h;
#define B2_SYM(field, name) ok = ok && ((*reinterpret_cast<void**>(&api.field) = dlsym(h, name)) != nullptr)
      B2_SYM(GetUniqueId, "ncclGetUniqueId");
      B2_SYM(CommInitRank, "ncclCommInitRank");
      B2_SYM(CommDestroy, "ncclCommDestroy");
      B2_SYM(AllReduce, "ncclAllReduce");
      B2_SYM(AllGather, "ncclAllGather");
      B2_SYM(GetErrorString, "ncclGetErrorString");
#undef B2_SYM
    }
    state.store(ok ? 1 : -1, std::memory_order_release);
  }
  busy.clear(std::memory_order_release);
  return state.load() == 1 ? &api : nullptr;
}
#define CKN(call)                                                                                         \
  do {                                                                                                    \
    ncclResult_t _r = (call);                                                                             \
    if (_r != ncclSuccess)                                                                                \
      return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 2-D row-major tensor [rows, cols] (cols contiguous), 128-byte-wide box rows, 128B swizzle.
// bf16: box = {64 columns, box_rows} (MMA operand tiles); fp32: box = {32 columns, box_rows} (epilogue slabs).
static int make_tmap_t(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows, bool f32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const uint64_t esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u ptr=%p f32=%d",
                static_cast<int>(r), (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems,
                box_rows, ptr, (int)f32);
  return B2_OK;
}
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  return make_tmap_t(m, ptr, rows, cols, ld_elems, box_rows, false);
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
struct Stats {
  int64_t launches = 0, gemm_launches = 0, attn_launches = 0;
  bool timing = false;     // b2_set_timing(on > 0)
  int every = 1;           // events are recorded on every `every`-th forward only
  bool sample = false;     // this forward records events
  int64_t forwards = 0, timed_forwards = 0;
  struct Ev {
    cudaEvent_t a, b;
    double flops;
    int kind;
  };
  std::vector<Ev> used;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
};

// Every kernel goes through this: cudaLaunchKernelEx with the programmatic-stream-serialization attribute (PDL),
// so a kernel's launch + prologue overlap its predecessor's tail (B2_NO_PDL=1 turns it off).
static bool pdl_enabled() {
  static const bool on = getenv("B2_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// SM count of the current device (queried once per device; atomics: one host thread per engine is allowed)
static std::atomic<int> g_num_sms[64];
static int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    const int c = g_num_sms[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) g_num_sms[dev].store(n, std::memory_order_relaxed);
  return n;
}

// cudaFuncSetAttribute(max dynamic smem) once per (kernel instantiation, device).  `mask` is the instantiation's own
// static bit set; a benign double-set under a race is harmless, the atomic keeps it data-race free.
template <typename Kern>
static int smem_attr_once(Kern kern, int bytes, std::atomic<uint64_t>& mask, bool max_carveout = false) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull)) return B2_OK;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (max_carveout) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  return B2_OK;
}

template <int BN, int EPI>
static int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                         cudaStream_t st) {
  auto kern = gemm_bf16_kernel<BN, EPI>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, GemmCfg<BN>::SMEM_BYTES, attr));
  const int m_blocks = (p.M + GEMM_BM - 1) / GEMM_BM, n_blocks = (p.N + BN - 1) / BN;
  const int total = m_blocks * n_blocks * p.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  CK(launch_k(kern, dim3(grid), dim3(GEMM_THREADS), GemmCfg<BN>::SMEM_BYTES, st, ta, tb, tc, p));
  return B2_OK;
}

// tc: fp32 output map ({32 x 32} boxes) for EPI_F32 / EPI_RESID; any valid map otherwise (unused)
// 2-CTA (cta_group::2) variant: 256 x 256 tiles per CTA pair; tb must have {64 x 128} boxes (each CTA loads half of W)
template <int EPI>
static int launch_gemm2_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, GemmParams p, cudaStream_t st) {
  auto kern = gemm2_bf16_kernel<EPI>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, Gemm2Cfg::SMEM_BYTES, attr));
  p.band_m = p.band_m > 1 ? p.band_m / 2 : 1;          // bands are counted in 256-row tiles here
  const int total = ((p.M + 255) / 256) * ((p.N + 255) / 256);
  int clusters = num_sms() / 2;
  if (total < clusters) clusters = total;
  CK(launch_k(kern, dim3(2 * clusters), dim3(GEMM_THREADS), Gemm2Cfg::SMEM_BYTES, st, ta, tb, tc, p));
  return B2_OK;
}
static int launch_gemm2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                        cudaStream_t st) {
  if (p.split_k != 1) return fail(B2_ERR_ARG, "2-CTA GEMM has no split-K");
  if (epi == EPI_BF16) return launch_gemm2_t<EPI_BF16>(ta, tb, tc, p, st);
  if (epi == EPI_F32) return launch_gemm2_t<EPI_F32>(ta, tb, tc, p, st);
  if (epi == EPI_RESID) return launch_gemm2_t<EPI_RESID>(ta, tb, tc, p, st);
  if (epi == EPI_QKV_ROPE) return launch_gemm2_t<EPI_QKV_ROPE>(ta, tb, tc, p, st);
  if (epi == EPI_GLU) return launch_gemm2_t<EPI_GLU>(ta, tb, tc, p, st);
  return fail(B2_ERR_ARG, "bad epilogue %d", epi);
}

// few-row cluster split-K GEMM (gemm_sk_tcgen05.cuh): grid = tiles x CK, cluster = the CK CTAs that share an output tile
template <int EPI>
static int launch_gemm_sk_t(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_sk_kernel<EPI>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, GemmSkCfg::SMEM_BYTES, attr));
  const int tiles = ((p.N + SK_BN - 1) / SK_BN);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(tiles * p.split_k);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GemmSkCfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (p.split_k > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = p.split_k;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  CK(cudaLaunchKernelEx(&cfg, kern, ta, tb, ta, p));
  return B2_OK;
}
static int launch_gemm_sk(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  if (p.M > SK_MAX_ROWS || p.M <= 0) return fail(B2_ERR_ARG, "cluster split-K GEMM handles 1..%d rows (got %d)", SK_MAX_ROWS, p.M);
  if (p.K % GEMM_BK || p.N % SK_BN) return fail(B2_ERR_ARG, "cluster split-K GEMM needs K %% 64 == 0 and N %% 64 == 0 (K=%d N=%d)", p.K, p.N);
  if (p.split_k < 1 || p.split_k > SK_MAX_CLUSTER || (p.K / GEMM_BK) % p.split_k)
    return fail(B2_ERR_ARG, "cluster size %d must be 1..%d and divide K/64 = %d", p.split_k, SK_MAX_CLUSTER, p.K / GEMM_BK);
  if (epi == EPI_QKV_ROPE) {
    if (p.head_dim != SK_BN) return fail(B2_ERR_ARG, "cluster split-K QKV epilogue needs head_dim == %d", SK_BN);
    return launch_gemm_sk_t<EPI_QKV_ROPE>(ta, tb, p, st);
  }
  if (epi == EPI_RESID_NORM) return launch_gemm_sk_t<EPI_RESID_NORM>(ta, tb, p, st);
  return fail(B2_ERR_ARG, "no cluster split-K instantiation for epilogue %d", epi);
}
// cluster size for a few-row GEMM with `tiles` output tiles and `kb` k-blocks: as many CTAs as fit one wave
static int pick_cluster(int tiles, int kb, int n_sms) {
  int best = 1;
  for (int ck = 2; ck <= SK_MAX_CLUSTER; ++ck)
    if (kb % ck == 0 && kb / ck >= 2 && tiles * ck <= n_sms) best = ck;
  return best;
}

static int launch_gemm(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const GemmParams& p, cudaStream_t st) {
  if (p.K % GEMM_BK != 0 || p.K <= 0) return fail(B2_ERR_ARG, "GEMM K=%d must be a positive multiple of 64", p.K);
  if (p.N % 32 != 0) return fail(B2_ERR_ARG, "GEMM N=%d must be a multiple of 32", p.N);
  if (epi == EPI_GLU && p.N % 64 != 0) return fail(B2_ERR_ARG, "GLU GEMM N=%d must be a multiple of 64", p.N);
  if (p.split_k < 1 || (p.K / GEMM_BK) % p.split_k != 0)
    return fail(B2_ERR_ARG, "split_k=%d must divide K/64=%d", p.split_k, p.K / GEMM_BK);
  if (p.split_k > 1 && epi != EPI_RESID && epi != EPI_F32) return fail(B2_ERR_ARG, "split_k needs the residual or fp32 epilogue");
  if (p.ss_in && p.M > GEMM_BM) return fail(B2_ERR_ARG, "the split-RMSNorm row scale is a few-row (M <= 128) feature");
  if ((epi == EPI_F32 || epi == EPI_RESID) && p.ldo % 4 != 0) return fail(B2_ERR_ARG, "fp32 output stride %d must be a multiple of 4", p.ldo);
#define GEMM_CASE(BN_, EPI_) \
  if (bn == BN_ && epi == EPI_) return launch_gemm_t<BN_, EPI_>(ta, tb, tc, p, st)
  GEMM_CASE(256, EPI_BF16);
  GEMM_CASE(256, EPI_F32);
  GEMM_CASE(256, EPI_RESID);
  GEMM_CASE(256, EPI_QKV_ROPE);
  GEMM_CASE(256, EPI_GLU);
  GEMM_CASE(128, EPI_BF16);
  GEMM_CASE(128, EPI_F32);
  GEMM_CASE(128, EPI_RESID);
  GEMM_CASE(128, EPI_QKV_ROPE);
  GEMM_CASE(128, EPI_GLU);
  GEMM_CASE(64, EPI_BF16);
  GEMM_CASE(64, EPI_F32);
  GEMM_CASE(64, EPI_RESID);
  GEMM_CASE(64, EPI_QKV_ROPE);
  GEMM_CASE(64, EPI_GLU);
  GEMM_CASE(32, EPI_RESID);
  GEMM_CASE(32, EPI_F32);
#undef GEMM_CASE
  return fail(B2_ERR_ARG, "no GEMM instantiation for bn=%d epi=%d", bn, epi);
}

template <int HD, int NP>
static int launch_attn_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                         cudaStream_t st) {
  auto kern = attn_prefill_kernel<HD, NP>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, AttnCfg<HD, NP>::SMEM_BYTES, attr));
  dim3 grid((p.S + ATT_BM - 1) / ATT_BM, p.NH, p.B);
  CK(launch_k(kern, grid, dim3(ATT_THREADS), AttnCfg<HD, NP>::SMEM_BYTES, st, tq, tk, tv, p));
  return B2_OK;
}

template <int HD>
static int launch_attn_co_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                            cudaStream_t st) {
  auto kern = attn_prefill_co_kernel<HD>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, AttnCoCfg<HD>::SMEM_BYTES, attr, true));
  dim3 grid((p.S + ATT_BM - 1) / ATT_BM, p.NH, p.B);
  CK(launch_k(kern, grid, dim3(ATT_CO_THREADS), AttnCoCfg<HD>::SMEM_BYTES, st, tq, tk, tv, p));
  return B2_OK;
}

template <int HD>
static int launch_attn_persist_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                                 cudaStream_t st) {
  auto kern = attn_prefill_persist_kernel<HD>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, AttnCoCfg<HD>::SMEM_BYTES, attr, true));
  const int items = ((p.S + ATT_BM - 1) / ATT_BM) * p.NH * p.B;
  const int slots = 2 * num_sms();                  // two co-resident CTAs per SM
  CK(launch_k(kern, dim3(items < slots ? items : slots), dim3(ATT_CO_THREADS), AttnCoCfg<HD>::SMEM_BYTES, st, tq, tk, tv, p, items));
  return B2_OK;
}

// prefill attention dispatch: the co-resident kernel (one q tile per CTA, two CTAs per SM) for head_dim 64/128 unless
// B2_ATTN_SINGLE is set or the sequence fits one q tile; the one-CTA-per-SM single-tile kernel otherwise (head_dim 256:
// TMEM/smem budget; precise mode: three MMAs per product).
static int attn_prefill(const bf16* q, int ldq, const bf16* kc, const bf16* vc, const AttnParams& p_in, int HD,
                        cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  AttnParams p = p_in;
#ifdef B2_DIAG   // bring-up / tracing knobs (tools/gpu_diag.py): compiled out of the product library
  if (const char* ev = getenv("B2_V_LBO")) p.v_lbo = atoi(ev);   // MN-major V descriptor overrides
  if (const char* ev = getenv("B2_V_SBO")) p.v_sbo = atoi(ev);
  if (const char* ev = getenv("B2_ATT_CTATRACE")) p.cta_trace = reinterpret_cast<unsigned long long*>(strtoull(ev, nullptr, 0));
  if (const char* ev = getenv("B2_ATT_TRACE")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(ev, nullptr, 0));   // device pointer
#endif
  static const bool force_single = getenv("B2_ATTN_SINGLE") != nullptr;   // read once
  const bool precise = p.q_lo_col != 0;
  const bool co = !precise && !force_single && (HD == 64 || HD == 128) && p.S > ATT_BM && !p.v_lbo && !p.v_sbo;
  const int planes = precise ? 2 : 1;
  const uint64_t plane_rows = p.kv_plane_rows ? static_cast<uint64_t>(p.kv_plane_rows) : static_cast<uint64_t>(p.B) * p.NKV * p.Tmax;
  const uint64_t kv_rows = static_cast<uint64_t>(planes) * (precise ? p.kv_lo_row : plane_rows);
  const int bnkv = precise ? (HD <= 128 ? 64 : 32) : (co ? (HD == 64 ? 128 : 64) : (HD <= 128 ? 128 : 64));
  CKR(make_tmap(&tq, q, static_cast<uint64_t>(p.B) * p.S, static_cast<uint64_t>(planes) * p.NH * HD, ldq, ATT_BM));
  CKR(make_tmap(&tk, kc, kv_rows, HD, HD, bnkv));
  CKR(make_tmap(&tv, vc, kv_rows, HD, HD, bnkv));
  if (precise) {
    if (HD == 64) return launch_attn_t<64, 2>(tq, tk, tv, p, st);
    if (HD == 128) return launch_attn_t<128, 2>(tq, tk, tv, p, st);
    if (HD == 256) return launch_attn_t<256, 2>(tq, tk, tv, p, st);
    return fail(B2_ERR_ARG, "head_dim %d unsupported (64/128/256)", HD);
  }
  static const bool one_tile_ctas = getenv("B2_ATTN_CO_ONESHOT") != nullptr;   // A/B: the one-tile-per-CTA kernel of round 1
  if (co && !one_tile_ctas) return HD == 64 ? launch_attn_persist_t<64>(tq, tk, tv, p, st) : launch_attn_persist_t<128>(tq, tk, tv, p, st);
  if (co) return HD == 64 ? launch_attn_co_t<64>(tq, tk, tv, p, st) : launch_attn_co_t<128>(tq, tk, tv, p, st);
  if (HD == 64) return launch_attn_t<64, 1>(tq, tk, tv, p, st);
  if (HD == 128) return launch_attn_t<128, 1>(tq, tk, tv, p, st);
  if (HD == 256) return launch_attn_t<256, 1>(tq, tk, tv, p, st);
  return fail(B2_ERR_ARG, "head_dim %d unsupported (64/128/256)", HD);
}

// number of T-slices of the split-KV decode kernel for a cache of Tmax positions (fixed per cache so the decode CUDA
// graph keeps its grid): ~8 tiles per CTA, at least ~2 CTAs per SM, never more slices than tiles.
// B2_DECODE_SPLIT_MIN_T=n sends caches of <= n positions to the one-warp-per-(batch, head) kernel instead (bring-up).
static int decode_splits(int B, int NKV, int Tmax) {
  static const int min_t = getenv("B2_DECODE_SPLIT_MIN_T") ? atoi(getenv("B2_DECODE_SPLIT_MIN_T")) : 0;
  if (Tmax <= min_t) return 0;
  const int tiles = (Tmax + 63) / 64;
  int ns = (tiles + 7) / 8;
  const int want = (2 * num_sms() + B * NKV - 1) / (B * NKV);
  if (ns < want) ns = want;
  if (ns > tiles) ns = tiles;
  if (ns > DEC_MAX_SPLITS) ns = DEC_MAX_SPLITS;
  return ns < 1 ? 1 : ns;
}
static size_t decode_part_floats(int B, int NH, int HD) { return static_cast<size_t>(B) * NH * DEC_MAX_SPLITS * (HD + 2); }

template <int HD>
static int launch_decode_split(const bf16* q, int ldq, const bf16* kc, const bf16* vc, bf16* out, int ldo, int B, int NKV,
                               int nrep, int Tmax, const int* t0_ptr, float scale, float softcap, int window, float* part,
                               int* cnt, int nsplit, cudaStream_t st, const KvPages& pages, int plane_rows) {
  auto kern = attn_decode_split_kernel<HD>;
  static std::atomic<uint64_t> attr{0};
  CKR(smem_attr_once(kern, DecSplitCfg<HD>::SMEM_BYTES, attr));
  CUtensorMap tk, tv;
  const uint64_t kv_rows = plane_rows ? static_cast<uint64_t>(plane_rows) : static_cast<uint64_t>(B) * NKV * Tmax;
  CKR(make_tmap(&tk, kc, kv_rows, HD, HD, 64));
  CKR(make_tmap(&tv, vc, kv_rows, HD, HD, 64));
  CK(launch_k(kern, dim3(nsplit, NKV, B), dim3(DEC_SPLIT_THREADS), DecSplitCfg<HD>::SMEM_BYTES, st, tk, tv, q, ldq, out, ldo,
              NKV, nrep, Tmax, t0_ptr, scale, softcap, window, part, cnt, pages));
  return B2_OK;
}

// decode attention dispatch: the split-KV GQA tensor-core kernel when scratch is supplied (plain bf16, n_rep <= 8,
// long cache), else the one-warp-per-(batch, head) kernel (precise mode, short caches, unusual head ratios)
static int attn_decode(const bf16* q, int ldq, const bf16* kc, const bf16* vc, bf16* out, int ldo, int B, int NH,
                       int NKV, int HD, int Tmax, const int* t0_ptr, float scale, float softcap, int window,
                       cudaStream_t st, int q_lo_col = 0, size_t kv_lo_off = 0, int out_lo_col = 0, float* part = nullptr,
                       int* cnt = nullptr, int nsplit = 0, KvPages pages = KvPages{nullptr, 0, 0}, int plane_rows = 0) {
  const int nrep = NH / NKV;
  if (part && cnt && nsplit > 0 && !q_lo_col && nrep >= 1 && nrep <= DEC_MAX_NREP) {
    if (HD == 64) return launch_decode_split<64>(q, ldq, kc, vc, out, ldo, B, NKV, nrep, Tmax, t0_ptr, scale, softcap, window, part, cnt, nsplit, st, pages, plane_rows);
    if (HD == 128) return launch_decode_split<128>(q, ldq, kc, vc, out, ldo, B, NKV, nrep, Tmax, t0_ptr, scale, softcap, window, part, cnt, nsplit, st, pages, plane_rows);
    if (HD == 256) return launch_decode_split<256>(q, ldq, kc, vc, out, ldo, B, NKV, nrep, Tmax, t0_ptr, scale, softcap, window, part, cnt, nsplit, st, pages, plane_rows);
  }
  const int grid = (B * NH + 3) / 4;
#define DEC(HD_)                                                                                              \
  if (HD == HD_) {                                                                                            \
    CK(launch_k(attn_decode_kernel<HD_>, dim3(grid), dim3(128), 0, st, q, ldq, kc, vc, out, ldo, B, NH, NKV, Tmax, t0_ptr, \
                scale, softcap, window, q_lo_col, kv_lo_off, out_lo_col, pages));                             \
    return B2_OK;                                                                                             \
  }
  DEC(64) DEC(128) DEC(256)
#undef DEC
  return fail(B2_ERR_ARG, "head_dim %d unsupported (64/128/256)", HD);
}

// host-side RoPE tables mirroring the reference's fp32 arithmetic (llama3.2_model_numpy.py:42-60):
// inv_freq = 1/(base^(2i/HD)) in fp32, angle = inv_freq*pos in fp32, cos/sin of that fp32 angle.
// Optional HF "llama3" frequency scaling (transformers modeling_rope_utils._compute_llama3_parameters, fp32 like
// there): long wavelengths divided by `factor`, a smooth blend in the medium band; factor <= 1 = off (reference).
static void rope_tables_host(int HD, float base, int max_pos, std::vector<float>& c, std::vector<float>& s,
                             float factor = 0.f, float low_ff = 1.f, float high_ff = 4.f, int old_ctx = 8192) {
  const int half = HD / 2;
  std::vector<float> inv(half);
  for (int i = 0; i < half; ++i) {
    const float e = static_cast<float>(2 * i) / static_cast<float>(HD);
    inv[i] = 1.0f / powf(base, e);
  }
  if (factor > 1.0f) {
    const float oc = static_cast<float>(old_ctx);
    const float low_wl = oc / low_ff, high_wl = oc / high_ff;
    for (int i = 0; i < half; ++i) {
      const float wl = 2.0f * 3.14159265358979323846f / inv[i];
      const float scaled = (wl > low_wl) ? inv[i] / factor : inv[i];
      const float smooth = (oc / wl - low_ff) / (high_ff - low_ff);
      const float blended = (1.0f - smooth) * scaled / factor + smooth * scaled;
      const bool medium = !(wl < high_wl) && !(wl > low_wl);
      inv[i] = medium ? blended : scaled;
    }
  }
  c.resize(static_cast<size_t>(max_pos) * half);
  s.resize(c.size());
  for (int p = 0; p < max_pos; ++p)
    for (int i = 0; i < half; ++i) {
      const float ang = inv[i] * static_cast<float>(p);
      c[static_cast<size_t>(p) * half + i] = static_cast<float>(cos(static_cast<double>(ang)));
      s[static_cast<size_t>(p) * half + i] = static_cast<float>(sin(static_cast<double>(ang)));
    }
}

static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static float f16_to_f32(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) u = sign;
    else {
      int e = -1;
      uint32_t m = man;
      do { ++e; m <<= 1; } while (!(m & 0x400));
      u = sign | ((127 - 15 - e) << 23) | ((m & 0x3ff) << 13);
    }
  } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
  else u = sign | ((exp - 15 + 127) << 23) | (man << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------
constexpr int MAX_SPLIT = 8;

struct Layer {
  bf16 *wqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wd = nullptr;
  float *g_in = nullptr, *g_post_attn = nullptr, *g_pre_ffn = nullptr, *g_post_ffn = nullptr;
  unsigned loaded = 0;  // bit per tensor
};

struct b2_cache {
  b2_engine* e;
  uint64_t uid;   // never reused (heap addresses are): keys the captured decode graphs
  int max_batch, max_tokens, len;
  bf16* kv;      // [L][2][max_batch][NKV][max_tokens][HD]
  int* t0_dev;
  size_t plane;  // elements per (layer, k|v) plane
  int pf;        // planes per tensor: 1, or 2 in precise mode (hi plane followed by lo plane)
  float* dec_part = nullptr;   // split-KV decode partials [max_batch][NH][DEC_MAX_SPLITS][HD+2]
  int* dec_cnt = nullptr;      // per (batch, kv head) arrival counters (self-resetting)
  int dec_splits = 0;
  // paged cache: each (layer, K|V) plane is [n_pages][NKV][page][HD]; table[b][lp] = physical page of logical page lp
  bool paged = false;
  int page = 0, max_lp = 0, n_pages = 0;
  int assigned_lp = 0, next_free = 0;      // logical pages handed out so far (every row alike), next unused physical page
  bool user_table = false;                 // the caller supplied the block table (b2_cache_set_block_table)
  int* table_dev = nullptr;
  std::vector<int> table_host;
  KvPages pages() const { return KvPages{paged ? table_dev : nullptr, page, max_lp}; }
  int plane_rows() const { return static_cast<int>(plane / static_cast<size_t>(e_head_dim)); }
  int e_head_dim = 1;
};

struct GraphKey {
  const void *ids, *logits;
  uint64_t cache_uid;
  int B, mode;
  bool operator<(const GraphKey& o) const {
    return std::tie(ids, logits, cache_uid, B, mode) < std::tie(o.ids, o.logits, o.cache_uid, o.B, o.mode);
  }
};
struct GraphEntry {
  cudaGraphExec_t exec;
  int64_t kernels;                       // kernel nodes in the graph (for the launch counter)
  std::list<GraphKey>::iterator lru;     // position in b2_engine::graph_lru (front = most recently used)
};
constexpr size_t MAX_GRAPHS = 32;        // callers that pass fresh ids/logits buffers every step would otherwise grow the map

struct b2_engine {
  b2_config cfg;
  int device;
  cudaStream_t stream = nullptr;
  bool finalized = false;
  // Shapes of THIS rank's shard (== the model's when tp_world == 1): heads, FFN columns and vocabulary rows are split
  // evenly over the tensor-parallel ranks; hidden_size is never split.
  int tp_rank = 0, tp_world = 1;
  int nh = 0, nkv = 0, I = 0, V = 0;    // local q heads, kv heads, intermediate columns, vocabulary rows
  int q_cols, kv_cols, qkv_n;           // local widths: nh*HD, nkv*HD, (nh + 2 nkv)*HD
  ncclComm_t comm = nullptr;
  float* tp_lg = nullptr;               // vocab-parallel lm_head: local [rows, V_local] logits, gathered [world][rows][V_local]
  float* tp_gather = nullptr;
  size_t tp_lg_elems = 0;
  int64_t comm_calls = 0;
  double comm_bytes = 0;
  // peer-memory path of the tensor-parallel decode step (b2_tp_ipc_handle / b2_tp_set_peers): one IPC-shared buffer per
  // rank = [2 sites][MAX_SPLIT][128][H] fp32 partial planes + flags; peers' buffers mapped with cudaIpcOpenMemHandle
  float* p2p_buf = nullptr;
  size_t p2p_site_floats = 0;
  unsigned* p2p_ctr = nullptr;           // [0..1] round counters, [2..3] blocks-done counters (local)
  void* p2p_peer[8] = {};                // mapped base pointers of the peers' buffers (nullptr for self)
  bool p2p_ready = false;
  int64_t p2p_fused = 0;                 // fused all-reduce+norm kernels launched since the last reset
  bf16* embed = nullptr;
  float* g_final = nullptr;
  bool embed_loaded = false, final_loaded = false;
  std::vector<Layer> layers;
  float *rope_cos = nullptr, *rope_sin = nullptr;
  // workspace
  int ws_tokens = 0, ws_batch = 0;
  float *x = nullptr, *y = nullptr, *xl = nullptr, *part = nullptr;   // part: [MAX_SPLIT][128][H] split-K partials (skinny GEMMs)
  bf16 *xn = nullptr, *q = nullptr, *att = nullptr, *hb = nullptr, *xnl = nullptr;
  float* ss = nullptr;                   // [64][H/32] per-row partial sums of squares (decode step with the norm split across GEMMs)
  int64_t* ids_ws = nullptr;
  float* logits_ws = nullptr;
  size_t logits_ws_elems = 0;
  b2_cache* scratch = nullptr;  // used when the caller passes no cache (use_cache=False)
  std::map<std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, bool>, CUtensorMap> tmaps;
  Stats st;
  bool decode_graph = false;
  std::map<GraphKey, GraphEntry> graphs;
  std::list<GraphKey> graph_lru;
  bool capturing = false;
  std::vector<b2_cache*> caches;         // live caches created on this engine (orphaned, not freed, by b2_engine_destroy)
  std::vector<uint8_t> sliding;          // per layer: 1 = sliding-window layer (HF layer_types); default: even layers
  int n_sms = 148;
};

static void drop_graph(b2_engine* e, std::map<GraphKey, GraphEntry>::iterator it) {
  cudaGraphExecDestroy(it->second.exec);
  e->graph_lru.erase(it->second.lru);
  e->graphs.erase(it);
}

// Descriptor cache.  The descriptor is returned BY VALUE: the map may be cleared (caller-provided output pointers keep
// changing in long-running processes) while a caller still holds descriptors from earlier lookups of the same launch.
static int get_tmap(b2_engine* e, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    CUtensorMap* out, bool f32 = false) {
  auto key = std::make_tuple(ptr, rows, cols, ld, box_rows, f32);
  auto it = e->tmaps.find(key);
  if (it == e->tmaps.end()) {
    if (e->tmaps.size() > 4096) e->tmaps.clear();
    CUtensorMap m;
    CKR(make_tmap_t(&m, ptr, rows, cols, ld, box_rows, f32));
    it = e->tmaps.emplace(key, m).first;
  }
  *out = it->second;
  return B2_OK;
}

static void time_begin(b2_engine* e, cudaStream_t st, double flops, int kind) {
  if (!e->st.sample || e->capturing) return;
  Stats::Ev ev;
  if (!e->st.pool.empty()) {
    ev.a = e->st.pool.back().first;
    ev.b = e->st.pool.back().second;
    e->st.pool.pop_back();
  } else {
    cudaEventCreate(&ev.a);
    cudaEventCreate(&ev.b);
  }
  ev.flops = flops;
  ev.kind = kind;
  cudaEventRecord(ev.a, st);
  e->st.used.push_back(ev);
}
static void time_end(b2_engine* e, cudaStream_t st) {
  if (!e->st.sample || e->capturing) return;
  cudaEventRecord(e->st.used.back().b, st);
}

// rasterisation band: if W fits comfortably in L2 keep bands short (A streams from HBM once); otherwise
// make them long enough to amortise re-reading W once per band
static int pick_band(int N, int K) {
  const double w_bytes = 2.0 * N * K;
  return w_bytes <= 48e6 ? 8 : 32;
}

// C = A . W^T through the engine (cached descriptors, stats)
static int eng_gemm(b2_engine* e, cudaStream_t st, int bn, int epi, const bf16* a, int lda, const bf16* w, GemmParams p,
                    int a_rows = 0, int c_rows = 0) {
  CUtensorMap ta, tb, tc;
  if (e->cfg.precise) {   // A rows are [hi | lo]: contraction runs over 2K with the W k-coordinate wrapping at K
    p.K = 2 * p.k_wrap;
    lda *= 2;
  }
  if (!a_rows) a_rows = e->ws_tokens > p.M ? e->ws_tokens : p.M;
  // prefill-size GEMMs run on CTA pairs (cta_group::2, 256 x 256 tiles): measured 1420 vs 1300 TFLOP/s in-step for the
  // 1-CTA kernel (B2_GEMM1=1 selects the latter)
  static const bool use2 = getenv("B2_GEMM1") == nullptr;
  const bool two_cta = use2 && bn == 256 && p.split_k == 1 && p.N % 256 == 0;
  // few-row GEMMs (decode): the A box covers only the rows that exist (32 / 64) -- A is 2/3 of the TMA bytes of a
  // BN = 64 weight-streaming CTA otherwise
  p.a_box_rows = (!two_cta && p.M <= 64) ? (p.M <= 32 ? 32 : 64) : 0;
  CKR(get_tmap(e, a, a_rows, p.K, lda, p.a_box_rows ? p.a_box_rows : GEMM_BM, &ta));
  CKR(get_tmap(e, w, p.N, p.k_wrap, p.k_wrap, two_cta ? 128 : bn, &tb));
  tc = ta;
  if (epi == EPI_F32 || epi == EPI_RESID) CKR(get_tmap(e, p.out, c_rows ? c_rows : p.M, p.N, p.ldo, 32, &tc, true));
  p.band_m = pick_band(p.N, p.k_wrap);
  // lm_head-like GEMM (fp32 output far larger than L2, W larger than L2), B2_LMHEAD_STREAM=1: one band over all of M with
  // A pinned in L2 (evict-last), W and C streamed (evict-first) -- each W tile then comes from HBM once.  Measured on
  // B200 (same-box A/B): 1-3 % SLOWER per step than the banded rasterisation (the kernel is not
  // DRAM-bound and the single band gives up W-tile reuse inside L2 across concurrent clusters), so it stays opt-in.
  static const bool lm_stream = getenv("B2_LMHEAD_STREAM") != nullptr;
  if (two_cta && epi == EPI_F32 && lm_stream && 2.0 * p.N * p.k_wrap > 96e6 && 2.0 * p.M * p.K <= 80e6) {
    p.stream_out = 1;
    p.band_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  }
  time_begin(e, st, 2.0 * p.M * static_cast<double>(p.N) * p.K, 0);
  int r = two_cta ? launch_gemm2(epi, ta, tb, tc, p, st) : launch_gemm(bn, epi, ta, tb, tc, p, st);
  time_end(e, st);
  e->st.launches++;
  e->st.gemm_launches++;
  return r;
}

// few-row cluster split-K GEMM through the engine (decode step)
static int eng_gemm_sk(b2_engine* e, cudaStream_t st, int epi, const bf16* a, int lda, const bf16* w, GemmParams p) {
  CUtensorMap ta, tb;
  p.a_box_rows = p.M <= 32 ? 32 : 64;
  p.split_k = pick_cluster(p.N / SK_BN, p.K / GEMM_BK, e->n_sms);
  CKR(get_tmap(e, a, e->ws_tokens > p.M ? e->ws_tokens : p.M, p.K, lda, p.a_box_rows, &ta));
  CKR(get_tmap(e, w, p.N, p.K, p.K, SK_BN, &tb));
  time_begin(e, st, 2.0 * p.M * static_cast<double>(p.N) * p.K, 0);
  int r = launch_gemm_sk(epi, ta, tb, p, st);
  time_end(e, st);
  e->st.launches++;
  e->st.gemm_launches++;
  return r;
}

extern "C" void b2_engine_destroy(b2_engine* e);
// device allocations of b2_engine_create (separate so that a failure half-way can be unwound by b2_engine_destroy)
static int engine_alloc(b2_engine* e) {
  const b2_config& c = e->cfg;
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->q_cols = e->nh * c.head_dim;
  e->kv_cols = e->nkv * c.head_dim;
  e->qkv_n = e->q_cols + 2 * e->kv_cols;
  if (e->qkv_n % 32) return fail(B2_ERR_ARG, "fused QKV width %d must be a multiple of 32", e->qkv_n);
  const size_t H = c.hidden_size, I = e->I;
  CK(cudaMalloc(&e->embed, static_cast<size_t>(c.vocab_size) * H * 2));
  CK(cudaMalloc(&e->g_final, H * 4));
  e->layers.resize(c.num_hidden_layers);
  for (auto& L : e->layers) {
    CK(cudaMalloc(&L.wqkv, static_cast<size_t>(e->qkv_n) * H * 2));
    CK(cudaMalloc(&L.wo, H * e->q_cols * 2));
    CK(cudaMalloc(&L.wgu, 2 * I * H * 2));
    CK(cudaMalloc(&L.wd, H * I * 2));
    CK(cudaMalloc(&L.g_in, H * 4));
    CK(cudaMalloc(&L.g_post_attn, H * 4));
    if (c.gemma) {
      CK(cudaMalloc(&L.g_pre_ffn, H * 4));
      CK(cudaMalloc(&L.g_post_ffn, H * 4));
    }
  }
  std::vector<float> cs, sn;
  rope_tables_host(c.head_dim, c.rope_theta, c.max_positions, cs, sn, c.rope_factor, c.rope_low_freq_factor,
                   c.rope_high_freq_factor, c.rope_original_max_positions);
  CK(cudaMalloc(&e->rope_cos, cs.size() * 4));
  CK(cudaMalloc(&e->rope_sin, sn.size() * 4));
  CK(cudaMemcpy(e->rope_cos, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->rope_sin, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
  return B2_OK;
}

extern "C" int b2_engine_create(const b2_config* cfg, int device, b2_engine** out) {
  if (!cfg || !out) return fail(B2_ERR_ARG, "null argument");
  CKR(require_gpu(device));
  const b2_config& c = *cfg;
  if (c.hidden_size % 128 || c.hidden_size > 128 * NORM_MAXV)
    return fail(B2_ERR_ARG, "hidden_size %d must be a multiple of 128 and <= %d", c.hidden_size, 128 * NORM_MAXV);
  if (c.head_dim != 64 && c.head_dim != 128 && c.head_dim != 256)
    return fail(B2_ERR_ARG, "head_dim %d unsupported (64/128/256)", c.head_dim);
  if (c.num_key_value_heads <= 0 || c.num_attention_heads % c.num_key_value_heads)
    return fail(B2_ERR_ARG, "num_attention_heads %d must be a multiple of num_key_value_heads %d",
                c.num_attention_heads, c.num_key_value_heads);
  // the reference raises ValueError on the same condition (llama3.2_model_numpy.py:361-365)
  if (c.hidden_size % c.num_attention_heads)
    return fail(B2_ERR_ARG, "hidden_size must be divisible by num_heads (got `hidden_size`: %d and `num_heads`: %d).",
                c.hidden_size, c.num_attention_heads);
  if (c.intermediate_size % 64) return fail(B2_ERR_ARG, "intermediate_size %d must be a multiple of 64", c.intermediate_size);
  if (c.vocab_size % 32) return fail(B2_ERR_ARG, "vocab_size %d must be a multiple of 32", c.vocab_size);
  if (c.max_positions <= 0) return fail(B2_ERR_ARG, "max_positions must be positive");
  const int tpw = c.tp_world > 1 ? c.tp_world : 1;
  if (tpw > 1) {
    if (c.tp_rank < 0 || c.tp_rank >= tpw) return fail(B2_ERR_ARG, "tp_rank %d outside [0, %d)", c.tp_rank, tpw);
    if (c.precise) return fail(B2_ERR_ARG, "tensor parallelism is not available in precise mode");
    if (c.num_attention_heads % tpw || c.num_key_value_heads % tpw)
      return fail(B2_ERR_ARG, "tp_world %d must divide the %d query heads and the %d key/value heads", tpw, c.num_attention_heads,
                  c.num_key_value_heads);
    if (c.intermediate_size % (64 * tpw)) return fail(B2_ERR_ARG, "tp_world %d: intermediate_size %d must be a multiple of %d", tpw, c.intermediate_size, 64 * tpw);
    if (c.vocab_size % (32 * tpw)) return fail(B2_ERR_ARG, "tp_world %d: vocab_size %d must be a multiple of %d", tpw, c.vocab_size, 32 * tpw);
  }
  CK(cudaSetDevice(device));
  b2_engine* e = new b2_engine();
  e->cfg = c;
  e->device = device;
  e->tp_world = tpw;
  e->tp_rank = tpw > 1 ? c.tp_rank : 0;
  e->nh = c.num_attention_heads / tpw;
  e->nkv = c.num_key_value_heads / tpw;
  e->I = c.intermediate_size / tpw;
  e->V = c.vocab_size / tpw;
  e->n_sms = num_sms();
  e->sliding.assign(c.num_hidden_layers, 0);
  for (int i = 0; i < c.num_hidden_layers; i += 2) e->sliding[i] = 1;   // HF Gemma-2 default layer_types: even layers slide
  const int rc = engine_alloc(e);
  if (rc != B2_OK) {                  // e.g. out of memory: free whatever was allocated
    b2_engine_destroy(e);
    return rc;
  }
  *out = e;
  return B2_OK;
}

static void free_ws(b2_engine* e) {
  cudaFree(e->x); cudaFree(e->y); cudaFree(e->xl); cudaFree(e->part); cudaFree(e->xn); cudaFree(e->q); cudaFree(e->att);
  cudaFree(e->hb); cudaFree(e->xnl); cudaFree(e->ids_ws); cudaFree(e->ss);
  e->ss = nullptr;
  e->x = e->y = e->xl = e->part = nullptr;
  e->xn = e->q = e->att = e->hb = e->xnl = nullptr;
  e->ids_ws = nullptr;
  e->tmaps.clear();
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.second.exec);   // graphs bake workspace pointers in
  e->graphs.clear();
  e->graph_lru.clear();
}

extern "C" void b2_cache_destroy(b2_cache* c) {
  if (!c) return;
  if (c->e) {   // graphs captured on this cache hold its (about to be freed) kv / t0 / scratch pointers
    cudaSetDevice(c->e->device);
    cudaDeviceSynchronize();
    auto& live = c->e->caches;
    for (size_t i = 0; i < live.size(); ++i)
      if (live[i] == c) { live.erase(live.begin() + i); break; }
    for (auto it = c->e->graphs.begin(); it != c->e->graphs.end();) {
      auto cur = it++;
      if (cur->first.cache_uid == c->uid) drop_graph(c->e, cur);
    }
  }
  cudaFree(c->kv);
  cudaFree(c->t0_dev);
  cudaFree(c->dec_part);
  cudaFree(c->dec_cnt);
  cudaFree(c->table_dev);
  delete c;
}

extern "C" void b2_engine_destroy(b2_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  free_ws(e);
  if (e->scratch) b2_cache_destroy(e->scratch);
  for (b2_cache* c : e->caches) c->e = nullptr;   // caches that outlive the engine (GC order) still free their own memory
  cudaFree(e->logits_ws);
  cudaFree(e->tp_lg); cudaFree(e->tp_gather);
  if (e->comm && nccl_api()) nccl_api()->CommDestroy(e->comm);
  for (int p = 0; p < 8; ++p)
    if (e->p2p_peer[p]) cudaIpcCloseMemHandle(e->p2p_peer[p]);
  cudaFree(e->p2p_buf);
  cudaFree(e->p2p_ctr);
  cudaFree(e->embed); cudaFree(e->g_final); cudaFree(e->rope_cos); cudaFree(e->rope_sin);
  for (auto& L : e->layers) {
    cudaFree(L.wqkv); cudaFree(L.wo); cudaFree(L.wgu); cudaFree(L.wd);
    cudaFree(L.g_in); cudaFree(L.g_post_attn); cudaFree(L.g_pre_ffn); cudaFree(L.g_post_ffn);
  }
  for (auto& ev : e->st.used) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto& ev : e->st.pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  cudaStreamDestroy(e->stream);
  delete e;
}

// convert a host tensor to bf16 bits
static void to_bf16_host(const void* src, size_t n, int dtype, std::vector<uint16_t>& out) {
  out.resize(n);
  if (dtype == B2_DTYPE_BF16) memcpy(out.data(), src, n * 2);
  else if (dtype == B2_DTYPE_F32) {
    const float* f = static_cast<const float*>(src);
    for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16(f[i]);
  } else {
    const uint16_t* h = static_cast<const uint16_t*>(src);
    for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16(f16_to_f32(h[i]));
  }
}
static void to_f32_host(const void* src, size_t n, int dtype, float add, std::vector<float>& out) {
  out.resize(n);
  for (size_t i = 0; i < n; ++i) {
    float v;
    if (dtype == B2_DTYPE_F32) v = static_cast<const float*>(src)[i];
    else if (dtype == B2_DTYPE_BF16) {
      uint32_t u = static_cast<uint32_t>(static_cast<const uint16_t*>(src)[i]) << 16;
      memcpy(&v, &u, 4);
    } else v = f16_to_f32(static_cast<const uint16_t*>(src)[i]);
    out[i] = v + add;
  }
}

enum { T_IN = 1, T_Q = 2, T_K = 4, T_V = 8, T_O = 16, T_POST = 32, T_GATE = 64, T_UP = 128, T_DOWN = 256,
       T_PRE_FFN = 512, T_POST_FFN = 1024 };

extern "C" int b2_load_weight(b2_engine* e, const char* name, const void* host, const int64_t* shape, int ndim,
                              int dtype) {
  if (!e || !name || !host || !shape) return fail(B2_ERR_ARG, "null argument");
  if (dtype < 0 || dtype > 2) return fail(B2_ERR_ARG, "unknown dtype %d", dtype);
  CK(cudaSetDevice(e->device));
  const b2_config& c = e->cfg;
  const int64_t H = c.hidden_size, I = c.intermediate_size;
  std::string n(name);
  auto want = [&](int64_t r, int64_t cc) -> int {
    if (cc == 0) {
      if (ndim != 1 || shape[0] != r) return fail(B2_ERR_ARG, "%s: expected shape [%lld]", name, (long long)r);
    } else if (ndim != 2 || shape[0] != r || shape[1] != cc)
      return fail(B2_ERR_ARG, "%s: expected shape [%lld,%lld]", name, (long long)r, (long long)cc);
    return B2_OK;
  };
  std::vector<uint16_t> hb;
  std::vector<float> hf;
  const float gadd = c.gemma ? 1.0f : 0.0f;  // Gemma2RMSNorm_np: weight = gamma + 1 (gemma2_model.py:334)
  if (n == "lm_head.weight") return B2_OK;   // aliased to the embedding, like the reference (:1025-1026)
  if (n == "model.embed_tokens.weight") {
    CKR(want(c.vocab_size, H));
    to_bf16_host(host, static_cast<size_t>(c.vocab_size) * H, dtype, hb);
    CK(cudaMemcpy(e->embed, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    e->embed_loaded = true;
    return B2_OK;
  }
  if (n == "model.norm.weight") {
    CKR(want(H, 0));
    to_f32_host(host, H, dtype, gadd, hf);
    CK(cudaMemcpy(e->g_final, hf.data(), H * 4, cudaMemcpyHostToDevice));
    e->final_loaded = true;
    return B2_OK;
  }
  int li = -1, off = 0;
  if (sscanf(name, "model.layers.%d.%n", &li, &off) < 1 || off == 0 || li < 0 || li >= c.num_hidden_layers)
    return fail(B2_ERR_ARG, "unknown tensor name %s", name);
  Layer& L = e->layers[li];
  std::string t(name + off);
  auto norm = [&](float* dst, unsigned bit) -> int {
    CKR(want(H, 0));
    to_f32_host(host, H, dtype, gadd, hf);
    CK(cudaMemcpy(dst, hf.data(), H * 4, cudaMemcpyHostToDevice));
    L.loaded |= bit;
    return B2_OK;
  };
  // The caller always passes the FULL HF tensor (the reference's load_weights(name) seam); under tensor parallelism this
  // rank keeps its slice: `rows` tensors (q/k/v/gate/up: output features) a contiguous row block, `cols` tensors
  // (o_proj/down_proj: input features) a column block.
  const int64_t tw = e->tp_world, tr = e->tp_rank;
  const size_t esz = dtype == B2_DTYPE_F32 ? 4 : 2;
  auto rows = [&](bf16* dst, int64_t r_full, int64_t cc, unsigned bit) -> int {   // dst gets rows [tr*r_full/tw, +r_full/tw)
    CKR(want(r_full, cc));
    const int64_t r = r_full / tw;
    to_bf16_host(static_cast<const uint8_t*>(host) + static_cast<size_t>(tr) * r * cc * esz, static_cast<size_t>(r) * cc, dtype, hb);
    CK(cudaMemcpy(dst, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    L.loaded |= bit;
    return B2_OK;
  };
  auto cols = [&](bf16* dst, int64_t r, int64_t c_full, unsigned bit) -> int {   // dst [r, c_full/tw] = columns [tr*c_full/tw, ...)
    CKR(want(r, c_full));
    const int64_t cc = c_full / tw;
    to_bf16_host(host, static_cast<size_t>(r) * c_full, dtype, hb);
    CK(cudaMemcpy2D(dst, cc * 2, hb.data() + tr * cc, c_full * 2, cc * 2, r, cudaMemcpyHostToDevice));
    L.loaded |= bit;
    return B2_OK;
  };
  auto glu = [&](int which, unsigned bit) -> int {  // 32-row groups: [gate g | up g]
    CKR(want(I, H));
    const int64_t Il = e->I;
    to_bf16_host(static_cast<const uint8_t*>(host) + static_cast<size_t>(tr) * Il * H * esz, static_cast<size_t>(Il) * H, dtype, hb);
    const size_t grp = 32 * H * 2;
    CK(cudaMemcpy2D(reinterpret_cast<uint8_t*>(L.wgu) + which * grp, 2 * grp, hb.data(), grp, grp, Il / 32,
                    cudaMemcpyHostToDevice));
    L.loaded |= bit;
    return B2_OK;
  };
  const int64_t qg = static_cast<int64_t>(c.num_attention_heads) * c.head_dim, kvg = static_cast<int64_t>(c.num_key_value_heads) * c.head_dim;
  if (t == "input_layernorm.weight") return norm(L.g_in, T_IN);
  if (t == "post_attention_layernorm.weight") return norm(L.g_post_attn, T_POST);
  if (t == "pre_feedforward_layernorm.weight" && c.gemma) return norm(L.g_pre_ffn, T_PRE_FFN);
  if (t == "post_feedforward_layernorm.weight" && c.gemma) return norm(L.g_post_ffn, T_POST_FFN);
  if (t == "self_attn.q_proj.weight") return rows(L.wqkv, qg, H, T_Q);
  if (t == "self_attn.k_proj.weight") return rows(L.wqkv + static_cast<size_t>(e->q_cols) * H, kvg, H, T_K);
  if (t == "self_attn.v_proj.weight") return rows(L.wqkv + static_cast<size_t>(e->q_cols + e->kv_cols) * H, kvg, H, T_V);
  if (t == "self_attn.o_proj.weight") return cols(L.wo, H, qg, T_O);
  if (t == "mlp.gate_proj.weight") return glu(0, T_GATE);
  if (t == "mlp.up_proj.weight") return glu(1, T_UP);
  if (t == "mlp.down_proj.weight") return cols(L.wd, H, I, T_DOWN);
  return fail(B2_ERR_ARG, "unknown tensor name %s", name);
}

extern "C" int b2_finalize(b2_engine* e) {
  if (!e) return fail(B2_ERR_ARG, "null engine");
  if (!e->embed_loaded) return fail(B2_ERR_STATE, "missing tensor model.embed_tokens.weight");
  if (!e->final_loaded) return fail(B2_ERR_STATE, "missing tensor model.norm.weight");
  unsigned need = T_IN | T_Q | T_K | T_V | T_O | T_POST | T_GATE | T_UP | T_DOWN;
  if (e->cfg.gemma) need |= T_PRE_FFN | T_POST_FFN;
  for (size_t i = 0; i < e->layers.size(); ++i)
    if ((e->layers[i].loaded & need) != need)
      return fail(B2_ERR_STATE, "layer %zu is missing tensors (have mask 0x%x, need 0x%x)", i, e->layers[i].loaded, need);
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());       // pageable H2D copies may still be in flight on the default stream
  e->finalized = true;
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// cache
// ------------------------------------------------------------------------------------------
static int cache_create_impl(b2_engine* e, int max_batch, int max_tokens, int page_size, int n_pages, b2_cache** out) {
  if (!e || !out || max_batch <= 0 || max_tokens <= 0) return fail(B2_ERR_ARG, "bad cache arguments");
  if (max_tokens > e->cfg.max_positions)
    return fail(B2_ERR_ARG, "cache max_tokens %d exceeds the engine's RoPE table (%d positions)", max_tokens, e->cfg.max_positions);
  const bool paged = page_size > 0;
  if (paged && (page_size % 128 || page_size > 4096)) return fail(B2_ERR_ARG, "page_size %d must be a multiple of 128 (<= 4096)", page_size);
  CK(cudaSetDevice(e->device));
  static std::atomic<uint64_t> next_uid{1};
  b2_cache* c = new b2_cache();
  c->e = e;
  c->uid = next_uid.fetch_add(1);
  c->max_batch = max_batch;
  c->max_tokens = max_tokens;
  c->len = 0;
  c->e_head_dim = e->cfg.head_dim;
  c->paged = paged;
  if (paged) {
    c->page = page_size;
    c->max_lp = (max_tokens + page_size - 1) / page_size;
    const long long full = static_cast<long long>(max_batch) * c->max_lp;
    c->n_pages = n_pages > 0 ? n_pages : static_cast<int>(full);
    if (c->n_pages < max_batch) c->n_pages = max_batch;               // at least one page per row
    if (c->n_pages > full) c->n_pages = static_cast<int>(full);
    c->plane = static_cast<size_t>(c->n_pages) * e->nkv * page_size * e->cfg.head_dim;
    c->table_host.assign(static_cast<size_t>(max_batch) * c->max_lp, 0);
  } else {
    c->plane = static_cast<size_t>(max_batch) * e->nkv * max_tokens * e->cfg.head_dim;
  }
  c->pf = e->cfg.precise ? 2 : 1;
  const size_t bytes = c->plane * c->pf * 2 * e->cfg.num_hidden_layers * 2;
  if (c->pf == 2 && c->plane > 0x7fffffffull) { delete c; return fail(B2_ERR_ARG, "precise-mode cache plane too large"); }
  if (c->plane / e->cfg.head_dim > 0x7fffffffull) { delete c; return fail(B2_ERR_ARG, "cache plane has more rows than a tensor-map coordinate covers"); }
  auto alloc = [&]() -> int {
    CK(cudaMalloc(&c->kv, bytes));
    CK(cudaMemset(c->kv, 0, bytes));   // rows past the valid length feed P.V with weight 0: must stay finite
    CK(cudaMalloc(&c->t0_dev, 4));
    CK(cudaMemset(c->t0_dev, 0, 4));
    if (paged) {
      CK(cudaMalloc(&c->table_dev, c->table_host.size() * 4));
      CK(cudaMemset(c->table_dev, 0, c->table_host.size() * 4));
    }
    if (!e->cfg.precise) {
      c->dec_splits = decode_splits(max_batch, e->nkv, max_tokens);
      CK(cudaMalloc(&c->dec_part, decode_part_floats(max_batch, e->nh, e->cfg.head_dim) * 4));
      CK(cudaMalloc(&c->dec_cnt, static_cast<size_t>(max_batch) * e->nkv * 4));
      CK(cudaMemset(c->dec_cnt, 0, static_cast<size_t>(max_batch) * e->nkv * 4));
    }
    CK(cudaDeviceSynchronize());       // cudaMemset is asynchronous; callers may use non-blocking streams
    return B2_OK;
  };
  const int rc = alloc();
  if (rc != B2_OK) {                   // e.g. out of memory: nothing is left behind
    b2_cache_destroy(c);
    return rc;
  }
  e->caches.push_back(c);
  *out = c;
  return B2_OK;
}
extern "C" int b2_cache_create(b2_engine* e, int max_batch, int max_tokens, b2_cache** out) {
  return cache_create_impl(e, max_batch, max_tokens, 0, 0, out);
}
extern "C" int b2_cache_create_paged(b2_engine* e, int max_batch, int max_tokens, int page_size, int n_pages, b2_cache** out) {
  if (page_size <= 0) return fail(B2_ERR_ARG, "page_size must be positive");
  return cache_create_impl(e, max_batch, max_tokens, page_size, n_pages, out);
}
extern "C" int b2_cache_set_block_table(b2_cache* c, const int32_t* table_host, int n_entries) {
  if (!c || !table_host) return fail(B2_ERR_ARG, "null argument");
  if (!c->paged) return fail(B2_ERR_ARG, "not a paged cache");
  if (!c->e) return fail(B2_ERR_STATE, "the cache's engine has been destroyed");
  if (c->len != 0) return fail(B2_ERR_STATE, "the block table can only be replaced while the cache is empty");
  if (n_entries != static_cast<int>(c->table_host.size()))
    return fail(B2_ERR_ARG, "block table has %d entries, expected max_batch * max_logical_pages = %zu", n_entries, c->table_host.size());
  std::vector<uint8_t> seen(c->n_pages, 0);
  for (int i = 0; i < n_entries; ++i) {
    const int pg = table_host[i];
    if (pg < 0 || pg >= c->n_pages || seen[pg]) return fail(B2_ERR_ARG, "block table entry %d = %d is out of range or used twice (pool of %d pages)", i, pg, c->n_pages);
    seen[pg] = 1;
  }
  CK(cudaSetDevice(c->e->device));
  c->table_host.assign(table_host, table_host + n_entries);
  CK(cudaMemcpy(c->table_dev, c->table_host.data(), c->table_host.size() * 4, cudaMemcpyHostToDevice));
  c->user_table = true;
  c->assigned_lp = c->max_lp;
  return B2_OK;
}
extern "C" int b2_cache_info(const b2_cache* c, int32_t* page_size, int32_t* n_pages, int32_t* pages_in_use) {
  if (!c) return fail(B2_ERR_ARG, "null cache");
  if (page_size) *page_size = c->paged ? c->page : 0;
  if (n_pages) *n_pages = c->paged ? c->n_pages : 0;
  if (pages_in_use) *pages_in_use = c->paged ? (c->user_table ? ((c->len + c->page - 1) / c->page) * c->max_batch : c->next_free) : 0;
  return B2_OK;
}

// Paged cache: make sure logical pages [0, ceil(tokens / page)) of every row have a physical page.  Pages are handed out
// row-interleaved (logical page lp of row b -> physical page lp * max_batch + b), so consecutive logical pages of a
// sequence are max_batch pages apart in the pool -- the block-table indirection is always exercised.  A pool that runs
// out grows by doubling (new allocation + plane-wise copy; the captured decode graphs of this cache are dropped).
static int cache_ensure_pages(b2_cache* c, int tokens, cudaStream_t st) {
  if (!c->paged) return B2_OK;
  const int need_lp = (tokens + c->page - 1) / c->page;
  if (need_lp <= c->assigned_lp) return B2_OK;
  if (c->user_table) return fail(B2_ERR_STATE, "internal: user block table covers every logical page");
  b2_engine* e = c->e;
  const long long need_pages = static_cast<long long>(need_lp) * c->max_batch;
  if (need_pages > c->n_pages) {
    long long nn = c->n_pages;
    while (nn < need_pages) nn *= 2;
    const long long full = static_cast<long long>(c->max_batch) * c->max_lp;
    if (nn > full) nn = full;
    const size_t new_plane = static_cast<size_t>(nn) * e->nkv * c->page * e->cfg.head_dim;
    if (c->pf == 2 && new_plane > 0x7fffffffull) return fail(B2_ERR_ARG, "precise-mode paged cache cannot grow to %lld pages", nn);
    if (new_plane / e->cfg.head_dim > 0x7fffffffull) return fail(B2_ERR_ARG, "paged cache cannot grow to %lld pages (tensor-map row range)", nn);
    const int nplanes = e->cfg.num_hidden_layers * 2 * c->pf;
    bf16* nkv = nullptr;
    CK(cudaStreamSynchronize(st));
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&nkv, new_plane * nplanes * 2));
    CK(cudaMemset(nkv, 0, new_plane * nplanes * 2));
    for (int i = 0; i < nplanes; ++i)
      CK(cudaMemcpy(nkv + static_cast<size_t>(i) * new_plane, c->kv + static_cast<size_t>(i) * c->plane, c->plane * 2, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    cudaFree(c->kv);
    c->kv = nkv;
    c->plane = new_plane;
    c->n_pages = static_cast<int>(nn);
    for (auto it = e->graphs.begin(); it != e->graphs.end();) { auto cur = it++; if (cur->first.cache_uid == c->uid) drop_graph(e, cur); }
  }
  for (int lp = c->assigned_lp; lp < need_lp; ++lp)
    for (int b = 0; b < c->max_batch; ++b) c->table_host[static_cast<size_t>(b) * c->max_lp + lp] = c->next_free++;
  c->assigned_lp = need_lp;
  CK(cudaMemcpyAsync(c->table_dev, c->table_host.data(), c->table_host.size() * 4, cudaMemcpyHostToDevice, st));
  return B2_OK;
}
extern "C" int b2_cache_len(const b2_cache* c) { return c ? c->len : fail(B2_ERR_ARG, "null cache"); }
extern "C" int b2_cache_reset(b2_cache* c) {
  if (!c) return fail(B2_ERR_ARG, "null cache");
  if (!c->e) return fail(B2_ERR_STATE, "the cache's engine has been destroyed");
  CK(cudaSetDevice(c->e->device));
  CK(cudaStreamSynchronize(c->e->stream));
  CK(cudaDeviceSynchronize());
  c->len = 0;
  if (c->paged && !c->user_table) c->assigned_lp = c->next_free = 0;   // every page goes back to the pool
  CK(cudaMemset(c->t0_dev, 0, 4));
  CK(cudaDeviceSynchronize());
  return B2_OK;
}
static bf16* cache_k(const b2_cache* c, int layer) { return c->kv + (static_cast<size_t>(layer) * 2) * c->pf * c->plane; }
static bf16* cache_v(const b2_cache* c, int layer) { return c->kv + (static_cast<size_t>(layer) * 2 + 1) * c->pf * c->plane; }

extern "C" int b2_cache_read(const b2_cache* c, int layer, int which, int batch, void* dev_out) {
  if (c && !c->e) return fail(B2_ERR_STATE, "the cache's engine has been destroyed");
  if (!c || !dev_out || layer < 0 || layer >= c->e->cfg.num_hidden_layers || batch <= 0 || batch > c->max_batch)
    return fail(B2_ERR_ARG, "bad cache_read arguments");
  const b2_config& g = c->e->cfg;
  CK(cudaSetDevice(c->e->device));
  if (c->len == 0) return B2_OK;                 // nothing cached: a zero-width copy is an error in CUDA
  CK(cudaDeviceSynchronize());                   // forwards may still be running on non-blocking streams
  const bf16* src = which ? cache_v(c, layer) : cache_k(c, layer);
  if (c->paged) {   // gather the pages of every row: one 2-D copy (NKV slabs) per (row, logical page)
    const size_t hd = g.head_dim;
    for (int b = 0; b < batch; ++b)
      for (int lp = 0; lp * c->page < c->len; ++lp) {
        const int rows = (c->len - lp * c->page) < c->page ? (c->len - lp * c->page) : c->page;
        const int phys = c->table_host[static_cast<size_t>(b) * c->max_lp + lp];
        CK(cudaMemcpy2DAsync(static_cast<bf16*>(dev_out) + (static_cast<size_t>(b) * c->e->nkv * c->len + static_cast<size_t>(lp) * c->page) * hd,
                             static_cast<size_t>(c->len) * hd * 2, src + static_cast<size_t>(phys) * c->e->nkv * c->page * hd,
                             static_cast<size_t>(c->page) * hd * 2, static_cast<size_t>(rows) * hd * 2, c->e->nkv, cudaMemcpyDeviceToDevice, 0));
      }
    CK(cudaDeviceSynchronize());
    return B2_OK;
  }
  // [batch*NKV] slabs of [len, HD] out of [max_tokens, HD]
  CK(cudaMemcpy2DAsync(dev_out, static_cast<size_t>(c->len) * g.head_dim * 2, src,
                       static_cast<size_t>(c->max_tokens) * g.head_dim * 2, static_cast<size_t>(c->len) * g.head_dim * 2,
                       static_cast<size_t>(batch) * c->e->nkv, cudaMemcpyDeviceToDevice, 0));
  CK(cudaDeviceSynchronize());
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
static int ensure_ws(b2_engine* e, int M, int B) {
  if (M <= e->ws_tokens && B <= e->ws_batch) return B2_OK;
  CK(cudaStreamSynchronize(e->stream));
  free_ws(e);
  const b2_config& c = e->cfg;
  const size_t H = c.hidden_size, I = e->I;
  const size_t Mp = M > e->ws_tokens ? M : e->ws_tokens, Bp = B > e->ws_batch ? B : e->ws_batch;
  CK(cudaMalloc(&e->x, Mp * H * 4));
  if (c.gemma || e->tp_world > 1) CK(cudaMalloc(&e->y, Mp * H * 4));   // tp: partial sums of the row-parallel GEMMs
  const size_t pf = c.precise ? 2 : 1;   // precise mode: every bf16 activation row is [hi | lo]
  CK(cudaMalloc(&e->xn, Mp * H * 2 * pf));
  CK(cudaMalloc(&e->q, Mp * e->q_cols * 2 * pf));
  CK(cudaMalloc(&e->att, Mp * e->q_cols * 2 * pf));
  CK(cudaMalloc(&e->hb, Mp * I * 2 * pf));
  CK(cudaMalloc(&e->ids_ws, Mp * 8));
  CK(cudaMalloc(&e->xl, Bp * H * 4));
  CK(cudaMalloc(&e->part, static_cast<size_t>(MAX_SPLIT) * 128 * H * 4));
  CK(cudaMalloc(&e->xnl, Bp * H * 2 * pf));
  CK(cudaMalloc(&e->ss, static_cast<size_t>(SK_MAX_ROWS) * (H / 32) * 4));
  e->ws_tokens = static_cast<int>(Mp);
  e->ws_batch = static_cast<int>(Bp);
  return B2_OK;
}

// nsplit > 0: first fold the pending split-K partials of the previous o_proj / down_proj into x
// Partial sums still to be folded into the residual stream x by the next norm kernel: `n` planes of [rows, H] fp32,
// `stride` floats apart, starting at `src` -- split-K partials of a skinny o_proj / down_proj (summed in plane order:
// deterministic), or the single all-reduced partial of a row-parallel GEMM under tensor parallelism.
struct Fold {
  const float* src = nullptr;
  int n = 0;
  size_t stride = 0;
  int tp_site = -1;      // >= 0: the planes live in the peer-mapped buffer of that site on EVERY rank: the norm kernel all-reduces
};

static int rmsnorm(b2_engine* e, cudaStream_t st, float* x, const float* g, bf16* ob, float* of, int M, Fold f = Fold()) {
  const int H = e->cfg.hidden_size, pf = e->cfg.precise ? 2 : 1;
  const int ld = pf * H, lo = e->cfg.precise ? H : 0;   // precise: [hi | lo] rows
  e->st.launches++;
  if (f.tp_site >= 0) {   // tensor-parallel decode: all-reduce over peer memory + residual + norm in one kernel
    if (M > 128 || lo) return fail(B2_ERR_STATE, "internal: peer-memory all-reduce is a few-row, non-precise path");
    TpPeers pr;
    memset(&pr, 0, sizeof(pr));
    pr.rank = e->tp_rank;
    pr.world = e->tp_world;
    const size_t flag_off = 2 * e->p2p_site_floats;
    for (int p = 0; p < e->tp_world; ++p) {
      float* base = p == e->tp_rank ? e->p2p_buf : static_cast<float*>(e->p2p_peer[p]);
      pr.part[p] = base + static_cast<size_t>(f.tp_site) * e->p2p_site_floats;
      pr.flags[p] = reinterpret_cast<unsigned*>(base + flag_off);
    }
    CK(launch_k(tp_allreduce_norm_row_kernel, dim3(M), dim3(256), 0, st, x, g, ob, ld, of, H, e->cfg.rms_norm_eps, pr, f.n, f.stride,
                e->p2p_ctr, e->p2p_ctr + 2, f.tp_site));
    e->p2p_fused++;
    return B2_OK;
  }
  if (M <= 128) {   // decode / tiny prompts: a CTA per row keeps every load of the row in flight
    CK(launch_k(rmsnorm_row_kernel, dim3(M), dim3(256), 0, st, x, g, ob, ld, lo, of, H, e->cfg.rms_norm_eps,
                f.src, f.n, f.stride));
    return B2_OK;
  }
  if (f.n > 1) return fail(B2_ERR_STATE, "split-K partials are only folded on the few-row path");
  CK(launch_k(rmsnorm_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x, f.n ? f.src : static_cast<const float*>(nullptr), g, ob, ld, lo,
              of, M, H, e->cfg.rms_norm_eps));
  return B2_OK;
}

// sum over the tensor-parallel ranks, in place (NCCL over NVLink / NVSwitch)
static int tp_allreduce(b2_engine* e, cudaStream_t st, float* buf, size_t count) {
  if (e->tp_world <= 1) return B2_OK;
  if (!e->comm) return fail(B2_ERR_STATE, "tensor-parallel engine used before b2_tp_init");
  time_begin(e, st, 0.0, 2);
  ncclResult_t r = nccl_api()->AllReduce(buf, buf, count, ncclFloat, ncclSum, e->comm, st);
  time_end(e, st);
  if (r != ncclSuccess) return fail(B2_ERR_CUDA, "ncclAllReduce failed: %s", nccl_api()->GetErrorString(r));
  e->comm_calls++;
  e->comm_bytes += static_cast<double>(count) * 4;
  e->st.launches++;
  return B2_OK;
}

static GemmParams base_params(int M, int N, int K) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.k_wrap = K;
  p.band_m = 32;
  p.split_k = 1;
  return p;
}

// split-K factor for a skinny GEMM (M <= 128): aim at >= ~128 CTAs with >= 8 k-blocks each
static int pick_split(int N, int K, int bn) {
  const int tiles = (N + bn - 1) / bn, kb = K / GEMM_BK;
  int best = 1;
  for (int sp = 2; sp <= MAX_SPLIT; sp *= 2)
    if (kb % sp == 0 && kb / sp >= 8 && tiles * sp <= num_sms()) best = sp;   // one wave: a second, partial wave costs more than it buys
  return best;
}

// tile width for a GEMM with M rows: wide tiles for prefill, narrow ones so skinny decode GEMMs spread over SMs
static int pick_bn(int M, int N, int min_bn) {
  int bn = 256;
  if (M <= 128) {
    bn = 64;
    if (bn < min_bn) bn = min_bn;
  }
  (void)N;
  return bn;
}

struct FwdArgs {
  const int64_t* ids = nullptr;     // [B,S] token ids, or
  const float* embeds = nullptr;    // [B,S,H] fp32 inputs_embeds (llama3.2_model_numpy.py:645-676)
  const int* pos_ids = nullptr;     // [S] RoPE positions (:657); nullptr = cache length + s
  int B = 0, S = 0, logits_mode = 0;
  float* logits = nullptr;
  float* hidden = nullptr;          // [L+1,B,S,H]
  float* attn_probs = nullptr;      // [L,B,NH,S,T] opt-in materialisation of the reference's `attentions` (:713-714)
};

// o_proj / down_proj (Linear_np + the residual add, llama3.2_model_numpy.py:514 + :569, :176-182 + :578): A [M,K] . W[H,K]^T.
// Four shapes of the same thing; what is left for the following norm kernel to fold into x comes back in `f`:
//   prefill, 1 GPU (Llama)   x += A.W^T by the GEMM's own TMA reduce-add epilogue                       f = {}
//   few rows (decode)        split-K partial planes in e->part (all SMs stream weights)                  f = {part, ns, stride}
//   tensor parallel          this rank's partial sum -> ncclAllReduce over the ranks                     f = {y or part, ...}
//   Gemma-2                  y is needed on its own (post-norm before the residual add), never EPI_RESID f = {y or part, ...}
static int row_parallel_proj(b2_engine* e, cudaStream_t st, const bf16* a, int K, const bf16* w, int M, bool skinny, Fold& f, int site) {
  const b2_config& c = e->cfg;
  const int H = c.hidden_size;
  const bool tp = e->tp_world > 1;
  GemmParams p = base_params(M, H, K);
  f = Fold();
  if (skinny) {
    const int sr = tp ? ((M + 31) / 32) * 32 : 128;      // tp: planes packed back to back so one all-reduce covers them
    const bool p2p = tp && e->p2p_ready && !c.gemma;     // Llama decode: the following norm kernel all-reduces over peer memory
    float* planes = p2p ? e->p2p_buf + static_cast<size_t>(site) * e->p2p_site_floats : e->part;
    p.out = planes; p.ldo = H; p.split_rows = sr;
    p.split_k = pick_split(H, K, 64);
    CKR(eng_gemm(e, st, 64, EPI_F32, a, K, w, p, 0, MAX_SPLIT * 128));
    f.src = planes; f.n = p.split_k; f.stride = static_cast<size_t>(sr) * H;
    if (p2p) f.tp_site = site;
    else if (tp) CKR(tp_allreduce(e, st, e->part, static_cast<size_t>(p.split_k) * sr * H));
    return B2_OK;
  }
  if (!c.gemma && !tp) {
    p.out = e->x; p.ldo = H;
    return eng_gemm(e, st, M <= 128 ? 32 : 256, EPI_RESID, a, K, w, p);
  }
  p.out = e->y; p.ldo = H;
  CKR(eng_gemm(e, st, pick_bn(M, H, 0), EPI_F32, a, K, w, p));
  if (tp) CKR(tp_allreduce(e, st, e->y, static_cast<size_t>(M) * H));
  f.src = e->y; f.n = 1; f.stride = 0;
  return B2_OK;
}

// One-token decode step (S == 1, <= 64 sequences, Llama) with the RMSNorm split across the GEMMs and the split-K
// reduction done inside thread-block clusters (gemm_sk_tcgen05.cuh): 5 kernels per layer
//   QKV(+1/rms, RoPE, append) -> attention -> o_proj(+residual, x*gamma, sum x^2) -> gate/up(+1/rms, SiLU*up) ->
//   down_proj(+residual, x*gamma', sum x^2)
// instead of 7 (two RMSNorm kernels folding split-K partial planes).  Same arithmetic as the reference's layer
// (llama3.2_model_numpy.py:554-578); the norm's 1/sqrt(mean(x^2)+eps) is applied to the fp32 accumulators of the
// consuming GEMM instead of to its bf16 operand.  B2_NO_FUSED_DECODE=1 selects the 7-kernel path.
static bool fused_decode_ok(const b2_engine* e, const FwdArgs& fa) {
  static const bool off = getenv("B2_NO_FUSED_DECODE") != nullptr;
  const b2_config& c = e->cfg;
  return !off && fa.S == 1 && fa.B <= SK_MAX_ROWS && fa.ids && !fa.hidden && !fa.attn_probs && !c.gemma && !c.precise &&
         e->tp_world == 1 && c.hidden_size % 64 == 0 && e->q_cols % 64 == 0 && e->I % 64 == 0 &&
         fa.logits_mode != B2_LOGITS_NONE;
}

static int run_decode_fused(b2_engine* e, b2_cache* cache, cudaStream_t st, const FwdArgs& fa) {
  const b2_config& c = e->cfg;
  const int M = fa.B, H = c.hidden_size, I = e->I, L = c.num_hidden_layers, HD = c.head_dim;
  const int ss_n = H / 32;
  auto consumer = [&](GemmParams& p) {   // this GEMM's A operand is x*gamma: scale its rows by 1/rms(x)
    p.ss_in = e->ss; p.ss_n = ss_n; p.norm_dim = H; p.norm_eps = c.rms_norm_eps;
  };
  CK(launch_k(embed_norm_row_kernel, dim3(M), dim3(256), 0, st, reinterpret_cast<const long long*>(fa.ids),
              static_cast<const bf16*>(e->embed), e->x, e->xn, static_cast<const float*>(e->layers[0].g_in), e->ss, ss_n, H,
              c.vocab_size, 1.0f));
  e->st.launches++;
  const float scale = c.attn_scale > 0.f ? c.attn_scale : 1.0f / sqrtf(static_cast<float>(HD));
  for (int li = 0; li < L; ++li) {
    Layer& Ly = e->layers[li];
    {
      GemmParams p = base_params(M, e->qkv_n, H);
      p.q_cols = e->q_cols; p.kv_cols = e->kv_cols; p.head_dim = HD; p.S = 1; p.Tmax = cache->max_tokens; p.nkv = e->nkv;
      p.q_out = e->q; p.ldq = e->q_cols;
      p.k_cache = cache_k(cache, li); p.v_cache = cache_v(cache, li);
      p.t0_ptr = cache->t0_dev; p.pages = cache->pages(); p.pos_ids = fa.pos_ids; p.rope_cos = e->rope_cos; p.rope_sin = e->rope_sin;
      consumer(p);
      if (HD == SK_BN) CKR(eng_gemm_sk(e, st, EPI_QKV_ROPE, e->xn, H, Ly.wqkv, p));
      else CKR(eng_gemm(e, st, HD, EPI_QKV_ROPE, e->xn, H, Ly.wqkv, p));
    }
    const int window = (c.sliding_window > 0 && e->sliding[li]) ? c.sliding_window : 0;
    CKR(attn_decode(e->q, e->q_cols, cache_k(cache, li), cache_v(cache, li), e->att, e->q_cols, M, e->nh, e->nkv, HD,
                    cache->max_tokens, cache->t0_dev, scale, c.attn_logit_softcap, window, st, 0, 0, 0, cache->dec_part,
                    cache->dec_cnt, cache->dec_splits, cache->pages(), cache->plane_rows()));
    e->st.launches++;
    {
      GemmParams p = base_params(M, H, e->q_cols);          // x += attn . Wo^T; xn = bf16(x * g_post); ss
      p.out = e->x; p.ldo = H; p.a_out = e->xn; p.gamma_next = Ly.g_post_attn; p.ss_out = e->ss;
      CKR(eng_gemm_sk(e, st, EPI_RESID_NORM, e->att, e->q_cols, Ly.wo, p));
    }
    {
      GemmParams p = base_params(M, 2 * I, H);
      p.out = e->hb; p.ldo = I; p.act = c.hidden_act;
      consumer(p);
      const int bn_glu = ((2 * I) / 128 <= e->n_sms) ? 128 : 64;
      CKR(eng_gemm(e, st, bn_glu, EPI_GLU, e->xn, H, Ly.wgu, p));
    }
    {
      GemmParams p = base_params(M, H, I);                  // x += h . Wd^T; xn = bf16(x * g_in[next] | g_final); ss
      p.out = e->x; p.ldo = H; p.a_out = e->xn; p.ss_out = e->ss;
      p.gamma_next = (li + 1 < L) ? e->layers[li + 1].g_in : e->g_final;
      CKR(eng_gemm_sk(e, st, EPI_RESID_NORM, e->hb, I, Ly.wd, p));
    }
  }
  {   // lm_head on the last down_proj's x * g_final (tied embedding, :1025-1026); S == 1: ALL and LAST are the same rows
    GemmParams p = base_params(M, e->V, H);
    p.out = fa.logits; p.ldo = e->V; p.softcap = c.final_logit_softcap;
    consumer(p);
    CKR(eng_gemm(e, st, 128, EPI_F32, e->xn, H, e->embed, p, e->ws_tokens));
  }
  CK(launch_k(add_scalar_kernel, dim3(1), dim3(32), 0, st, cache->t0_dev, 1));
  e->st.launches++;
  return B2_OK;
}

static int run_layers(b2_engine* e, b2_cache* cache, cudaStream_t st, const FwdArgs& fa) {
  if (fused_decode_ok(e, fa)) return run_decode_fused(e, cache, st, fa);
  const b2_config& c = e->cfg;
  const int B = fa.B, S = fa.S, logits_mode = fa.logits_mode;
  float* const logits = fa.logits;
  float* const hidden = fa.hidden;
  const int M = B * S, H = c.hidden_size, I = e->I, L = c.num_hidden_layers, HD = c.head_dim;
  const bool decode = (S == 1);
  const bool tp = e->tp_world > 1;
  const int pf = c.precise ? 2 : 1;            // bf16 activation rows are [hi | lo] in precise mode
  const bool skinny = M <= 128 && !hidden;     // few rows: split-K so the weight stream covers all SMs
  Fold pend;                                   // partial sums of the last down_proj not yet folded into x
  const float escale = c.gemma ? sqrtf(static_cast<float>(H)) : 1.0f;
  if (fa.embeds) {   // hidden_states = inputs_embeds (* sqrt(H) for Gemma-2, gemma2_model.py:738-739)
    CK(launch_k(scale_copy_kernel, dim3(2 * e->n_sms), dim3(256), 0, st, fa.embeds, e->x, static_cast<size_t>(M) * H / 4, escale));
  } else {
    CK(launch_k(embed_kernel, dim3((M + 7) / 8), dim3(256), 0, st, reinterpret_cast<const long long*>(fa.ids),
                static_cast<const bf16*>(e->embed), e->x, M, H, c.vocab_size, escale));
  }
  e->st.launches++;
  const int causal = (c.always_causal || S > 2) ? 1 : 0;   // llama3.2_model_numpy.py:479
  const float scale = c.attn_scale > 0.f ? c.attn_scale : 1.0f / sqrtf(static_cast<float>(HD));
  const int T_total = cache->len + S;
  for (int li = 0; li < L; ++li) {
    Layer& Ly = e->layers[li];
    if (hidden) {
      if (pend.n) return fail(B2_ERR_STATE, "internal: hidden-state capture with unfolded partial sums");
      CK(cudaMemcpyAsync(hidden + static_cast<size_t>(li) * M * H, e->x, static_cast<size_t>(M) * H * 4, cudaMemcpyDeviceToDevice, st));
    }
    CKR(rmsnorm(e, st, e->x, Ly.g_in, e->xn, nullptr, M, pend));
    pend = Fold();
    {  // fused QKV projection + RoPE + KV-cache append (column-parallel under tp: this rank's heads)
      GemmParams p = base_params(M, e->qkv_n, H);
      p.q_cols = e->q_cols;
      p.kv_cols = e->kv_cols;
      p.head_dim = HD;
      p.S = S;
      p.Tmax = cache->max_tokens;
      p.nkv = e->nkv;
      p.q_out = e->q;
      p.ldq = pf * e->q_cols;
      p.lo_off = c.precise ? e->q_cols : 0;
      p.kv_lo_off = c.precise ? static_cast<int>(cache->plane) : 0;
      p.k_cache = cache_k(cache, li);
      p.v_cache = cache_v(cache, li);
      p.t0_ptr = cache->t0_dev;
      p.pages = cache->pages();
      p.pos_ids = fa.pos_ids;
      p.rope_cos = e->rope_cos;
      p.rope_sin = e->rope_sin;
      static const int qkv_bn = getenv("B2_QKV_BN") ? atoi(getenv("B2_QKV_BN")) : 0;   // bring-up knob, read once
      const int bn_qkv = (qkv_bn >= HD && M > 128) ? qkv_bn : pick_bn(M, p.N, HD);
      CKR(eng_gemm(e, st, bn_qkv, EPI_QKV_ROPE, e->xn, H, Ly.wqkv, p));
    }
    // HF Gemma-2: only the layers whose layer_types entry is "sliding_attention" slide (default pattern: even layers)
    const int window = (c.sliding_window > 0 && e->sliding[li]) ? c.sliding_window : 0;
    if (fa.attn_probs) {   // debug: this layer's [B,NH,S,T] softmax probabilities, recomputed from q and the K cache
      const size_t rows = static_cast<size_t>(B) * e->nh * S;
      CK(launch_k(attn_probs_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, st, static_cast<const bf16*>(e->q),
                  pf * e->q_cols, static_cast<const bf16*>(cache_k(cache, li)), fa.attn_probs + static_cast<size_t>(li) * rows * T_total,
                  B, S, e->nh, e->nkv, HD, cache->max_tokens, static_cast<const int*>(cache->t0_dev),
                  scale, c.attn_logit_softcap, causal, window, cache->pages()));
      e->st.launches++;
    }
    if (decode) {
      CKR(attn_decode(e->q, pf * e->q_cols, cache_k(cache, li), cache_v(cache, li), e->att, pf * e->q_cols, B, e->nh,
                      e->nkv, HD, cache->max_tokens, cache->t0_dev, scale, c.attn_logit_softcap, window, st,
                      c.precise ? e->q_cols : 0, c.precise ? cache->plane : 0, c.precise ? e->q_cols : 0, cache->dec_part,
                      cache->dec_cnt, cache->dec_splits, cache->pages(), cache->plane_rows()));
      e->st.launches++;
    } else {
      AttnParams a;
      memset(&a, 0, sizeof(a));
      a.B = B; a.S = S; a.NH = e->nh; a.NKV = e->nkv; a.Tmax = cache->max_tokens;
      a.t0_ptr = cache->t0_dev; a.scale = scale; a.softcap = c.attn_logit_softcap; a.causal = causal; a.window = window;
      a.out = e->att; a.ldo = pf * e->q_cols;
      a.pages = cache->pages();
      a.kv_plane_rows = cache->plane_rows();
      if (c.precise) {
        a.q_lo_col = e->q_cols;
        a.out_lo_col = e->q_cols;
        a.kv_lo_row = cache->plane_rows();
      }
      const double T = cache->len + S;
      const double pairs = causal ? (static_cast<double>(S) * (S + 1) / 2 + static_cast<double>(S) * cache->len) : S * T;
      time_begin(e, st, 4.0 * B * e->nh * HD * pairs, 1);
      int r = attn_prefill(e->q, pf * e->q_cols, cache_k(cache, li), cache_v(cache, li), a, HD, st);
      time_end(e, st);
      CKR(r);
      e->st.launches++;
      e->st.attn_launches++;
    }
    Fold f;
    CKR(row_parallel_proj(e, st, e->att, e->q_cols, Ly.wo, M, skinny, f, 0));             // attn . Wo^T
    if (!c.gemma) {
      CKR(rmsnorm(e, st, e->x, Ly.g_post_attn, e->xn, nullptr, M, f));                     // x += ...; xn = norm(x)
    } else {   // Gemma-2: x += post_attention_norm(attn_out); xn = pre_feedforward_norm(x)  (gemma2_model.py:623-641)
      CK(launch_k(add_rmsnorm_kernel, dim3((M + 7) / 8), dim3(256), 0, st, e->x, f.src, static_cast<const float*>(Ly.g_post_attn), M, H,
                  c.rms_norm_eps, f.n, f.stride));
      e->st.launches++;
      CKR(rmsnorm(e, st, e->x, Ly.g_pre_ffn, e->xn, nullptr, M));
    }
    {
      GemmParams p = base_params(M, 2 * I, H);
      p.out = e->hb; p.ldo = pf * I; p.act = c.hidden_act;
      p.lo_off = c.precise ? I : 0;
      // few rows: 128-wide tiles give exactly one wave of weight-streaming CTAs for the usual 2I / 128 <= #SMs
      const int bn_glu = (M <= 128 && (2 * I) / 128 <= e->n_sms) ? 128 : pick_bn(M, 2 * I, 0);
      CKR(eng_gemm(e, st, bn_glu, EPI_GLU, e->xn, H, Ly.wgu, p));                          // act(gate)*up
    }
    CKR(row_parallel_proj(e, st, e->hb, I, Ly.wd, M, skinny, f, 1));                       // h . Wd^T
    if (!c.gemma) {
      pend = f;                  // folded into x by the next layer's input norm / the final norm
      if (hidden && pend.n) {    // hidden-state capture copies x at the top of the next layer: fold now
        CKR(rmsnorm(e, st, e->x, Ly.g_post_attn, e->xn, nullptr, M, pend));
        pend = Fold();
      }
    } else {
      CK(launch_k(add_rmsnorm_kernel, dim3((M + 7) / 8), dim3(256), 0, st, e->x, f.src, static_cast<const float*>(Ly.g_post_ffn), M, H,
                  c.rms_norm_eps, f.n, f.stride));
      e->st.launches++;
    }
  }
  // final norm + lm_head (tied to the embedding, llama3.2_model_numpy.py:1025-1026); vocabulary-parallel under tp
  float* hid_last = hidden ? hidden + static_cast<size_t>(L) * M * H : nullptr;
  const bf16* wlm = e->embed + static_cast<size_t>(e->tp_rank) * e->V * H;
  const int Vl = e->V;
  auto lm_head = [&](const bf16* a, int rows, int a_rows) -> int {
    GemmParams p = base_params(rows, Vl, H);
    p.softcap = c.final_logit_softcap;
    if (!tp) {
      p.out = logits; p.ldo = Vl;
      return eng_gemm(e, st, rows <= 128 ? 128 : 256, EPI_F32, a, H, wlm, p, a_rows);
    }
    const size_t loc = static_cast<size_t>(rows) * Vl;
    if (loc * (1 + e->tp_world) > e->tp_lg_elems) {
      if (e->capturing) return fail(B2_ERR_STATE, "tp logits scratch must be sized before graph capture");
      CK(cudaStreamSynchronize(st));
      cudaFree(e->tp_lg);
      e->tp_lg = nullptr;
      e->tp_lg_elems = 0;
      CK(cudaMalloc(&e->tp_lg, loc * (1 + e->tp_world) * 4));
      e->tp_lg_elems = loc * (1 + e->tp_world);
    }
    float* gathered = e->tp_lg + loc;
    p.out = e->tp_lg; p.ldo = Vl;
    CKR(eng_gemm(e, st, rows <= 128 ? 128 : 256, EPI_F32, a, H, wlm, p, a_rows));
    time_begin(e, st, 0.0, 2);
    ncclResult_t r = nccl_api()->AllGather(e->tp_lg, gathered, loc, ncclFloat, e->comm, st);
    time_end(e, st);
    if (r != ncclSuccess) return fail(B2_ERR_CUDA, "ncclAllGather failed: %s", nccl_api()->GetErrorString(r));
    e->comm_calls++;
    e->comm_bytes += static_cast<double>(loc) * 4 * e->tp_world;
    CK(launch_k(tp_unshard_logits_kernel, dim3(4 * e->n_sms), dim3(256), 0, st, static_cast<const float*>(gathered), logits, rows, Vl,
                e->tp_world));
    e->st.launches += 2;
    return B2_OK;
  };
  if (logits_mode == B2_LOGITS_ALL || hidden) {
    CKR(rmsnorm(e, st, e->x, e->g_final, e->xn, hid_last, M, pend));
    pend = Fold();
  }
  if (logits_mode == B2_LOGITS_ALL) {
    CKR(lm_head(e->xn, M, 0));
  } else if (logits_mode == B2_LOGITS_LAST) {
    const bf16* a = e->xn;
    int lda_rows = 0;
    if (S == 1) {
      if (!hidden) CKR(rmsnorm(e, st, e->x, e->g_final, e->xn, nullptr, M, pend));
    } else {
      if (pend.n) {   // fold the last down_proj partials into x before gathering rows (xn output unused here)
        CKR(rmsnorm(e, st, e->x, e->g_final, e->xn, nullptr, M, pend));
        pend = Fold();
      }
      CK(launch_k(gather_last_kernel, dim3(B), dim3(256), 0, st, static_cast<const float*>(e->x), e->xl, B, S, H));
      e->st.launches++;
      CKR(rmsnorm(e, st, e->xl, e->g_final, e->xnl, nullptr, B));
      a = e->xnl;
      lda_rows = 1;
    }
    CKR(lm_head(a, B, lda_rows ? e->ws_batch : e->ws_tokens));
  }
  CK(launch_k(add_scalar_kernel, dim3(1), dim3(32), 0, st, cache->t0_dev, S));
  e->st.launches++;
  return B2_OK;
}

extern "C" int b2_forward_ex(b2_engine* e, b2_cache* cache, const b2_forward_args* args) {
  if (!e || !args) return fail(B2_ERR_ARG, "null argument");
  const b2_forward_args& ua = *args;
  if (!ua.ids && !ua.inputs_embeds) return fail(B2_ERR_ARG, "You have to specify either input_ids or inputs_embeds");
  if (ua.ids && ua.inputs_embeds) return fail(B2_ERR_ARG, "You cannot specify both input_ids and inputs_embeds at the same time");
  if (!e->finalized) return fail(B2_ERR_STATE, "b2_finalize has not been called");
  const int B = ua.B, S = ua.S, logits_mode = ua.logits_mode;
  if (B <= 0 || S <= 0) return fail(B2_ERR_ARG, "B and S must be positive");
  if (logits_mode != B2_LOGITS_NONE && logits_mode != B2_LOGITS_ALL && logits_mode != B2_LOGITS_LAST)
    return fail(B2_ERR_ARG, "bad logits_mode %d", logits_mode);
  if (logits_mode != B2_LOGITS_NONE && !ua.logits) return fail(B2_ERR_ARG, "logits buffer is null");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(ua.stream);   // NULL = the CUDA (legacy) default stream
  const b2_config& c = e->cfg;
  b2_cache* kc = cache;
  if (!kc) {  // no cache object: attention over this call's tokens only
    if (!e->scratch || e->scratch->max_batch < B || e->scratch->max_tokens < S) {
      if (e->scratch) { CK(cudaStreamSynchronize(st)); b2_cache_destroy(e->scratch); e->scratch = nullptr; }
      if (S > c.max_positions) return fail(B2_ERR_ARG, "sequence length %d exceeds max_positions %d", S, c.max_positions);
      CKR(b2_cache_create(e, B, S, &e->scratch));
    }
    kc = e->scratch;
    kc->len = 0;
    CK(cudaMemsetAsync(kc->t0_dev, 0, 4, st));
  } else {
    if (kc->e != e) return fail(B2_ERR_ARG, "cache belongs to a different engine");
    if (B > kc->max_batch) return fail(B2_ERR_ARG, "batch %d exceeds the cache's max_batch %d", B, kc->max_batch);
    if (kc->len + S > kc->max_tokens)
      return fail(B2_ERR_ARG, "cache overflow: %d cached + %d new > max_tokens %d", kc->len, S, kc->max_tokens);
    if (kc->len > 0 && S > 1 && !c.always_causal)
      // the reference's [q,q] mask cannot broadcast onto [q,kv] (llama3.2_model_numpy.py:481-486);
      // a 2-token chunk is unmasked there (q_len <= 2) and is allowed.
      if (S > 2)
        return fail(B2_ERR_ARG, "operands could not be broadcast together: a %d-token call on a non-empty cache (%d items) "
                                "is not supported by the reference path (set always_causal for chunked prefill)", S, kc->len);
  }
  CKR(ensure_ws(e, B * S, B));
  CKR(cache_ensure_pages(kc, kc->len + S, st));        // paged cache: physical pages for the positions this call appends
  FwdArgs fa;
  fa.ids = ua.ids; fa.embeds = ua.inputs_embeds; fa.pos_ids = ua.position_ids;
  fa.B = B; fa.S = S; fa.logits_mode = logits_mode; fa.logits = ua.logits; fa.hidden = ua.hidden; fa.attn_probs = ua.attn_probs;
  // per-launch CUDA events on every st.every-th forward only: the other forwards take the plain launch path (events
  // between launches defeat the programmatic-dependent-launch overlap)
  e->st.sample = e->st.timing && (e->st.forwards % e->st.every == 0);
  e->st.forwards++;
  if (e->st.sample) e->st.timed_forwards++;

  const bool use_graph = e->decode_graph && S == 1 && cache && ua.ids && !ua.position_ids && !ua.hidden && !ua.attn_probs &&
                         !e->st.sample;
  if (use_graph) {
    GraphKey key{ua.ids, ua.logits, cache->uid, B, logits_mode};
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
      // warm-up launch outside capture creates every descriptor / function attribute, then restore t0
      CKR(run_layers(e, kc, st, fa));
      CK(launch_k(add_scalar_kernel, dim3(1), dim3(32), 0, st, kc->t0_dev, -S));
      CK(cudaStreamSynchronize(st));
      // the legacy default stream cannot be captured: record on the engine's own stream, replay on `st`
      cudaStream_t cs = e->stream;
      cudaGraph_t g;
      const int64_t launches_before = e->st.launches;
      e->capturing = true;
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      int r = run_layers(e, kc, cs, fa);
      cudaError_t ce = cudaStreamEndCapture(cs, &g);
      e->capturing = false;
      CKR(r);
      CK(ce);
      cudaGraphExec_t ge;
      CK(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      const int64_t kernels = e->st.launches - launches_before;
      e->st.launches = launches_before;
      if (e->graphs.size() >= MAX_GRAPHS) {   // evict the least recently used graph
        CK(cudaStreamSynchronize(st));
        drop_graph(e, e->graphs.find(e->graph_lru.back()));
      }
      e->graph_lru.push_front(key);
      it = e->graphs.emplace(key, GraphEntry{ge, kernels, e->graph_lru.begin()}).first;
    } else if (it->second.lru != e->graph_lru.begin()) {
      e->graph_lru.splice(e->graph_lru.begin(), e->graph_lru, it->second.lru);
    }
    CK(cudaGraphLaunch(it->second.exec, st));
    e->st.launches += it->second.kernels;
  } else {
    CKR(run_layers(e, kc, st, fa));
  }
  if (cache) cache->len += S;
  return B2_OK;
}

extern "C" int b2_forward(b2_engine* e, b2_cache* cache, const int64_t* ids, int B, int S, int logits_mode,
                          float* logits, float* hidden, void* stream) {
  if (!ids) return fail(B2_ERR_ARG, "null argument");
  b2_forward_args a;
  memset(&a, 0, sizeof(a));
  a.ids = ids; a.B = B; a.S = S; a.logits_mode = logits_mode; a.logits = logits; a.hidden = hidden; a.stream = stream;
  return b2_forward_ex(e, cache, &a);
}

// ------------------------------------------------------------------------------------------
// tensor parallelism (optional): NCCL communicator over the tp_world engines
// ------------------------------------------------------------------------------------------
extern "C" int b2_tp_unique_id(void* id_out) {
  if (!id_out) return fail(B2_ERR_ARG, "null argument");
  const NcclApi* n = nccl_api();
  if (!n) return fail(B2_ERR_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
  static_assert(sizeof(ncclUniqueId) == B2_NCCL_ID_BYTES, "ncclUniqueId size");
  CKN(n->GetUniqueId(static_cast<ncclUniqueId*>(id_out)));
  return B2_OK;
}
extern "C" int b2_tp_init(b2_engine* e, const void* id) {
  if (!e || !id) return fail(B2_ERR_ARG, "null argument");
  if (e->tp_world <= 1) return fail(B2_ERR_ARG, "engine was created with tp_world <= 1");
  if (e->comm) return fail(B2_ERR_STATE, "b2_tp_init called twice");
  const NcclApi* n = nccl_api();
  if (!n) return fail(B2_ERR_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
  CK(cudaSetDevice(e->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  CKN(n->CommInitRank(&e->comm, e->tp_world, uid, e->tp_rank));
  return B2_OK;
}

// Peer-memory path of the tensor-parallel decode step: every rank exports one buffer (b2_tp_ipc_handle), the caller
// all-gathers the 64-byte handles, every rank maps its peers' buffers (b2_tp_set_peers).  From then on the all-reduces of
// the few-row (decode) Llama path run inside tp_allreduce_norm_row_kernel over NVLink instead of through NCCL.
extern "C" int b2_tp_ipc_handle(b2_engine* e, void* handle_out) {
  if (!e || !handle_out) return fail(B2_ERR_ARG, "null argument");
  if (e->tp_world <= 1) return fail(B2_ERR_ARG, "engine was created with tp_world <= 1");
  static_assert(sizeof(cudaIpcMemHandle_t) == B2_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  CK(cudaSetDevice(e->device));
  if (!e->p2p_buf) {
    e->p2p_site_floats = static_cast<size_t>(MAX_SPLIT) * 128 * e->cfg.hidden_size;
    const size_t bytes = 2 * e->p2p_site_floats * 4 + 4096;                 // planes of both sites + flags
    CK(cudaMalloc(&e->p2p_buf, bytes));
    CK(cudaMemset(e->p2p_buf, 0, bytes));
    CK(cudaMalloc(&e->p2p_ctr, 16));
    CK(cudaMemset(e->p2p_ctr, 0, 16));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), e->p2p_buf));
  return B2_OK;
}
extern "C" int b2_tp_set_peers(b2_engine* e, const void* handles, int n_handles) {
  if (!e || !handles) return fail(B2_ERR_ARG, "null argument");
  if (n_handles != e->tp_world || e->tp_world > TP_MAX_WORLD) return fail(B2_ERR_ARG, "expected %d handles (tp_world <= %d)", e->tp_world, TP_MAX_WORLD);
  if (!e->p2p_buf) return fail(B2_ERR_STATE, "b2_tp_ipc_handle must be called first");
  CK(cudaSetDevice(e->device));
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int p = 0; p < e->tp_world; ++p) {
    if (p == e->tp_rank || e->p2p_peer[p]) continue;
    CK(cudaIpcOpenMemHandle(&e->p2p_peer[p], h[p], cudaIpcMemLazyEnablePeerAccess));
  }
  e->p2p_ready = true;
  for (auto it = e->graphs.begin(); it != e->graphs.end();) { auto cur = it++; drop_graph(e, cur); }
  return B2_OK;
}

extern "C" int b2_set_layer_types(b2_engine* e, const int32_t* is_sliding, int n) {
  if (!e || !is_sliding) return fail(B2_ERR_ARG, "null argument");
  if (n != e->cfg.num_hidden_layers) return fail(B2_ERR_ARG, "layer_types has %d entries, the model has %d layers", n, e->cfg.num_hidden_layers);
  for (int i = 0; i < n; ++i) e->sliding[i] = is_sliding[i] ? 1 : 0;
  for (auto it = e->graphs.begin(); it != e->graphs.end();) { auto cur = it++; drop_graph(e, cur); }   // windows are baked in
  return B2_OK;
}

extern "C" int b2_forward_host(b2_engine* e, b2_cache* cache, const int64_t* ids_host, int B, int S, int logits_mode,
                               float* logits_host, float* hidden_host) {
  if (!e || !ids_host) return fail(B2_ERR_ARG, "null argument");
  if (B <= 0 || S <= 0) return fail(B2_ERR_ARG, "B and S must be positive");
  CK(cudaSetDevice(e->device));
  CKR(ensure_ws(e, B * S, B));
  const b2_config& c = e->cfg;
  const size_t M = static_cast<size_t>(B) * S, V = c.vocab_size;
  const bool all_dev = logits_mode == B2_LOGITS_ALL || logits_mode == B2_LOGITS_ALL_LAST_TO_HOST;
  const size_t need = all_dev ? M * V : (logits_mode == B2_LOGITS_LAST ? static_cast<size_t>(B) * V : 0);
  if (need > e->logits_ws_elems) {
    CK(cudaDeviceSynchronize());
    cudaFree(e->logits_ws);
    e->logits_ws = nullptr;
    e->logits_ws_elems = 0;
    CK(cudaMalloc(&e->logits_ws, need * 4));
    e->logits_ws_elems = need;
  }
  float* hidden_dev = nullptr;
  const size_t hid_elems = static_cast<size_t>(c.num_hidden_layers + 1) * M * c.hidden_size;
  if (hidden_host) CK(cudaMalloc(&hidden_dev, hid_elems * 4));
  CK(cudaMemcpyAsync(e->ids_ws, ids_host, M * 8, cudaMemcpyHostToDevice, e->stream));
  int r = b2_forward(e, cache, e->ids_ws, B, S, all_dev ? B2_LOGITS_ALL : logits_mode, e->logits_ws, hidden_dev, e->stream);
  if (r == B2_OK && need && logits_host) {
    cudaError_t ce;
    if (logits_mode == B2_LOGITS_ALL_LAST_TO_HOST)   // all positions computed on the device, last rows read back
      ce = cudaMemcpy2DAsync(logits_host, V * 4, e->logits_ws + (static_cast<size_t>(S) - 1) * V, static_cast<size_t>(S) * V * 4,
                             V * 4, B, cudaMemcpyDeviceToHost, e->stream);
    else
      ce = cudaMemcpyAsync(logits_host, e->logits_ws, need * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce != cudaSuccess) r = fail(B2_ERR_CUDA, "logits D2H failed: %s", cudaGetErrorString(ce));
  }
  if (r == B2_OK && hidden_host) {
    cudaError_t ce = cudaMemcpyAsync(hidden_host, hidden_dev, hid_elems * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce != cudaSuccess) r = fail(B2_ERR_CUDA, "hidden D2H failed: %s", cudaGetErrorString(ce));
  }
  cudaError_t se = cudaStreamSynchronize(e->stream);
  if (hidden_dev) cudaFree(hidden_dev);
  if (r == B2_OK && se != cudaSuccess) r = fail(B2_ERR_CUDA, "forward failed: %s", cudaGetErrorString(se));
  return r;
}

extern "C" int b2_set_decode_graph(b2_engine* e, int enable) {
  if (!e) return fail(B2_ERR_ARG, "null engine");
  e->decode_graph = enable != 0;
  return B2_OK;
}

extern "C" int b2_argmax(b2_engine* e, const float* logits, int B, int ld, int64_t* out_ids, void* stream) {
  if (!e || !logits || !out_ids || B <= 0) return fail(B2_ERR_ARG, "bad argmax arguments");
  CK(cudaSetDevice(e->device));
  argmax_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, e->cfg.vocab_size, reinterpret_cast<long long*>(out_ids));
  CK(cudaGetLastError());
  e->st.launches++;
  return B2_OK;
}

extern "C" int b2_sample_min_p(b2_engine* e, const float* logits, int B, int ld, float p_base, const float* u,
                               int64_t* out_ids, float* probs, void* stream) {
  if (!e || !logits || !u || !out_ids || B <= 0) return fail(B2_ERR_ARG, "bad sample arguments");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);   // NULL = the CUDA (legacy) default stream
  min_p_sample_kernel<<<B, 1024, 0, st>>>(logits, ld, e->cfg.vocab_size, p_base, u, reinterpret_cast<long long*>(out_ids), probs);
  CK(cudaGetLastError());
  e->st.launches++;
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------
extern "C" int b2_set_timing(b2_engine* e, int on) {
  if (!e) return fail(B2_ERR_ARG, "null engine");
  e->st.timing = on > 0;       // on = k > 1: per-launch events on every k-th forward only
  e->st.every = on > 1 ? on : 1;
  e->st.sample = false;
  e->st.forwards = 0;
  return B2_OK;
}
extern "C" int b2_reset_stats(b2_engine* e) {
  if (!e) return fail(B2_ERR_ARG, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  for (auto& ev : e->st.used) e->st.pool.emplace_back(ev.a, ev.b);
  e->st.used.clear();
  e->st.launches = e->st.gemm_launches = e->st.attn_launches = 0;
  e->st.forwards = e->st.timed_forwards = 0;
  e->comm_calls = 0;
  e->comm_bytes = 0;
  e->p2p_fused = 0;
  return B2_OK;
}
extern "C" int b2_get_stats(b2_engine* e, b2_stats* out) {
  if (!e || !out) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(e->device));
  memset(out, 0, sizeof(*out));
  out->launches = e->st.launches;
  out->gemm_launches = e->st.gemm_launches;
  out->attn_launches = e->st.attn_launches;
  out->timed_forwards = e->st.timed_forwards;
  out->comm_calls = e->comm_calls;
  out->p2p_fused = e->p2p_fused;
  out->comm_bytes = e->comm_bytes;
  if (!e->st.used.empty()) {
    CK(cudaDeviceSynchronize());
    for (auto& ev : e->st.used) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
      if (ev.kind == 0) { out->gemm_ms += ms; out->gemm_flops += ev.flops; }
      else if (ev.kind == 1) { out->attn_ms += ms; out->attn_flops += ev.flops; }
      else out->comm_ms += ms;
    }
  }
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// single-kernel entry points
// ------------------------------------------------------------------------------------------
extern "C" int b2_op_rmsnorm(const float* x, const float* g, void* ob, float* of, int M, int H, float eps, void* stream) {
  CKR(require_gpu(0));
  if (H % 128 || H > 128 * NORM_MAXV) return fail(B2_ERR_ARG, "H=%d must be a multiple of 128 and <= %d", H, 128 * NORM_MAXV);
  rmsnorm_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(const_cast<float*>(x), nullptr, g, static_cast<bf16*>(ob), H, 0, of, M, H, eps);
  CK(cudaGetLastError());
  return B2_OK;
}
extern "C" int b2_op_add_rmsnorm(float* x, const float* y, const float* g, int M, int H, float eps, void* stream) {
  CKR(require_gpu(0));
  if (H % 128 || H > 128 * NORM_MAXV) return fail(B2_ERR_ARG, "H=%d must be a multiple of 128 and <= %d", H, 128 * NORM_MAXV);
  add_rmsnorm_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, g, M, H, eps, 1, 0);
  CK(cudaGetLastError());
  return B2_OK;
}

extern "C" int b2_op_gemm(const void* a, const void* w, void* out, int M, int N, int K, int epi, int bn, int split_k,
                          float softcap, int act, void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  CKR(require_gpu(dev));
  if (epi == EPI_QKV_ROPE) return fail(B2_ERR_ARG, "use b2_op_qkv_rope");
  CUtensorMap ta, tb;
  const bool two_cta = bn == 512;                     // test hook: bn = 512 selects the 2-CTA kernel (256-wide tiles)
  GemmParams p = base_params(M, N, K);
  p.a_box_rows = (!two_cta && M <= 64) ? (M <= 32 ? 32 : 64) : 0;     // same rule as eng_gemm
  CKR(make_tmap(&ta, a, M, K, K, p.a_box_rows ? p.a_box_rows : GEMM_BM));
  CKR(make_tmap(&tb, w, N, K, K, two_cta ? 128 : bn));
  p.split_k = split_k < 1 ? 1 : split_k;
  p.out = out;
  p.ldo = (epi == EPI_GLU) ? N / 2 : N;
  p.softcap = softcap;
  p.act = act;
  p.band_m = pick_band(N, K);
  CUtensorMap tc = ta;
  if (epi == EPI_F32 || epi == EPI_RESID) CKR(make_tmap_t(&tc, out, M, N, N, 32, true));
  if (two_cta) return launch_gemm2(epi, ta, tb, tc, p, static_cast<cudaStream_t>(stream));
  return launch_gemm(bn, epi, ta, tb, tc, p, static_cast<cudaStream_t>(stream));
}

// few-row cluster split-K GEMM with the residual + half-norm epilogue (decode step): x[M,N] += a[M,K] . w[N,K]^T;
// a_out = bf16(x * gamma); ss[M, N/32] = per-32-column sums of x^2.  cluster = CTAs per output tile (0 = engine's choice).
extern "C" int b2_op_gemm_sk(const void* a, const void* w, float* x, void* a_out, const float* gamma, float* ss, int M, int N, int K,
                             int cluster, void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  CKR(require_gpu(dev));
  GemmParams p = base_params(M, N, K);
  p.a_box_rows = M <= 32 ? 32 : 64;
  p.split_k = cluster > 0 ? cluster : pick_cluster(N / SK_BN, K / GEMM_BK, num_sms());
  p.out = x; p.ldo = N; p.a_out = static_cast<bf16*>(a_out); p.gamma_next = gamma; p.ss_out = ss;
  CUtensorMap ta, tb;
  CKR(make_tmap(&ta, a, M, K, K, p.a_box_rows));
  CKR(make_tmap(&tb, w, N, K, K, SK_BN));
  return launch_gemm_sk(EPI_RESID_NORM, ta, tb, p, static_cast<cudaStream_t>(stream));
}

extern "C" int b2_op_qkv_rope(const void* a, const void* w, void* q, void* kc, void* vc, int B, int S, int H, int NH,
                              int NKV, int HD, int Tmax, int t0, float rope_theta, int bn, void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  CKR(require_gpu(dev));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * S, N = (NH + 2 * NKV) * HD;
  if (bn < HD || bn % HD) return fail(B2_ERR_ARG, "bn=%d must be a multiple of head_dim=%d", bn, HD);
  std::vector<float> cs, sn;
  rope_tables_host(HD, rope_theta, Tmax, cs, sn);
  float *dc = nullptr, *ds = nullptr;
  int* dt0 = nullptr;
  CK(cudaMalloc(&dc, cs.size() * 4));
  CK(cudaMalloc(&ds, sn.size() * 4));
  CK(cudaMalloc(&dt0, 4));
  CK(cudaMemcpy(dc, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ds, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dt0, &t0, 4, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  CKR(make_tmap(&ta, a, M, H, H, GEMM_BM));
  CKR(make_tmap(&tb, w, N, H, H, bn));
  GemmParams p = base_params(M, N, H);
  p.q_cols = NH * HD; p.kv_cols = NKV * HD; p.head_dim = HD; p.S = S; p.Tmax = Tmax; p.nkv = NKV;
  p.q_out = static_cast<bf16*>(q); p.ldq = NH * HD;
  p.k_cache = static_cast<bf16*>(kc); p.v_cache = static_cast<bf16*>(vc);
  p.t0_ptr = dt0; p.rope_cos = dc; p.rope_sin = ds;
  int r = launch_gemm(bn, EPI_QKV_ROPE, ta, tb, ta, p, st);
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(dc); cudaFree(ds); cudaFree(dt0);
  if (r == B2_OK && se != cudaSuccess) return fail(B2_ERR_CUDA, "qkv_rope kernel failed: %s", cudaGetErrorString(se));
  return r;
}

extern "C" int b2_op_attention(const void* q, const void* kc, const void* vc, void* out, int B, int S, int NH, int NKV,
                               int HD, int Tmax, int t0, float scale, float softcap, int causal, int window, int decode,
                               void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  CKR(require_gpu(dev));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* dt0 = nullptr;
  CK(cudaMalloc(&dt0, 4));
  CK(cudaMemcpy(dt0, &t0, 4, cudaMemcpyHostToDevice));
  int r;
  if (decode) {
    if (S != 1) { cudaFree(dt0); return fail(B2_ERR_ARG, "decode attention needs S == 1"); }
    float* part = nullptr;
    int* cnt = nullptr;
    CK(cudaMalloc(&part, decode_part_floats(B, NH, HD) * 4));
    CK(cudaMalloc(&cnt, static_cast<size_t>(B) * NKV * 4));
    CK(cudaMemset(cnt, 0, static_cast<size_t>(B) * NKV * 4));
    CK(cudaDeviceSynchronize());
    r = attn_decode(static_cast<const bf16*>(q), NH * HD, static_cast<const bf16*>(kc), static_cast<const bf16*>(vc),
                    static_cast<bf16*>(out), NH * HD, B, NH, NKV, HD, Tmax, dt0, scale, softcap, window, st, 0, 0, 0, part, cnt,
                    decode_splits(B, NKV, Tmax));
    cudaStreamSynchronize(st);
    cudaFree(part);
    cudaFree(cnt);
  } else {
    AttnParams a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.S = S; a.NH = NH; a.NKV = NKV; a.Tmax = Tmax; a.t0_ptr = dt0; a.scale = scale; a.softcap = softcap;
    a.causal = causal; a.window = window; a.out = static_cast<bf16*>(out); a.ldo = NH * HD;
    r = attn_prefill(static_cast<const bf16*>(q), NH * HD, static_cast<const bf16*>(kc), static_cast<const bf16*>(vc), a, HD, st);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(dt0);
  if (r == B2_OK && se != cudaSuccess) return fail(B2_ERR_CUDA, "attention kernel failed: %s", cudaGetErrorString(se));
  return r;
}

extern "C" int b2_interleave_gate_up(const uint16_t* gate, const uint16_t* up, uint16_t* out, int I, int H) {
  if (!gate || !up || !out || I % 32) return fail(B2_ERR_ARG, "bad interleave arguments (I must be a multiple of 32)");
  const size_t grp = static_cast<size_t>(32) * H;
  for (int g = 0; g < I / 32; ++g) {
    memcpy(out + (2 * g) * grp, gate + g * grp, grp * 2);
    memcpy(out + (2 * g + 1) * grp, up + g * grp, grp * 2);
  }
  return B2_OK;
}
