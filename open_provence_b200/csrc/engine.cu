// libopv_sm100.so -- C ABI (include/opv.h) and host-side launch sequence of the ModernBERT forward.
//
// One engine per (process, device).  Everything is enqueued on the caller's stream; nothing here
// synchronises or allocates device memory (weights, workspace and outputs belong to the caller).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/opv.h"
#include "attention_simt.cuh"
#include "attention_tcgen05.cuh"
#include "attention_tcgen05_local.cuh"
#include "attention_tcgen05_pp.cuh"
#include "attention_tcgen05_q4.cuh"
#include "attention_tcgen05_v3.cuh"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_rowln.cuh"
#include "gemm_tcgen05.cuh"
#include "pointwise.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define OPV_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t err__ = (expr);                                                                     \
    if (err__ != cudaSuccess)                                                                       \
      return fail(OPV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                  __LINE__);                                                                        \
  } while (0)

#define OPV_LAUNCH_CHECK(name)                                                                     \
  do {                                                                                             \
    cudaError_t err__ = cudaGetLastError();                                                        \
    if (err__ != cudaSuccess)                                                                      \
      return fail(OPV_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err__));       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// TMA descriptors: cuTensorMapEncodeTiled is fetched through the runtime so the library does not
// link libcuda (the build box has no driver).
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    OPV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
      return fail(OPV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return OPV_OK;
}

// 2D row-major [rows, cols] tensor, box = [box_rows, 128 B of columns] with 128 B swizzle
// (64 bf16 or 32 fp32 columns per box row).
int make_tmap_2d(CUtensorMap* map, const void* ptr, bool f32, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn encode;
  if (int rc = get_encode_fn(&encode)) return rc;
  const uint64_t elt = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(OPV_ERR_INVALID_ARGUMENT, "TMA base not 16 B aligned");
  if ((cols * elt) % 16 != 0) return fail(OPV_ERR_INVALID_ARGUMENT, "TMA row pitch must be a multiple of 16 B");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * elt};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elt), box_rows};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(ptr), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OPV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return OPV_OK;
}
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_2d(map, ptr, false, rows, cols, box_rows);
}

// Tuning switches.  ``g_defaults`` is what opv_set_option() edits (under g_mu); every engine takes a snapshot at
// opv_create() and can be changed alone with opv_engine_set_option(); the single-op entry points (opv_op_*) read the
// defaults at call time.  Nothing below reads a process-global while launching: each extern "C" entry point copies
// the options of ITS engine (or the defaults) and the SM count of the CURRENT device into thread-local state.
struct Options {
  // bf16 attention kernel: 1 = default (four-Q-tile kernel for global layers; one-pass kernel for sliding-window
  // layers with window <= 128, two-threads-per-row kernel for wider windows), 2 = one-thread-per-row with P in smem, 3 = two-threads-per-row everywhere,
  // 4 = one-thread-per-row (P in TMEM, 2 CTAs / SM) everywhere, 5 = two-Q-tile kernel (1 CTA / SM) everywhere,
  // 6 = one-pass sliding-window kernel (window <= 128; what 1 uses for such layers), 7 = four-Q-tile kernel (global
  // layers only; what 1 uses for them)
  int attention_impl = 1;
  int attention_q4 = 1;                  // 1 = attention_impl 1 runs global layers on the four-Q-tile kernel (default), 0 = on the 2-CTA kernel
  // four-Q-tile kernel: every n-th probability pair of a row on the FMA pipe (exp2_poly2) instead of MUFU.EX2; 0 = none.
  // 64 x 2048 x 8 heads, per global layer: 0 -> 0.645 ms, 16 / 12 -> 0.623, 8 -> 0.606, 6 -> 0.604, 4 -> 0.619, 3 -> 0.622,
  // 2 -> 0.698 (issue-bound); in situ 6.04 -> 5.71 ms per step with 6.
  int attention_q4_poly = 6;
  int attention_debug = 0;               // timing experiments of the two-Q-tile kernel (attention_tcgen05_pp.cuh), 0 = off
  long long* attention_trace = nullptr;  // device buffer for clock64() stamps (tools/attn_check.py); nullptr in the product
  // bf16 GEMM with N % 256 == 0: 1 = CTA-pair kernel (cta_group::2, 256 x 256 tiles), 0 = single-CTA kernel
  int gemm_pair = 1;
  // ROPE GEMM: 1 = a cluster finishes all column tiles of a row block before the next row block (cos|sin staged once)
  int gemm_group_rows = 1;
  // Where the LayerNorms between the projections run (bf16 forwards with at least one 256-row block per CTA pair):
  // 0 = standalone layernorm_kernel launches;
  // 1 = inside the residual GEMM that completes the rows, RE-READING them (RESIDUAL_LN, gemm_tcgen05.cuh): bit-identical,
  //     measured slower (2122-2204 against 2221-2260 pairs/s: the re-reads miss L2);
  // 2 = (default) full-row residual GEMM with the LayerNorm computed from TMEM (gemm_rowln.cuh) for attn.Wo when the
  //     hidden size is 256 or 512, and for mlp.Wo when it is 256: base-130M 28.42 -> 28.0-28.3 ms per step, xsmall-30M
  //     4.97 -> 4.66 ms; other widths keep the standalone launches.
  int ln_fuse = 2;
  // ln_fuse = 2 applies to forwards with at least this many 256-row blocks; -1 = measured crossover (B200,
  // tools/lnfuse_sweep.sh): hidden size 256 always (2048 .. 65536 tokens: 3-5 % faster at every size), 512 from 48
  // blocks (2048 / 4096 / 8192 tokens: 7.5 / 6 / 3.4 % slower -- half as many CTA pairs busy as with 256-wide tiles
  // -- 16384 tokens and up: 0.3-1.4 % faster).  Below the crossover the LayerNorm statistics are summed in
  // layernorm_kernel's order, above in gemm_rowln's: x can differ by one bf16 ulp in a few elements per 10^5.
  int ln_fuse_min_blocks = -1;
  // 1 = the kernels that call pdl_wait() (GEMMs, attention, LayerNorm family) are launched with programmatic stream
  // serialization, so each one's prologue overlaps its predecessor's tail; 0 = plain stream order
  int pdl = 1;
  // Measured on B200 (profiles/r1t_pdl.md): PDL takes 18-23 % off a 481-token forward, 12 % off a 2048-token one and
  // 1.5 % off 32768 tokens, but costs 1.4 % at 65536 and 0.9 % at 131072 tokens (dependents parked on the SMs while
  // 200-us kernels run), so early release is applied to forwards of at most this many tokens.
  long long pdl_max_tokens = 32768;
  // 1 = forwards ABOVE pdl_max_tokens also get the attribute, but their GEMM / attention kernels do not release the
  // dependent early (it launches when the last CTA exits), so nothing is parked on the SMs and only the launch gap is
  // hidden: 30.50 / 30.38 -> 30.23 / 30.28 ms per 131072-token step (profiles/r1t_pdl.md).
  int pdl_late = 1;
};

constexpr int kMaxDevices = 64;
struct DeviceState {
  bool attrs_set = false;  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is PER DEVICE
  int num_sms = 0;
};
std::mutex g_mu;  // guards g_defaults and g_devices
Options g_defaults;
DeviceState g_devices[kMaxDevices];

thread_local Options t_opt;             // options of the call being enqueued
thread_local int t_num_sms = 0;         // SM count of the device the call runs on
thread_local long long t_forward_tokens = 0;  // tokens of the forward being enqueued (0 outside opv_forward_packed)
#define g_num_sms t_num_sms
inline bool pdl_attr_on() { return t_opt.pdl && (t_forward_tokens <= t_opt.pdl_max_tokens || t_opt.pdl_late); }
inline int pdl_late_flag() { return t_forward_tokens > t_opt.pdl_max_tokens ? 1 : 0; }

int set_option_in(Options& o, const char* name, int64_t value);

// Launch through cudaLaunchKernelEx so that the PDL attribute can ride along.  Only for kernels that call
// opv::pdl_wait() before their first global-memory access.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr_on() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <int BLOCK_N, int EPI>
int set_gemm_attr() {
  OPV_CUDA(cudaFuncSetAttribute(opv::gemm_bf16_tcgen05_kernel<BLOCK_N, EPI>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, opv::GemmSmemLayout<BLOCK_N, EPI>::kTotal));
  return OPV_OK;
}

template <int EPI>
int set_gemm_pair_attr() {
  OPV_CUDA(cudaFuncSetAttribute(opv::gemm_bf16_tcgen05_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                opv::GemmPairSmemLayout<EPI>::kTotal));
  return OPV_OK;
}

// Per-device one-time setup (function attributes are per device) + the calling thread's launch context.
int ensure_device_setup() {
  int dev = 0;
  OPV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(OPV_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceState& ds = g_devices[dev];
  if (!ds.attrs_set) {
    cudaDeviceProp prop;
    OPV_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
      return fail(OPV_ERR_UNSUPPORTED, "libopv_sm100 needs an sm_100 (B200) device, found sm_%d%d", prop.major,
                  prop.minor);
    if (int rc = set_gemm_attr<256, opv::kEpiStore>()) return rc;
    if (int rc = set_gemm_attr<256, opv::kEpiRope>()) return rc;
    if (int rc = set_gemm_attr<256, opv::kEpiResidual>()) return rc;
    if (int rc = set_gemm_attr<256, opv::kEpiGeglu>()) return rc;
    if (int rc = set_gemm_attr<128, opv::kEpiStore>()) return rc;
    if (int rc = set_gemm_attr<128, opv::kEpiRope>()) return rc;
    if (int rc = set_gemm_attr<128, opv::kEpiResidual>()) return rc;
    if (int rc = set_gemm_pair_attr<opv::kEpiStore>()) return rc;
    if (int rc = set_gemm_pair_attr<opv::kEpiRope>()) return rc;
    if (int rc = set_gemm_pair_attr<opv::kEpiResidual>()) return rc;
    if (int rc = set_gemm_pair_attr<opv::kEpiGeglu>()) return rc;
    if (int rc = set_gemm_pair_attr<opv::kEpiResidualLn>()) return rc;
    OPV_CUDA(cudaFuncSetAttribute(opv::gemm_rowln_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::RowLnSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::FaSmemLayout<true>::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::FaSmemLayout<false>::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Fa3SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_q4_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::Q4SmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_local_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::LoSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_pp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::PpSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::PpSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_pp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::PpSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_pp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::PpSmemLayout::kTotal));
    OPV_CUDA(cudaFuncSetAttribute(opv::attention_tcgen05_pp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  opv::PpSmemLayout::kTotal));
    ds.num_sms = prop.multiProcessorCount;
    ds.attrs_set = true;
  }
  t_num_sms = ds.num_sms;
  return OPV_OK;
}

// Makes `device` current for the duration of an entry point (and restores the caller's device afterwards).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(int device) {
    cudaError_t err = cudaGetDevice(&prev);
    if (err != cudaSuccess) return err;
    if (prev != device) {
      err = cudaSetDevice(device);
      switched = err == cudaSuccess;
    }
    return err;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

inline Options defaults_snapshot() {
  std::lock_guard<std::mutex> lock(g_mu);
  return g_defaults;
}

template <int BLOCK_N, int EPI>
int launch_gemm_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_c,
                   const opv::GemmEpilogueArgs& ep, int64_t M, int N, int K, cudaStream_t stream) {
  const int64_t tiles = ((M + opv::kGemmBlockM - 1) / opv::kGemmBlockM) * (N / BLOCK_N);
  const int grid = static_cast<int>(tiles < g_num_sms ? tiles : g_num_sms);
  opv::GemmEpilogueArgs ep_launch = ep;
  ep_launch.pdl_late = pdl_late_flag();
  launch_pdl(opv::gemm_bf16_tcgen05_kernel<BLOCK_N, EPI>, dim3(grid), dim3(opv::gemm_threads(EPI)),
             opv::GemmSmemLayout<BLOCK_N, EPI>::kTotal, stream, tm_a, tm_b, tm_c, ep_launch, (int)M, N, K);
  OPV_LAUNCH_CHECK("gemm_bf16_tcgen05_kernel");
  return OPV_OK;
}

template <int EPI>
int launch_gemm_pair(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_c,
                     const opv::GemmEpilogueArgs& ep, int64_t M, int N, int K, cudaStream_t stream) {
  const int64_t pairs_m = (M + 2 * opv::kGemmBlockM - 1) / (2 * opv::kGemmBlockM);
  const int64_t tiles = pairs_m * (N / 256);
  const int max_clusters = g_num_sms / 2;
  const int clusters = static_cast<int>(tiles < max_clusters ? tiles : max_clusters);
  opv::GemmEpilogueArgs ep_launch = ep;
  ep_launch.group_rows = (EPI == opv::kEpiRope && t_opt.gemm_group_rows && pairs_m >= 4 * clusters) ? 1 : 0;
  if (EPI == opv::kEpiResidualLn) ep_launch.group_rows = 1;  // the LayerNorm needs a row block's column tiles in one CTA pair
  ep_launch.pdl_late = pdl_late_flag();
  launch_pdl(opv::gemm_bf16_tcgen05_pair_kernel<EPI>, dim3(2 * clusters), dim3(opv::gemm_threads(EPI)),
             opv::GemmPairSmemLayout<EPI>::kTotal, stream, tm_a, tm_b, tm_c, ep_launch, (int)M, N, K);
  OPV_LAUNCH_CHECK("gemm_bf16_tcgen05_pair_kernel");
  return OPV_OK;
}

// Residual GEMM + the following LayerNorm from on-chip data (gemm_rowln.cuh): N = hidden size in {256, 512}.
int launch_gemm_rowln(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_r, const CUtensorMap& tm_x,
                      const float* ln_w, float eps, int64_t M, int N, int K, cudaStream_t stream) {
  if (M <= 0) return OPV_OK;
  if ((N != 256 && N != 512) || K % opv::kGemmBlockK != 0)
    return fail(OPV_ERR_UNSUPPORTED, "row-LN GEMM needs N in {256, 512} and K %% 64 == 0 (N = %d, K = %d)", N, K);
  if (M > 0x7fffff00LL) return fail(OPV_ERR_UNSUPPORTED, "M = %lld rows exceeds the int32 tile range", (long long)M);
  const int64_t pairs_m = (M + 2 * opv::kGemmBlockM - 1) / (2 * opv::kGemmBlockM);
  const int max_clusters = g_num_sms / 2;
  const int clusters = static_cast<int>(pairs_m < max_clusters ? pairs_m : max_clusters);
  launch_pdl(opv::gemm_rowln_pair_kernel, dim3(2 * clusters), dim3(opv::kRowLnThreads), opv::RowLnSmemLayout::kTotal, stream, tm_a, tm_b,
             tm_r, tm_x, ln_w, eps, pdl_late_flag(), (int)M, N, K);
  OPV_LAUNCH_CHECK("gemm_rowln_pair_kernel");
  return OPV_OK;
}

int gemm_block_n(int N, int epi) {
  if (epi == opv::kEpiGeglu) return (N % 256 == 0) ? 256 : 0;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  return 0;
}

// A: [M, K] bf16 activations, W: [N, K] bf16 weight (tensor map with box rows = BLOCK_N), C through tm_c:
// bf16 [M, N] (STORE/ROPE) or [M, N/2] (GEGLU) with a 64-column box, fp32 [M, N] (RESIDUAL) with a 32-column box.
// pair: W's tensor map was built with a 128-row box (half tile per CTA) for the CTA-pair kernel.
int gemm_bf16(int epi, bool pair, const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_c,
              const opv::GemmEpilogueArgs& ep, int64_t M, int N, int K, cudaStream_t stream) {
  if (M <= 0) return OPV_OK;
  if (M > 0x7fffff00LL) return fail(OPV_ERR_UNSUPPORTED, "M = %lld rows exceeds the int32 tile range", (long long)M);
  if (K % opv::kGemmBlockK != 0) return fail(OPV_ERR_UNSUPPORTED, "K = %d must be a multiple of 64", K);
  const int bn = gemm_block_n(N, epi);
  if (bn == 0) return fail(OPV_ERR_UNSUPPORTED, "N = %d must be a multiple of 128 (256 for GeGLU)", N);
  if (bn == 256 && pair) {
    switch (epi) {
      case opv::kEpiStore: return launch_gemm_pair<opv::kEpiStore>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiRope: return launch_gemm_pair<opv::kEpiRope>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiResidual: return launch_gemm_pair<opv::kEpiResidual>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiGeglu: return launch_gemm_pair<opv::kEpiGeglu>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiResidualLn:
        if (N > 1024 || !ep.ln_w || !ep.ln_r || !ep.ln_x)
          return fail(OPV_ERR_INVALID_ARGUMENT, "RESIDUAL_LN epilogue needs N in {256,512,768,1024} and the LayerNorm operands");
        return launch_gemm_pair<opv::kEpiResidualLn>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
    }
  } else if (bn == 256) {
    switch (epi) {
      case opv::kEpiStore: return launch_gemm_tc<256, opv::kEpiStore>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiRope: return launch_gemm_tc<256, opv::kEpiRope>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiResidual: return launch_gemm_tc<256, opv::kEpiResidual>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiGeglu: return launch_gemm_tc<256, opv::kEpiGeglu>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
    }
  } else {
    switch (epi) {
      case opv::kEpiStore: return launch_gemm_tc<128, opv::kEpiStore>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiRope: return launch_gemm_tc<128, opv::kEpiRope>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
      case opv::kEpiResidual: return launch_gemm_tc<128, opv::kEpiResidual>(tm_a, tm_b, tm_c, ep, M, N, K, stream);
    }
  }
  if (epi == opv::kEpiResidualLn) return fail(OPV_ERR_UNSUPPORTED, "RESIDUAL_LN epilogue: CTA-pair kernel only (N %% 256 == 0)");
  return fail(OPV_ERR_INVALID_ARGUMENT, "unknown epilogue %d", epi);
}

// rows of W per TMA box: the CTA-pair kernel loads half of the 256-row tile per CTA
inline int weight_box_rows(int bn, bool pair) { return (bn == 256 && pair) ? 128 : bn; }

int gemm_f32(bool accumulate, const float* a, const float* w, float* c, int64_t M, int N, int K, int64_t ldc,
             cudaStream_t stream) {
  if (M <= 0) return OPV_OK;
  if (N % 64 != 0 || K % 16 != 0) return fail(OPV_ERR_UNSUPPORTED, "fp32 GEMM needs N %% 64 == 0 and K %% 16 == 0");
  dim3 grid(N / 64, static_cast<unsigned>((M + 63) / 64));
  if (grid.y > 65535u) return fail(OPV_ERR_UNSUPPORTED, "fp32 GEMM: too many rows for one launch (%lld)", (long long)M);
  if (accumulate)
    opv::gemm_f32_simt_kernel<true><<<grid, 256, 0, stream>>>(a, w, c, (int)M, N, K, ldc);
  else
    opv::gemm_f32_simt_kernel<false><<<grid, 256, 0, stream>>>(a, w, c, (int)M, N, K, ldc);
  OPV_LAUNCH_CHECK("gemm_f32_simt_kernel");
  return OPV_OK;
}

// ------------------------------------------------------------------------------------------------
// row-kernel dispatch on H / 128
// ------------------------------------------------------------------------------------------------
#define OPV_DISPATCH_VEC(H, ...)                                                       \
  switch ((H) / 128) {                                                                 \
    case 1: { constexpr int VEC = 1; __VA_ARGS__; } break;                             \
    case 2: { constexpr int VEC = 2; __VA_ARGS__; } break;                             \
    case 3: { constexpr int VEC = 3; __VA_ARGS__; } break;                             \
    case 4: { constexpr int VEC = 4; __VA_ARGS__; } break;                             \
    case 6: { constexpr int VEC = 6; __VA_ARGS__; } break;                             \
    case 8: { constexpr int VEC = 8; __VA_ARGS__; } break;                             \
    default: return fail(OPV_ERR_UNSUPPORTED, "hidden_size %d not in {128,256,384,512,768,1024}", (H)); \
  }

inline unsigned row_grid(int64_t M) { return static_cast<unsigned>((M + opv::kRowWarps - 1) / opv::kRowWarps); }

template <typename OutT>
int launch_layernorm(const float* h, const float* w, OutT* x, int64_t M, int H, float eps, cudaStream_t s) {
  if (M <= 0) return OPV_OK;
  OPV_DISPATCH_VEC(H, launch_pdl(opv::layernorm_kernel<OutT, VEC>, dim3(row_grid(M)), dim3(opv::kRowWarps * 32), 0, s, h,
                                 w, x, M, eps));
  OPV_LAUNCH_CHECK("layernorm_kernel");
  return OPV_OK;
}

template <typename EmbT, typename OutT>
int launch_embed_ln(const int32_t* ids, const EmbT* emb, const float* w, float* h, OutT* x, int64_t M, int H, int V,
                    float eps, cudaStream_t s) {
  if (M <= 0) return OPV_OK;
  OPV_DISPATCH_VEC(H, launch_pdl(opv::embed_ln_kernel<EmbT, OutT, VEC>, dim3(row_grid(M)), dim3(opv::kRowWarps * 32), 0,
                                 s, ids, emb, w, h, x, M, V, eps));
  OPV_LAUNCH_CHECK("embed_ln_kernel");
  return OPV_OK;
}

int launch_final_prune(const float* h, const float* w, const float* wp, const float* bp, float* logits, int64_t M,
                       int H, float eps, cudaStream_t s) {
  if (M <= 0) return OPV_OK;
  OPV_DISPATCH_VEC(H, launch_pdl(opv::final_ln_prune_kernel<VEC>, dim3(row_grid(M)), dim3(opv::kRowWarps * 32), 0, s, h,
                                 w, wp, bp, logits, M, eps));
  OPV_LAUNCH_CHECK("final_ln_prune_kernel");
  return OPV_OK;
}

// tm_qkv: TMA map of the packed qkv [T, 3H] buffer with a 128-row x 64-column box (tcgen05 kernels only)
int launch_attention(int dtype, const void* qkv, void* out, const int32_t* cu, int n_seqs, int max_seqlen, int heads,
                     int half_window, const CUtensorMap* tm_qkv, const CUtensorMap* tm_qkv64, cudaStream_t s) {
  if (n_seqs <= 0 || max_seqlen <= 0) return OPV_OK;
  if (n_seqs > 65535) return fail(OPV_ERR_UNSUPPORTED, "at most 65535 sequences per launch");
  const int H = heads * 64;
  if (dtype == OPV_DTYPE_BF16) {
    if (!tm_qkv) return fail(OPV_ERR_INVALID_ARGUMENT, "tcgen05 attention needs the qkv tensor map");
    const int impl = t_opt.attention_impl;
    if (impl == 7 && half_window >= 0)
      return fail(OPV_ERR_UNSUPPORTED, "attention_impl 7 (four-Q-tile kernel) implements global attention only");
    if (impl == 7 || (impl == 1 && half_window < 0 && t_opt.attention_q4)) {
      // global layers: ONE CTA per SM, four 128-row query tiles in flight, 64-key blocks (attention_tcgen05_q4.cuh)
      if (!tm_qkv64) return fail(OPV_ERR_INVALID_ARGUMENT, "the four-Q-tile attention kernel needs the 64-row qkv tensor map");
      const int supers_per_seq = (max_seqlen + opv::kQ4SuperM - 1) / opv::kQ4SuperM;
      const int64_t total = static_cast<int64_t>(n_seqs) * heads * supers_per_seq;
      if (total > 0x7fffffffLL) return fail(OPV_ERR_UNSUPPORTED, "too many attention tiles for one launch");
      const int grid = static_cast<int>(total < g_num_sms ? total : g_num_sms);
      auto* kernel = opv::attention_tcgen05_q4_kernel<0>;
      switch (t_opt.attention_q4_poly) {  // every n-th probability pair on the FMA pipe instead of MUFU.EX2
        case 2: kernel = opv::attention_tcgen05_q4_kernel<2>; break;
        case 3: kernel = opv::attention_tcgen05_q4_kernel<3>; break;
        case 4: kernel = opv::attention_tcgen05_q4_kernel<4>; break;
        case 6: kernel = opv::attention_tcgen05_q4_kernel<6>; break;
        case 8: kernel = opv::attention_tcgen05_q4_kernel<8>; break;
        case 12: kernel = opv::attention_tcgen05_q4_kernel<12>; break;
        case 16: kernel = opv::attention_tcgen05_q4_kernel<16>; break;
        default: break;
      }
      launch_pdl(kernel, dim3(grid), dim3(opv::kQ4Threads), opv::Q4SmemLayout::kTotal, s, *tm_qkv, *tm_qkv64,
                 static_cast<__nv_bfloat16*>(out), cu, H, n_seqs, supers_per_seq, pdl_late_flag());
      OPV_LAUNCH_CHECK("attention_tcgen05_q4_kernel");
      return OPV_OK;
    }
    if (impl == 5) {
      // persistent, ONE CTA per SM: (sequence, head, 256-query super tile) list with a grid stride, two Q tiles in flight
      const int supers_per_seq = (max_seqlen + opv::kPpSuperM - 1) / opv::kPpSuperM;
      const int64_t total = static_cast<int64_t>(n_seqs) * heads * supers_per_seq;
      if (total > 0x7fffffffLL) return fail(OPV_ERR_UNSUPPORTED, "too many attention tiles for one launch");
      const int grid = static_cast<int>(total < g_num_sms ? total : g_num_sms);
      auto* kernel = opv::attention_tcgen05_pp_kernel<0>;
      switch (t_opt.attention_debug) {  // timing experiments of tools/attn_check.py; 0 in the product
        case 1: kernel = opv::attention_tcgen05_pp_kernel<1>; break;
        case 2: kernel = opv::attention_tcgen05_pp_kernel<2>; break;
        case 3: kernel = opv::attention_tcgen05_pp_kernel<3>; break;
        case 4: kernel = opv::attention_tcgen05_pp_kernel<4>; break;
        default: break;
      }
      launch_pdl(kernel, dim3(grid), dim3(opv::kPpThreads), opv::PpSmemLayout::kTotal, s, *tm_qkv,
                 static_cast<__nv_bfloat16*>(out), cu, H, half_window, n_seqs, supers_per_seq, pdl_late_flag(),
                 t_opt.attention_trace);
      OPV_LAUNCH_CHECK("attention_tcgen05_pp_kernel");
      return OPV_OK;
    }
    // persistent: two CTAs per SM walk the (sequence, head, query tile) list with a grid stride
    const int tiles_per_seq = (max_seqlen + opv::kFaBlockM - 1) / opv::kFaBlockM;
    const int64_t total_tiles = static_cast<int64_t>(n_seqs) * heads * tiles_per_seq;
    if (total_tiles > 0x7fffffffLL) return fail(OPV_ERR_UNSUPPORTED, "too many attention tiles for one launch");
    const int grid = static_cast<int>(total_tiles < 2 * g_num_sms ? total_tiles : 2 * g_num_sms);
    // sliding-window layers whose band fits one 256-key score tile (window <= 128, every published checkpoint): the
    // one-pass kernel; 6 forces it (and fails for wider windows), 3 / 4 force one of the online-softmax kernels
    if (impl == 6 && (half_window < 0 || half_window > opv::kLoMaxHalfWindow))
      return fail(OPV_ERR_UNSUPPORTED, "attention_impl 6 (one-pass sliding window) needs 0 <= half_window <= %d",
                  opv::kLoMaxHalfWindow);
    if (impl == 6 || (impl == 1 && half_window >= 0 && half_window <= opv::kLoMaxHalfWindow)) {
      launch_pdl(opv::attention_local_onepass_kernel, dim3(grid), dim3(opv::kLoThreads), opv::LoSmemLayout::kTotal, s,
                 *tm_qkv, static_cast<__nv_bfloat16*>(out), cu, H, half_window, n_seqs, tiles_per_seq, pdl_late_flag());
      OPV_LAUNCH_CHECK("attention_local_onepass_kernel");
      return OPV_OK;
    }
    // wider windows: two softmax threads per row (0.156 vs 0.176 ms per layer at 64 x 2048)
    if (impl == 3 || (impl == 1 && half_window >= 0))
      launch_pdl(opv::attention_tcgen05_v3_kernel, dim3(grid), dim3(opv::kFa3Threads), opv::Fa3SmemLayout::kTotal, s,
                 *tm_qkv, static_cast<__nv_bfloat16*>(out), cu, H, half_window, n_seqs, tiles_per_seq,
                 pdl_late_flag());
    else if (impl == 4 || impl == 1)
      launch_pdl(opv::attention_tcgen05_kernel<true>, dim3(grid), dim3(opv::kFaThreads), opv::FaSmemLayout<true>::kTotal,
                 s, *tm_qkv, static_cast<__nv_bfloat16*>(out), cu, H, half_window, n_seqs, tiles_per_seq,
                 t_opt.attention_trace, pdl_late_flag());
    else
      launch_pdl(opv::attention_tcgen05_kernel<false>, dim3(grid), dim3(opv::kFaThreads),
                 opv::FaSmemLayout<false>::kTotal, s, *tm_qkv, static_cast<__nv_bfloat16*>(out), cu, H, half_window,
                 n_seqs, tiles_per_seq, t_opt.attention_trace, pdl_late_flag());
    OPV_LAUNCH_CHECK("attention_tcgen05_kernel");
  } else {
    dim3 grid((max_seqlen + 3) / 4, heads, n_seqs);
    opv::attention_simt_kernel<float>
        <<<grid, 128, 0, s>>>(static_cast<const float*>(qkv), static_cast<float*>(out), cu, H, half_window);
    OPV_LAUNCH_CHECK("attention_simt_kernel");
  }
  return OPV_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int set_option_in(Options& o, const char* name, int64_t value) {
  if (strcmp(name, "attention_impl") == 0) {
    if (value < 1 || value > 7) return fail(OPV_ERR_INVALID_ARGUMENT, "attention_impl must be 1 .. 7");
    o.attention_impl = static_cast<int>(value);
  } else if (strcmp(name, "attention_q4") == 0) {
    o.attention_q4 = value != 0;
  } else if (strcmp(name, "attention_q4_poly") == 0) {
    if (value != 0 && value != 2 && value != 3 && value != 4 && value != 6 && value != 8 && value != 12 && value != 16)
      return fail(OPV_ERR_INVALID_ARGUMENT, "attention_q4_poly must be 0, 2, 3, 4, 6, 8, 12 or 16");
    o.attention_q4_poly = static_cast<int>(value);
  } else if (strcmp(name, "attention_debug") == 0) {
    if (value < 0 || value > 4) return fail(OPV_ERR_INVALID_ARGUMENT, "attention_debug must be 0 .. 4");
    o.attention_debug = static_cast<int>(value);
  } else if (strcmp(name, "attention_trace_ptr") == 0) {
    o.attention_trace = reinterpret_cast<long long*>(static_cast<intptr_t>(value));
  } else if (strcmp(name, "gemm_group_rows") == 0) {
    o.gemm_group_rows = value != 0;
  } else if (strcmp(name, "ln_fuse") == 0) {
    if (value < 0 || value > 2) return fail(OPV_ERR_INVALID_ARGUMENT, "ln_fuse must be 0, 1 or 2");
    o.ln_fuse = static_cast<int>(value);
  } else if (strcmp(name, "ln_fuse_min_blocks") == 0) {
    o.ln_fuse_min_blocks = static_cast<int>(value);
  } else if (strcmp(name, "pdl") == 0) {
    o.pdl = value != 0;
  } else if (strcmp(name, "pdl_late") == 0) {
    o.pdl_late = value != 0;
  } else if (strcmp(name, "pdl_max_tokens") == 0) {
    o.pdl_max_tokens = value;
  } else if (strcmp(name, "gemm_pair") == 0) {
    o.gemm_pair = value != 0;
  } else {
    return fail(OPV_ERR_INVALID_ARGUMENT, "unknown option '%s'", name);
  }
  return OPV_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
struct ProfEvent {
  int cls, n;
  cudaEvent_t a, b;
};

struct opv_engine {
  bool profiling = false;
  int64_t launches = 0;
  std::vector<ProfEvent> prof;
  opv_config cfg;
  opv_weights w;
  std::vector<opv_layer_weights> layers;
  std::vector<CUtensorMap> tm_wqkv, tm_wo, tm_wi, tm_wo2;  // bf16 mode: one per layer; F32_TC: three (hi | mid | lo) per layer
  bool gemm_pair = true;                                   // CTA-pair GEMM kernel for the 256-wide tiles
  Options opt;                                             // snapshot of the defaults at opv_create (+ opv_engine_set_option)
  int device;
  size_t elt;  // bytes per operand element
  // Tensor maps of the activation buffers depend on (workspace address, token count) only: encoded once per
  // distinct pair instead of on every forward (5 cuTensorMapEncodeTiled calls ~ 10 us of a 1 ms single-pair call).
  struct ActMaps {
    const void* ws = nullptr;
    int64_t tokens = -1;
    uint64_t stamp = 0;
    CUtensorMap x, attn, act, qkv, qkv64, h, u;
  };
  ActMaps act_maps[4];
  uint64_t act_maps_clock = 0;
  std::mutex mu;  // act_maps, prof, launches
};

// Counts kernel launches and, when profiling is on, brackets them with CUDA events on the launch stream.
struct LaunchScope {
  opv_engine* e;
  cudaStream_t s;
  int cls, n;
  cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(opv_engine* e_, cudaStream_t s_, int cls_, int n_ = 1) : e(e_), s(s_), cls(cls_), n(n_) {
    e->launches += n;
    if (e->profiling) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
    }
  }
  ~LaunchScope() {
    if (a) {
      cudaEventRecord(b, s);
      e->prof.push_back({cls, n, a, b});
    }
  }
};

struct WorkspaceLayout {
  size_t h, x, qkv, attn, act, u, pos, cu, status, cls, y, pool, split, total;
};

static WorkspaceLayout workspace_layout(const opv_engine* e, int64_t T, int64_t n_seqs) {
  const size_t H = e->cfg.hidden_size, I = e->cfg.intermediate_size, elt = e->elt;
  const bool unfused = e->cfg.dtype != OPV_DTYPE_BF16 || !e->cfg.fuse_epilogues;
  WorkspaceLayout l;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off += align_up(bytes, 1024);
    return at;
  };
  const size_t rows = static_cast<size_t>(T > 0 ? T : 1);
  l.h = take(rows * H * 4);
  l.x = take(rows * H * elt);
  l.qkv = take(rows * 3 * H * elt);
  l.attn = take(rows * H * elt);
  l.act = take(rows * I * elt);
  l.u = take(unfused ? rows * 2 * I * elt : 0);
  l.pos = take(rows * 4);
  const size_t seqs = static_cast<size_t>(n_seqs > 0 ? n_seqs : 1);
  l.cu = take((seqs + 1) * 4);  // cu_seqlens clamped into [0, T] (positions_checked_kernel)
  l.status = take(seqs * 4);    // per sequence: 1 = its boundaries were not what the contract says
  l.cls = take(seqs * H * 4);  // rank head: LN(CLS row)
  l.y = take(seqs * H * 4);    // rank head: gelu(dense(cls))
  // mean pooling: per-chunk sums of the final-normed rows (slot = begin / chunk + s + c, pointwise.cuh)
  l.pool = take(e->cfg.classifier_pooling ? (rows / opv::kPoolChunkRows + seqs + 2) * H * 4 : 0);
  // F32_TC: hi | mid | lo bf16 planes of the current GEMM's A operand (widest: the [T, I] GeGLU output)
  l.split = take(e->cfg.dtype == OPV_DTYPE_F32_TC ? 3 * rows * (I > H ? I : H) * 2 : 0);
  l.total = off;
  return l;
}

// host_pack.cu reports errors through the same thread-local string
void opv_detail_set_error(const char* message) { g_last_error = message ? message : ""; }

extern "C" {

const char* opv_last_error(void) { return g_last_error.c_str(); }
int opv_abi_version(void) { return OPV_ABI_VERSION; }

int opv_create(const opv_config* cfg, const opv_weights* w, int device, opv_handle* out) {
  if (!cfg || !w || !out) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_create: null argument");
  if (cfg->abi_version != OPV_ABI_VERSION)
    return fail(OPV_ERR_INVALID_ARGUMENT, "ABI version mismatch: header %d, caller %d", OPV_ABI_VERSION,
                cfg->abi_version);
  if (cfg->num_layers <= 0 || cfg->num_layers > OPV_MAX_LAYERS)
    return fail(OPV_ERR_UNSUPPORTED, "num_layers %d out of range", cfg->num_layers);
  if (cfg->hidden_size != cfg->num_heads * 64)
    return fail(OPV_ERR_UNSUPPORTED, "head_dim must be 64 (hidden_size %d, heads %d)", cfg->hidden_size,
                cfg->num_heads);
  if (cfg->hidden_size % 128 != 0 || cfg->hidden_size > 1024)
    return fail(OPV_ERR_UNSUPPORTED, "hidden_size %d must be a multiple of 128 and <= 1024", cfg->hidden_size);
  if (cfg->intermediate_size % 128 != 0)
    return fail(OPV_ERR_UNSUPPORTED, "intermediate_size %d must be a multiple of 128", cfg->intermediate_size);
  if (cfg->num_labels < 1) return fail(OPV_ERR_INVALID_ARGUMENT, "num_labels must be >= 1");
  if (cfg->classifier_pooling != 0 && cfg->classifier_pooling != 1)
    return fail(OPV_ERR_INVALID_ARGUMENT, "classifier_pooling must be 0 (cls) or 1 (mean), got %d",
                cfg->classifier_pooling);
  if (cfg->dtype != OPV_DTYPE_BF16 && cfg->dtype != OPV_DTYPE_F32 && cfg->dtype != OPV_DTYPE_F32_TC)
    return fail(OPV_ERR_INVALID_ARGUMENT, "unknown dtype %d", cfg->dtype);
  if (!w->h_layers) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_create: weights.h_layers is null");
  DeviceGuard guard;
  OPV_CUDA(guard.enter(device));
  if (int rc = ensure_device_setup()) return rc;

  opv_engine* e = new opv_engine();
  e->cfg = *cfg;
  e->w = *w;
  e->device = device;
  e->elt = cfg->dtype == OPV_DTYPE_BF16 ? 2 : 4;  // activations; F32_TC keeps fp32 activations
  e->opt = defaults_snapshot();
  e->gemm_pair = e->opt.gemm_pair != 0;
  e->layers.assign(w->h_layers, w->h_layers + cfg->num_layers);
  e->w.h_layers = e->layers.data();
  const int H = cfg->hidden_size, I = cfg->intermediate_size;
  for (int l = 0; l < cfg->num_layers; ++l) {
    const opv_layer_weights& lw = e->layers[l];
    if (!lw.d_wqkv || !lw.d_wo || !lw.d_wi || !lw.d_wo2 || !lw.d_mlp_norm || (l > 0 && !lw.d_attn_norm)) {
      delete e;
      return fail(OPV_ERR_INVALID_ARGUMENT, "layer %d: missing weight pointer", l);
    }
  }
  if (cfg->dtype == OPV_DTYPE_BF16) {
    e->tm_wqkv.resize(cfg->num_layers);
    e->tm_wo.resize(cfg->num_layers);
    e->tm_wi.resize(cfg->num_layers);
    e->tm_wo2.resize(cfg->num_layers);
    const bool fused = cfg->fuse_epilogues != 0;
    for (int l = 0; l < cfg->num_layers; ++l) {
      const opv_layer_weights& lw = e->layers[l];
      int rc = 0;
      const int bn_qkv = gemm_block_n(3 * H, fused ? opv::kEpiRope : opv::kEpiStore);
      const int bn_h = gemm_block_n(H, opv::kEpiResidual);
      const int bn_wi = gemm_block_n(2 * I, fused ? opv::kEpiGeglu : opv::kEpiStore);
      if (!bn_qkv || !bn_h || !bn_wi) {
        delete e;
        return fail(OPV_ERR_UNSUPPORTED, "projection widths (3H=%d, H=%d, 2I=%d) do not tile", 3 * H, H, 2 * I);
      }
      const bool pair = e->gemm_pair;
      rc = rc ? rc : make_tmap_bf16(&e->tm_wqkv[l], lw.d_wqkv, 3 * H, H, weight_box_rows(bn_qkv, pair));
      rc = rc ? rc : make_tmap_bf16(&e->tm_wo[l], lw.d_wo, H, H, weight_box_rows(bn_h, pair));
      rc = rc ? rc : make_tmap_bf16(&e->tm_wi[l], lw.d_wi, 2 * I, H, weight_box_rows(bn_wi, pair));
      rc = rc ? rc : make_tmap_bf16(&e->tm_wo2[l], lw.d_wo2, H, I, weight_box_rows(bn_h, pair));
      if (rc) {
        delete e;
        return rc;
      }
    }
  }
  if (cfg->dtype == OPV_DTYPE_F32_TC) {
    const int bn_qkv = gemm_block_n(3 * H, opv::kEpiResidual), bn_h = gemm_block_n(H, opv::kEpiResidual);
    const int bn_wi = gemm_block_n(2 * I, opv::kEpiResidual);
    if (!bn_qkv || !bn_h || !bn_wi) {
      delete e;
      return fail(OPV_ERR_UNSUPPORTED, "projection widths (3H=%d, H=%d, 2I=%d) do not tile", 3 * H, H, 2 * I);
    }
    const bool pair = e->gemm_pair;
    e->tm_wqkv.resize(3 * cfg->num_layers);
    e->tm_wo.resize(3 * cfg->num_layers);
    e->tm_wi.resize(3 * cfg->num_layers);
    e->tm_wo2.resize(3 * cfg->num_layers);
    for (int l = 0; l < cfg->num_layers; ++l) {
      const opv_layer_weights& lw = e->layers[l];
      for (int p = 0; p < 3; ++p) {  // hi | mid | lo planes, [out][in] bf16 each
        using bf16 = __nv_bfloat16;
        int rc = 0;
        rc = rc ? rc : make_tmap_bf16(&e->tm_wqkv[3 * l + p], static_cast<const bf16*>(lw.d_wqkv) + (size_t)p * 3 * H * H, 3 * H, H, weight_box_rows(bn_qkv, pair));
        rc = rc ? rc : make_tmap_bf16(&e->tm_wo[3 * l + p], static_cast<const bf16*>(lw.d_wo) + (size_t)p * H * H, H, H, weight_box_rows(bn_h, pair));
        rc = rc ? rc : make_tmap_bf16(&e->tm_wi[3 * l + p], static_cast<const bf16*>(lw.d_wi) + (size_t)p * 2 * I * H, 2 * I, H, weight_box_rows(bn_wi, pair));
        rc = rc ? rc : make_tmap_bf16(&e->tm_wo2[3 * l + p], static_cast<const bf16*>(lw.d_wo2) + (size_t)p * H * I, H, I, weight_box_rows(bn_h, pair));
        if (rc) {
          delete e;
          return rc;
        }
      }
    }
  }
  *out = e;
  return OPV_OK;
}

int opv_destroy(opv_handle h) {
  delete h;
  return OPV_OK;
}

size_t opv_workspace_bytes(opv_handle h, int64_t max_tokens, int32_t max_seqs) {
  if (!h) return 0;
  return workspace_layout(h, max_tokens, max_seqs).total + 1024;
}

int opv_forward_packed(opv_handle e, const int32_t* d_ids, const int32_t* d_cu_seqlens, int32_t n_seqs,
                       int64_t n_tokens, int32_t max_seqlen, float* d_prune_logits, float* d_rank_logits,
                       void* d_workspace, size_t workspace_bytes, void* stream_) {
  if (!e) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_forward_packed: null engine");
  if (n_seqs < 0 || n_tokens < 0) return fail(OPV_ERR_INVALID_ARGUMENT, "negative sizes");
  if (n_seqs == 0 || n_tokens == 0) return OPV_OK;
  if (!d_ids) return fail(OPV_ERR_INVALID_ARGUMENT, "input_ids must be provided");
  if (!d_cu_seqlens || !d_prune_logits || !d_rank_logits || !d_workspace)
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_forward_packed: null buffer");
  if (max_seqlen <= 0 || max_seqlen > e->cfg.max_positions)
    return fail(OPV_ERR_INVALID_ARGUMENT, "max_seqlen %d outside (0, max_positions=%d]", max_seqlen,
                e->cfg.max_positions);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard;  // the engine's weights, tensor maps and function attributes belong to e->device
  OPV_CUDA(guard.enter(e->device));
  if (int rc = ensure_device_setup()) return rc;
  t_opt = e->opt;
  const opv_config& c = e->cfg;
  const int H = c.hidden_size, I = c.intermediate_size, L = c.num_layers, heads = c.num_heads;
  const int64_t T = n_tokens;
  const WorkspaceLayout wl = workspace_layout(e, T, n_seqs);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 1023) & ~uintptr_t(1023));
  if (static_cast<size_t>(ws - static_cast<uint8_t*>(d_workspace)) + wl.total > workspace_bytes)
    return fail(OPV_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", wl.total + 1024, workspace_bytes);
  float* h = reinterpret_cast<float*>(ws + wl.h);
  void* x = ws + wl.x;
  void* qkv = ws + wl.qkv;
  void* attn = ws + wl.attn;
  void* act = ws + wl.act;
  void* u = ws + wl.u;
  int32_t* pos = reinterpret_cast<int32_t*>(ws + wl.pos);
  const int half_window = c.local_window / 2;
  int rc = OPV_OK;
  struct ForwardTokens {  // launch_pdl() decides per forward whether the PDL attribute pays (pdl_max_tokens)
    explicit ForwardTokens(long long t) { t_forward_tokens = t; }
    ~ForwardTokens() { t_forward_tokens = 0; }
  } forward_tokens(T);

  // every later kernel reads the clamped copy: a malformed cu_seqlens gives wrong results, never an out-of-bounds access
  int32_t* cu_clean = reinterpret_cast<int32_t*>(ws + wl.cu);
  {
    LaunchScope sc(e, stream, OPV_PROF_MISC);
    opv::positions_checked_kernel<<<n_seqs, 256, 0, stream>>>(d_cu_seqlens, cu_clean, reinterpret_cast<int32_t*>(ws + wl.status),
                                                             pos, T, max_seqlen, c.max_positions);
    OPV_LAUNCH_CHECK("positions_checked_kernel");
  }
  d_cu_seqlens = cu_clean;

  if (c.dtype == OPV_DTYPE_BF16) {
    using bf16 = __nv_bfloat16;
    const bool fused = c.fuse_epilogues != 0;
    CUtensorMap tm_x, tm_attn, tm_act, tm_qkv, tm_qkv64, tm_h, tm_u;
    {
      std::lock_guard<std::mutex> lock(e->mu);
      opv_engine::ActMaps* slot = nullptr;
      opv_engine::ActMaps* victim = &e->act_maps[0];
      for (auto& m : e->act_maps) {
        if (m.ws == ws && m.tokens == T) slot = &m;
        if (m.stamp < victim->stamp) victim = &m;
      }
      if (!slot) {
        slot = victim;
        slot->tokens = -1;  // invalid until every map below is encoded
        if ((rc = make_tmap_bf16(&slot->qkv, qkv, T, 3 * H, opv::kFaBlockM))) return rc;  // GEMM store + attention loads
        if ((rc = make_tmap_bf16(&slot->qkv64, qkv, T, 3 * H, opv::kQ4BlockN))) return rc;  // 64-key K / V blocks
        if ((rc = make_tmap_2d(&slot->h, h, true, T, H, opv::kGemmBlockM))) return rc;     // residual reduce-add
        if (!fused && (rc = make_tmap_bf16(&slot->u, u, T, 2 * I, opv::kGemmBlockM))) return rc;
        if ((rc = make_tmap_bf16(&slot->x, x, T, H, opv::kGemmBlockM))) return rc;
        if ((rc = make_tmap_bf16(&slot->attn, attn, T, H, opv::kGemmBlockM))) return rc;
        if ((rc = make_tmap_bf16(&slot->act, act, T, I, opv::kGemmBlockM))) return rc;
        slot->ws = ws;
        slot->tokens = T;
      }
      slot->stamp = ++e->act_maps_clock;
      tm_x = slot->x, tm_attn = slot->attn, tm_act = slot->act, tm_qkv = slot->qkv, tm_qkv64 = slot->qkv64, tm_h = slot->h, tm_u = slot->u;
    }
    {
      LaunchScope sc(e, stream, OPV_PROF_EMBED);
      rc = launch_embed_ln<bf16, bf16>(d_ids, static_cast<const bf16*>(e->w.d_tok_embeddings), e->w.d_emb_norm, h,
                                       static_cast<bf16*>(x), T, H, c.vocab_size, c.norm_eps, stream);
    }
    if (rc) return rc;
    // LayerNorm in the residual GEMMs' epilogue: CTA-pair kernel (H % 256 == 0) and enough 256-row blocks that
    // giving every CTA pair whole row blocks leaves no pair idle
    const bool ln_rows_ok = e->gemm_pair && (T + 255) / 256 >= static_cast<int64_t>(g_num_sms / 2);
    const bool ln_fused = t_opt.ln_fuse == 1 && ln_rows_ok && H % 256 == 0 && H <= 1024;
    // ln_fuse = 2: full-row residual GEMM with the LayerNorm from TMEM (gemm_rowln.cuh).  attn.Wo for H = 256 / 512;
    // mlp.Wo only for H = 256, where the second accumulator keeps the MMAs running during the LayerNorm passes.
    const int64_t ln_min_blocks = t_opt.ln_fuse_min_blocks < 0 ? (H == 256 ? 0 : 48) : t_opt.ln_fuse_min_blocks;
    const bool ln_row_wo = t_opt.ln_fuse == 2 && e->gemm_pair && (T + 255) / 256 >= ln_min_blocks && (H == 256 || H == 512);
    const bool ln_row_wo2 = ln_row_wo && H == 256;
    for (int l = 0; l < L; ++l) {
      const opv_layer_weights& lw = e->layers[l];
      const bool global = c.layer_is_global[l] != 0;
      if (l > 0 && !ln_fused && !ln_row_wo2) {
        LaunchScope sc(e, stream, OPV_PROF_LAYERNORM);
        rc = launch_layernorm<bf16>(h, lw.d_attn_norm, static_cast<bf16*>(x), T, H, c.norm_eps, stream);
      }
      if (rc) return rc;
      opv::GemmEpilogueArgs ep{};
      ep.pos = pos, ep.rope_cols = 2 * H, ep.rope_rows = c.max_positions;
      ep.cos = global ? e->w.d_rope_cos_global : e->w.d_rope_cos_local;
      ep.sin = global ? e->w.d_rope_sin_global : e->w.d_rope_sin_local;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_QKV);
        rc = gemm_bf16(fused ? opv::kEpiRope : opv::kEpiStore, e->gemm_pair, tm_x, e->tm_wqkv[l], tm_qkv, ep, T, 3 * H, H, stream);
      }
      if (rc) return rc;
      if (!fused) {
        LaunchScope sc(e, stream, OPV_PROF_MISC);
        opv::rope_inplace_kernel<bf16><<<g_num_sms * 8, 256, 0, stream>>>(static_cast<bf16*>(qkv), pos, ep.cos, ep.sin, T, H,
                                                                          c.max_positions);
        OPV_LAUNCH_CHECK("rope_inplace_kernel");
      }
      {
        LaunchScope sc(e, stream, global ? OPV_PROF_ATTN_GLOBAL : OPV_PROF_ATTN_LOCAL);
        rc = launch_attention(OPV_DTYPE_BF16, qkv, attn, d_cu_seqlens, n_seqs, max_seqlen, heads,
                              global ? -1 : half_window, &tm_qkv, &tm_qkv64, stream);
      }
      if (rc) return rc;
      opv::GemmEpilogueArgs er{};
      opv::GemmEpilogueArgs eln{};  // RESIDUAL_LN: the LayerNorm that follows the projection runs in its epilogue
      eln.ln_r = h, eln.ln_x = static_cast<bf16*>(x), eln.ln_eps = c.norm_eps;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WO);
        eln.ln_w = lw.d_mlp_norm;
        if (ln_row_wo)
          rc = launch_gemm_rowln(tm_attn, e->tm_wo[l], tm_h, tm_x, lw.d_mlp_norm, c.norm_eps, T, H, H, stream);
        else
          rc = ln_fused ? gemm_bf16(opv::kEpiResidualLn, true, tm_attn, e->tm_wo[l], tm_h, eln, T, H, H, stream)
                        : gemm_bf16(opv::kEpiResidual, e->gemm_pair, tm_attn, e->tm_wo[l], tm_h, er, T, H, H, stream);
      }
      if (rc) return rc;
      if (!ln_fused && !ln_row_wo) {
        LaunchScope sc(e, stream, OPV_PROF_LAYERNORM);
        rc = launch_layernorm<bf16>(h, lw.d_mlp_norm, static_cast<bf16*>(x), T, H, c.norm_eps, stream);
      }
      if (rc) return rc;
      if (fused) {
        opv::GemmEpilogueArgs eg{};
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WI);
        rc = gemm_bf16(opv::kEpiGeglu, e->gemm_pair, tm_x, e->tm_wi[l], tm_act, eg, T, 2 * I, H, stream);
      } else {
        opv::GemmEpilogueArgs es{};
        {
          LaunchScope sc(e, stream, OPV_PROF_GEMM_WI);
          rc = gemm_bf16(opv::kEpiStore, e->gemm_pair, tm_x, e->tm_wi[l], tm_u, es, T, 2 * I, H, stream);
        }
        if (rc) return rc;
        LaunchScope sc(e, stream, OPV_PROF_MISC);
        opv::geglu_kernel<bf16><<<g_num_sms * 8, 256, 0, stream>>>(static_cast<const bf16*>(u), static_cast<bf16*>(act), T, I);
        OPV_LAUNCH_CHECK("geglu_kernel");
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WO2);
        if (ln_row_wo2 && l + 1 < L) {
          rc = launch_gemm_rowln(tm_act, e->tm_wo2[l], tm_h, tm_x, e->layers[l + 1].d_attn_norm, c.norm_eps, T, H, I, stream);
        } else if (ln_fused && l + 1 < L) {  // the next layer's attn_norm; the final norm stays with the prune head kernel
          eln.ln_w = e->layers[l + 1].d_attn_norm;
          rc = gemm_bf16(opv::kEpiResidualLn, true, tm_act, e->tm_wo2[l], tm_h, eln, T, H, I, stream);
        } else {
          rc = gemm_bf16(opv::kEpiResidual, e->gemm_pair, tm_act, e->tm_wo2[l], tm_h, er, T, H, I, stream);
        }
      }
      if (rc) return rc;
    }
  } else {
    // fp32 activations.  OPV_DTYPE_F32: FFMA GEMMs.  OPV_DTYPE_F32_TC: the SAME launch sequence, but every projection is
    // six passes of the product's tcgen05 GEMM (RESIDUAL epilogue = fp32 TMA reduce-add) over 3-way bf16 splits.
    const bool tc = c.dtype == OPV_DTYPE_F32_TC;
    __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(ws + wl.split);
    auto gemm32 = [&](bool accumulate, const float* a, const float* w_f32, const CUtensorMap* w_planes, float* out_c,
                      int N, int K) -> int {
      if (!tc) return gemm_f32(accumulate, a, w_f32, out_c, T, N, K, N, stream);
      const int64_t n_elems = T * K;
      opv::split3_bf16_kernel<<<g_num_sms * 8, 256, 0, stream>>>(a, planes, n_elems);
      OPV_LAUNCH_CHECK("split3_bf16_kernel");
      if (!accumulate) OPV_CUDA(cudaMemsetAsync(out_c, 0, static_cast<size_t>(T) * N * sizeof(float), stream));
      CUtensorMap tm_a[3], tm_c;
      for (int p = 0; p < 3; ++p)
        if (int r = make_tmap_bf16(&tm_a[p], planes + p * n_elems, T, K, opv::kGemmBlockM)) return r;
      if (int r = make_tmap_2d(&tm_c, out_c, true, T, N, opv::kGemmBlockM)) return r;
      opv::GemmEpilogueArgs er{};
      // smallest terms first: mid.mid, hi.lo, lo.hi, hi.mid, mid.hi, hi.hi  (a plane, w plane)
      static const int order[6][2] = {{1, 1}, {0, 2}, {2, 0}, {0, 1}, {1, 0}, {0, 0}};
      for (const auto& pq : order)
        if (int r = gemm_bf16(opv::kEpiResidual, e->gemm_pair, tm_a[pq[0]], w_planes[pq[1]], tm_c, er, T, N, K, stream)) return r;
      e->launches += 6;  // + the split (the scope counts one)
      return OPV_OK;
    };
    float* xf = static_cast<float*>(x);
    float* qf = static_cast<float*>(qkv);
    float* af = static_cast<float*>(attn);
    float* actf = static_cast<float*>(act);
    float* uf = static_cast<float*>(u);
    {
      LaunchScope sc(e, stream, OPV_PROF_EMBED);
      rc = launch_embed_ln<float, float>(d_ids, static_cast<const float*>(e->w.d_tok_embeddings), e->w.d_emb_norm, h,
                                         xf, T, H, c.vocab_size, c.norm_eps, stream);
    }
    if (rc) return rc;
    for (int l = 0; l < L; ++l) {
      const opv_layer_weights& lw = e->layers[l];
      const bool global = c.layer_is_global[l] != 0;
      if (l > 0) {
        LaunchScope sc(e, stream, OPV_PROF_LAYERNORM);
        rc = launch_layernorm<float>(h, lw.d_attn_norm, xf, T, H, c.norm_eps, stream);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_QKV);
        rc = gemm32(false, xf, static_cast<const float*>(lw.d_wqkv), tc ? &e->tm_wqkv[3 * l] : nullptr, qf, 3 * H, H);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_MISC);
        opv::rope_inplace_kernel<float><<<g_num_sms * 8, 256, 0, stream>>>(
            qf, pos, global ? e->w.d_rope_cos_global : e->w.d_rope_cos_local,
            global ? e->w.d_rope_sin_global : e->w.d_rope_sin_local, T, H, c.max_positions);
        OPV_LAUNCH_CHECK("rope_inplace_kernel");
      }
      {
        LaunchScope sc(e, stream, global ? OPV_PROF_ATTN_GLOBAL : OPV_PROF_ATTN_LOCAL);
        rc = launch_attention(OPV_DTYPE_F32, qf, af, d_cu_seqlens, n_seqs, max_seqlen, heads,
                              global ? -1 : half_window, nullptr, nullptr, stream);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WO);
        rc = gemm32(true, af, static_cast<const float*>(lw.d_wo), tc ? &e->tm_wo[3 * l] : nullptr, h, H, H);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_LAYERNORM);
        rc = launch_layernorm<float>(h, lw.d_mlp_norm, xf, T, H, c.norm_eps, stream);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WI);
        rc = gemm32(false, xf, static_cast<const float*>(lw.d_wi), tc ? &e->tm_wi[3 * l] : nullptr, uf, 2 * I, H);
      }
      if (rc) return rc;
      {
        LaunchScope sc(e, stream, OPV_PROF_MISC);
        opv::geglu_kernel<float><<<g_num_sms * 8, 256, 0, stream>>>(uf, actf, T, I);
        OPV_LAUNCH_CHECK("geglu_kernel");
      }
      {
        LaunchScope sc(e, stream, OPV_PROF_GEMM_WO2);
        rc = gemm32(true, actf, static_cast<const float*>(lw.d_wo2), tc ? &e->tm_wo2[3 * l] : nullptr, h, H, I);
      }
      if (rc) return rc;
    }
  }

  {
    LaunchScope sc(e, stream, OPV_PROF_HEADS, 4);
    rc = launch_final_prune(h, e->w.d_final_norm, e->w.d_prune_weight, e->w.d_prune_bias, d_prune_logits, T, H,
                            c.norm_eps, stream);
    if (rc) return rc;
    float* cls = reinterpret_cast<float*>(ws + wl.cls);
    float* y = reinterpret_cast<float*>(ws + wl.y);
    if (c.classifier_pooling) {
      float* partial = reinterpret_cast<float*>(ws + wl.pool);
      dim3 grid((max_seqlen + opv::kPoolChunkRows - 1) / opv::kPoolChunkRows, n_seqs);
      OPV_DISPATCH_VEC(H, opv::rank_head_mean_partial_kernel<VEC><<<grid, opv::kRowWarps * 32, 0, stream>>>(
                              h, d_cu_seqlens, e->w.d_final_norm, partial, c.norm_eps));
      OPV_LAUNCH_CHECK("rank_head_mean_partial_kernel");
      opv::rank_head_mean_finish_kernel<<<n_seqs, 128, 0, stream>>>(partial, d_cu_seqlens, cls, H);
      OPV_LAUNCH_CHECK("rank_head_mean_finish_kernel");
      e->launches += 1;  // one launch more than the "cls" head
    } else {
      opv::rank_head_cls_ln_kernel<<<n_seqs, 32, 0, stream>>>(h, d_cu_seqlens, e->w.d_final_norm, cls, H, c.norm_eps, T);
      OPV_LAUNCH_CHECK("rank_head_cls_ln_kernel");
    }
    {
      dim3 grid(H / opv::kRankFeatTile, (n_seqs + opv::kRankSeqTile - 1) / opv::kRankSeqTile);
      const size_t smem = static_cast<size_t>(opv::kRankSeqTile) * H * sizeof(float);
      OPV_DISPATCH_VEC(H, opv::rank_head_dense_kernel<VEC>
                           <<<grid, opv::kRankFeatTile * 32, smem, stream>>>(cls, e->w.d_head_dense, y, n_seqs));
      OPV_LAUNCH_CHECK("rank_head_dense_kernel");
    }
    opv::rank_head_out_kernel<<<n_seqs, 32, 0, stream>>>(y, e->w.d_head_norm, e->w.d_cls_weight, e->w.d_cls_bias,
                                                        d_rank_logits, H, c.num_labels, c.norm_eps);
    OPV_LAUNCH_CHECK("rank_head_out_kernel");
  }
  return OPV_OK;
}

int opv_profile_enable(opv_handle e, int32_t on) {
  if (!e) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_profile_enable: null engine");
  e->profiling = on != 0;
  return OPV_OK;
}

int opv_profile_collect(opv_handle e, float* h_ms, int32_t* h_launches, int32_t n_classes) {
  if (!e || !h_ms || !h_launches) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_profile_collect: null argument");
  for (int i = 0; i < n_classes; ++i) h_ms[i] = 0.f, h_launches[i] = 0;
  for (auto& ev : e->prof) {
    OPV_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0.f;
    OPV_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (ev.cls >= 0 && ev.cls < n_classes) h_ms[ev.cls] += ms, h_launches[ev.cls] += ev.n;
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  e->prof.clear();
  return OPV_OK;
}

int64_t opv_launch_count(opv_handle e) { return e ? e->launches : 0; }

int opv_fragment_means(const float* d_prune_logits, int64_t n_tokens, const int32_t* d_frag_ranges, int32_t n_frags,
                       float* d_frag_mean, const float* d_rank_logits, int32_t n_seqs, int32_t num_labels,
                       float* d_rank_score, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_frags > 0) {
    if (!d_prune_logits || !d_frag_ranges || !d_frag_mean)
      return fail(OPV_ERR_INVALID_ARGUMENT, "opv_fragment_means: null buffer");
    opv::fragment_mean_kernel<<<(n_frags + opv::kRowWarps - 1) / opv::kRowWarps, opv::kRowWarps * 32, 0, stream>>>(
        d_prune_logits, n_tokens, d_frag_ranges, n_frags, d_frag_mean);
    OPV_LAUNCH_CHECK("fragment_mean_kernel");
  }
  if (n_seqs > 0 && d_rank_score) {
    if (!d_rank_logits || num_labels < 1) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_fragment_means: rank logits");
    opv::rank_score_kernel<<<(n_seqs + 255) / 256, 256, 0, stream>>>(d_rank_logits, n_seqs, num_labels, d_rank_score);
    OPV_LAUNCH_CHECK("rank_score_kernel");
  }
  return OPV_OK;
}

int opv_token_keep_probs(const float* d_prune_logits, int64_t n_tokens, float* d_keep_prob, void* stream_) {
  if (n_tokens <= 0) return OPV_OK;
  if (!d_prune_logits || !d_keep_prob) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_token_keep_probs: null buffer");
  const int64_t blocks = (n_tokens + 255) / 256;
  opv::token_keep_prob_kernel<<<static_cast<unsigned>(blocks < 4736 ? blocks : 4736), 256, 0,
                                static_cast<cudaStream_t>(stream_)>>>(d_prune_logits, n_tokens, d_keep_prob);
  OPV_LAUNCH_CHECK("token_keep_prob_kernel");
  return OPV_OK;
}

int opv_sentence_prune(const float* d_frag_mean, const int32_t* d_sent_offsets, const int32_t* d_sent_frag_index,
                       int32_t n_sents, double threshold, double guard, double* d_sent_prob, uint8_t* d_keep,
                       uint8_t* d_near, void* stream_) {
  if (n_sents <= 0) return OPV_OK;
  if (!d_frag_mean || !d_sent_offsets || !d_sent_frag_index || !d_sent_prob || !d_keep || !d_near)
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_sentence_prune: null buffer");
  opv::sentence_prune_kernel<<<(n_sents + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      d_frag_mean, d_sent_offsets, d_sent_frag_index, n_sents, threshold, guard, d_sent_prob, d_keep, d_near);
  OPV_LAUNCH_CHECK("sentence_prune_kernel");
  return OPV_OK;
}

// ---- single-op entry points -----------------------------------------------------------------------

int opv_forward_status(opv_handle e, const void* d_workspace, int32_t n_seqs, int64_t n_tokens, int32_t* n_bad,
                       void* stream_) {
  if (!e || !d_workspace || !n_bad) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_forward_status: null argument");
  *n_bad = 0;
  if (n_seqs <= 0 || n_tokens <= 0) return OPV_OK;
  DeviceGuard guard;
  OPV_CUDA(guard.enter(e->device));
  const WorkspaceLayout wl = workspace_layout(e, n_tokens, n_seqs);
  const uint8_t* ws = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 1023) & ~uintptr_t(1023));
  std::vector<int32_t> status(static_cast<size_t>(n_seqs));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OPV_CUDA(cudaMemcpyAsync(status.data(), ws + wl.status, status.size() * 4, cudaMemcpyDeviceToHost, stream));
  OPV_CUDA(cudaStreamSynchronize(stream));
  int bad = 0;
  for (int32_t v : status) bad += v != 0;
  *n_bad = bad;
  return OPV_OK;
}

int opv_op_gemm(int32_t dtype, int32_t epilogue, const void* d_a, const void* d_w, void* d_c, int64_t m, int32_t n,
                int32_t k, const int32_t* d_pos, const float* d_cos, const float* d_sin, int32_t hidden_size,
                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d_a || !d_w || !d_c) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_op_gemm: null buffer");
  if (int rc = ensure_device_setup()) return rc;
  t_opt = defaults_snapshot();
  if (dtype == OPV_DTYPE_F32) {
    if (epilogue == OPV_EPI_STORE)
      return gemm_f32(false, static_cast<const float*>(d_a), static_cast<const float*>(d_w), static_cast<float*>(d_c),
                      m, n, k, n, stream);
    if (epilogue == OPV_EPI_RESIDUAL)
      return gemm_f32(true, static_cast<const float*>(d_a), static_cast<const float*>(d_w), static_cast<float*>(d_c),
                      m, n, k, n, stream);
    return fail(OPV_ERR_UNSUPPORTED, "fp32 GEMM has STORE and RESIDUAL epilogues only");
  }
  if (m <= 0) return OPV_OK;
  const int bn = gemm_block_n(n, epilogue);
  if (bn == 0) return fail(OPV_ERR_UNSUPPORTED, "N = %d does not tile for epilogue %d", n, epilogue);
  CUtensorMap tm_a, tm_b, tm_c;
  if (int rc = make_tmap_bf16(&tm_a, d_a, m, k, opv::kGemmBlockM)) return rc;
  const bool pair = t_opt.gemm_pair != 0;
  if (int rc = make_tmap_bf16(&tm_b, d_w, n, k, weight_box_rows(bn, pair))) return rc;
  const bool c_f32 = epilogue == OPV_EPI_RESIDUAL;
  if (int rc = make_tmap_2d(&tm_c, d_c, c_f32, m, epilogue == OPV_EPI_GEGLU ? n / 2 : n, opv::kGemmBlockM)) return rc;
  opv::GemmEpilogueArgs ep{};
  ep.pos = d_pos, ep.cos = d_cos, ep.sin = d_sin, ep.rope_cols = 2 * hidden_size;
  if (epilogue == OPV_EPI_ROPE && (!d_pos || !d_cos || !d_sin || n != 3 * hidden_size))
    return fail(OPV_ERR_INVALID_ARGUMENT, "ROPE epilogue needs pos/cos/sin and N == 3 * hidden_size");
  return gemm_bf16(epilogue, pair, tm_a, tm_b, tm_c, ep, m, n, k, stream);
}

int opv_op_gemm_residual_ln(const void* d_a, const void* d_w, float* d_r, void* d_x, const float* d_ln_w, float eps,
                            int64_t m, int32_t n, int32_t k, void* stream_) {
  if (!d_a || !d_w || !d_r || !d_x || !d_ln_w) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_op_gemm_residual_ln: null buffer");
  if (int rc = ensure_device_setup()) return rc;
  t_opt = defaults_snapshot();
  if (m <= 0) return OPV_OK;
  CUtensorMap tm_a, tm_b, tm_r, tm_x;
  if (int rc = make_tmap_bf16(&tm_a, d_a, m, k, opv::kGemmBlockM)) return rc;
  if (int rc = make_tmap_bf16(&tm_b, d_w, n, k, 128)) return rc;  // half of the 256-row W tile per CTA
  if (int rc = make_tmap_2d(&tm_r, d_r, true, m, n, opv::kGemmBlockM)) return rc;
  if (int rc = make_tmap_bf16(&tm_x, d_x, m, n, opv::kGemmBlockM)) return rc;
  return launch_gemm_rowln(tm_a, tm_b, tm_r, tm_x, d_ln_w, eps, m, n, k, static_cast<cudaStream_t>(stream_));
}

int opv_op_layernorm(int32_t dtype, const float* d_h, const float* d_w, void* d_out, int64_t m, int32_t hidden,
                     float eps, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (int rc = ensure_device_setup()) return rc;
  t_opt = defaults_snapshot();
  if (dtype == OPV_DTYPE_BF16) return launch_layernorm<__nv_bfloat16>(d_h, d_w, static_cast<__nv_bfloat16*>(d_out), m, hidden, eps, s);
  return launch_layernorm<float>(d_h, d_w, static_cast<float*>(d_out), m, hidden, eps, s);
}

int opv_op_embed_ln(int32_t dtype, const int32_t* d_ids, const void* d_emb, const float* d_w, float* d_h, void* d_x,
                    int64_t m, int32_t hidden, int32_t vocab, float eps, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (int rc = ensure_device_setup()) return rc;
  t_opt = defaults_snapshot();
  if (dtype == OPV_DTYPE_BF16)
    return launch_embed_ln<__nv_bfloat16, __nv_bfloat16>(d_ids, static_cast<const __nv_bfloat16*>(d_emb), d_w, d_h,
                                                         static_cast<__nv_bfloat16*>(d_x), m, hidden, vocab, eps, s);
  return launch_embed_ln<float, float>(d_ids, static_cast<const float*>(d_emb), d_w, d_h, static_cast<float*>(d_x), m,
                                       hidden, vocab, eps, s);
}

int opv_op_attention(int32_t dtype, const void* d_qkv, void* d_out, const int32_t* d_cu_seqlens, int32_t n_seqs,
                     int64_t n_tokens, int32_t max_seqlen, int32_t num_heads, int32_t half_window, void* stream_) {
  if (!d_qkv || !d_out || !d_cu_seqlens) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_op_attention: null buffer");
  if (int rc = ensure_device_setup()) return rc;
  t_opt = defaults_snapshot();
  CUtensorMap tm_qkv, tm_qkv64;
  const bool tc = dtype == OPV_DTYPE_BF16;
  if (tc) {
    if (n_tokens <= 0) return OPV_OK;
    if (int rc = make_tmap_bf16(&tm_qkv, d_qkv, n_tokens, 3 * num_heads * 64, opv::kFaBlockM)) return rc;
    if (int rc = make_tmap_bf16(&tm_qkv64, d_qkv, n_tokens, 3 * num_heads * 64, opv::kQ4BlockN)) return rc;
  }
  return launch_attention(dtype, d_qkv, d_out, d_cu_seqlens, n_seqs, max_seqlen, num_heads, half_window,
                          tc ? &tm_qkv : nullptr, tc ? &tm_qkv64 : nullptr, static_cast<cudaStream_t>(stream_));
}

int opv_set_option(const char* name, int64_t value) {
  if (!name) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_set_option: null name");
  std::lock_guard<std::mutex> lock(g_mu);
  return set_option_in(g_defaults, name, value);
}

int opv_engine_set_option(opv_handle e, const char* name, int64_t value) {
  if (!e || !name) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_engine_set_option: null argument");
  if (strcmp(name, "gemm_pair") == 0)
    return fail(OPV_ERR_INVALID_ARGUMENT, "gemm_pair is fixed at opv_create (the weight tensor maps depend on it)");
  std::lock_guard<std::mutex> lock(e->mu);
  return set_option_in(e->opt, name, value);
}

int opv_op_rope(int32_t dtype, void* d_qkv, const int32_t* d_pos, const float* d_cos, const float* d_sin, int64_t m,
                int32_t hidden, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (m <= 0) return OPV_OK;
  if (dtype == OPV_DTYPE_BF16)
    opv::rope_inplace_kernel<__nv_bfloat16><<<1184, 256, 0, s>>>(static_cast<__nv_bfloat16*>(d_qkv), d_pos, d_cos, d_sin, m, hidden);
  else
    opv::rope_inplace_kernel<float><<<1184, 256, 0, s>>>(static_cast<float*>(d_qkv), d_pos, d_cos, d_sin, m, hidden);
  OPV_LAUNCH_CHECK("rope_inplace_kernel");
  return OPV_OK;
}

int opv_op_geglu(int32_t dtype, const void* d_u, void* d_act, int64_t m, int32_t intermediate, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (m <= 0) return OPV_OK;
  if (dtype == OPV_DTYPE_BF16)
    opv::geglu_kernel<__nv_bfloat16><<<1184, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(d_u), static_cast<__nv_bfloat16*>(d_act), m, intermediate);
  else
    opv::geglu_kernel<float><<<1184, 256, 0, s>>>(static_cast<const float*>(d_u), static_cast<float*>(d_act), m, intermediate);
  OPV_LAUNCH_CHECK("geglu_kernel");
  return OPV_OK;
}

int opv_op_positions(const int32_t* d_cu_seqlens, int32_t n_seqs, int32_t* d_pos, void* stream_) {
  if (n_seqs <= 0) return OPV_OK;
  opv::positions_kernel<<<n_seqs, 256, 0, static_cast<cudaStream_t>(stream_)>>>(d_cu_seqlens, d_pos);
  OPV_LAUNCH_CHECK("positions_kernel");
  return OPV_OK;
}

}  // extern "C"
