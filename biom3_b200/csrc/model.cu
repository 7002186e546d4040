// C-ABI implementation: model handle, strict weight loading, per-step forward, persistent decode loop.
// See include/biom3_b200.h for the contract and the reference interfaces each entry point replaces.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/biom3_b200.h"
#include "attention.cuh"
#include "fp32_path.cuh"
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_OK(expr)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return fail(BIOM3_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                      std::to_string(__LINE__) + ")");                               \
  } while (0)

using bf16 = __nv_bfloat16;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// bf16 row-major [rows][cols] -> tiles of [box_rows][64 cols], 128-byte swizzle
// bf16 row-major [rows][cols] -> store boxes of [32 rows][32 cols] (64 bytes), 64-byte swizzle
int make_store_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(bf16)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled (store) failed: " + std::to_string(int(r)));
  return BIOM3_OK;
}

// bf16 [rows][32] (64-byte rows, e.g. the head-major qkv buffer) -> tiles of [128 rows][32], 64-byte swizzle
int make_tmap_sw64(CUtensorMap* tm, const void* base, uint64_t rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {32, rows};
  cuuint64_t strides[1] = {32 * sizeof(bf16)};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled (sw64) failed: " + std::to_string(int(r)));
  return BIOM3_OK;
}

int make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(bf16)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BIOM3_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(int(r)));
  return BIOM3_OK;
}

// Programmatic dependent launch (see ptx::pdl_sync): set by run_step for the launches of one decode step.
thread_local bool g_pdl = false;

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute when g_pdl is set
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Operand-ring depth.  Pair tiling, 256-wide tiles: 5 stages of 32 KB (3, 4 and 6 were measured: profiles/r01_ab_resid_ring_depth.jsonl;
// a 6-stage QKV ring again in round 2: no change)
#ifndef BIOM3_RESID_STAGES
#define BIOM3_RESID_STAGES 3
#endif
template <int BN, bool CG2, int EPI>
constexpr int gemm_stages() {
  // epilogue 6 spends 96 KB on its residual slots; it serves the K = 512 out-projection, whose mainloop is far from
  // binding (3 stages measured equal to 5 for that GEMM in round 1)
  if (CG2 && BN == 256 && EPI == gemm::EPI_BIAS_RESID_SPLIT_TMA) return BIOM3_RESID_STAGES;
  if (CG2 && BN == 256 && EPI == gemm::EPI_BIAS_GELU_SPLIT) return 4;      // two staging blocks per epilogue warp
  if (CG2 && BN == 256) return 5;
  return CG2 ? 7 : (BN == 256 ? 3 : 5);
}

// launch only; the caller checks cudaGetLastError()
template <int BN, int EPI, bool CG2>
void launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                   const gemm::Params& p_in, int num_sms, cudaStream_t st) {
  gemm::Params p = p_in;
  gemm::fill_shifts(p, BN);
  constexpr int STAGES = gemm_stages<BN, CG2, EPI>();
  const int smem = gemm::SmemLayout<BN, STAGES, CG2, EPI>::TOTAL;
  const int tiles = (p.M / (CG2 ? 256 : 128)) * (p.N / BN);
  const int workers = CG2 ? num_sms / 2 : num_sms;
  const int grid = (tiles < workers ? tiles : workers) * (CG2 ? 2 : 1);
  const int threads = 32 * (int(gemm::EPI_WARP0) + gemm::epi_warps(EPI));
  if constexpr (CG2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 2 : 1;
    cudaLaunchKernelEx(&cfg, gemm::gemm_bf16_tcgen05<BN, STAGES, EPI, true>, ta, tb, tc, td, p);
  } else {
    launch_k(gemm::gemm_bf16_tcgen05<BN, STAGES, EPI, false>, dim3(grid), dim3(threads), size_t(smem), st, ta, tb, tc, td, p);
  }
}

// bn: 128 or 256 columns per tile.  pair: CTA-pair (cta_group::2) tiling, needs bn == 256 and M % 256 == 0;
// `tb` must then be the 128-row-box weight map (each CTA stages half of the 256 weight rows).
template <int EPI>
void launch_gemm(int bn, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                 const gemm::Params& p, int num_sms, cudaStream_t st) {
  (void)bn;                                     // tiles are 256 columns wide (128-wide tiles were measured slower in round 1)
  if (pair) launch_gemm_t<256, EPI, true>(ta, tb, tc, tc, p, num_sms, st);
  else launch_gemm_t<256, EPI, false>(ta, tb, tc, tc, p, num_sms, st);
}

constexpr int HEAD_SMEM_MAX = 32 * 1024 * 4;   // num_classes <= 32, dim <= 1024, fp32

// opt in to large dynamic shared memory for every kernel once per process, outside any stream capture
cudaError_t init_kernel_attributes_impl() {
  cudaError_t e;
#define SET_GEMM1(BN, EPI, CG2)                                                                                  \
  e = cudaFuncSetAttribute(gemm::gemm_bf16_tcgen05<BN, gemm_stages<BN, CG2, EPI>(), EPI, CG2>,                   \
                           cudaFuncAttributeMaxDynamicSharedMemorySize,                                          \
                           gemm::SmemLayout<BN, gemm_stages<BN, CG2, EPI>(), CG2, EPI>::TOTAL);                  \
  if (e != cudaSuccess) return e;
#define SET_GEMM(EPI) SET_GEMM1(256, EPI, false) SET_GEMM1(256, EPI, true)
  SET_GEMM(gemm::EPI_QKV_HEADMAJOR)
  SET_GEMM(gemm::EPI_BIAS_RESID_F32)
  SET_GEMM(gemm::EPI_BIAS_RESID_SPLIT)
  SET_GEMM(gemm::EPI_BIAS_GELU_BF16)
  SET_GEMM(gemm::EPI_STORE_BF16)
  SET_GEMM(gemm::EPI_STORE_F32)
  SET_GEMM1(256, gemm::EPI_BIAS_RESID_SPLIT_TMA, true)      // pair tiling only
  SET_GEMM(gemm::EPI_BIAS_GELU_SPLIT)
#undef SET_GEMM1
#undef SET_GEMM
  e = cudaFuncSetAttribute(attn::local_attention_kernel<attn::LOCAL_NST, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           attn::LocalCfg<attn::LOCAL_NST>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attn::local_attention_kernel<attn::LOCAL_NST, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           attn::LocalCfg<attn::LOCAL_NST>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attn::local_attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::L3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attn::linear_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::LIN_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(attn::linear_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::LIN_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(f32p::local_attention_f32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f32p::LAT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(f32p::local_attention_f32_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f32p::LAF_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(f32p::linear_attention_f32_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f32p::LINF_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k::head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM_MAX);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k::sample_all_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           2 * 32 * k::SAMPLE_TILE * int(sizeof(float)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k::random_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12);
  return e;
}

// cudaFuncSetAttribute applies to the CURRENT device only, so the opt-ins are tracked per device (a process may hold
// engines on several GPUs); a failure is returned to the caller and not cached, so a later call retries.
cudaError_t init_kernel_attributes() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  e = init_kernel_attributes_impl();
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

// windowed attention (attention.cuh): persistent, one CTA per SM; trace = 1 records CTA 0's clock64 timeline
int g_attn3 = 1;                   // BIOM3_ATTN3=0: the two-stream windowed-attention kernel (the three-stream one is 7 % faster)
void launch_local(const CUtensorMap& tm, bf16* out, int B, int H, int L, int NL, float scale_log2e, int reverse, int num_sms,
                  int trace, cudaStream_t st) {
  const int grid = std::min(num_sms, (L / attn::WIN) * NL * B);
  if (g_attn3 && !trace) {
    launch_k(attn::local_attention3_kernel, dim3(grid), dim3(attn::L3_THREADS), size_t(attn::L3_SMEM_BYTES), st, tm, out, B, H, L, NL,
             scale_log2e, reverse);
    return;
  }
  constexpr size_t smem = attn::LocalCfg<attn::LOCAL_NST>::SMEM_BYTES;
  if (trace)
    launch_k(attn::local_attention_kernel<attn::LOCAL_NST, 1>, dim3(grid), dim3(attn::LOCAL_THREADS), smem, st, tm, out, B, H, L, NL,
             scale_log2e, reverse);
  else
    launch_k(attn::local_attention_kernel<attn::LOCAL_NST, 0>, dim3(grid), dim3(attn::LOCAL_THREADS), smem, st, tm, out, B, H, L, NL,
             scale_log2e, reverse);
}

enum Cat { C_QKV, C_OUT, C_FF1, C_FF2, C_LOCAL, C_LINEAR, C_LN, C_EMBED, C_HEAD, C_OTHER, C_COUNT };

struct Profiler {
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  cudaStream_t st;
  void begin(int c) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    ev.push_back(a);
    ev.push_back(b);
    cat.push_back(c);
    cudaEventRecord(a, st);
  }
  void end() { cudaEventRecord(ev.back(), st); }
};

}  // namespace

struct biom3_model {
  biom3_config cfg{};
  int device = 0, max_batch = 0, num_sms = 0;
  int bn_wide = 256, bn_narrow = 256;          // tile widths for the N = 3D/4D and N = D GEMMs
  bool use_pair = true;                         // CTA-pair (cta_group::2) GEMM tiling when M % 256 == 0
  bool finalized = false;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;   // as passed to biom3_set_weight (error messages)
  // weights
  bf16 *Wqkv = nullptr, *Wo = nullptr, *W1 = nullptr, *W2 = nullptr;
  float *ln_s_qkv = nullptr, *ln_t_qkv = nullptr, *ln_s_ff = nullptr, *ln_t_ff = nullptr;   // folded LayerNorms
  float* stats = nullptr;                       // [M][ln_parts][2] row partial (sum, sumsq)
  int ln_parts = 2;
  float *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *emb = nullptr, *ax0 = nullptr, *ax1 = nullptr, *norm_g = nullptr, *norm_b = nullptr;
  float *w_out = nullptr, *b_out = nullptr;
  float *Ttab = nullptr;
  float *y_w0 = nullptr, *y_b0 = nullptr, *y_w2 = nullptr, *y_b2 = nullptr;
  // workspace
  float* u = nullptr;                           // fp32 residual stream (fp32-class mode, or BIOM3_SPLIT_RESID=0)
  bf16* u_lo = nullptr;                         // split residual stream: u = a (hi, bf16) + u_lo (bf16)
  bool split_resid = true;
  bf16 *a = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr;
  float *Yh = nullptr, *Ytmp = nullptr, *Y = nullptr, *cvec = nullptr;
  uint8_t* state = nullptr;
  int *inv_path = nullptr, *t_i32 = nullptr;
  int* err_flags = nullptr;                     // sticky input-range bits set by the load kernels (biom3_input_errors)
  k::DecodeCtl* ctl = nullptr;
  unsigned long long* stamps = nullptr;         // [L][2] per-step %globaltimer stamps of the last decode (debug_copy "stamps")
  CUtensorMap tm_a{}, tm_att{}, tm_hid{};
  CUtensorMap tm_qkv_attn{};                   // qkv as [3*B*H*L][32], 128-row boxes, 64B swizzle (tcgen05 attention loads)
  CUtensorMap tm_st_qkv{}, tm_st_hid{};        // TMA-store maps: qkv as [3*B*H*L][32], hid as [M][4D]
  bool use_pdl = true;                          // programmatic dependent launch between the kernels of a step
  bool serpentine = true;                       // alternate the row walking direction kernel to kernel (L2 reuse)
  // Last-layer row compaction (decode, split residual): after the last block's attention only the B * group token
  // rows the sampler consumes are carried through out-proj / MLP / head (k::gather_rows_kernel).  BIOM3_COMPACT=0
  // computes every row like the reference does.
  bool compact_last = true;
  unsigned short qkv_chunk[gemm::QKV_MAX_CHUNKS] = {};   // gemm::Params::qkv_chunk: output order of the QKV GEMM, set by finalize
  bool qsoft_epi = true;        // softmax(q) of the linear-attention heads in the QKV GEMM epilogue (BIOM3_QSOFT_EPI=0: in the attention kernel)
  int compact_rows_max = 0;                     // rows allocated for the compact buffers (multiple of 256), 0 = none
  int last_compact_rows = 0;                    // rows the last run_step() carried through the last layer's MLP (0 = all)
  bf16 *att_c = nullptr, *a_c = nullptr, *ulo_c = nullptr, *hid_c = nullptr;
  float* stats_c = nullptr;
  CUtensorMap tm_att_c{}, tm_a_c{}, tm_hid_c{}, tm_st_hid_c{};
  CUtensorMap tm_wqkv[2]{}, tm_wo[2]{}, tm_w1[2]{}, tm_w2[2]{};   // [0]: box 128 rows, [1]: box 256 rows
  // Residual planes as 32 x 32 TMA boxes (epilogue 6).  resid_tma (BIOM3_RESID_TMA): 0 = epilogue 5 everywhere, 1 = the
  // out-projection (K = 512, bound by its residual epilogue) uses epilogue 6, 2 = FF2 as well (A/B only: FF2 is bound
  // by its K = 2048 mainloop and wants the 5-stage operand ring epilogue 6 has no room for)
  int resid_tma = 1;
  CUtensorMap tm_rs_hi{}, tm_rs_lo{}, tm_rs_hi_c{}, tm_rs_lo_c{};
  // step graph cache
  cudaStream_t cap_stream = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_B = -1, graph_group = -1;
  // forward API: the ~100 launches of one forward replayed as a graph too (B = 1 is launch bound: 1.64 ms un-graphed); the graph
  // writes logits into an internal buffer that is then copied to the caller's pointer.  BIOM3_FWD_GRAPH=0: plain launches.
  bool fwd_graph = true;
  cudaGraphExec_t fwd_exec = nullptr;
  int fwd_B = -1;
  float* logits_buf = nullptr;                  // [max_batch][C][L], allocated at the first forward
  std::vector<void*> allocs;
  int launches_per_step = 0;
  // fp32-class mode (biom3_set_precision(m, 1) before finalize): split [hi | lo] weights and fp32 activations
  int precision = 0;
  bool f32_fused_gelu = true;                   // BIOM3_F32_FUSED_GELU=0: fp32 hidden activation + a separate bias / GELU / split pass
  CUtensorMap tm_st_hid2{};                     // TMA-store map of hid2 [M][8D] (gemm::EPI_BIAS_GELU_SPLIT)
  int f32_attn_mma = 2;                         // BIOM3_F32_ATTN_MMA: 2 = windowed attention on tcgen05 + linear on mma.sync, 1 = both on mma.sync,
                                                // 0 = the CUDA-core fp32 attention kernels of round 1
  bf16 *Wqkv2 = nullptr, *Wo2 = nullptr, *W1s = nullptr, *W2s = nullptr;       // [N][2K] per layer, stacked
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;  // [depth][D]
  bf16 *a2 = nullptr, *hid2 = nullptr;                                            // [M][2D], [M][8D]
  float *qkv32 = nullptr, *hid32 = nullptr;                                       // [M][3D], [M][4D]
  CUtensorMap tm_a2{}, tm_hid2{};
  CUtensorMap tm_wqkv2[2]{}, tm_wo2[2]{}, tm_w1s[2]{}, tm_w2s[2]{};
};

struct biom3_facilitator_t {
  int device = 0, in_dim = 0, hid_dim = 0, out_dim = 0, hid_rows = 0;
  float *w0 = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr, *hid = nullptr;   // weight norm folded
};

namespace {

template <typename T>
int dev_alloc(biom3_model* m, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(BIOM3_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  m->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return BIOM3_OK;
}

int get_w(biom3_model* m, const std::string& key, size_t numel, const std::vector<float>** out) {
  auto it = m->host_w.find(key);
  if (it == m->host_w.end()) return fail(BIOM3_ERR_STATE, "missing weight: " + key);
  if (it->second.size() != numel)
  {
    std::string shp;
    for (int64_t d : m->host_shape[key]) shp += (shp.empty() ? "[" : ", ") + std::to_string(d);
    return fail(BIOM3_ERR_STATE, "size mismatch for " + key + ": got " + std::to_string(it->second.size()) + " elements " + shp +
                                     "], expected " + std::to_string(numel));
  }
  *out = &it->second;
  return BIOM3_OK;
}

int upload_f32(biom3_model* m, const std::string& key, size_t numel, float* dst) {
  const std::vector<float>* v;
  int r = get_w(m, key, numel, &v);
  if (r) return r;
  CU_OK(cudaMemcpy(dst, v->data(), numel * sizeof(float), cudaMemcpyHostToDevice));
  return BIOM3_OK;
}

int upload_bf16(biom3_model* m, const std::string& key, size_t numel, bf16* dst) {
  const std::vector<float>* v;
  int r = get_w(m, key, numel, &v);
  if (r) return r;
  std::vector<bf16> tmp(numel);
  for (size_t i = 0; i < numel; ++i) tmp[i] = __float2bfloat16_rn((*v)[i]);
  CU_OK(cudaMemcpy(dst, tmp.data(), numel * sizeof(bf16), cudaMemcpyHostToDevice));
  return BIOM3_OK;
}

// fp32-class mode: W fp32 [N][K] -> bf16 [N][2K] = [bf16(W) | bf16(W - bf16(W))]
int upload_split(biom3_model* m, const std::string& key, size_t N, size_t K, bf16* dst) {
  const std::vector<float>* v;
  int r = get_w(m, key, N * K, &v);
  if (r) return r;
  std::vector<bf16> tmp(N * 2 * K);
  for (size_t n = 0; n < N; ++n)
    for (size_t kk = 0; kk < K; ++kk) {
      const float w = (*v)[n * K + kk];
      const bf16 hi = __float2bfloat16_rn(w);
      tmp[n * 2 * K + kk] = hi;
      tmp[n * 2 * K + K + kk] = __float2bfloat16_rn(w - __bfloat162float(hi));
    }
  CU_OK(cudaMemcpy(dst, tmp.data(), tmp.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  return BIOM3_OK;
}

// LayerNorm folded into the following linear layer (see gemm_tcgen05.cuh): uploads W' = gamma (.) W as bf16
// [N][K], s[n] = sum_k bf16(W'[n][k]) and t[n] = sum_k beta[k] W[n][k] (+ bias[n] if bias_key is non-empty).
// t_scale: 0.5 for the GELU consumer, whose epilogue works on x / 2 (gemm::Params::ln_t); the halving is exact.
int upload_folded(biom3_model* m, const std::string& g_key, const std::string& b_key, const std::string& w_key,
                  const std::string& bias_key, size_t N, size_t K, bf16* w_dst, float* s_dst, float* t_dst,
                  float t_scale = 1.0f) {
  const std::vector<float>*g, *b, *w, *bias = nullptr;
  int r;
  if ((r = get_w(m, g_key, K, &g))) return r;
  if ((r = get_w(m, b_key, K, &b))) return r;
  if ((r = get_w(m, w_key, N * K, &w))) return r;
  if (!bias_key.empty() && (r = get_w(m, bias_key, N, &bias))) return r;
  std::vector<bf16> wf(N * K);
  std::vector<float> s(N), t(N);
  for (size_t n = 0; n < N; ++n) {
    double sa = 0.0, ta = 0.0;
    for (size_t kk = 0; kk < K; ++kk) {
      const float wv = (*w)[n * K + kk];
      const bf16 q = __float2bfloat16_rn((*g)[kk] * wv);
      wf[n * K + kk] = q;
      sa += double(__bfloat162float(q));
      ta += double((*b)[kk]) * double(wv);
    }
    s[n] = float(sa);
    t[n] = float(ta + (bias ? double((*bias)[n]) : 0.0)) * t_scale;
  }
  CU_OK(cudaMemcpy(w_dst, wf.data(), N * K * sizeof(bf16), cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(s_dst, s.data(), N * sizeof(float), cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(t_dst, t.data(), N * sizeof(float), cudaMemcpyHostToDevice));
  return BIOM3_OK;
}

// Output order of the QKV GEMM in 32-column chunks (one chunk = one head of q, k or v).  With the q softmax of the
// linear-attention heads done in the GEMM epilogue (qsoft_epi) those chunks cost the epilogue warps about twice the plain
// ones; in the natural order they fill a whole 256-column tile (heads NL.. of q are columns 256-511 at the stage3 shape)
// whose epilogue then outlasts the next tile's mainloop (profiles/r02_ab_qk_epilogue_unpermuted.jsonl: QKV +0.13 ms per
// step, all of the linear-attention kernel's gain).  The weight rows are permuted so that they are spread evenly
// (profiles/r02_ab_qk_epilogue.jsonl: QKV unchanged, step -0.07 ms); gemm::Params::qkv_chunk tells the epilogue what
// each chunk is and where it goes, so the qkv layout in memory does not change.
void build_qkv_chunk_order(biom3_model* m) {
  const int H = m->cfg.heads, NL = m->cfg.local_heads;
  std::vector<unsigned short> heavy, light;
  for (int which = 0; which < 3; ++which)
    for (int h = 0; h < H; ++h) {
      int kind = 0;
      if (m->precision == 0 && h >= NL && which == 0 && m->qsoft_epi) kind = 1;
      (kind ? heavy : light).push_back((unsigned short)(h | (which << 12) | (kind << 14)));
    }
  const size_t n = size_t(3) * H, nh = heavy.size();
  size_t ih = 0, il = 0;
  for (size_t i = 0; i < n; ++i)
    m->qkv_chunk[i] = ((i + 1) * nh / n > i * nh / n) ? heavy[ih++] : light[il++];
}

// rows of layer j's stacked [q | k | v] weight (and its folded-LayerNorm vectors) from the natural order into qkv_chunk's
int permute_qkv_chunks(biom3_model* m, size_t j) {
  const size_t D = m->cfg.dim, H = m->cfg.heads, n = 3 * H;
  bool identity = true;
  for (size_t i = 0; i < n; ++i) identity = identity && size_t(((m->qkv_chunk[i] >> 12) & 3) * H + (m->qkv_chunk[i] & 0xfff)) == i;
  if (identity) return BIOM3_OK;
  bf16* w = m->Wqkv + j * 3 * D * D;
  float* vec[2] = {m->ln_s_qkv + j * 3 * D, m->ln_t_qkv + j * 3 * D};
  bf16* wtmp = nullptr;
  float* vtmp = nullptr;
  CU_OK(cudaMalloc(&wtmp, 3 * D * D * sizeof(bf16)));
  CU_OK(cudaMalloc(&vtmp, 3 * D * sizeof(float)));
  CU_OK(cudaMemcpy(wtmp, w, 3 * D * D * sizeof(bf16), cudaMemcpyDeviceToDevice));
  for (size_t i = 0; i < n; ++i) {
    const size_t src = size_t((m->qkv_chunk[i] >> 12) & 3) * H + (m->qkv_chunk[i] & 0xfff);
    CU_OK(cudaMemcpy(w + i * 32 * D, wtmp + src * 32 * D, 32 * D * sizeof(bf16), cudaMemcpyDeviceToDevice));
  }
  for (float* v : vec) {
    CU_OK(cudaMemcpy(vtmp, v, 3 * D * sizeof(float), cudaMemcpyDeviceToDevice));
    for (size_t i = 0; i < n; ++i) {
      const size_t src = size_t((m->qkv_chunk[i] >> 12) & 3) * H + (m->qkv_chunk[i] & 0xfff);
      CU_OK(cudaMemcpy(v + i * 32, vtmp + src * 32, 32 * sizeof(float), cudaMemcpyDeviceToDevice));
    }
  }
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaFree(wtmp));
  CU_OK(cudaFree(vtmp));
  return BIOM3_OK;
}

// C = act(A W^T + b), fp32 (conditioning MLPs only)
void sgemm(const float* A, const float* W, const float* b, float* C, int M, int N, int K, int act, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  k::sgemm_bias_act_kernel<<<grid, 256, 0, st>>>(A, W, b, C, M, N, K, act);
}

// y_mlp(y_c) -> Y[b][j][d]
void run_y_mlp(biom3_model* m, const float* y_c, int B, cudaStream_t st) {
  const int D = m->cfg.dim, depth = m->cfg.depth, E = m->cfg.text_emb_dim;
  sgemm(y_c, m->y_w0, m->y_b0, m->Yh, B, 4 * D, E, 1, st);
  sgemm(m->Yh, m->y_w2, m->y_b2, m->Ytmp, B, D * depth, 4 * D, 0, st);
  const size_t n = size_t(B) * D * depth;
  k::cond_transpose_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(m->Ytmp, m->Y, B, D, depth);
}

// Residual update of the split (hi, lo) stream: gemm::EPI_BIAS_RESID_SPLIT, or, with `tma_resid` (pair tiling; the caller
// picks it per GEMM, see biom3_model::resid_tma), gemm::EPI_BIAS_RESID_SPLIT_TMA over the (hi, lo) plane maps `th`, `tl`.
void launch_resid_split(biom3_model* m, bool pair, bool tma_resid, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& th,
                        const CUtensorMap& tl, const gemm::Params& r, cudaStream_t st) {
  if (tma_resid && pair) launch_gemm_t<256, gemm::EPI_BIAS_RESID_SPLIT_TMA, true>(ta, tb, th, tl, r, m->num_sms, st);
  else launch_gemm<gemm::EPI_BIAS_RESID_SPLIT>(m->bn_narrow, pair, ta, tb, m->tm_st_hid, r, m->num_sms, st);
}

// One per-step forward over the resident state.  sample: draw + unmask (decode); else write logits.
// t_per_sample != nullptr -> forward API (arbitrary step per sample); else the device step counter.
cudaError_t run_step(biom3_model* m, int B, int group, const int* t_per_sample, float* logits_out, bool sample,
                     bool advance, cudaStream_t st, Profiler* prof, int* n_launch) {
  const biom3_config& c = m->cfg;
  const int D = c.dim, L = c.seq_len, H = c.heads, depth = c.depth, NL = c.local_heads, C = c.num_classes;
  const int M = B * L;
  const int JD = depth * D;
  int launches = 0;
  cudaError_t err = cudaSuccess;
  g_pdl = m->use_pdl && !prof && m->precision == 0;
#define LAUNCH(cat, ...)                     \
  do {                                       \
    if (prof) prof->begin(cat);              \
    __VA_ARGS__;                             \
    if (prof) prof->end();                   \
    ++launches;                              \
    err = cudaGetLastError();                \
    if (err != cudaSuccess) { g_pdl = false; return err; } \
  } while (0)

  const int row_blocks = std::min((M + 7) / 8, m->num_sms * 8);
  const bool split = m->precision == 0 && m->split_resid;      // residual stream stored as bf16 hi + lo
  LAUNCH(C_OTHER, launch_k(k::cond_build_kernel, dim3(std::max(1, JD / 4 / 256), B), dim3(256), 0, st,
                      m->Ttab, m->Y, t_per_sample, m->ctl, m->cvec, B, JD));
  LAUNCH(C_EMBED, launch_k(k::embed_kernel, dim3(row_blocks), dim3(256), 0, st, m->state, m->emb, m->ax0, m->ax1, m->cvec, JD, m->u, m->a, split ? m->u_lo : nullptr,
                                                              m->stats, m->ln_parts, M, L, c.local_window, D));
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(attn::DH));
  const float q_scale = 1.0f / sqrtf(float(attn::DH));
  const bool pw = m->use_pair && m->bn_wide == 256 && M % 256 == 0, pn = m->use_pair && m->bn_narrow == 256 && M % 256 == 0;
  const int iw = (m->bn_wide == 256 && !pw) ? 1 : 0, in = (m->bn_narrow == 256 && !pn) ? 1 : 0;   // weight map: 256- or 128-row box
  // last-layer row compaction: rows of the compact buffers, 0 = every row goes through the last layer's MLP
  int Mc = 0;
  if (sample && split && group > 0 && m->compact_rows_max > 0) {
    const int want = (B * group + 255) / 256 * 256;
    if (want <= m->compact_rows_max && want * 2 <= M) Mc = want;
  }
  m->last_compact_rows = Mc;
  int dir = 0;                                  // row walking direction of the next launch (see Params::reverse)
  auto next_dir = [&]() { const int d = dir; if (m->serpentine) dir ^= 1; return d; };
  next_dir();                                   // the embed kernel walked forward
  if (m->precision == 1) {
    // fp32-class layer loop (fp32_path.cuh): explicit fp32 LayerNorm / GELU / attention, bf16x3 split GEMMs
    const int ew_blocks = m->num_sms * 8;
    for (int j = 0; j < depth; ++j) {
      gemm::Params p{};
      p.L = L; p.H = H; p.Bsz = B; p.M = M; p.split3 = 1;
      LAUNCH(C_LN, f32p::ln_split_kernel<<<row_blocks, 256, 0, st>>>(m->u, m->ln1_g + size_t(j) * D, m->ln1_b + size_t(j) * D,
                                                                    m->a2, M, D));
      p.N = 3 * D; p.K = D; p.b_row_offset = j * 3 * D; p.out = m->qkv32; p.reverse = 0;
      LAUNCH(C_QKV, launch_gemm<gemm::EPI_STORE_F32>(m->bn_wide, pw, m->tm_a2, m->tm_wqkv2[iw], m->tm_st_hid, p, m->num_sms, st));
      if (H - NL > 0 && m->f32_attn_mma)
        LAUNCH(C_LINEAR, f32p::linear_attention_f32_mma_kernel<<<dim3(H - NL, B), 256, f32p::LINF_SMEM_BYTES, st>>>(m->qkv32, m->a2, B, H, L, NL, q_scale));
      else if (H - NL > 0)
        LAUNCH(C_LINEAR, f32p::linear_attention_f32_kernel<<<dim3(H - NL, B), 256, 0, st>>>(m->qkv32, m->a2, B, H, L, NL, q_scale));
      if (NL > 0 && m->f32_attn_mma >= 2)
        LAUNCH(C_LOCAL, f32p::local_attention_f32_tc_kernel<<<dim3(L / attn::WIN, NL, B), 128, f32p::LAT_SMEM_BYTES, st>>>(m->qkv32, m->a2, B, H, L, q_scale));
      else if (NL > 0 && m->f32_attn_mma)
        LAUNCH(C_LOCAL, f32p::local_attention_f32_mma_kernel<<<dim3(L / attn::WIN, NL, B), 256, f32p::LAF_SMEM_BYTES, st>>>(m->qkv32, m->a2, B, H, L, q_scale));
      else if (NL > 0)
        LAUNCH(C_LOCAL, f32p::local_attention_f32_kernel<<<dim3(L / attn::WIN, NL, B), 128, 0, st>>>(m->qkv32, m->a2, B, H, L, q_scale));
      gemm::Params r{};
      r.L = L; r.H = H; r.Bsz = B; r.M = M; r.split3 = 1;
      r.N = D; r.K = D; r.b_row_offset = j * D; r.out = m->u; r.bias = m->bo + size_t(j) * D;
      LAUNCH(C_OUT, launch_gemm<gemm::EPI_BIAS_RESID_F32>(m->bn_narrow, pn, m->tm_a2, m->tm_wo2[in], m->tm_st_hid, r, m->num_sms, st));
      LAUNCH(C_LN, f32p::ln_split_kernel<<<row_blocks, 256, 0, st>>>(m->u, m->ln2_g + size_t(j) * D, m->ln2_b + size_t(j) * D,
                                                                    m->a2, M, D));
      p.N = 4 * D; p.K = D; p.b_row_offset = j * 4 * D; p.out = m->hid32;
      if (m->f32_fused_gelu) {
        // bias + exact-erf GELU + [hi | lo] split in the GEMM epilogue: no fp32 hidden activation in HBM
        p.bias = m->b1 + size_t(j) * 4 * D;
        LAUNCH(C_FF1, launch_gemm<gemm::EPI_BIAS_GELU_SPLIT>(m->bn_wide, pw, m->tm_a2, m->tm_w1s[iw], m->tm_st_hid2, p, m->num_sms, st));
        p.bias = nullptr;
      } else {
        LAUNCH(C_FF1, launch_gemm<gemm::EPI_STORE_F32>(m->bn_wide, pw, m->tm_a2, m->tm_w1s[iw], m->tm_st_hid, p, m->num_sms, st));
        LAUNCH(C_OTHER, f32p::bias_gelu_split_kernel<<<ew_blocks, 256, 0, st>>>(m->hid32, m->b1 + size_t(j) * 4 * D, m->hid2, size_t(M), 4 * D));
      }
      r.N = D; r.K = 4 * D; r.b_row_offset = j * D; r.bias = m->b2 + size_t(j) * D;
      r.cond = (j + 1 < depth) ? m->cvec + size_t(j + 1) * D : nullptr;
      r.cond_stride = JD;
      LAUNCH(C_FF2, launch_gemm<gemm::EPI_BIAS_RESID_F32>(m->bn_narrow, pn, m->tm_hid2, m->tm_w2s[in], m->tm_st_hid, r, m->num_sms, st));
    }
  } else
  for (int j = 0; j < depth; ++j) {
    gemm::Params p{};
    p.L = L; p.H = H; p.Bsz = B; p.M = M;
    // q, k, v = LN1(u) Wqkv^T (no bias; LayerNorm folded) -> head-major bf16
    p.N = 3 * D; p.K = D; p.b_row_offset = j * 3 * D; p.out = m->qkv;
    p.ln_stats = m->stats; p.ln_parts = m->ln_parts;
    p.ln_s = m->ln_s_qkv + size_t(j) * 3 * D; p.ln_t = m->ln_t_qkv + size_t(j) * 3 * D;
    p.reverse = next_dir();
    memcpy(p.qkv_chunk, m->qkv_chunk, sizeof(p.qkv_chunk));
    LAUNCH(C_QKV, launch_gemm<gemm::EPI_QKV_HEADMAJOR>(m->bn_wide, pw, m->tm_a, m->tm_wqkv[iw], m->tm_st_qkv, p, m->num_sms, st));
    const int adir = next_dir();                // both attention kernels read the same QKV output
    // the two attention kernels read the same QKV output and write disjoint column ranges of `att`
    if (H - NL > 0) {
      if (m->qsoft_epi)
        LAUNCH(C_LINEAR, launch_k(attn::linear_attention_kernel<false>, dim3(H - NL, B), dim3(128), size_t(attn::LIN_SMEM_BYTES), st, m->qkv,
                                  m->att, B, H, L, NL, q_scale, adir));
      else
        LAUNCH(C_LINEAR, launch_k(attn::linear_attention_kernel<true>, dim3(H - NL, B), dim3(128), size_t(attn::LIN_SMEM_BYTES), st, m->qkv,
                                  m->att, B, H, L, NL, q_scale, adir));
    }
    if (NL > 0)
      LAUNCH(C_LOCAL, launch_local(m->tm_qkv_attn, m->att, B, H, L, NL, scale_log2e, adir, m->num_sms, 0, st));
    // u += att . Wo^T + bo ; also emits bf16(u) and its row statistics for the next folded LayerNorm
    gemm::Params r{};
    r.L = L; r.H = H; r.Bsz = B; r.M = M;
    r.N = D; r.K = D; r.b_row_offset = j * D; r.out = m->u; r.bias = m->bo + size_t(j) * D;
    r.out_bf16 = m->a; r.stats_out = m->stats; r.reverse = next_dir();
    const bool cl = Mc > 0 && j == depth - 1;   // compact last layer: the rest of the step runs on the selected rows only
    if (cl) {
      LAUNCH(C_OTHER, launch_k(k::gather_rows_kernel, dim3(std::min(Mc / 8, m->num_sms * 8)), dim3(256), 0, st, m->att, m->a, m->u_lo,
                               m->att_c, m->a_c, m->ulo_c, m->inv_path, m->ctl, L, D, group, B * group, Mc));
      r.M = Mc; r.out = m->ulo_c; r.out_bf16 = m->a_c; r.stats_out = m->stats_c; r.reverse = 0;
      LAUNCH(C_OUT, launch_resid_split(m, pn, m->resid_tma >= 1, m->tm_att_c, m->tm_wo[in], m->tm_rs_hi_c, m->tm_rs_lo_c, r, st));
      p.M = Mc; p.a_row_offset = 0; p.reverse = 0;
      p.N = 4 * D; p.K = D; p.b_row_offset = j * 4 * D; p.out = m->hid_c;
      p.ln_stats = m->stats_c;
      p.ln_s = m->ln_s_ff + size_t(j) * 4 * D; p.ln_t = m->ln_t_ff + size_t(j) * 4 * D;
      LAUNCH(C_FF1, launch_gemm<gemm::EPI_BIAS_GELU_BF16>(m->bn_wide, pw, m->tm_a_c, m->tm_w1[iw], m->tm_st_hid_c, p, m->num_sms, st));
      r.N = D; r.K = 4 * D; r.b_row_offset = j * D; r.bias = m->b2 + size_t(j) * D;
      r.cond = nullptr; r.cond_stride = JD;
      LAUNCH(C_FF2, launch_resid_split(m, pn, m->resid_tma >= 2, m->tm_hid_c, m->tm_w2[in], m->tm_rs_hi_c, m->tm_rs_lo_c, r, st));
      continue;
    }
    if (split) {
      r.out = m->u_lo;
      LAUNCH(C_OUT, launch_resid_split(m, pn, m->resid_tma >= 1, m->tm_att, m->tm_wo[in], m->tm_rs_hi, m->tm_rs_lo, r, st));
    } else {
      LAUNCH(C_OUT, launch_gemm<gemm::EPI_BIAS_RESID_F32>(m->bn_narrow, pn, m->tm_att, m->tm_wo[in], m->tm_st_hid, r, m->num_sms, st));
    }
    // hid = gelu(LN2(u) W1^T + b1)   (LayerNorm and bias folded into ln_s / ln_t), then
    // u += hid . W2^T + b2 (+ next layer's conditioning vector)
    p.M = M; p.a_row_offset = 0; p.reverse = next_dir();
    p.N = 4 * D; p.K = D; p.b_row_offset = j * 4 * D; p.out = m->hid;
    p.ln_stats = m->stats;
    p.ln_s = m->ln_s_ff + size_t(j) * 4 * D; p.ln_t = m->ln_t_ff + size_t(j) * 4 * D;
    LAUNCH(C_FF1, launch_gemm<gemm::EPI_BIAS_GELU_BF16>(m->bn_wide, pw, m->tm_a, m->tm_w1[iw], m->tm_st_hid, p, m->num_sms, st));
    r.M = M; r.a_row_offset = 0; r.reverse = next_dir();
    r.N = D; r.K = 4 * D; r.b_row_offset = j * D; r.bias = m->b2 + size_t(j) * D;
    r.out = m->u; r.out_bf16 = m->a; r.stats_out = m->stats;
    r.cond = (j + 1 < depth) ? m->cvec + size_t(j + 1) * D : nullptr;
    r.cond_stride = JD;
    if (split) {
      r.out = m->u_lo;
      LAUNCH(C_FF2, launch_resid_split(m, pn, m->resid_tma >= 2, m->tm_hid, m->tm_w2[in], m->tm_rs_hi, m->tm_rs_lo, r, st));
    } else {
      LAUNCH(C_FF2, launch_gemm<gemm::EPI_BIAS_RESID_F32>(m->bn_narrow, pn, m->tm_hid, m->tm_w2[in], m->tm_st_hid, r, m->num_sms, st));
    }
  }
  k::HeadArgs ha{};
  ha.u = split ? nullptr : m->u; ha.u_hi = Mc ? m->a_c : m->a; ha.u_lo = Mc ? m->ulo_c : m->u_lo; ha.compact = Mc ? 1 : 0; ha.gamma = m->norm_g; ha.beta = m->norm_b; ha.w_out = m->w_out; ha.b_out = m->b_out;
  ha.logits_out = logits_out; ha.state = sample ? m->state : nullptr; ha.inv_path = m->inv_path; ha.ctl = m->ctl;
  ha.B = B; ha.L = L; ha.D = D; ha.C = C; ha.group = sample ? group : 0;
  const int ntok = sample ? B * group : M;
  const int head_blocks = std::min((ntok + 7) / 8, m->num_sms * 2);
  LAUNCH(C_HEAD, launch_k(k::head_kernel, dim3(head_blocks), dim3(256), size_t(C) * D * sizeof(float), st, ha));
  if (advance) {
    const int n = M;
    const int blocks = std::min(std::max(1, n / (256 * 16)), m->num_sms);
    LAUNCH(C_OTHER, launch_k(k::advance_kernel, dim3(blocks), dim3(256), 0, st, m->ctl, m->state, n));
  }
#undef LAUNCH
  g_pdl = false;
  if (n_launch) *n_launch = launches;
  return cudaSuccess;
}

__global__ void set_ctl_kernel(k::DecodeCtl* ctl, int step, int start, const float* noise, uint8_t* traj,
                               unsigned long long seed, const unsigned long long* group_seeds, unsigned long long* stamps) {
  ctl->step = step;
  ctl->start = start;
  ctl->done = 0;
  ctl->pad = 0;
  ctl->noise = noise;
  ctl->traj = traj;
  ctl->seed = seed;
  ctl->group_seeds = group_seeds;
  ctl->stamps = stamps;
}

int check_ready(biom3_model* m, int B) {
  if (!m) return fail(BIOM3_ERR_INVALID, "null model");
  if (!m->finalized) return fail(BIOM3_ERR_STATE, "weights not finalized");
  if (B < 1 || B > m->max_batch)
    return fail(BIOM3_ERR_INVALID, "batch " + std::to_string(B) + " outside [1, max_batch=" +
                                       std::to_string(m->max_batch) + "]");
  return BIOM3_OK;
}

}  // namespace

extern "C" {

const char* biom3_last_error(void) { return g_err.c_str(); }

int biom3_create(const biom3_config* cfg, int device, int max_batch, biom3_model** out) {
  if (!cfg || !out) return fail(BIOM3_ERR_INVALID, "null argument");
  const biom3_config& c = *cfg;
  if (c.reversible) return fail(BIOM3_ERR_INVALID, "transformer_reversible=true is not supported");
  if (c.n_blocks != 1) return fail(BIOM3_ERR_INVALID, "transformer_blocks must be 1");
  if (c.heads <= 0 || c.dim != c.heads * attn::DH)
    return fail(BIOM3_ERR_INVALID, "transformer_dim / transformer_heads must be 32");
  if (3 * c.heads > gemm::QKV_MAX_CHUNKS) return fail(BIOM3_ERR_INVALID, "transformer_heads must be <= 85");
  if (c.dim % 256 != 0 || c.dim > 1024) return fail(BIOM3_ERR_INVALID, "transformer_dim must be a multiple of 256, <= 1024");
  if (c.local_window != attn::WIN) return fail(BIOM3_ERR_INVALID, "transformer_local_size must be 128");
  if (c.seq_len <= 0 || c.seq_len % c.local_window != 0)
    return fail(BIOM3_ERR_INVALID, "diffusion_steps must be a positive multiple of the local window");
  if (c.local_heads < 0 || c.local_heads > c.heads) return fail(BIOM3_ERR_INVALID, "bad transformer_local_heads");
  if (c.num_classes < 2 || c.num_classes > 32) return fail(BIOM3_ERR_INVALID, "num_classes must be in [2, 32]");
  if (c.depth < 1 || c.text_emb_dim < 1 || max_batch < 1) return fail(BIOM3_ERR_INVALID, "bad depth/text_emb_dim/max_batch");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BIOM3_ERR_CUDA, "no CUDA device: biom3_b200 has no CPU path");
  CU_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(BIOM3_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                    "; this library is built for sm_100a only");
  biom3_model* m = new biom3_model();
  m->cfg = c;
  m->device = device;
  m->max_batch = max_batch;
  m->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("BIOM3_PAIR")) m->use_pair = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_PDL")) m->use_pdl = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_SPLIT_RESID")) m->split_resid = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_SERPENTINE")) m->serpentine = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_RESID_TMA")) m->resid_tma = atoi(e);
  if (const char* e = getenv("BIOM3_ATTN3")) g_attn3 = atoi(e);
  if (const char* e = getenv("BIOM3_F32_ATTN_MMA")) m->f32_attn_mma = atoi(e);
  if (const char* e = getenv("BIOM3_F32_FUSED_GELU")) m->f32_fused_gelu = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_COMPACT")) m->compact_last = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_QSOFT_EPI")) m->qsoft_epi = atoi(e) != 0;
  if (const char* e = getenv("BIOM3_FWD_GRAPH")) m->fwd_graph = atoi(e) != 0;
  CU_OK(init_kernel_attributes());
  CU_OK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
  *out = m;
  return BIOM3_OK;
}

void biom3_destroy(biom3_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  if (m->fwd_exec) cudaGraphExecDestroy(m->fwd_exec);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

int biom3_set_weight(biom3_model* m, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
  if (!m || !key || !data || !shape || ndim < 0 || ndim > 8) return fail(BIOM3_ERR_INVALID, "bad set_weight argument");
  if (m->finalized) return fail(BIOM3_ERR_STATE, "weights already finalized");
  size_t esz = 0;
  switch (dtype) {
    case BIOM3_DTYPE_F32: esz = 4; break;
    case BIOM3_DTYPE_BF16: case BIOM3_DTYPE_F16: esz = 2; break;
    case BIOM3_DTYPE_F64: esz = 8; break;
    default: return fail(BIOM3_ERR_INVALID, std::string("unsupported dtype code for ") + key);
  }
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) return fail(BIOM3_ERR_INVALID, std::string("negative extent for ") + key);
    numel *= shape[i];
  }
  if (numel <= 0) return fail(BIOM3_ERR_INVALID, std::string("empty tensor for ") + key);
  // host or device pointer: device (and managed) memory is staged through one host copy
  cudaPointerAttributes pa{};
  const bool on_device = cudaPointerGetAttributes(&pa, data) == cudaSuccess &&
                         (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
  cudaGetLastError();                                   // an unregistered host pointer is not an error
  std::vector<uint8_t> staged;
  const uint8_t* src = static_cast<const uint8_t*>(data);
  if (on_device) {
    staged.resize(size_t(numel) * esz);
    CU_OK(cudaMemcpy(staged.data(), data, staged.size(), cudaMemcpyDeviceToHost));
    src = staged.data();
  }
  std::vector<float> w(static_cast<size_t>(numel));
  for (int64_t i = 0; i < numel; ++i) {
    switch (dtype) {
      case BIOM3_DTYPE_F32: w[i] = reinterpret_cast<const float*>(src)[i]; break;
      case BIOM3_DTYPE_F64: w[i] = float(reinterpret_cast<const double*>(src)[i]); break;
      case BIOM3_DTYPE_BF16: {
        const uint32_t bits = uint32_t(reinterpret_cast<const uint16_t*>(src)[i]) << 16;
        memcpy(&w[i], &bits, 4);
        break;
      }
      default: w[i] = __half2float(reinterpret_cast<const __half*>(src)[i]); break;
    }
  }
  m->host_w[key] = std::move(w);
  m->host_shape[key] = std::vector<int64_t>(shape, shape + ndim);
  return BIOM3_OK;
}

int biom3_set_precision(biom3_model* m, int precision) {
  if (!m) return fail(BIOM3_ERR_INVALID, "null model");
  if (precision != 0 && precision != 1) return fail(BIOM3_ERR_INVALID, "precision must be 0 (bf16) or 1 (fp32-class)");
  if (m->finalized) return fail(BIOM3_ERR_STATE, "precision must be chosen before biom3_finalize_weights");
  m->precision = precision;
  return BIOM3_OK;
}

int biom3_finalize_weights(biom3_model* m) {
  if (!m) return fail(BIOM3_ERR_INVALID, "null model");
  if (m->finalized) return fail(BIOM3_ERR_STATE, "weights already finalized");
  CU_OK(cudaSetDevice(m->device));
  const biom3_config& c = m->cfg;
  const size_t D = c.dim, depth = c.depth, L = c.seq_len, C = c.num_classes, E = c.text_emb_dim, W = c.local_window;
  const size_t Bm = m->max_batch, M = Bm * L;
  int r;
#define TRY(x) if ((r = (x)) != BIOM3_OK) return r
  TRY(dev_alloc(m, &m->Wqkv, depth * 3 * D * D));
  TRY(dev_alloc(m, &m->Wo, depth * D * D));
  TRY(dev_alloc(m, &m->W1, depth * 4 * D * D));
  TRY(dev_alloc(m, &m->W2, depth * 4 * D * D));
  TRY(dev_alloc(m, &m->ln_s_qkv, depth * 3 * D));
  TRY(dev_alloc(m, &m->ln_t_qkv, depth * 3 * D));
  TRY(dev_alloc(m, &m->ln_s_ff, depth * 4 * D));
  TRY(dev_alloc(m, &m->ln_t_ff, depth * 4 * D));
  TRY(dev_alloc(m, &m->bo, depth * D));
  TRY(dev_alloc(m, &m->b1, depth * 4 * D));
  TRY(dev_alloc(m, &m->b2, depth * D));
  TRY(dev_alloc(m, &m->emb, C * D));
  TRY(dev_alloc(m, &m->ax0, (L / W) * D));
  TRY(dev_alloc(m, &m->ax1, W * D));
  TRY(dev_alloc(m, &m->norm_g, D));
  TRY(dev_alloc(m, &m->norm_b, D));
  TRY(dev_alloc(m, &m->w_out, C * D));
  TRY(dev_alloc(m, &m->b_out, C));
  TRY(dev_alloc(m, &m->Ttab, L * depth * D));
  TRY(dev_alloc(m, &m->y_w0, 4 * D * E));
  TRY(dev_alloc(m, &m->y_b0, 4 * D));
  TRY(dev_alloc(m, &m->y_w2, D * depth * 4 * D));
  TRY(dev_alloc(m, &m->y_b2, D * depth));

  const std::string T = "transformer.";
  TRY(upload_f32(m, T + "x_emb_NN.weight", C * D, m->emb));
  TRY(upload_f32(m, T + "axial_pos_emb.weights_0", (L / W) * D, m->ax0));
  TRY(upload_f32(m, T + "axial_pos_emb.weights_1", W * D, m->ax1));
  TRY(upload_f32(m, T + "norm.weight", D, m->norm_g));
  TRY(upload_f32(m, T + "norm.bias", D, m->norm_b));
  TRY(upload_f32(m, T + "out.weight", C * D, m->w_out));
  TRY(upload_f32(m, T + "out.bias", C, m->b_out));
  TRY(upload_f32(m, T + "y_mlp.0.weight", 4 * D * E, m->y_w0));
  TRY(upload_f32(m, T + "y_mlp.0.bias", 4 * D, m->y_b0));
  TRY(upload_f32(m, T + "y_mlp.2.weight", D * depth * 4 * D, m->y_w2));
  TRY(upload_f32(m, T + "y_mlp.2.bias", D * depth, m->y_b2));
  build_qkv_chunk_order(m);
  for (size_t j = 0; j < depth; ++j) {
    const std::string P = T + "transformer_blocks.0." + std::to_string(j) + ".layers.layers.0.";
    // Folded LayerNorms: W' = gamma (.) W in bf16, s_n = sum_k bf16(W'_nk), t_n = sum_k beta_k W_nk (+ bias)
    TRY(upload_folded(m, P + "0.norm.weight", P + "0.norm.bias", P + "0.fn.to_q.weight", "", D, D,
                      m->Wqkv + (j * 3 + 0) * D * D, m->ln_s_qkv + (j * 3 + 0) * D, m->ln_t_qkv + (j * 3 + 0) * D));
    TRY(upload_folded(m, P + "0.norm.weight", P + "0.norm.bias", P + "0.fn.to_k.weight", "", D, D,
                      m->Wqkv + (j * 3 + 1) * D * D, m->ln_s_qkv + (j * 3 + 1) * D, m->ln_t_qkv + (j * 3 + 1) * D));
    TRY(upload_folded(m, P + "0.norm.weight", P + "0.norm.bias", P + "0.fn.to_v.weight", "", D, D,
                      m->Wqkv + (j * 3 + 2) * D * D, m->ln_s_qkv + (j * 3 + 2) * D, m->ln_t_qkv + (j * 3 + 2) * D));
    TRY(permute_qkv_chunks(m, j));
    TRY(upload_bf16(m, P + "0.fn.to_out.weight", D * D, m->Wo + j * D * D));
    TRY(upload_f32(m, P + "0.fn.to_out.bias", D, m->bo + j * D));
    TRY(upload_folded(m, P + "1.norm.weight", P + "1.norm.bias", P + "1.fn.fn.w1.weight", P + "1.fn.fn.w1.bias", 4 * D, D,
                      m->W1 + j * 4 * D * D, m->ln_s_ff + j * 4 * D, m->ln_t_ff + j * 4 * D, 0.5f));
    TRY(upload_bf16(m, P + "1.fn.fn.w2.weight", 4 * D * D, m->W2 + j * 4 * D * D));
    TRY(upload_f32(m, P + "1.fn.fn.w2.bias", D, m->b2 + j * D));
  }
  if (m->precision == 1) {
    TRY(dev_alloc(m, &m->Wqkv2, depth * 3 * D * 2 * D));
    TRY(dev_alloc(m, &m->Wo2, depth * D * 2 * D));
    TRY(dev_alloc(m, &m->W1s, depth * 4 * D * 2 * D));
    TRY(dev_alloc(m, &m->W2s, depth * D * 8 * D));
    TRY(dev_alloc(m, &m->ln1_g, depth * D));
    TRY(dev_alloc(m, &m->ln1_b, depth * D));
    TRY(dev_alloc(m, &m->ln2_g, depth * D));
    TRY(dev_alloc(m, &m->ln2_b, depth * D));
    for (size_t j = 0; j < depth; ++j) {
      const std::string P = T + "transformer_blocks.0." + std::to_string(j) + ".layers.layers.0.";
      TRY(upload_split(m, P + "0.fn.to_q.weight", D, D, m->Wqkv2 + (j * 3 + 0) * D * 2 * D));
      TRY(upload_split(m, P + "0.fn.to_k.weight", D, D, m->Wqkv2 + (j * 3 + 1) * D * 2 * D));
      TRY(upload_split(m, P + "0.fn.to_v.weight", D, D, m->Wqkv2 + (j * 3 + 2) * D * 2 * D));
      TRY(upload_split(m, P + "0.fn.to_out.weight", D, D, m->Wo2 + j * D * 2 * D));
      TRY(upload_split(m, P + "1.fn.fn.w1.weight", 4 * D, D, m->W1s + j * 4 * D * 2 * D));
      TRY(upload_split(m, P + "1.fn.fn.w2.weight", D, 4 * D, m->W2s + j * D * 8 * D));
      TRY(upload_f32(m, P + "1.fn.fn.w1.bias", 4 * D, m->b1 + j * 4 * D));
      TRY(upload_f32(m, P + "0.norm.weight", D, m->ln1_g + j * D));
      TRY(upload_f32(m, P + "0.norm.bias", D, m->ln1_b + j * D));
      TRY(upload_f32(m, P + "1.norm.weight", D, m->ln2_g + j * D));
      TRY(upload_f32(m, P + "1.norm.bias", D, m->ln2_b + j * D));
    }
  }
  // time-conditioning table: depends on the step only -> once per model load
  {
    float *te, *h1, *tt, *w0, *b0, *w2, *b2;
    CU_OK(cudaMalloc(&te, L * D * sizeof(float)));
    CU_OK(cudaMalloc(&h1, L * 4 * D * sizeof(float)));
    CU_OK(cudaMalloc(&tt, L * D * depth * sizeof(float)));
    CU_OK(cudaMalloc(&w0, 4 * D * D * sizeof(float)));
    CU_OK(cudaMalloc(&b0, 4 * D * sizeof(float)));
    CU_OK(cudaMalloc(&w2, D * depth * 4 * D * sizeof(float)));
    CU_OK(cudaMalloc(&b2, D * depth * sizeof(float)));
    TRY(upload_f32(m, T + "mlp.0.weight", 4 * D * D, w0));
    TRY(upload_f32(m, T + "mlp.0.bias", 4 * D, b0));
    TRY(upload_f32(m, T + "mlp.2.weight", D * depth * 4 * D, w2));
    TRY(upload_f32(m, T + "mlp.2.bias", D * depth, b2));
    k::time_embedding_kernel<<<unsigned((L * D + 255) / 256), 256>>>(te, int(L), int(D), float(L));
    sgemm(te, w0, b0, h1, int(L), int(4 * D), int(D), 1, 0);
    sgemm(h1, w2, b2, tt, int(L), int(D * depth), int(4 * D), 0, 0);
    const size_t n = L * D * depth;
    k::cond_transpose_kernel<<<unsigned((n + 255) / 256), 256>>>(tt, m->Ttab, int(L), int(D), int(depth));
    CU_OK(cudaDeviceSynchronize());
    cudaFree(te); cudaFree(h1); cudaFree(tt); cudaFree(w0); cudaFree(b0); cudaFree(w2); cudaFree(b2);
  }
  // expected number of keys: anything else in the map is an unexpected key (strict load)
  const size_t expected = 15 + 13 * depth;
  if (m->host_w.size() != expected)
    return fail(BIOM3_ERR_STATE, "unexpected keys in state dict: got " + std::to_string(m->host_w.size()) +
                                     " tensors, expected " + std::to_string(expected));
  m->host_w.clear();
  m->host_shape.clear();

  // workspace, sized for max_batch
  if (m->precision == 0 && m->split_resid) {
    TRY(dev_alloc(m, &m->u_lo, M * D));
  } else {
    TRY(dev_alloc(m, &m->u, M * D));
  }
  TRY(dev_alloc(m, &m->a, M * D));
  TRY(dev_alloc(m, &m->qkv, M * 3 * D));
  TRY(dev_alloc(m, &m->att, M * D));
  TRY(dev_alloc(m, &m->hid, M * 4 * D));
  TRY(dev_alloc(m, &m->Yh, Bm * 4 * D));
  TRY(dev_alloc(m, &m->Ytmp, Bm * D * depth));
  TRY(dev_alloc(m, &m->Y, Bm * D * depth));
  TRY(dev_alloc(m, &m->cvec, Bm * D * depth));
  m->ln_parts = int(D) / m->bn_narrow * 2;
  TRY(dev_alloc(m, &m->stats, M * size_t(m->ln_parts) * 2));
  TRY(dev_alloc(m, &m->state, M));
  TRY(dev_alloc(m, &m->inv_path, M));
  TRY(dev_alloc(m, &m->t_i32, Bm));
  TRY(dev_alloc(m, &m->err_flags, 1));
  CU_OK(cudaMemset(m->err_flags, 0, sizeof(int)));
  TRY(dev_alloc(m, &m->ctl, 1));
  TRY(dev_alloc(m, &m->stamps, 2 * L));
  CU_OK(cudaMemset(m->stamps, 0, 2 * L * sizeof(unsigned long long)));
  CU_OK(cudaMemset(m->ctl, 0, sizeof(k::DecodeCtl)));
  CU_OK(cudaMemset(m->inv_path, 0, M * sizeof(int)));

  if (m->precision == 1) {
    TRY(dev_alloc(m, &m->a2, M * 2 * D));
    TRY(dev_alloc(m, &m->hid2, M * 8 * D));
    TRY(dev_alloc(m, &m->qkv32, M * 3 * D));
    TRY(dev_alloc(m, &m->hid32, M * 4 * D));
    TRY(make_tmap(&m->tm_a2, m->a2, M, 2 * D, 128));
    TRY(make_tmap(&m->tm_hid2, m->hid2, M, 8 * D, 128));
    TRY(make_store_tmap(&m->tm_st_hid2, m->hid2, M, 8 * D));
    for (int i = 0; i < 2; ++i) {
      const uint32_t box = i ? 256 : 128;
      TRY(make_tmap(&m->tm_wqkv2[i], m->Wqkv2, depth * 3 * D, 2 * D, box));
      TRY(make_tmap(&m->tm_wo2[i], m->Wo2, depth * D, 2 * D, box));
      TRY(make_tmap(&m->tm_w1s[i], m->W1s, depth * 4 * D, 2 * D, box));
      TRY(make_tmap(&m->tm_w2s[i], m->W2s, depth * D, 8 * D, box));
    }
  }
  if (m->precision == 0 && m->split_resid && m->compact_last) {
    // at most max_batch^2 selected tokens; worth it only while that is a fraction of the B * L rows
    const size_t want = (Bm * Bm + 255) / 256 * 256;
    if (want * 2 <= M) {
      m->compact_rows_max = int(want);
      TRY(dev_alloc(m, &m->att_c, want * D));
      TRY(dev_alloc(m, &m->a_c, want * D));
      TRY(dev_alloc(m, &m->ulo_c, want * D));
      TRY(dev_alloc(m, &m->hid_c, want * 4 * D));
      TRY(dev_alloc(m, &m->stats_c, want * size_t(m->ln_parts) * 2));
      TRY(make_tmap(&m->tm_att_c, m->att_c, want, D, 128));
      TRY(make_tmap(&m->tm_a_c, m->a_c, want, D, 128));
      TRY(make_tmap(&m->tm_hid_c, m->hid_c, want, 4 * D, 128));
      TRY(make_store_tmap(&m->tm_st_hid_c, m->hid_c, want, 4 * D));
      TRY(make_store_tmap(&m->tm_rs_hi_c, m->a_c, want, D));
      TRY(make_store_tmap(&m->tm_rs_lo_c, m->ulo_c, want, D));
    }
  }
  if (m->u_lo) {
    TRY(make_store_tmap(&m->tm_rs_hi, m->a, M, D));
    TRY(make_store_tmap(&m->tm_rs_lo, m->u_lo, M, D));
  } else {
    m->resid_tma = 0;
  }
  TRY(make_tmap(&m->tm_a, m->a, M, D, 128));
  TRY(make_tmap(&m->tm_att, m->att, M, D, 128));
  TRY(make_tmap(&m->tm_hid, m->hid, M, 4 * D, 128));
  TRY(make_store_tmap(&m->tm_st_qkv, m->qkv, M * 3 * D / 32, 32));
  TRY(make_tmap_sw64(&m->tm_qkv_attn, m->qkv, M * 3 * D / 32));
  TRY(make_store_tmap(&m->tm_st_hid, m->hid, M, 4 * D));
  for (int i = 0; i < 2; ++i) {
    const uint32_t box = i ? 256 : 128;
    TRY(make_tmap(&m->tm_wqkv[i], m->Wqkv, depth * 3 * D, D, box));
    TRY(make_tmap(&m->tm_wo[i], m->Wo, depth * D, D, box));
    TRY(make_tmap(&m->tm_w1[i], m->W1, depth * 4 * D, D, box));
    TRY(make_tmap(&m->tm_w2[i], m->W2, depth * D, 4 * D, box));
  }
#undef TRY
  m->finalized = true;
  m->launches_per_step = 2 + int(depth) * (m->precision == 1 ? (m->f32_fused_gelu ? 8 : 9) : 6) + 2 - (c.local_heads == 0 ? int(depth) : 0) -
                         (c.heads == c.local_heads ? int(depth) : 0);
  return BIOM3_OK;
}

int biom3_launches_per_step(const biom3_model* m) { return m ? m->launches_per_step : 0; }

int biom3_input_errors(biom3_model* m, int* flags) {
  if (!m || !flags || !m->finalized) return fail(BIOM3_ERR_INVALID, "bad input_errors argument");
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(flags, m->err_flags, sizeof(int), cudaMemcpyDeviceToHost));
  if (*flags) CU_OK(cudaMemset(m->err_flags, 0, sizeof(int)));
  return BIOM3_OK;
}

int biom3_forward(biom3_model* m, const int64_t* x, const int64_t* t, const float* y_c, int B, float* logits,
                  void* stream) {
  int r = check_ready(m, B);
  if (r) return r;
  if (!x || !t || !y_c || !logits) return fail(BIOM3_ERR_INVALID, "null argument");
  CU_OK(cudaSetDevice(m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int L = m->cfg.seq_len, n = B * L;
  k::i64_to_u8_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(x), m->state, n, m->cfg.num_classes, m->err_flags);
  k::i64_to_i32_kernel<<<(B + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(t), m->t_i32, B, L, m->err_flags);
  run_y_mlp(m, y_c, B, st);
  if (!m->fwd_graph) {
    CU_OK(run_step(m, B, 0, m->t_i32, logits, false, false, st, nullptr, nullptr));
    CU_OK(cudaGetLastError());
    return BIOM3_OK;
  }
  const size_t C = m->cfg.num_classes;
  if (!m->logits_buf) {
    int r2 = dev_alloc(m, &m->logits_buf, size_t(m->max_batch) * C * L);
    if (r2) return r2;
  }
  if (!m->fwd_exec || m->fwd_B != B) {
    if (m->fwd_exec) {
      cudaGraphExecDestroy(m->fwd_exec);
      m->fwd_exec = nullptr;
    }
    cudaGraph_t graph;
    CU_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = run_step(m, B, 0, m->t_i32, m->logits_buf, false, false, m->cap_stream, nullptr, nullptr);
    cudaError_t e2 = cudaStreamEndCapture(m->cap_stream, &graph);
    if (e != cudaSuccess) return fail(BIOM3_ERR_CUDA, std::string("forward capture: ") + cudaGetErrorString(e));
    CU_OK(e2);
    CU_OK(cudaGraphInstantiate(&m->fwd_exec, graph, 0));
    cudaGraphDestroy(graph);
    m->fwd_B = B;
  }
  CU_OK(cudaGraphLaunch(m->fwd_exec, st));
  CU_OK(cudaMemcpyAsync(logits, m->logits_buf, size_t(B) * C * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_decode(biom3_model* m, const float* y_c, const int64_t* path, const int64_t* state0, int start_step,
                 int num_steps, int group, const float* noise, uint64_t seed, const uint64_t* group_seeds, int64_t* tokens,
                 uint8_t* traj, int B, void* stream) {
  int r = check_ready(m, B);
  if (r) return r;
  if (!y_c || !path || !tokens) return fail(BIOM3_ERR_INVALID, "null argument");
  const int L = m->cfg.seq_len;
  if (group < 1 || B % group != 0) return fail(BIOM3_ERR_INVALID, "B must be a multiple of group");
  if (start_step < 0 || num_steps < 0 || start_step + num_steps > L)
    return fail(BIOM3_ERR_INVALID, "step range outside [0, diffusion_steps]");
  CU_OK(cudaSetDevice(m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = B * L;
  if (state0)
    k::i64_to_u8_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(state0), m->state, n, m->cfg.num_classes, m->err_flags);
  else
    CU_OK(cudaMemsetAsync(m->state, 0, n, st));
  CU_OK(cudaMemsetAsync(m->inv_path, 0, size_t(n) * sizeof(int), st));   // a non-permutation row falls back to location 0, like argmax of an all-false mask
  k::inverse_path_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(path), m->inv_path, B, L, m->err_flags);
  run_y_mlp(m, y_c, B, st);
  set_ctl_kernel<<<1, 1, 0, st>>>(m->ctl, start_step, start_step, noise, traj, seed,
                                  reinterpret_cast<const unsigned long long*>(group_seeds), m->stamps);
  CU_OK(cudaGetLastError());

  if (num_steps > 0) {
    if (!m->graph_exec || m->graph_B != B || m->graph_group != group) {
      if (m->graph_exec) {
        cudaGraphExecDestroy(m->graph_exec);
        m->graph_exec = nullptr;
      }
      cudaGraph_t graph;
      CU_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
      int n_launch = 0;
      cudaError_t e = run_step(m, B, group, nullptr, nullptr, true, true, m->cap_stream, nullptr, &n_launch);
      cudaError_t e2 = cudaStreamEndCapture(m->cap_stream, &graph);
      if (e == cudaSuccess) m->launches_per_step = n_launch;
      if (e != cudaSuccess) return fail(BIOM3_ERR_CUDA, std::string("step capture: ") + cudaGetErrorString(e));
      CU_OK(e2);
      CU_OK(cudaGraphInstantiate(&m->graph_exec, graph, 0));
      cudaGraphDestroy(graph);
      m->graph_B = B;
      m->graph_group = group;
    }
    for (int s = 0; s < num_steps; ++s) CU_OK(cudaGraphLaunch(m->graph_exec, st));
  }
  k::u8_to_i64_kernel<<<(n + 255) / 256, 256, 0, st>>>(m->state, reinterpret_cast<long long*>(tokens), n);
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_profile_step(biom3_model* m, int B, int group, biom3_step_profile* out) {
  int r = check_ready(m, B);
  if (r) return r;
  if (!out) return fail(BIOM3_ERR_INVALID, "null argument");
  if (group < 1 || B % group != 0) return fail(BIOM3_ERR_INVALID, "B must be a multiple of group");
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());             // a decode still running on the caller's stream owns the resident state
  cudaStream_t st = m->cap_stream;
  // a valid resident state is assumed (call after a decode); run one warm step then a timed one
  set_ctl_kernel<<<1, 1, 0, st>>>(m->ctl, 0, 0, nullptr, nullptr, 1234ull, nullptr, nullptr);
  CU_OK(run_step(m, B, group, nullptr, nullptr, true, true, st, nullptr, nullptr));
  Profiler prof;
  prof.st = st;
  int launches = 0;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  cudaEventRecord(t0, st);
  CU_OK(run_step(m, B, group, nullptr, nullptr, true, true, st, &prof, &launches));
  cudaEventRecord(t1, st);
  CU_OK(cudaStreamSynchronize(st));
  float acc[C_COUNT] = {};
  for (size_t i = 0; i < prof.cat.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, prof.ev[2 * i], prof.ev[2 * i + 1]);
    acc[prof.cat[i]] += ms;
    cudaEventDestroy(prof.ev[2 * i]);
    cudaEventDestroy(prof.ev[2 * i + 1]);
  }
  cudaEventElapsedTime(&out->total_ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  out->gemm_qkv_ms = acc[C_QKV]; out->gemm_out_ms = acc[C_OUT]; out->gemm_ff1_ms = acc[C_FF1];
  out->gemm_ff2_ms = acc[C_FF2]; out->local_attn_ms = acc[C_LOCAL]; out->linear_attn_ms = acc[C_LINEAR];
  out->layernorm_ms = acc[C_LN]; out->embed_ms = acc[C_EMBED]; out->head_ms = acc[C_HEAD];
  out->other_ms = acc[C_OTHER];
  out->launches = launches;
  out->compact_rows = m->last_compact_rows;
  return BIOM3_OK;
}

int biom3_debug_copy(biom3_model* m, const char* name, void* host_dst, int64_t nbytes) {
  if (!m || !name || !host_dst || !m->finalized) return fail(BIOM3_ERR_INVALID, "bad debug_copy argument");
  const size_t D = m->cfg.dim, depth = m->cfg.depth, L = m->cfg.seq_len, M = size_t(m->max_batch) * L;
  const std::string n(name);
  const void* src = nullptr;
  size_t sz = 0;
  if (n == "u") { src = m->u; sz = m->u ? M * D * 4 : 0; }
  else if (n == "u_lo") { src = m->u_lo; sz = m->u_lo ? M * D * 2 : 0; }
  else if (n == "a") { src = m->a; sz = M * D * 2; }
  else if (n == "qkv") { src = m->qkv; sz = M * 3 * D * 2; }
  else if (n == "att") { src = m->att; sz = M * D * 2; }
  else if (n == "hid") { src = m->hid; sz = M * 4 * D * 2; }
  else if (n == "cvec") { src = m->cvec; sz = size_t(m->max_batch) * depth * D * 4; }
  else if (n == "Y") { src = m->Y; sz = size_t(m->max_batch) * depth * D * 4; }
  else if (n == "Ttab") { src = m->Ttab; sz = L * depth * D * 4; }
  else if (n == "state") { src = m->state; sz = M; }
  else if (n == "stamps") { src = m->stamps; sz = 2 * L * sizeof(unsigned long long); }
  else return fail(BIOM3_ERR_INVALID, "unknown buffer " + n);
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(host_dst, src, std::min(sz, size_t(nbytes)), cudaMemcpyDeviceToHost));
  return BIOM3_OK;
}

int biom3_facilitator_create(int in_dim, int hid_dim, int out_dim, const float* w0_v, float w0_g, const float* b0,
                             const float* w1_v, float w1_g, const float* b1, int device, biom3_facilitator_t** out) {
  if (!w0_v || !b0 || !w1_v || !b1 || !out || in_dim < 1 || hid_dim < 1 || out_dim < 1)
    return fail(BIOM3_ERR_INVALID, "bad facilitator argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BIOM3_ERR_CUDA, "no CUDA device: biom3_b200 has no CPU path");
  CU_OK(cudaSetDevice(device));
  // weight_norm(dim=None): W = g * V / ||V||_F with a scalar g -> folded into the weights once, here
  auto fold = [](const float* v, float g, size_t n) {
    double ss = 0.0;
    for (size_t i = 0; i < n; ++i) ss += double(v[i]) * double(v[i]);
    const float scale = float(double(g) / std::sqrt(ss));
    std::vector<float> w(n);
    for (size_t i = 0; i < n; ++i) w[i] = v[i] * scale;
    return w;
  };
  const std::vector<float> w0 = fold(w0_v, w0_g, size_t(hid_dim) * in_dim);
  const std::vector<float> w1 = fold(w1_v, w1_g, size_t(out_dim) * hid_dim);
  biom3_facilitator_t* f = new biom3_facilitator_t();
  f->device = device; f->in_dim = in_dim; f->hid_dim = hid_dim; f->out_dim = out_dim;
  cudaError_t e = cudaMalloc(&f->w0, w0.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->b0, hid_dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->w1, w1.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->b1, out_dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(f->w0, w0.data(), w0.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->b0, b0, hid_dim * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->w1, w1.data(), w1.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(f->b1, b1, out_dim * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    biom3_facilitator_destroy(f);
    return fail(BIOM3_ERR_CUDA, std::string("facilitator upload: ") + cudaGetErrorString(e));
  }
  *out = f;
  return BIOM3_OK;
}

void biom3_facilitator_destroy(biom3_facilitator_t* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  cudaFree(f->w0); cudaFree(f->b0); cudaFree(f->w1); cudaFree(f->b1); cudaFree(f->hid);
  delete f;
}

int biom3_facilitator_forward(biom3_facilitator_t* f, const float* z_t, int P, float* z_c, void* stream) {
  if (!f || !z_t || !z_c || P < 1) return fail(BIOM3_ERR_INVALID, "bad facilitator argument");
  CU_OK(cudaSetDevice(f->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P > f->hid_rows) {                                    // hidden activation [P][hid], grown on demand
    CU_OK(cudaStreamSynchronize(st));
    cudaFree(f->hid);
    f->hid = nullptr;
    f->hid_rows = 0;
    CU_OK(cudaMalloc(&f->hid, size_t(P) * f->hid_dim * sizeof(float)));
    f->hid_rows = P;
  }
  sgemm(z_t, f->w0, f->b0, f->hid, P, f->hid_dim, f->in_dim, 2, st);      // Linear + exact erf-GELU (dropout p = 0 at eval)
  sgemm(f->hid, f->w1, f->b1, z_c, P, f->out_dim, f->hid_dim, 0, st);
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_facilitator(const float* z_t, int P, int in_dim, int hid_dim, int out_dim, const float* w0_v, float w0_g,
                      const float* b0, const float* w1_v, float w1_g, const float* b1, float* z_c, void* stream) {
  int dev = 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BIOM3_ERR_CUDA, "no CUDA device: biom3_b200 has no CPU path");
  CU_OK(cudaGetDevice(&dev));
  biom3_facilitator_t* f = nullptr;
  int r = biom3_facilitator_create(in_dim, hid_dim, out_dim, w0_v, w0_g, b0, w1_v, w1_g, b1, dev, &f);
  if (r) return r;
  r = biom3_facilitator_forward(f, z_t, P, z_c, stream);
  if (r == BIOM3_OK && cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess)
    r = fail(BIOM3_ERR_CUDA, "facilitator: stream synchronize failed");
  biom3_facilitator_destroy(f);
  return r;
}

int biom3_debug_trace(int which, void* host_dst, int64_t nbytes) {
  if (!host_dst || nbytes <= 0 || which < 0 || which > 1) return fail(BIOM3_ERR_INVALID, "bad debug_trace argument");
  CU_OK(cudaDeviceSynchronize());
  if (which == 0) CU_OK(cudaMemcpyFromSymbol(host_dst, attn::g_ms_trace, std::min(size_t(nbytes), sizeof(attn::g_ms_trace))));
  else CU_OK(cudaMemcpyFromSymbol(host_dst, gemm::g_gemm_trace, std::min(size_t(nbytes), sizeof(gemm::g_gemm_trace))));
  return BIOM3_OK;
}

int biom3_debug_noise(uint64_t seed, int step, int B, int L, int C, float* out, void* stream) {
  if (!out || step < 0 || B < 1 || L < 1 || C < 1 || C > 32) return fail(BIOM3_ERR_INVALID, "bad debug_noise argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BIOM3_ERR_CUDA, "no CUDA device: biom3_b200 has no CPU path");
  const long long n = static_cast<long long>(B) * L * C;
  if (n > INT_MAX) return fail(BIOM3_ERR_INVALID, "debug_noise: B * L * C too large");
  k::debug_noise_kernel<<<unsigned((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned long long>(seed), step, B * L, C, out);
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_random_paths(uint64_t seed, int B, int L, int64_t* path, void* stream) {
  if (!path || B < 1 || L < 1 || L > 8192) return fail(BIOM3_ERR_INVALID, "bad random_paths argument (1 <= L <= 8192)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BIOM3_ERR_CUDA, "no CUDA device: biom3_b200 has no CPU path");
  int N = 1;
  while (N < L) N <<= 1;
  const size_t smem = size_t(N) * 12;
  CU_OK(init_kernel_attributes());
  k::random_paths_kernel<<<B, std::min(1024, std::max(32, N / 2)), smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned long long>(seed), reinterpret_cast<long long*>(path), L, N);
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_sample_all(const float* logits, const float* noise, int64_t* tok, int B, int L, int C, void* stream) {
  if (!logits || !noise || !tok || B < 1 || L < 1 || C < 2 || C > 32) return fail(BIOM3_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = B * L;
  if (L % k::SAMPLE_TILE == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(noise)) & 15) == 0) {
    CU_OK(init_kernel_attributes());
    k::sample_all_tiled_kernel<<<n / k::SAMPLE_TILE, k::SAMPLE_TILE, size_t(2) * C * k::SAMPLE_TILE * sizeof(float), st>>>(
        logits, noise, reinterpret_cast<long long*>(tok), B, L, C);
  } else {
    k::sample_all_kernel<<<(n + 255) / 256, 256, 0, st>>>(logits, noise, reinterpret_cast<long long*>(tok), B, L, C);
  }
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_unmask(const int64_t* tok, const int64_t* path, int64_t* state, int B, int L, int group, int step,
                 void* stream) {
  if (!tok || !path || !state || B < 1 || L < 1 || group < 1 || B % group != 0 || step < 0 || step >= L)
    return fail(BIOM3_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  k::unmask_scan_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const long long*>(tok), reinterpret_cast<const long long*>(path),
                                           reinterpret_cast<long long*>(state), B, L, group, step);
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_attention_test(const void* qkv, void* out, int B, int H, int L, int NL, int variant, void* stream) {
  if (!qkv || !out || B < 1 || H < 1 || NL < 0 || NL > H || L < attn::WIN || L % attn::WIN)
    return fail(BIOM3_ERR_INVALID, "bad attention_test argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU_OK(init_kernel_attributes());
  if (const char* e = getenv("BIOM3_ATTN3")) g_attn3 = atoi(e);
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(attn::DH));
  const float q_scale = 1.0f / sqrtf(float(attn::DH));
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  bf16* o = reinterpret_cast<bf16*>(out);
  if (NL > 0) {
    CUtensorMap tm;
    int r = make_tmap_sw64(&tm, qkv, uint64_t(3) * B * H * L);
    if (r) return r;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    launch_local(tm, o, B, H, L, NL, scale_log2e, 0, sms, (variant & 1) != 0, st);
  }
  if (H - NL > 0) {                      // variant bit 1: the caller's q of the linear heads is already softmaxed over its features
    if (variant & 2) attn::linear_attention_kernel<false><<<dim3(H - NL, B), 128, attn::LIN_SMEM_BYTES, st>>>(q, o, B, H, L, NL, q_scale, 0);
    else attn::linear_attention_kernel<true><<<dim3(H - NL, B), 128, attn::LIN_SMEM_BYTES, st>>>(q, o, B, H, L, NL, q_scale, 0);
  }
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

int biom3_gemm_test(const void* A, const void* W, const float* bias, void* out, int M, int N, int K, int epi,
                    int block_n, int pair, void* stream) {
  if (!A || !W || !out) return fail(BIOM3_ERR_INVALID, "null argument");
  if (block_n != 256) return fail(BIOM3_ERR_INVALID, "block_n must be 256");
  if (M % 128 || N % block_n || K % 64) return fail(BIOM3_ERR_INVALID, "M%128, N%block_n, K%64 must be 0");
  const int split3 = (pair >> 1) & 1;          // bit 1: A and W are [hi | lo] halves of width 2K (fp32-class schedule)
  pair &= 1;
  if (pair && (block_n != 256 || M % 256)) return fail(BIOM3_ERR_INVALID, "pair tiling needs block_n == 256 and M % 256 == 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int gemm_dev = 0, sms = 0;
  CU_OK(cudaGetDevice(&gemm_dev));
  CU_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gemm_dev));
  CUtensorMap ta, tb;
  int r;
  if ((r = make_tmap(&ta, A, M, split3 ? 2 * K : K, 128))) return r;
  if ((r = make_tmap(&tb, W, N, split3 ? 2 * K : K, pair ? block_n / 2 : block_n))) return r;
  gemm::Params p{};
  p.split3 = split3;
  p.M = M; p.N = N; p.K = K; p.b_row_offset = 0; p.out = out; p.bias = bias; p.cond = nullptr; p.cond_stride = 0;
  p.L = M; p.H = 1; p.Bsz = 1;
  if (const char* e = getenv("BIOM3_GEMM_TRACE")) p.trace = atoi(e);
  CUtensorMap tc;
  memset(&tc, 0, sizeof(tc));
  if (epi == gemm::EPI_STORE_BF16 || epi == gemm::EPI_BIAS_GELU_BF16) {
    if ((r = make_store_tmap(&tc, out, M, N))) return r;
  }
  CU_OK(init_kernel_attributes());
  switch (epi) {
    case gemm::EPI_STORE_BF16:
      launch_gemm<gemm::EPI_STORE_BF16>(block_n, pair != 0, ta, tb, tc, p, sms, st);
      break;
    case gemm::EPI_BIAS_GELU_BF16:
      if (!bias) return fail(BIOM3_ERR_INVALID, "bias required");
      launch_gemm<gemm::EPI_BIAS_GELU_BF16>(block_n, pair != 0, ta, tb, tc, p, sms, st);
      break;
    case gemm::EPI_BIAS_RESID_F32:
      if (!bias) return fail(BIOM3_ERR_INVALID, "bias required");
      launch_gemm<gemm::EPI_BIAS_RESID_F32>(block_n, pair != 0, ta, tb, tc, p, sms, st); break;
    case gemm::EPI_STORE_F32: launch_gemm<gemm::EPI_STORE_F32>(block_n, pair != 0, ta, tb, tc, p, sms, st); break;
    case gemm::EPI_BIAS_RESID_SPLIT:      // out = bf16 [2][M][N]: hi plane then lo plane, updated in place
      if (!bias) return fail(BIOM3_ERR_INVALID, "bias required");
      p.out_bf16 = reinterpret_cast<bf16*>(out);
      p.out = reinterpret_cast<bf16*>(out) + size_t(M) * N;
      launch_gemm<gemm::EPI_BIAS_RESID_SPLIT>(block_n, pair != 0, ta, tb, tc, p, sms, st); break;
    case gemm::EPI_BIAS_GELU_SPLIT: {       // out = bf16 [M][2N]: gelu_erf(A W^T + bias) as hi (columns [0, N)) and lo (columns [N, 2N))
      if (!bias) return fail(BIOM3_ERR_INVALID, "bias required");
      CUtensorMap ts;
      if ((r = make_store_tmap(&ts, out, M, 2 * uint64_t(N)))) return r;
      launch_gemm<gemm::EPI_BIAS_GELU_SPLIT>(block_n, pair != 0, ta, tb, ts, p, sms, st);
      break;
    }
    case gemm::EPI_BIAS_RESID_SPLIT_TMA: {  // same planes, TMA-fed residual ring (pair tiling only)
      if (!bias) return fail(BIOM3_ERR_INVALID, "bias required");
      if (!pair) return fail(BIOM3_ERR_INVALID, "epilogue 6 needs pair tiling");
      CUtensorMap th, tl;
      if ((r = make_store_tmap(&th, out, M, N))) return r;
      if ((r = make_store_tmap(&tl, reinterpret_cast<bf16*>(out) + size_t(M) * N, M, N))) return r;
      p.out_bf16 = reinterpret_cast<bf16*>(out);
      p.out = reinterpret_cast<bf16*>(out) + size_t(M) * N;
      launch_gemm_t<256, gemm::EPI_BIAS_RESID_SPLIT_TMA, true>(ta, tb, th, tl, p, sms, st);
      break;
    }
    default: return fail(BIOM3_ERR_INVALID, "unknown epilogue");
  }
  CU_OK(cudaGetLastError());
  return BIOM3_OK;
}

}  // extern "C"
