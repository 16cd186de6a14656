// Host launchers + C ABI for the tcgen05 GEMM family (see gemm_sm100.cuh).
#include "gemm_sm100.cuh"

#include <atomic>

#include "host.h"

namespace tavsr {

thread_local char g_last_error[512] = "";
int g_debug[16] = {0};
void* g_debug_ptr = nullptr;  // optional device buffer for kernel phase timestamps
std::atomic<long long> g_launches{0};

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols;
  int elem_bytes, flags;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld &&
           box_rows == o.box_rows && box_cols == o.box_cols && elem_bytes == o.elem_bytes &&
           flags == o.flags;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_rows); mix(k.box_cols);
    mix(static_cast<uint64_t>(k.elem_bytes)); mix(static_cast<uint64_t>(k.flags));
    return h;
  }
};

int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, bool is_bf16, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                 bool is_load) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const bool l2_promote_256 = is_load;
  // Operand loads use the TFLOAT32 map type: TMA then rounds fp32 -> tf32 to nearest while
  // copying into shared memory (measured: GEMM error 2.0e-4 vs 5.3e-4 with truncation), so
  // producers do not have to pre-round.  g_debug[1] = 1 switches back to plain FLOAT32.
  const int tf32_type = (is_load && !is_bf16 && g_debug[1] == 0) ? 1 : 0;
  TmapKey key{ptr, rows, cols, ld, box_rows, box_cols, elem_bytes,
              (is_bf16 ? 1 : 0) | (l2_promote_256 ? 2 : 0) | (tf32_type << 2)};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(TAVSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * elem_bytes) % 16 != 0)
    return set_error(TAVSR_ERR_INVALID,
                     "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch "
                     "(ptr=%p ld=%llu elem=%d)",
                     ptr, static_cast<unsigned long long>(ld), elem_bytes);
  if (box_cols * elem_bytes != 128 || box_rows > 256)
    return set_error(TAVSR_ERR_INVALID, "bad TMA box %u x %u", box_rows, box_cols);
  CUtensorMapDataType dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : (tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   l2_promote_256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                  : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TAVSR_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u",
                     static_cast<int>(r), static_cast<unsigned long long>(rows),
                     static_cast<unsigned long long>(cols), static_cast<unsigned long long>(ld),
                     box_rows, box_cols);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
template <bool kTf32, int kBlockN, int kMode, bool kDual, int kCtas, int kAct, bool kSeq = false,
          bool kWideEpi = false, int kAct2 = kNoGroup>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kTf32, kBlockN, kMode, kDual, kCtas, kSeq, kWideEpi>;
  auto kern = gemm_sm100_kernel<kTf32, kBlockN, kMode, kDual, kCtas, kAct, kSeq, kWideEpi, kAct2>;
  static unsigned long long configured = 0;  // one bit per device ordinal
  if (first_use_on_device(configured))
    TAVSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  int units = p.num_m_tiles * (p.num_n_tiles + (kAct2 != kNoGroup ? p.num_n_tiles2 : 0));
  if (kMode == kModeTiled && kAct2 == kNoGroup && p.ksplit > 1) units *= p.ksplit;
  const int max_units = num_sms() / kCtas;
  const int grid = (units < max_units ? units : max_units) * kCtas;
  TAVSR_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, kCtas, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Tiled GEMM dispatch over (operand type, tile width, epilogue width, activation): CTA pairs only.
template <bool kTf32, int kBlockN, bool kWide>
static int launch_tiled(const GemmParams& p, cudaStream_t s) {
  switch (p.act) {
    case ACT_NONE: return launch_gemm<kTf32, kBlockN, kModeTiled, false, 2, ACT_NONE, false, kWide>(p, s);
    case ACT_SWISH: return launch_gemm<kTf32, kBlockN, kModeTiled, false, 2, ACT_SWISH, false, kWide>(p, s);
    case ACT_GELU: return launch_gemm<kTf32, kBlockN, kModeTiled, false, 2, ACT_GELU, false, kWide>(p, s);
    default: return launch_gemm<kTf32, kBlockN, kModeTiled, false, 2, -1, false, kWide>(p, s);
  }
}

}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_version(void) { return TAVSR_VERSION; }
extern "C" const char* tavsr_last_error(void) { return g_last_error; }
extern "C" int tavsr_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return TAVSR_ERR_INVALID;
  g_debug[key] = value;
  return 0;
}
extern "C" long long tavsr_launch_count(void) { return g_launches.load(); }
extern "C" int tavsr_debug_set_ptr(void* p) {
  g_debug_ptr = p;
  return 0;
}

extern "C" int tavsr_gemm_bias_act(const void* x, long long ldx, const void* w, long long ldw,
                                   const float* bias, void* y, long long ldy, int M, int N, int K,
                                   int act, int round_out, int dtype, void* stream) {
  TAVSR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  const int op = dtype & TAVSR_DT_MASK;
  TAVSR_REQUIRE(op == TAVSR_DT_TF32 || op == TAVSR_DT_BF16,
                "gemm: operand dtype must be TAVSR_DT_TF32 or TAVSR_DT_BF16 (tf32x3 is composed by "
                "the caller from tavsr_split_tf32 + a K-tripled TF32 product)");
  const bool bf16 = op == TAVSR_DT_BF16;
  const bool out_bf16 = (dtype & TAVSR_DT_OUT_BF16) != 0;
  TAVSR_REQUIRE(bf16 || !out_bf16, "gemm: bf16 output only with bf16 operands");
  const int eb = bf16 ? 2 : 4;  // operand element bytes; a 128-byte swizzle row holds 128 / eb
  TAVSR_REQUIRE(K % (16 / eb) == 0 && N % (out_bf16 ? 8 : 4) == 0,
                "gemm: K must be a multiple of %d and N of %d (K=%d N=%d)", 16 / eb,
                out_bf16 ? 8 : 4, K, N);
  TAVSR_REQUIRE(!bf16 || !round_out, "gemm: TF32 output rounding is a tf32-mode option");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.act = act; p.round_c = round_out;
  p.c_bf16 = out_bf16;
  // CTA pairs (cta_group::2): a pair covers 256 rows and shares the B tile, which halves the
  // per-SM weight traffic and makes room for a deeper pipeline.
  const int ctas = 2;
  const int mt = (M + 255) / 256;
  const int slots = num_sms() / ctas;
  // tile-width heuristic: fewer waves wins (a 128-wide tile costs about half a 256-wide one)
  const int t256 = mt * ((N + 255) / 256), t128 = mt * ((N + 127) / 128);
  const double cost256 = static_cast<double>((t256 + slots - 1) / slots) * 2.0;
  // (a 128-wide tile moves 1.5x the operand bytes per output element, hence the 1.25 penalty)
  const double cost128 = static_cast<double>((t128 + slots - 1) / slots) * 1.25;
  const bool use128 = (N <= 128) || (cost128 < cost256) || g_debug[2] == 128;
  const int bn = (use128 && g_debug[2] != 256) ? 128 : 256;
  p.num_m_tiles = mt;
  p.num_n_tiles = (N + bn - 1) / bn;
  int rc;
  if ((rc = make_tmap_2d(&p.tmA, x, eb, bf16, M, K, ldx, 128, 128 / eb))) return rc;
  if ((rc = make_tmap_2d(&p.tmB, w, eb, bf16, N, K, ldw, bn / ctas, 128 / eb))) return rc;
  if (out_bf16) {
    if ((rc = make_tmap_2d(&p.tmC, y, 2, true, M, N, ldy, 32, 64, false))) return rc;
  } else {
    if ((rc = make_tmap_2d(&p.tmC, y, 4, false, M, N, ldy, 32, 32, false))) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // 16-warp epilogue (GemmCfg::kWideEpi) on the 256-wide tiles: the K = 256 projections are bound
  // by their epilogue's latency (+3.7 % on the C2 step).  g_debug[13] = 1 falls back to 8 warps.
  const bool wide = bn == 256 && g_debug[13] == 0;
  if (bf16) {
    if (bn == 128) return launch_tiled<false, 128, false>(p, s);
    return wide ? launch_tiled<false, 256, true>(p, s) : launch_tiled<false, 256, false>(p, s);
  }
  if (bn == 128) return launch_tiled<true, 128, false>(p, s);
  return wide ? launch_tiled<true, 256, true>(p, s) : launch_tiled<true, 256, false>(p, s);
}

// ------------------------------------------------------------------------------------------------
// Weight-gradient product: out[M, N] = A[M, K] . B[N, K]^T with a LONG reduction axis (K = all frames
// of the batch) and few output tiles (M, N = layer widths): the K axis is split over `ksplit` CTA
// pairs per output tile (each writes its partial tile into a slab of the workspace), then
// sum_slabs_kernel adds the slabs in a fixed order (bit-reproducible, no atomics).  Without the
// split a 2048 x 256 gradient ran on 32 of 148 SMs for 82 us, a 256 x 256 one on 4.
// ------------------------------------------------------------------------------------------------
namespace tavsr {
__global__ void __launch_bounds__(256)
sum_slabs_kernel(const float* __restrict__ ws, long long slab_stride, int nslab, float* __restrict__ out,
                 long long ldo, int M, int N4) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * N4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / N4), q = static_cast<int>(i % N4);
    const float4* src = reinterpret_cast<const float4*>(ws + static_cast<long long>(m) * (4 * N4)) + q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sidx = 0; sidx < nslab; ++sidx) {
      const float4 v = ld_act4(reinterpret_cast<const float4*>(
          reinterpret_cast<const float*>(src) + sidx * slab_stride));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out + m * ldo)[q] = acc;
  }
}

// split plan shared by the workspace query and the launch
static void wgrad_plan(int M, int N, int K, int* ksplit, int* split_kb, int* slab_rows, int* bn) {
  const int mt = (M + 255) / 256;
  const int slots = num_sms() / 2;
  *bn = N <= 128 ? 128 : 256;
  const int tiles = mt * ((N + *bn - 1) / *bn);
  const int num_kb = (K + 31) / 32;
  int want = tiles >= slots ? 1 : slots / tiles;          // fill the machine once
  const int max_by_k = num_kb / 8 > 0 ? num_kb / 8 : 1;   // >= 8 k-blocks (256 frames) per split
  if (want > max_by_k) want = max_by_k;
  if (want > 64) want = 64;
  const int kb = (num_kb + want - 1) / want;
  *split_kb = kb;
  *ksplit = (num_kb + kb - 1) / kb;
  *slab_rows = mt * 256;
}
}  // namespace tavsr

extern "C" size_t tavsr_gemm_wgrad_workspace_bytes(int M, int N, int K) {
  int ks, kb, rows, bn;
  wgrad_plan(M, N, K, &ks, &kb, &rows, &bn);
  return ks > 1 ? static_cast<size_t>(ks) * rows * N * sizeof(float) : 0;
}

extern "C" int tavsr_gemm_wgrad(const float* a, long long lda, const float* b, long long ldb,
                                float* out, long long ldo, int M, int N, int K, void* workspace,
                                long long workspace_bytes, void* stream) {
  TAVSR_REQUIRE(M > 0 && N > 0 && K > 0 && a && b && out, "gemm_wgrad: bad arguments");
  TAVSR_REQUIRE(K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldo % 4 == 0,
                "gemm_wgrad: K, N and the pitches must be multiples of 4 (K=%d N=%d)", K, N);
  int ks, kb, rows, bn;
  wgrad_plan(M, N, K, &ks, &kb, &rows, &bn);
  TAVSR_REQUIRE(ks == 1 || (workspace && static_cast<size_t>(workspace_bytes) >=
                                             tavsr_gemm_wgrad_workspace_bytes(M, N, K)),
                "gemm_wgrad: workspace too small");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.act = ACT_NONE;
  p.num_m_tiles = (M + 255) / 256;
  p.num_n_tiles = (N + bn - 1) / bn;
  p.ksplit = ks; p.split_kb = kb; p.slab_rows = rows;
  int rc;
  if ((rc = make_tmap_2d(&p.tmA, a, 4, false, M, K, lda, 128, 32))) return rc;
  if ((rc = make_tmap_2d(&p.tmB, b, 4, false, N, K, ldb, bn / 2, 32))) return rc;
  float* c = ks > 1 ? static_cast<float*>(workspace) : out;
  const long long ldc = ks > 1 ? N : ldo;
  const unsigned long long c_rows = ks > 1 ? static_cast<unsigned long long>(ks) * rows : M;
  if ((rc = make_tmap_2d(&p.tmC, c, 4, false, c_rows, N, ldc, 32, 32, false))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = bn == 128 ? launch_tiled<true, 128, false>(p, s)
                 : (g_debug[13] == 0 ? launch_tiled<true, 256, true>(p, s) : launch_tiled<true, 256, false>(p, s));
  if (rc || ks == 1) return rc;
  const long long total = static_cast<long long>(M) * (N / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
  TAVSR_CUDA_OK(launch_kernel(sum_slabs_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, 0,
                              static_cast<const float*>(c), static_cast<long long>(rows) * N, ks, out,
                              ldo, M, N / 4));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Two projections of the same rows in ONE launch (the fused QKV projection and channel_proj1 + GELU
// of a two-branch block both read LayerNorm outputs of the same x): the tile scheduler walks both
// problems' 256-wide tiles, so the tail of the first GEMM and the head of the second share a wave
// and one launch + prologue disappears per block.  Activations: problem 1 none, problem 2 GELU
// (the only grouped pair the layer needs); same M, same K.
extern "C" int tavsr_gemm_group2(const void* x1, long long ldx1, const void* w1, long long ldw1,
                                 const float* bias1, void* y1, long long ldy1, int N1, const void* x2,
                                 long long ldx2, const void* w2, long long ldw2, const float* bias2,
                                 void* y2, long long ldy2, int N2, int M, int K, int dtype,
                                 void* stream) {
  TAVSR_REQUIRE(M > 0 && N1 > 0 && N2 > 0 && K > 0, "gemm_group2: empty problem");
  const int op = dtype & TAVSR_DT_MASK;
  TAVSR_REQUIRE(op == TAVSR_DT_TF32 || op == TAVSR_DT_BF16, "gemm_group2: dtype must be tf32 or bf16");
  const bool bf16 = op == TAVSR_DT_BF16;
  const bool out_bf16 = (dtype & TAVSR_DT_OUT_BF16) != 0;
  TAVSR_REQUIRE(bf16 || !out_bf16, "gemm_group2: bf16 output only with bf16 operands");
  const int eb = bf16 ? 2 : 4;
  TAVSR_REQUIRE(K % (16 / eb) == 0 && N1 % 8 == 0 && N2 % 8 == 0, "gemm_group2: K / N alignment");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N1; p.K = K; p.N2 = N2;
  p.bias = bias1; p.bias2 = bias2; p.act = ACT_NONE;
  p.c_bf16 = out_bf16;
  p.num_m_tiles = (M + 255) / 256;
  p.num_n_tiles = (N1 + 255) / 256;
  p.num_n_tiles2 = (N2 + 255) / 256;
  int rc;
  if ((rc = make_tmap_2d(&p.tmA, x1, eb, bf16, M, K, ldx1, 128, 128 / eb))) return rc;
  if ((rc = make_tmap_2d(&p.tmA2, x2, eb, bf16, M, K, ldx2, 128, 128 / eb))) return rc;
  if ((rc = make_tmap_2d(&p.tmB, w1, eb, bf16, N1, K, ldw1, 128, 128 / eb))) return rc;
  if ((rc = make_tmap_2d(&p.tmB2, w2, eb, bf16, N2, K, ldw2, 128, 128 / eb))) return rc;
  if (out_bf16) {
    if ((rc = make_tmap_2d(&p.tmC, y1, 2, true, M, N1, ldy1, 32, 64, false))) return rc;
    if ((rc = make_tmap_2d(&p.tmC2, y2, 2, true, M, N2, ldy2, 32, 64, false))) return rc;
  } else {
    if ((rc = make_tmap_2d(&p.tmC, y1, 4, false, M, N1, ldy1, 32, 32, false))) return rc;
    if ((rc = make_tmap_2d(&p.tmC2, y2, 4, false, M, N2, ldy2, 32, 32, false))) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bf16 ? launch_gemm<false, 256, kModeTiled, false, 2, ACT_NONE, false, true, ACT_GELU>(p, s)
              : launch_gemm<true, 256, kModeTiled, false, 2, ACT_NONE, false, true, ACT_GELU>(p, s);
}

extern "C" size_t tavsr_rowln_workspace_bytes(int M) {
  const size_t units = (static_cast<size_t>(M) + 255) / 256;
  return 4096 /* flag words (split-K is only used for <= 37 row units = 296 flags) */ +
         units * 256 * 256 * sizeof(float);
}

namespace tavsr {
// Validates the epilogue part of a rowln description and fills the matching GemmParams fields
// (shared by tavsr_gemm_rowln and tavsr_ffn_fused).
int fill_rowln_epilogue(GemmParams& p, const tavsr_rowln_args* a, const char* who) {
  TAVSR_REQUIRE(!(a->ln0_g && !a->ln0_b) && !(a->lnA_g && !(a->lnA_b && a->out_lnA)) &&
                    !(a->lnB_g && !(a->lnB_b && a->out_lnB)),
                "%s: LayerNorm stages need gamma, beta and an output", who);
  TAVSR_REQUIRE(!a->dots_out || (a->dot1 && a->dot2), "%s: dots_out needs dot1 and dot2", who);
  TAVSR_REQUIRE(!a->residual || a->ldr % 4 == 0, "%s: residual pitch must be a multiple of 4", who);
  const bool bf16 = (a->dtype & TAVSR_DT_MASK) == TAVSR_DT_BF16;
  TAVSR_REQUIRE(bf16 || !(a->dtype & (TAVSR_DT_OUT_BF16 | TAVSR_DT_LNA_BF16 | TAVSR_DT_LNB_BF16)),
                "%s: bf16 outputs only with bf16 operands", who);
  p.M = a->M; p.N = 256;
  p.bias = a->bias;
  p.act = ACT_NONE;
  p.round_c = a->round_main;
  p.residual = a->residual; p.ldr = a->ldr; p.alpha = a->alpha;
  p.ln0_g = a->ln0_g; p.ln0_b = a->ln0_b;
  p.lnA_g = a->lnA_g; p.lnA_b = a->lnA_b; p.lnB_g = a->lnB_g; p.lnB_b = a->lnB_b;
  p.has_main = a->out_main != nullptr;
  p.round_lnA = a->round_lnA; p.round_lnB = a->round_lnB;
  p.dot1 = a->dot1; p.dot2 = a->dot2; p.dots_out = a->dots_out;
  p.eps = a->eps;
  p.eps0 = a->eps0;
  p.out_main = a->out_main; p.ld_main = a->ld_main;
  p.out_lnA = a->out_lnA; p.ld_lnA = a->ld_lnA;
  p.out_lnB = a->out_lnB; p.ld_lnB = a->ld_lnB;
  p.c_bf16 = (a->dtype & TAVSR_DT_OUT_BF16) != 0;
  p.lnA_bf16 = (a->dtype & TAVSR_DT_LNA_BF16) != 0;
  p.lnB_bf16 = (a->dtype & TAVSR_DT_LNB_BF16) != 0;
  auto out_map = [&](CUtensorMap* tm, void* ptr, long long ld, bool as_bf16) {
    return as_bf16 ? make_tmap_2d(tm, ptr, 2, true, a->M, 256, ld, 32, 64, false)
                   : make_tmap_2d(tm, ptr, 4, false, a->M, 256, ld, 32, 32, false);
  };
  int rc;
  if (a->out_main && (rc = out_map(&p.tmC, a->out_main, a->ld_main, p.c_bf16))) return rc;
  if (a->lnA_g && (rc = out_map(&p.tmLnA, a->out_lnA, a->ld_lnA, p.lnA_bf16))) return rc;
  if (a->lnB_g && (rc = out_map(&p.tmLnB, a->out_lnB, a->ld_lnB, p.lnB_bf16))) return rc;
  return 0;
}
}  // namespace tavsr

extern "C" int tavsr_gemm_rowln(const tavsr_rowln_args* a, void* stream) {
  TAVSR_REQUIRE(a != nullptr && a->struct_size == static_cast<int>(sizeof(tavsr_rowln_args)),
                "rowln: bad args struct (size %d, expected %d)", a ? a->struct_size : -1,
                static_cast<int>(sizeof(tavsr_rowln_args)));
  const int op = a->dtype & TAVSR_DT_MASK;
  TAVSR_REQUIRE(op == TAVSR_DT_TF32 || op == TAVSR_DT_BF16,
                "rowln: operand dtype must be TAVSR_DT_TF32 or TAVSR_DT_BF16");
  const bool bf16 = op == TAVSR_DT_BF16;
  const int eb = bf16 ? 2 : 4;
  const int bk = 128 / eb;  // elements per 128-byte k-block
  TAVSR_REQUIRE(a->M > 0 && a->K > 0 && a->K % (16 / eb) == 0, "rowln: bad shape M=%d K=%d", a->M, a->K);
  TAVSR_REQUIRE(a->x && a->w, "rowln: x and w are required");
  const bool dual = a->x2 != nullptr;
  TAVSR_REQUIRE(!dual || (a->rowscale1 && a->rowscale2 && a->rows_per_seg > 0),
                "rowln: dual mode needs rowscale1/2 and rows_per_seg");
  const bool seq = dual && a->k1 > 0;
  TAVSR_REQUIRE(a->k1 == 0 || (dual && a->k1 % bk == 0 && a->k1 < a->K && (a->K - a->k1) % bk == 0),
                "rowln: sequential dual mode needs x2 and k1, K-k1 multiples of %d (k1=%d K=%d)", bk,
                a->k1, a->K);
  TAVSR_REQUIRE(!(a->segbias1 || a->segbias2) || (seq && a->segbias1 && a->segbias2 && !a->dots_out),
                "rowln: segbias1/2 come as a pair, only in sequential dual mode, without dots");
  TAVSR_REQUIRE(!bf16 || !dual || seq, "rowln: the bf16 kernel has the sequential dual mode only");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = fill_rowln_epilogue(p, a, "rowln"))) return rc;
  const int ctas = 2;
  p.K = a->K;
  p.num_m_tiles = (a->M + 128 * ctas - 1) / (128 * ctas);
  p.num_n_tiles = 1;
  p.rowscale1 = a->rowscale1; p.rowscale2 = a->rowscale2; p.rows_per_seg = a->rows_per_seg;
  // split-K over two CTA pairs when the row tiles alone would leave >= half of the SMs idle
  // (measured slower than the unsplit kernel at M = 8000 in round 1 - the fused FFN kernel is the
  // real fix for K = 2048 - so it is opt-in: g_debug[4] = 1)
  if (!dual && a->workspace != nullptr && g_debug[4] == 1 && a->K >= 1024 &&
      a->K % (2 * bk) == 0 && 2 * p.num_m_tiles <= num_sms() / 2 &&
      static_cast<size_t>(a->workspace_bytes) >= tavsr_rowln_workspace_bytes(a->M)) {
    p.num_n_tiles = 2;
    // flag words live at a FIXED offset (start of the scratch) so that calls with different M
    // never see an earlier call's partial sums where they expect zeroed flags
    p.flags = static_cast<unsigned int*>(a->workspace);
    p.partial = reinterpret_cast<float*>(static_cast<char*>(a->workspace) + 4096);
  }
  const int ka1 = seq ? a->k1 : a->K;
  const int ka2 = seq ? a->K - a->k1 : a->K;
  if (seq) {
    p.seq_kb1 = a->k1 / bk;
    if (a->segbias1) {
      p.seg_bias = 1;
      p.dot1 = a->segbias1;
      p.dot2 = a->segbias2;
    }
  }
  if ((rc = make_tmap_2d(&p.tmA, a->x, eb, bf16, a->M, ka1, a->ldx, 128, bk))) return rc;
  if (dual && (rc = make_tmap_2d(&p.tmA2, a->x2, eb, bf16, a->M, ka2, a->ldx2, 128, bk))) return rc;
  if ((rc = make_tmap_2d(&p.tmB, a->w, eb, bf16, 256, a->K, a->ldw, 256 / ctas, bk))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bf16) {
    if (seq) return launch_gemm<false, 256, kModeRowLN, true, 2, 0, true>(p, s);
    return launch_gemm<false, 256, kModeRowLN, false, 2, 0>(p, s);
  }
  if (seq) return launch_gemm<true, 256, kModeRowLN, true, 2, 0, true>(p, s);
  if (dual) return launch_gemm<true, 256, kModeRowLN, true, 2, 0>(p, s);
  return launch_gemm<true, 256, kModeRowLN, false, 2, 0>(p, s);
}
