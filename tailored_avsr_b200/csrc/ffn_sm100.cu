// Host launcher + C ABI of the fused feed-forward kernel (see ffn_sm100.cuh).
#include "ffn_sm100.cuh"

#include <atomic>

#include "host.h"

namespace tavsr {
extern std::atomic<long long> g_launches;
extern void* g_debug_ptr;
int fill_rowln_epilogue(GemmParams& p, const tavsr_rowln_args* a, const char* who);

template <int kAct>
static int launch_ffn(const FfnParams& p, int m_tiles, cudaStream_t stream) {
  auto kern = ffn_fused_kernel<kAct>;
  static bool configured = false;
  if (!configured) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ffn::kSmemBytesV1));
    configured = true;
  }
  TAVSR_CUDA_OK(launch_kernel(kern, dim3(2 * m_tiles), dim3(ffn::kThreadsV1), ffn::kSmemBytesV1, stream, 0, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
template <int kAct>
static int launch_ffn2(const Ffn2Params& p, int m_units, cudaStream_t stream) {
  auto kern = ffn_fused_pair_kernel<kAct>;
  static bool configured = false;
  if (!configured) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ffn2::kSmemBytes));
    configured = true;
  }
  if (g_debug[6]) {
    int ncl = 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(4 * m_units);
    cfg.blockDim = dim3(ffn::kThreads);
    cfg.dynamicSmemBytes = ffn2::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    printf("ffn2: max active clusters of 4 = %d (err %d)\n", ncl, static_cast<int>(e));
    g_debug[6] = 0;
  }
  TAVSR_CUDA_OK(launch_kernel(kern, dim3(4 * m_units), dim3(ffn::kThreads), ffn2::kSmemBytes, stream, 0, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_ffn_fused(const tavsr_ffn_args* a, void* stream) {
  TAVSR_REQUIRE(a != nullptr && a->struct_size == static_cast<int>(sizeof(tavsr_ffn_args)),
                "ffn: bad args struct (size %d, expected %d)", a ? a->struct_size : -1,
                static_cast<int>(sizeof(tavsr_ffn_args)));
  TAVSR_REQUIRE(a->hidden == ffn::kHid, "ffn: only hidden = 2048 is built (got %d)", a->hidden);
  TAVSR_REQUIRE(a->xn && a->w1 && a->w2 && a->ep.M > 0, "ffn: xn, w1, w2 and M are required");
  TAVSR_REQUIRE(a->ep.x2 == nullptr, "ffn: dual operands are not supported");
  int rc;
  if (g_debug[5] == 2) {
    // v2 (opt-in, measured slower than v1 in round 1: its N=128 pair MMAs are issue-bound, see
    // tools/mma_bench.cu): CTA pairs, cluster of 4
    Ffn2Params q;
    memset(&q, 0, sizeof(q));
    if ((rc = fill_rowln_epilogue(q.ep, &a->ep, "ffn"))) return rc;
    const int M2 = a->ep.M;
    if ((rc = make_tmap_2d(&q.tmX, a->xn, 4, false, M2, ffn::kD, a->ldxn, 128, 32))) return rc;
    if ((rc = make_tmap_2d(&q.tmW1, a->w1, 4, false, ffn::kHid, ffn::kD, a->ldw1, 64, 32))) return rc;
    if ((rc = make_tmap_2d(&q.tmW2, a->w2, 4, false, ffn::kD, ffn::kHid, a->ldw2, 64, 32))) return rc;
    q.b1 = a->b1;
    q.act = a->act;
    q.dbg = reinterpret_cast<long long*>(g_debug_ptr);
    const int m_units = (M2 + 255) / 256;
    cudaStream_t s2 = static_cast<cudaStream_t>(stream);
    switch (a->act) {
      case ACT_SWISH: return launch_ffn2<ACT_SWISH>(q, m_units, s2);
      case ACT_RELU: return launch_ffn2<ACT_RELU>(q, m_units, s2);
      case ACT_GELU: return launch_ffn2<ACT_GELU>(q, m_units, s2);
      default: return launch_ffn2<ACT_NONE>(q, m_units, s2);
    }
  }
  FfnParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = fill_rowln_epilogue(p.ep, &a->ep, "ffn"))) return rc;
  const int M = a->ep.M;
  if ((rc = make_tmap_2d(&p.tmX, a->xn, 4, false, M, ffn::kD, a->ldxn, 128, 32))) return rc;
  if ((rc = make_tmap_2d(&p.tmW1, a->w1, 4, false, ffn::kHid, ffn::kD, a->ldw1, 128, 32))) return rc;
  if ((rc = make_tmap_2d(&p.tmW2, a->w2, 4, false, ffn::kD, ffn::kHid, a->ldw2, 128, 32))) return rc;
  p.b1 = a->b1;
  p.act = a->act;
  p.dbg = reinterpret_cast<long long*>(g_debug_ptr);
  const int m_tiles = (M + 127) / 128;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->act) {
    case ACT_SWISH: return launch_ffn<ACT_SWISH>(p, m_tiles, s);
    case ACT_RELU: return launch_ffn<ACT_RELU>(p, m_tiles, s);
    case ACT_GELU: return launch_ffn<ACT_GELU>(p, m_tiles, s);
    default: return launch_ffn<ACT_NONE>(p, m_tiles, s);
  }
}
