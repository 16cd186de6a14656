// Memory-bound row kernels: stand-alone LayerNorm, the cgMLP convolutional spatial gating unit
// (LayerNorm statistics + depthwise conv k=31 over time + gate multiply) and the learned_ave merge
// weights.  All are coalesced along the channel dimension and use warp-shuffle reductions.
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (D <= 2048 -> <= 16 float4 per lane).
// ------------------------------------------------------------------------------------------------
template <int kVec>  // number of float4 per lane: D = 128 * kVec
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long ldx, int M, float eps,
                 const float* __restrict__ gA, const float* __restrict__ bA,
                 float* __restrict__ outA, long long ldA, int roundA,
                 const float* __restrict__ gB, const float* __restrict__ bB,
                 float* __restrict__ outB, long long ldB, int roundB, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint32_t lane = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
  float4 v[kVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = ld_act4(xr + lane + 32 * i);
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float inv_d = 1.0f / (128.0f * kVec);
  const float mean = warp_sum(sum) * inv_d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
  auto emit = [&](const float* g, const float* b, float* out, long long ld, int rnd) {
    float4* o = reinterpret_cast<float4*>(out + static_cast<long long>(row) * ld);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
      float4 y;
      y.x = (v[i].x * rstd * gg.x + bb.x) * scale;
      y.y = (v[i].y * rstd * gg.y + bb.y) * scale;
      y.z = (v[i].z * rstd * gg.z + bb.z) * scale;
      y.w = (v[i].w * rstd * gg.w + bb.w) * scale;
      if (rnd) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
      o[lane + 32 * i] = y;
    }
  };
  if (outA) emit(gA, bA, outA, ldA, roundA);
  if (outB) emit(gB, bB, outB, ldB, roundB);
}

// ------------------------------------------------------------------------------------------------
// CSGU pass 1: LayerNorm statistics of the gate half, one warp per frame.
// ------------------------------------------------------------------------------------------------
template <int kVec>  // Ch = 128 * kVec
__global__ void __launch_bounds__(256)
csgu_stats_kernel(const float* __restrict__ h, long long ldh, int M, int Ch, float eps,
                  float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint32_t lane = lane_id();
  const float4* g = reinterpret_cast<const float4*>(h + static_cast<long long>(row) * ldh + Ch);
  float4 v[kVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = ld_act4(g + lane + 32 * i);
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float inv_d = 1.0f / (128.0f * kVec);
  const float mean = warp_sum(sum) * inv_d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// ------------------------------------------------------------------------------------------------
// CSGU pass 2: out[b,t,c] = r[b,t,c] * (sum_k w[c,k] * LN(g)[b,t+k-15,c] + cb[c])
// CTA = 128 channels x kSeg output frames of one utterance.  The (kSeg+30) x 128 gate tile is
// brought in with cp.async (every request in flight at once: the kernel is bandwidth-, not
// latency-bound), then one thread per channel slides a register window over it: 16 outputs per
// pass from 46 shared-memory reads (conflict-free: lanes <-> consecutive channels) and 496 FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kTaps = 31;
constexpr int kHalo = 15;
constexpr int kSeg = 64;   // output frames per CTA
constexpr int kGrp = 16;   // outputs per register pass
constexpr int kCh = 128;   // channels per CTA
constexpr int kRows = kSeg + 2 * kHalo;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(sz)
               : "memory");
}

__global__ void __launch_bounds__(kCh, 4)
csgu_conv_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ norm_g,
                 const float* __restrict__ norm_b, const float* __restrict__ conv_w,
                 const float* __restrict__ conv_b, const float2* __restrict__ stats,
                 float* __restrict__ out, long long ldo, int T, int Ch, int round_out) {
  __shared__ __align__(16) float s_tile[kRows][kCh];
  __shared__ float2 s_ab[kRows];  // per frame: (rstd, -mean*rstd); (0,0) outside [0,T)
  pdl_launch_dependents();
  // weights first (not produced by the predecessor), then wait for the gate activations / stats
  const int c_early = blockIdx.x * kCh + threadIdx.x;
  float w[kTaps];
  float gam = 0.f, bet = 0.f, cb = 0.f;
  if (c_early < Ch) {
#pragma unroll
    for (int k = 0; k < kTaps; ++k) w[k] = __ldg(conv_w + static_cast<long long>(c_early) * kTaps + k);
    gam = __ldg(norm_g + c_early);
    bet = __ldg(norm_b + c_early);
    cb = __ldg(conv_b + c_early);
  }
  pdl_wait();
  const int c0 = blockIdx.x * kCh;
  const int t0 = blockIdx.y * kSeg;
  const int b = blockIdx.z;
  const long long row0 = static_cast<long long>(b) * T;
  const float* gbase = h + Ch + c0;
  for (int idx = threadIdx.x; idx < kRows * (kCh / 4); idx += kCh) {
    const int r = idx / (kCh / 4), q = idx % (kCh / 4);
    const int t = t0 - kHalo + r;
    const bool ok = t >= 0 && t < T;
    cp_async16_zfill(&s_tile[r][q * 4], gbase + (row0 + (ok ? t : 0)) * ldh + q * 4, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i < kRows; i += kCh) {
    const int t = t0 - kHalo + i;
    float2 ab = make_float2(0.f, 0.f);
    if (t >= 0 && t < T) {
      const float2 st = stats[row0 + t];
      ab = make_float2(st.y, -st.x * st.y);
    }
    s_ab[i] = ab;
  }
  const int c = c0 + threadIdx.x;
  const float* rcol = h + c;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  for (int g0 = 0; g0 < kSeg; g0 += kGrp) {
    const int tb = t0 + g0;  // first output frame of this group
    if (tb >= T) break;
    float rv[kGrp];          // carried half, loaded early so the latency hides behind the FMAs
#pragma unroll
    for (int o = 0; o < kGrp; ++o) rv[o] = tb + o < T ? ld_act(rcol + (row0 + tb + o) * ldh) : 0.f;
    float acc[kGrp];
#pragma unroll
    for (int o = 0; o < kGrp; ++o) acc[o] = 0.f;
#pragma unroll
    for (int ii = 0; ii < kGrp + kTaps - 1; ++ii) {
      const float2 ab = s_ab[g0 + ii];
      // LN(g) = (x - mean) * rstd * gamma + beta; exactly 0 outside [0,T) (conv zero padding)
      const float xh = fmaf(s_tile[g0 + ii][threadIdx.x], ab.x, ab.y);
      const float xn = ab.x != 0.f ? fmaf(xh, gam, bet) : 0.f;
#pragma unroll
      for (int o = 0; o < kGrp; ++o) {
        const int k = ii - o;
        if (k >= 0 && k < kTaps) acc[o] = fmaf(w[k], xn, acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < kGrp; ++o) {
      const int t = tb + o;
      if (t < T) {
        const float y = rv[o] * (acc[o] + cb);
        out[(row0 + t) * ldo + c] = round_out ? round_tf32(y) : y;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row dots: out[m] = (a[m,:] . va, a[m,:] . vb) for up to two activation matrices in one launch -
// the pooling_proj / weight_proj scores of the learned_ave merge (encoder_layer.py:243,258) taken
// directly on the attention context and the gated cgMLP activations, with the branch output
// projections folded into va / vb.  One warp per row, 16-byte coalesced loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_dots_kernel(const float* __restrict__ a1, long long ld1, int K1, const float* __restrict__ va1,
                const float* __restrict__ vb1, float2* __restrict__ out1,
                const float* __restrict__ a2, long long ld2, int K2, const float* __restrict__ va2,
                const float* __restrict__ vb2, float2* __restrict__ out2, int M) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // all loads of a row are issued before the first FMA (8 x 16 B per lane in flight): the kernel
  // is a pure L2 -> register stream, so exposed load latency is the only thing that can slow it
  auto dots2 = [&](const float* a, int K, const float* va, const float* vb, float& sa, float& sb) {
    const float4* r = reinterpret_cast<const float4*>(a);
    const float4* pa = reinterpret_cast<const float4*>(va);
    const float4* pb = reinterpret_cast<const float4*>(vb);
    const int n4 = K / 4;
    for (int k0 = lane; k0 < n4; k0 += 256) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        x[j] = k0 + 32 * j < n4 ? ld_act4(r + k0 + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k0 + 32 * j < n4) {
          const float4 wa = __ldg(pa + k0 + 32 * j);
          const float4 wb = __ldg(pb + k0 + 32 * j);
          sa += x[j].x * wa.x + x[j].y * wa.y + x[j].z * wa.z + x[j].w * wa.w;
          sb += x[j].x * wb.x + x[j].y * wb.y + x[j].z * wb.z + x[j].w * wb.w;
        }
      }
    }
  };
  for (int m = blockIdx.x * 8 + warp; m < M; m += gridDim.x * 8) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    dots2(a1 + static_cast<long long>(m) * ld1, K1, va1, vb1, s[0], s[1]);
    if (a2 != nullptr) dots2(a2 + static_cast<long long>(m) * ld2, K2, va2, vb2, s[2], s[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = warp_sum(s[i]);
    if (lane == 0) {
      out1[m] = make_float2(s[0], s[1]);
      if (a2 != nullptr) out2[m] = make_float2(s[2], s[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// learned_ave merge weights: one warp per utterance.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
merge_weights_kernel(const float2* __restrict__ dots1, const float2* __restrict__ dots2,
                     const int32_t* __restrict__ lens, float pool_b1, float pool_b2, float wproj_b1,
                     float wproj_b2, float inv_sqrt, float* __restrict__ w1, float* __restrict__ w2,
                     int B, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const uint32_t lane = lane_id();
  int len = lens ? lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  float omega[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const float2* d = (br == 0 ? dots1 : dots2) + static_cast<long long>(b) * T;
    const float pb = br == 0 ? pool_b1 : pool_b2;
    float mx = -INFINITY;
    for (int t = lane; t < len; t += 32) mx = fmaxf(mx, (d[t].x + pb) * inv_sqrt);
    mx = warp_max(mx);
    float se = 0.f, sz = 0.f;
    for (int t = lane; t < len; t += 32) {
      const float2 v = d[t];
      const float e = expf((v.x + pb) * inv_sqrt - mx);
      se += e;
      sz += e * v.y;
    }
    se = warp_sum(se);
    sz = warp_sum(sz);
    // len == 0: every score is masked, softmax-then-zero gives an all-zero pooling vector
    omega[br] = (len > 0 ? sz / se : 0.f) + (br == 0 ? wproj_b1 : wproj_b2);
  }
  if (lane == 0) {
    const float m = fmaxf(omega[0], omega[1]);
    const float e0 = expf(omega[0] - m), e1 = expf(omega[1] - m);
    w1[b] = e0 / (e0 + e1);
    w2[b] = e1 / (e0 + e1);
  }
}

}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_layernorm(const float* x, long long ldx, int M, int D, float eps,
                               const float* gA, const float* bA, float* outA, long long ldA,
                               int roundA, const float* gB, const float* bB, float* outB,
                               long long ldB, int roundB, float scale, void* stream) {
  TAVSR_REQUIRE(M > 0 && D > 0 && D % 128 == 0 && D <= 2048, "layernorm: D=%d unsupported", D);
  TAVSR_REQUIRE(ldx % 4 == 0 && (!outA || ldA % 4 == 0) && (!outB || ldB % 4 == 0),
                "layernorm: pitches must be multiples of 4");
  TAVSR_REQUIRE((!outA || (gA && bA)) && (!outB || (gB && bB)), "layernorm: missing affine");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_block = 8;
  const int grid = (M + rows_per_block - 1) / rows_per_block;
#define TAVSR_LN_CASE(V)                                                                          \
  case V:                                                                                         \
    TAVSR_CUDA_OK(launch_kernel(layernorm_kernel<V>, dim3(grid), dim3(256), 0, s, 0, x, ldx, M, eps, \
                                gA, bA, outA, ldA, roundA, gB, bB, outB, ldB, roundB, scale));    \
    break;
  switch (D / 128) {
    TAVSR_LN_CASE(1) TAVSR_LN_CASE(2) TAVSR_LN_CASE(3) TAVSR_LN_CASE(4) TAVSR_LN_CASE(6)
    TAVSR_LN_CASE(8) TAVSR_LN_CASE(16)
    default:
      return set_error(TAVSR_ERR_UNSUPPORTED, "layernorm: D=%d not instantiated", D);
  }
#undef TAVSR_LN_CASE
  TAVSR_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_csgu_fwd(const float* h, long long ldh, const float* norm_g,
                              const float* norm_b, const float* conv_w, const float* conv_b,
                              float* out, long long ldo, float* stats, int B, int T, int Ch,
                              int ksize, float eps, int round_out, void* stream) {
  TAVSR_REQUIRE(ksize == kTaps, "csgu: only kernel size 31 is built (got %d)", ksize);
  TAVSR_REQUIRE(B > 0 && T > 0 && Ch > 0 && Ch % 128 == 0 && Ch <= 2048, "csgu: bad shape");
  TAVSR_REQUIRE(ldh % 4 == 0 && stats != nullptr, "csgu: bad pitch / missing stats scratch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int M = B * T;
  const int grid1 = (M + 7) / 8;
  float2* st = reinterpret_cast<float2*>(stats);
  switch (Ch / 128) {
    case 1: TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<1>, dim3(grid1), dim3(256), 0, s, 0, h, ldh, M, Ch, eps, st)); break;
    case 2: TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<2>, dim3(grid1), dim3(256), 0, s, 0, h, ldh, M, Ch, eps, st)); break;
    case 4: TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<4>, dim3(grid1), dim3(256), 0, s, 0, h, ldh, M, Ch, eps, st)); break;
    case 8: TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<8>, dim3(grid1), dim3(256), 0, s, 0, h, ldh, M, Ch, eps, st)); break;
    case 16: TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<16>, dim3(grid1), dim3(256), 0, s, 0, h, ldh, M, Ch, eps, st)); break;
    default: return set_error(TAVSR_ERR_UNSUPPORTED, "csgu: Ch=%d not instantiated", Ch);
  }
  TAVSR_CUDA_OK(cudaGetLastError());
  dim3 grid2(Ch / kCh, (T + kSeg - 1) / kSeg, B);
  TAVSR_CUDA_OK(launch_kernel(csgu_conv_kernel, grid2, dim3(kCh), 0, s, 0, h, ldh, norm_g, norm_b, conv_w,
                              conv_b, static_cast<const float2*>(st), out, ldo, T, Ch, round_out));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_row_dots(const float* a1, long long ld1, int K1, const float* va1,
                              const float* vb1, float* out1, const float* a2, long long ld2, int K2,
                              const float* va2, const float* vb2, float* out2, int M, void* stream) {
  TAVSR_REQUIRE(M > 0 && a1 && va1 && vb1 && out1 && K1 > 0 && K1 % 4 == 0 && ld1 % 4 == 0,
                "row_dots: bad first operand (K=%d)", K1);
  TAVSR_REQUIRE(!a2 || (va2 && vb2 && out2 && K2 > 0 && K2 % 4 == 0 && ld2 % 4 == 0),
                "row_dots: bad second operand (K=%d)", K2);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = (M + 7) / 8;
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  TAVSR_CUDA_OK(launch_kernel(row_dots_kernel, dim3(grid), dim3(256), 0, s, 0, a1, ld1, K1, va1, vb1,
                              reinterpret_cast<float2*>(out1), a2, ld2, K2, va2, vb2,
                              reinterpret_cast<float2*>(out2), M));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_merge_learned_ave_weights(const float* dots1, const float* dots2,
                                               const int32_t* lens, float pool_b1, float pool_b2,
                                               float wproj_b1, float wproj_b2, float inv_sqrt_size,
                                               float* w1, float* w2, int B, int T, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && dots1 && dots2 && w1 && w2, "merge_weights: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TAVSR_CUDA_OK(launch_kernel(merge_weights_kernel, dim3((B + 3) / 4), dim3(128), 0, s, 0,
                              reinterpret_cast<const float2*>(dots1),
                              reinterpret_cast<const float2*>(dots2), lens, pool_b1, pool_b2,
                              wproj_b1, wproj_b2, inv_sqrt_size, w1, w2, B, T));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
