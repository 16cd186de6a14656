// Memory-bound row kernels: stand-alone LayerNorm, the cgMLP convolutional spatial gating unit
// (LayerNorm statistics + depthwise conv k=31 over time + gate multiply) and the learned_ave merge
// weights.  All are coalesced along the channel dimension and use warp-shuffle reductions.
#include <atomic>
#include <type_traits>

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
                 void* __restrict__ outA, long long ldA, int roundA,
                 const float* __restrict__ gB, const float* __restrict__ bB,
                 void* __restrict__ outB, long long ldB, int roundB, float scale, int bf16_mask) {
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
  auto emit = [&](const float* g, const float* b, void* out, long long ld, int rnd, bool as_bf16) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(out) + static_cast<long long>(row) * ld);
    uint2* ob = reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + static_cast<long long>(row) * ld);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
      float4 y;
      y.x = (v[i].x * rstd * gg.x + bb.x) * scale;
      y.y = (v[i].y * rstd * gg.y + bb.y) * scale;
      y.z = (v[i].z * rstd * gg.z + bb.z) * scale;
      y.w = (v[i].w * rstd * gg.w + bb.w) * scale;
      if (as_bf16) {
        ob[lane + 32 * i] = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
        continue;
      }
      if (rnd) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
      o[lane + 32 * i] = y;
    }
  };
  if (outA) emit(gA, bA, outA, ldA, roundA, (bf16_mask & 1) != 0);
  if (outB) emit(gB, bB, outB, ldB, roundB, (bf16_mask & 2) != 0);
}

// 4 consecutive activation elements as float4, fp32 or bf16 storage (index in units of 4 elements)
template <bool kBf16>
__device__ __forceinline__ float4 ld_act_vec4(const void* base, long long idx4) {
  if constexpr (kBf16) {
    const uint2 w = ld_act_u2(static_cast<const uint2*>(base) + idx4);
    return make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
  } else {
    return ld_act4(static_cast<const float4*>(base) + idx4);
  }
}
template <bool kBf16>
__device__ __forceinline__ const void* elem_ptr(const void* base, long long elem) {
  return kBf16 ? static_cast<const void*>(static_cast<const uint16_t*>(base) + elem)
               : static_cast<const void*>(static_cast<const float*>(base) + elem);
}

// ------------------------------------------------------------------------------------------------
// CSGU pass 1: LayerNorm statistics of the gate half, one warp per frame.
// ------------------------------------------------------------------------------------------------
template <int kVec, bool kBf16>  // Ch = 128 * kVec
__global__ void __launch_bounds__(256)
csgu_stats_kernel(const void* __restrict__ h, long long ldh, int M, int Ch, float eps,
                  float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint32_t lane = lane_id();
  const void* g = elem_ptr<kBf16>(h, static_cast<long long>(row) * ldh + Ch);
  float4 v[kVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = ld_act_vec4<kBf16>(g, lane + 32 * i);
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

// kBf16: h and out are bf16 (the gate tile sits in shared memory as bf16, 24 KB instead of 47 KB);
// LayerNorm, the convolution and the gate product stay fp32.
template <bool kBf16>
__global__ void __launch_bounds__(kCh, 4)
csgu_conv_kernel(const void* __restrict__ h, long long ldh, const float* __restrict__ norm_g,
                 const float* __restrict__ norm_b, const float* __restrict__ conv_w,
                 const float* __restrict__ conv_b, const float2* __restrict__ stats,
                 void* __restrict__ out, long long ldo, int T, int Ch, int round_out) {
  using elem_t = typename std::conditional<kBf16, uint16_t, float>::type;
  constexpr int kPer16 = 16 / static_cast<int>(sizeof(elem_t));  // elements per 16-byte request
  __shared__ __align__(16) elem_t s_tile[kRows][kCh];
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
  const elem_t* hb = static_cast<const elem_t*>(h);
  const elem_t* gbase = hb + Ch + c0;
  for (int idx = threadIdx.x; idx < kRows * (kCh / kPer16); idx += kCh) {
    const int r = idx / (kCh / kPer16), q = idx % (kCh / kPer16);
    const int t = t0 - kHalo + r;
    const bool ok = t >= 0 && t < T;
    cp_async16_zfill(&s_tile[r][q * kPer16], gbase + (row0 + (ok ? t : 0)) * ldh + q * kPer16, ok);
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
  const elem_t* rcol = hb + c;
  elem_t* ocol = static_cast<elem_t*>(out) + c;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  for (int g0 = 0; g0 < kSeg; g0 += kGrp) {
    const int tb = t0 + g0;  // first output frame of this group
    if (tb >= T) break;
    float rv[kGrp];          // carried half, loaded early so the latency hides behind the FMAs
#pragma unroll
    for (int o = 0; o < kGrp; ++o) {
      if constexpr (kBf16)
        rv[o] = tb + o < T ? __uint_as_float(static_cast<uint32_t>(ld_act_u16(rcol + (row0 + tb + o) * ldh)) << 16) : 0.f;
      else
        rv[o] = tb + o < T ? ld_act(rcol + (row0 + tb + o) * ldh) : 0.f;
    }
    float acc[kGrp];
#pragma unroll
    for (int o = 0; o < kGrp; ++o) acc[o] = 0.f;
#pragma unroll
    for (int ii = 0; ii < kGrp + kTaps - 1; ++ii) {
      const float2 ab = s_ab[g0 + ii];
      float xv;
      if constexpr (kBf16) xv = __uint_as_float(static_cast<uint32_t>(s_tile[g0 + ii][threadIdx.x]) << 16);
      else xv = s_tile[g0 + ii][threadIdx.x];
      // LN(g) = (x - mean) * rstd * gamma + beta; exactly 0 outside [0,T) (conv zero padding)
      const float xh = fmaf(xv, ab.x, ab.y);
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
        if constexpr (kBf16) {
          ocol[(row0 + t) * ldo] = static_cast<uint16_t>(pack_bf16x2(y, 0.f) & 0xFFFFu);
        } else {
          ocol[(row0 + t) * ldo] = round_out ? round_tf32(y) : y;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One-pass CSGU: the LayerNorm statistics come from the SAME shared-memory tile the convolution
// reads, so the gate half is fetched from HBM once and the stand-alone statistics launch disappears
// (it re-read 32.8 MB per block at C2 and cost a launch on the critical path).  The statistics need
// all Ch channels of a frame while a CTA holds 128: the Ch / 128 CTAs of a frame segment form a
// thread-block CLUSTER, each computes two-pass partials (mean_c, M2_c) of its 128 channels for its
// kRows frames, pushes them into every peer's shared memory over DSMEM, and after one cluster
// barrier combines the partials (Chan's parallel variance: equal counts per part).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSlabs = 8;   // portable cluster size: Ch <= 1024

template <bool kBf16>
__global__ void __launch_bounds__(kCh, 4)
csgu_onepass_kernel(const void* __restrict__ h, long long ldh, const float* __restrict__ norm_g,
                    const float* __restrict__ norm_b, const float* __restrict__ conv_w,
                    const float* __restrict__ conv_b, float2* __restrict__ stats_out, float eps,
                    void* __restrict__ out, long long ldo, int T, int Ch, int round_out) {
  using elem_t = typename std::conditional<kBf16, uint16_t, float>::type;
  constexpr int kPer16 = 16 / static_cast<int>(sizeof(elem_t));
  // dynamic shared memory (the fp32 tile alone is 47 KB): tile | partials | per-frame (a, b)
  extern __shared__ __align__(16) uint8_t s_dyn[];
  elem_t (*s_tile)[kCh] = reinterpret_cast<elem_t (*)[kCh]>(s_dyn);
  float2 (*s_part)[kRows] = reinterpret_cast<float2 (*)[kRows]>(s_dyn + kRows * kCh * sizeof(elem_t));
  float2* s_ab = reinterpret_cast<float2*>(s_dyn + kRows * kCh * sizeof(elem_t) +
                                           kMaxSlabs * kRows * sizeof(float2));
  pdl_launch_dependents();
  const int c_early = blockIdx.x * kCh + threadIdx.x;
  float w[kTaps];
#pragma unroll
  for (int k = 0; k < kTaps; ++k) w[k] = __ldg(conv_w + static_cast<long long>(c_early) * kTaps + k);
  const float gam = __ldg(norm_g + c_early), bet = __ldg(norm_b + c_early), cb = __ldg(conv_b + c_early);
  const uint32_t crank = cluster_ctarank();
  const int nslab = gridDim.x;                  // == cluster size == Ch / 128
  pdl_wait();
  const int c0 = blockIdx.x * kCh;
  const int t0 = blockIdx.y * kSeg;
  const int b = blockIdx.z;
  const long long row0 = static_cast<long long>(b) * T;
  const elem_t* hb = static_cast<const elem_t*>(h);
  const elem_t* gbase = hb + Ch + c0;
  for (int idx = threadIdx.x; idx < kRows * (kCh / kPer16); idx += kCh) {
    const int r = idx / (kCh / kPer16), q = idx % (kCh / kPer16);
    const int t = t0 - kHalo + r;
    const bool ok = t >= 0 && t < T;
    cp_async16_zfill(&s_tile[r][q * kPer16], gbase + (row0 + (ok ? t : 0)) * ldh + q * kPer16, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- partial statistics of this CTA's 128 channels, warp per frame, pushed to every peer ----
  {
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    for (int r = warp; r < kRows; r += kCh / 32) {
      float x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (kBf16) x[k] = __uint_as_float(static_cast<uint32_t>(s_tile[r][lane + 32 * k]) << 16);
        else x[k] = s_tile[r][lane + 32 * k];
      }
      const float mean_c = warp_sum(x[0] + x[1] + x[2] + x[3]) * (1.0f / kCh);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) q = fmaf(x[k] - mean_c, x[k] - mean_c, q);
      q = warp_sum(q);
      // lanes 0 .. nslab-1 each deliver the pair to one CTA of the cluster
      if (static_cast<int>(lane) < nslab) {
        uint32_t dst;
        const uint32_t local = smem_u32(&s_part[crank][r]);
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(local), "r"(lane));
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(mean_c), "f"(q)
                     : "memory");
      }
    }
  }
  cluster_sync_all();   // release / acquire: every slab's partials are in this CTA's s_part
  for (int i = threadIdx.x; i < kRows; i += kCh) {
    const int t = t0 - kHalo + i;
    float2 ab = make_float2(0.f, 0.f);
    if (t >= 0 && t < T) {
      float msum = 0.f, m2 = 0.f;
      for (int c = 0; c < nslab; ++c) {
        msum += s_part[c][i].x;
        m2 += s_part[c][i].y;
      }
      const float mean = msum / static_cast<float>(nslab);
      float dev = 0.f;
      for (int c = 0; c < nslab; ++c) {
        const float dq = s_part[c][i].x - mean;
        dev = fmaf(dq, dq, dev);
      }
      const float var = (m2 + static_cast<float>(kCh) * dev) / static_cast<float>(Ch);
      const float rstd = rsqrtf(var + eps);
      ab = make_float2(rstd, -mean * rstd);
      // the training forward keeps (mean, rstd) for the backward: one CTA of the cluster, its own
      // output frames only
      if (stats_out != nullptr && crank == 0 && i >= kHalo && i < kHalo + kSeg)
        stats_out[row0 + t] = make_float2(mean, rstd);
    }
    s_ab[i] = ab;
  }
  const int c = c0 + threadIdx.x;
  const elem_t* rcol = hb + c;
  elem_t* ocol = static_cast<elem_t*>(out) + c;
  __syncthreads();

  for (int g0 = 0; g0 < kSeg; g0 += kGrp) {
    const int tb = t0 + g0;  // first output frame of this group
    if (tb >= T) break;
    float rv[kGrp];
#pragma unroll
    for (int o = 0; o < kGrp; ++o) {
      if constexpr (kBf16)
        rv[o] = tb + o < T ? __uint_as_float(static_cast<uint32_t>(ld_act_u16(rcol + (row0 + tb + o) * ldh)) << 16) : 0.f;
      else
        rv[o] = tb + o < T ? ld_act(rcol + (row0 + tb + o) * ldh) : 0.f;
    }
    float acc[kGrp];
#pragma unroll
    for (int o = 0; o < kGrp; ++o) acc[o] = 0.f;
#pragma unroll
    for (int ii = 0; ii < kGrp + kTaps - 1; ++ii) {
      const float2 ab = s_ab[g0 + ii];
      float xv;
      if constexpr (kBf16) xv = __uint_as_float(static_cast<uint32_t>(s_tile[g0 + ii][threadIdx.x]) << 16);
      else xv = s_tile[g0 + ii][threadIdx.x];
      const float xh = fmaf(xv, ab.x, ab.y);
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
        if constexpr (kBf16) ocol[(row0 + t) * ldo] = static_cast<uint16_t>(pack_bf16x2(y, 0.f) & 0xFFFFu);
        else ocol[(row0 + t) * ldo] = round_out ? round_tf32(y) : y;
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
template <bool kBf16>
__global__ void __launch_bounds__(256)
row_dots_kernel(const void* __restrict__ a1, long long ld1, int K1, const float* __restrict__ va1,
                const float* __restrict__ vb1, float2* __restrict__ out1,
                const void* __restrict__ a2, long long ld2, int K2, const float* __restrict__ va2,
                const float* __restrict__ vb2, float2* __restrict__ out2, int M) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // all loads of a row are issued before the first FMA (8 requests per lane in flight): the kernel
  // is a pure L2 -> register stream, so exposed load latency is the only thing that can slow it
  auto dots2 = [&](const void* a, int K, const float* va, const float* vb, float& sa, float& sb) {
    const float4* pa = reinterpret_cast<const float4*>(va);
    const float4* pb = reinterpret_cast<const float4*>(vb);
    const int n4 = K / 4;
    for (int k0 = lane; k0 < n4; k0 += 256) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        x[j] = k0 + 32 * j < n4 ? ld_act_vec4<kBf16>(a, k0 + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    dots2(elem_ptr<kBf16>(a1, static_cast<long long>(m) * ld1), K1, va1, vb1, s[0], s[1]);
    if (a2 != nullptr) dots2(elem_ptr<kBf16>(a2, static_cast<long long>(m) * ld2), K2, va2, vb2, s[2], s[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = warp_sum(s[i]);
    if (lane == 0) {
      out1[m] = make_float2(s[0], s[1]);
      if (a2 != nullptr) out2[m] = make_float2(s[2], s[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// learned_ave merge weights in ONE launch (the hot two-branch block path): row dots + masked
// softmax pooling over time + 2-way softmax, i.e. row_dots_kernel and merge_weights_plain_kernel
// fused.  One CLUSTER of kMsCtas CTAs per utterance: each CTA takes a contiguous quarter of the
// frames, its 8 warps compute the four dots of a frame (scores s_i = x_i . va_i, values z_i = x_i .
// vb_i for the two branches; the vectors sit in registers), then two warps reduce the CTA's frames
// to online-softmax partials (max, sum e, sum e z) per branch, the partials cross to the cluster's
// first CTA over distributed shared memory and one thread combines them.  Between the last branch
// kernel and the merge GEMM the step used to pay two dependent launches (14.9 + 9.4 us at C2,
// 8 % of the step) for 20 MB of reads and a few hundred flops per frame.
// ------------------------------------------------------------------------------------------------
constexpr int kMsCtas = 4;
constexpr int kMsMaxRows = 512;   // frames per CTA held in shared memory (T <= 2048)

template <bool kBf16>
__global__ void __launch_bounds__(256)
merge_scores_kernel(const void* __restrict__ a1, long long ld1, const void* __restrict__ a2,
                    long long ld2, const float* __restrict__ va1, const float* __restrict__ vb1,
                    const float* __restrict__ va2, const float* __restrict__ vb2,
                    const int32_t* __restrict__ lens, float pool_b1, float pool_b2, float wproj_b1,
                    float wproj_b2, float inv_sqrt, float* __restrict__ w1, float* __restrict__ w2,
                    int T) {
  // branch 1 rows are 256 wide (attention context), branch 2 rows 1024 wide (gated cgMLP hidden)
  constexpr int kV1 = 2, kV2 = 8;   // float4 groups per lane: 256 / 128, 1024 / 128
  __shared__ float s_sc[4][kMsMaxRows];       // s1, z1, s2, z2 per local frame
  __shared__ float s_part[kMsCtas][6];        // leader only: (m, se, sz) x 2 branches per CTA
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t crank = cluster_ctarank();
  const int b = blockIdx.x / kMsCtas;
  float4 ra1[kV1], rb1[kV1], ra2[kV2], rb2[kV2];
#pragma unroll
  for (int i = 0; i < kV1; ++i) {
    ra1[i] = __ldg(reinterpret_cast<const float4*>(va1) + lane + 32 * i);
    rb1[i] = __ldg(reinterpret_cast<const float4*>(vb1) + lane + 32 * i);
  }
#pragma unroll
  for (int i = 0; i < kV2; ++i) {
    ra2[i] = __ldg(reinterpret_cast<const float4*>(va2) + lane + 32 * i);
    rb2[i] = __ldg(reinterpret_cast<const float4*>(vb2) + lane + 32 * i);
  }
  pdl_wait();
  int len = lens ? lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  const int per = (T + kMsCtas - 1) / kMsCtas;
  const int t0 = static_cast<int>(crank) * per;
  const int t1 = min(min(t0 + per, T), len);   // frames >= len are masked: never read
  const long long row0 = static_cast<long long>(b) * T;
  for (int t = t0 + warp; t < t1; t += 8) {
    const void* r1 = elem_ptr<kBf16>(a1, (row0 + t) * ld1);
    const void* r2 = elem_ptr<kBf16>(a2, (row0 + t) * ld2);
    float4 x1[kV1], x2[kV2];
#pragma unroll
    for (int i = 0; i < kV1; ++i) x1[i] = ld_act_vec4<kBf16>(r1, lane + 32 * i);
#pragma unroll
    for (int i = 0; i < kV2; ++i) x2[i] = ld_act_vec4<kBf16>(r2, lane + 32 * i);
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < kV1; ++i) {
      d[0] += x1[i].x * ra1[i].x + x1[i].y * ra1[i].y + x1[i].z * ra1[i].z + x1[i].w * ra1[i].w;
      d[1] += x1[i].x * rb1[i].x + x1[i].y * rb1[i].y + x1[i].z * rb1[i].z + x1[i].w * rb1[i].w;
    }
#pragma unroll
    for (int i = 0; i < kV2; ++i) {
      d[2] += x2[i].x * ra2[i].x + x2[i].y * ra2[i].y + x2[i].z * ra2[i].z + x2[i].w * ra2[i].w;
      d[3] += x2[i].x * rb2[i].x + x2[i].y * rb2[i].y + x2[i].z * rb2[i].z + x2[i].w * rb2[i].w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = warp_sum(d[k]);
    if (lane == 0) {
      s_sc[0][t - t0] = (d[0] + pool_b1) * inv_sqrt;
      s_sc[1][t - t0] = d[1];
      s_sc[2][t - t0] = (d[2] + pool_b2) * inv_sqrt;
      s_sc[3][t - t0] = d[3];
    }
  }
  __syncthreads();
  // ---- per-CTA online-softmax partials: warp 0 -> branch 1, warp 1 -> branch 2 ----
  if (warp < 2) {
    const float* sc = s_sc[2 * warp];
    const float* zz = s_sc[2 * warp + 1];
    const int n = t1 > t0 ? t1 - t0 : 0;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sc[i]);
    mx = warp_max(mx);
    float se = 0.f, sz = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float e = expf(sc[i] - mx);
      se += e;
      sz += e * zz[i];
    }
    se = warp_sum(se);
    sz = warp_sum(sz);
    if (lane == 0) {
      // into the cluster leader's s_part[crank][3 warp ..]
      uint32_t dst;
      const uint32_t local = smem_u32(&s_part[crank][3 * warp]);
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(local), "r"(0));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(mx) : "memory");
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst + 4), "f"(se) : "memory");
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst + 8), "f"(sz) : "memory");
    }
  }
  cluster_sync_all();   // release / acquire: every CTA's partials are visible in the leader
  if (crank == 0 && threadIdx.x == 0) {
    float omega[2];
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float M = -INFINITY;
      for (int c = 0; c < kMsCtas; ++c) M = fmaxf(M, s_part[c][3 * br]);
      float se = 0.f, sz = 0.f;
      for (int c = 0; c < kMsCtas; ++c) {
        const float m = s_part[c][3 * br];
        if (m == -INFINITY) continue;   // a CTA without valid frames
        const float f = expf(m - M);
        se += s_part[c][3 * br + 1] * f;
        sz += s_part[c][3 * br + 2] * f;
      }
      // len == 0: every score is masked, softmax-then-zero gives an all-zero pooling vector
      omega[br] = (len > 0 ? sz / se : 0.f) + (br == 0 ? wproj_b1 : wproj_b2);
    }
    const float m = fmaxf(omega[0], omega[1]);
    const float e0 = expf(omega[0] - m), e1 = expf(omega[1] - m);
    w1[b] = e0 / (e0 + e1);
    w2[b] = e1 / (e0 + e1);
  }
}

// ------------------------------------------------------------------------------------------------
// learned_ave merge weights: one warp per utterance.  dotsK holds npK partial (score, z) pairs per
// frame (np = 1: the row dots of a row-complete GEMM / row_dots; np > 1: the partials the
// attention and CSGU kernels emit per head half / channel slab), summed on the fly.  (A CTA per
// utterance with block reductions measured 6 us SLOWER per launch inside the PDL-chained graph.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 sum_partials(const float2* d, int np) {
  float2 acc = ld_act2(d);
  for (int q = 1; q < np; ++q) {
    const float2 v = ld_act2(d + q);
    acc.x += v.x;
    acc.y += v.y;
  }
  return acc;
}

__global__ void __launch_bounds__(128)
merge_weights_kernel(const float2* __restrict__ dots1, int np1, const float2* __restrict__ dots2,
                     int np2, const int32_t* __restrict__ lens1, const int32_t* __restrict__ lens2,
                     float pool_b1, float pool_b2, float wproj_b1, float wproj_b2, float inv_sqrt,
                     float* __restrict__ w1, float* __restrict__ w2, int B, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const uint32_t lane = lane_id();
  float omega[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const int32_t* lens = br == 0 ? lens1 : (lens2 ? lens2 : lens1);
    int len = lens ? lens[b] : T;
    len = len < 0 ? 0 : (len > T ? T : len);
    const int np = br == 0 ? np1 : np2;
    const float2* d = (br == 0 ? dots1 : dots2) + static_cast<long long>(b) * T * np;
    const float pb = br == 0 ? pool_b1 : pool_b2;
    float mx = -INFINITY;
    for (int t = lane; t < len; t += 32)
      mx = fmaxf(mx, (sum_partials(d + static_cast<long long>(t) * np, np).x + pb) * inv_sqrt);
    mx = warp_max(mx);
    float se = 0.f, sz = 0.f;
    for (int t = lane; t < len; t += 32) {
      const float2 v = sum_partials(d + static_cast<long long>(t) * np, np);
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

// np == 1 and one length array: the hot two-branch block path, kept exactly as tuned (this
// kernel sits alone on the critical path between the branches and the merge GEMM)
__global__ void __launch_bounds__(128)
merge_weights_plain_kernel(const float2* __restrict__ dots1, const float2* __restrict__ dots2,
                     const int32_t* __restrict__ lens, float pool_b1, float pool_b2, float wproj_b1,
                     float wproj_b2, float inv_sqrt, float* __restrict__ w1, float* __restrict__ w2,
                     int B, int T, const float* __restrict__ scal,
                     const int32_t* __restrict__ lens2) {
  pdl_launch_dependents();
  pdl_wait();
  if (scal != nullptr) {   // training: the biases live on the device
    pool_b1 = ld_act(scal); pool_b2 = ld_act(scal + 1);
    wproj_b1 = ld_act(scal + 2); wproj_b2 = ld_act(scal + 3);
  }
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const uint32_t lane = lane_id();
  int len_a = lens ? lens[b] : T;
  len_a = len_a < 0 ? 0 : (len_a > T ? T : len_a);
  int len_b = lens2 ? lens2[b] : len_a;   // the AV fusion masks each modality with its own lengths
  len_b = len_b < 0 ? 0 : (len_b > T ? T : len_b);
  float omega[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const int len = br == 0 ? len_a : len_b;
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

// ------------------------------------------------------------------------------------------------
// out[m,:] = w1[m / rows_per_seg] * a[m,:] + w2[m / rows_per_seg] * b[m,:]   (the weighted modality
// average in front of the fusion FFN, adaptive_audiovisual_fusion.py:187-194)
// ------------------------------------------------------------------------------------------------
template <bool kOutBf16>
__global__ void __launch_bounds__(256)
scale_add_rows_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                      long long ldb, const float* __restrict__ w1, const float* __restrict__ w2,
                      int rows_per_seg, void* __restrict__ out, long long ldo, int M, int D4) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * D4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / D4), q = static_cast<int>(idx % D4);
    const int seg = m / rows_per_seg;
    const float s1 = ld_act(w1 + seg), s2 = ld_act(w2 + seg);
    const float4 x = ld_act4(reinterpret_cast<const float4*>(a + m * lda) + q);
    const float4 y = ld_act4(reinterpret_cast<const float4*>(b + m * ldb) + q);
    const float4 r =
        make_float4(s1 * x.x + s2 * y.x, s1 * x.y + s2 * y.y, s1 * x.z + s2 * y.z, s1 * x.w + s2 * y.w);
    if constexpr (kOutBf16)
      reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + m * ldo)[q] =
          make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
    else
      reinterpret_cast<float4*>(static_cast<float*>(out) + m * ldo)[q] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 operand split ("tf32x3" mode): x = hi + lo with hi = tf32(x), lo = x - hi (exact in fp32;
// TMA rounds it to TF32 on load).  The product  a.b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi  then runs as
// ONE TF32 GEMM over a tripled reduction axis: activations are written as [hi | hi | lo] (pattern
// 0), weights as [hi | lo | hi] (pattern 1), K -> 3K.  Error ~2^-21 instead of 2^-11 per product.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ in, long long ld, float* __restrict__ out, long long ldo,
                  int M, int K4, int pattern) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * K4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / K4), q = static_cast<int>(idx % K4);
    const float4 x = ld_act4(reinterpret_cast<const float4*>(in + m * ld) + q);
    const float4 hi = make_float4(round_tf32(x.x), round_tf32(x.y), round_tf32(x.z), round_tf32(x.w));
    const float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    float4* o = reinterpret_cast<float4*>(out + m * ldo);
    o[q] = hi;
    o[K4 + q] = pattern == 0 ? hi : lo;
    o[2 * K4 + q] = pattern == 0 ? lo : hi;
  }
}

// ------------------------------------------------------------------------------------------------
// Conv2dSubsampling front end (espnet Conv2dSubsampling, reached from encoder.py:149-155):
//   conv1 = relu(Conv2d(1, C, 3, stride 2)(x)),  conv2 = relu(Conv2d(C, C, 3, stride 2)(conv1)).
// This kernel evaluates conv1 ON THE FLY (9 FMAs per value) while writing the im2col operand of
// conv2, so the (B, C, T1, F1) intermediate never exists:
//   A[(b, t2, f2), (i*3 + j) * C + c] = conv1[b, c, 2 t2 + i, 2 f2 + j]
// conv2 is then one tcgen05 GEMM  relu(A . W2r^T + b2)  with W2r[c2, (i*3+j)*C + c1] = w2[c2,c1,i,j],
// whose (B*T2, F2*C) channels-last output feeds the 4864 -> 256 projection with permuted columns.
// CTA = one (b, t2); thread = conv1 channel c: every store is a coalesced C*4-byte run.
// ------------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(256)
conv2d_sub_im2col_kernel(const float* __restrict__ x, int Tin, int F, const float* __restrict__ w1,
                         const float* __restrict__ b1, int C, int T2, int F2,
                         void* __restrict__ A) {
  using elem_t = typename std::conditional<kBf16, uint16_t, float>::type;
  extern __shared__ float s_x[];  // 7 input rows x F
  pdl_launch_dependents();
  const int b = blockIdx.x / T2, t2 = blockIdx.x % T2;
  pdl_wait();
  const float* xb = x + (static_cast<long long>(b) * Tin + 4 * t2) * F;
  for (int i = threadIdx.x; i < 7 * F; i += blockDim.x) s_x[i] = ld_act(xb + i);  // rows 4t2..4t2+6
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = __ldg(w1 + c * 9 + k);
    const float bias = __ldg(b1 + c);
    elem_t* arow = static_cast<elem_t*>(A) + (static_cast<long long>(b) * T2 + t2) * F2 * 9 * C + c;
    for (int f2 = 0; f2 < F2; ++f2) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float* xp = s_x + (2 * i) * F + 4 * f2 + 2 * j;
          float v = bias;
#pragma unroll
          for (int pp = 0; pp < 3; ++pp)
#pragma unroll
            for (int q = 0; q < 3; ++q) v = fmaf(w[pp * 3 + q], xp[pp * F + q], v);
          if constexpr (kBf16)
            arow[(static_cast<long long>(f2) * 9 + i * 3 + j) * C] =
                static_cast<uint16_t>(pack_bf16x2(fmaxf(v, 0.f), 0.f) & 0xFFFFu);
          else
            arow[(static_cast<long long>(f2) * 9 + i * 3 + j) * C] = fmaxf(v, 0.f);
        }
      }
    }
  }
}

}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_layernorm(const float* x, long long ldx, int M, int D, float eps,
                               const float* gA, const float* bA, void* outA, long long ldA,
                               int roundA, const float* gB, const float* bB, void* outB,
                               long long ldB, int roundB, float scale, int dtype, void* stream) {
  const int bf16_mask = ((dtype & TAVSR_DT_LNA_BF16) ? 1 : 0) | ((dtype & TAVSR_DT_LNB_BF16) ? 2 : 0);
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
                                gA, bA, outA, ldA, roundA, gB, bB, outB, ldB, roundB, scale,      \
                                bf16_mask));                                                      \
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

template <bool kBf16>
static int csgu_launch(const void* h, long long ldh, const float* norm_g, const float* norm_b,
                       const float* conv_w, const float* conv_b, void* out, long long ldo,
                       float* stats, int B, int T, int Ch, float eps, int round_out, cudaStream_t s) {
  const int M = B * T;
  // one-pass cluster kernel (g_debug[10] = 1; needs Ch <= 1024).  OPT-IN: measured SLOWER than the
  // two-kernel sequence on the C2 step (56 vs 44 us per block, 2.615 vs 2.424 ms per step in bf16):
  // the statistics phase sits in front of every CTA's convolution, the 8-CTA clusters constrain
  // the 4-CTAs-per-SM schedule, and the launch it removes was overlapped by PDL anyway.
  if (Ch / kCh <= kMaxSlabs && g_debug[10] == 1) {
    dim3 grid(Ch / kCh, (T + kSeg - 1) / kSeg, B);
    constexpr int smem = kRows * kCh * (kBf16 ? 2 : 4) + kMaxSlabs * kRows * 8 + kRows * 8;
    static unsigned long long configured = 0;
    if (first_use_on_device(configured))
      TAVSR_CUDA_OK(cudaFuncSetAttribute(csgu_onepass_kernel<kBf16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TAVSR_CUDA_OK(launch_kernel(csgu_onepass_kernel<kBf16>, grid, dim3(kCh), smem, s, Ch / kCh, h, ldh, norm_g,
                                norm_b, conv_w, conv_b, reinterpret_cast<float2*>(stats), eps, out, ldo, T,
                                Ch, round_out));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  const int grid1 = (M + 7) / 8;
  float2* st = reinterpret_cast<float2*>(stats);
#define TAVSR_STATS_CASE(V)                                                                        \
  case V:                                                                                          \
    TAVSR_CUDA_OK(launch_kernel(csgu_stats_kernel<V, kBf16>, dim3(grid1), dim3(256), 0, s, 0, h,   \
                                ldh, M, Ch, eps, st));                                             \
    break;
  switch (Ch / 128) {
    TAVSR_STATS_CASE(1) TAVSR_STATS_CASE(2) TAVSR_STATS_CASE(4) TAVSR_STATS_CASE(8) TAVSR_STATS_CASE(16)
    default: return set_error(TAVSR_ERR_UNSUPPORTED, "csgu: Ch=%d not instantiated", Ch);
  }
#undef TAVSR_STATS_CASE
  dim3 grid2(Ch / kCh, (T + kSeg - 1) / kSeg, B);
  TAVSR_CUDA_OK(launch_kernel(csgu_conv_kernel<kBf16>, grid2, dim3(kCh), 0, s, 0, h, ldh, norm_g, norm_b,
                              conv_w, conv_b, static_cast<const float2*>(st), out, ldo, T, Ch,
                              round_out));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_csgu_fwd(const void* h, long long ldh, const float* norm_g,
                              const float* norm_b, const float* conv_w, const float* conv_b,
                              void* out, long long ldo, float* stats, int B, int T, int Ch,
                              int ksize, float eps, int round_out, int dtype, void* stream) {
  TAVSR_REQUIRE(ksize == kTaps, "csgu: only kernel size 31 is built (got %d)", ksize);
  TAVSR_REQUIRE(B > 0 && T > 0 && Ch > 0 && Ch % 128 == 0 && Ch <= 2048, "csgu: bad shape");
  const int op = dtype & TAVSR_DT_MASK;
  TAVSR_REQUIRE(op == TAVSR_DT_TF32 || (op == TAVSR_DT_BF16 && (dtype & TAVSR_DT_OUT_BF16)),
                "csgu: dtype is TAVSR_DT_TF32 (fp32 h / out) or TAVSR_DT_BF16 | TAVSR_DT_OUT_BF16");
  const bool bf16 = op == TAVSR_DT_BF16;
  TAVSR_REQUIRE(ldh % (bf16 ? 8 : 4) == 0, "csgu: bad pitch");
  TAVSR_REQUIRE(stats != nullptr || (Ch / 128 <= 8 && g_debug[10] == 1),
                "csgu: the two-kernel path needs the stats scratch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bf16 ? csgu_launch<true>(h, ldh, norm_g, norm_b, conv_w, conv_b, out, ldo, stats, B, T, Ch,
                                  eps, round_out, s)
              : csgu_launch<false>(h, ldh, norm_g, norm_b, conv_w, conv_b, out, ldo, stats, B, T, Ch,
                                   eps, round_out, s);
}

extern "C" int tavsr_row_dots(const void* a1, long long ld1, int K1, const float* va1,
                              const float* vb1, float* out1, const void* a2, long long ld2, int K2,
                              const float* va2, const float* vb2, float* out2, int M, int dtype,
                              void* stream) {
  const bool bf16 = (dtype & TAVSR_DT_MASK) == TAVSR_DT_BF16;
  const int al = bf16 ? 8 : 4;
  TAVSR_REQUIRE(M > 0 && a1 && va1 && vb1 && out1 && K1 > 0 && K1 % 4 == 0 && ld1 % al == 0,
                "row_dots: bad first operand (K=%d)", K1);
  TAVSR_REQUIRE(!a2 || (va2 && vb2 && out2 && K2 > 0 && K2 % 4 == 0 && ld2 % al == 0),
                "row_dots: bad second operand (K=%d)", K2);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = (M + 7) / 8;
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  if (bf16)
    TAVSR_CUDA_OK(launch_kernel(row_dots_kernel<true>, dim3(grid), dim3(256), 0, s, 0, a1, ld1, K1, va1,
                                vb1, reinterpret_cast<float2*>(out1), a2, ld2, K2, va2, vb2,
                                reinterpret_cast<float2*>(out2), M));
  else
    TAVSR_CUDA_OK(launch_kernel(row_dots_kernel<false>, dim3(grid), dim3(256), 0, s, 0, a1, ld1, K1, va1,
                                vb1, reinterpret_cast<float2*>(out1), a2, ld2, K2, va2, vb2,
                                reinterpret_cast<float2*>(out2), M));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_merge_learned_ave_weights2(const float* dots1, int np1, const float* dots2,
                                                int np2, const int32_t* lens1, const int32_t* lens2,
                                                float pool_b1, float pool_b2, float wproj_b1,
                                                float wproj_b2, float inv_sqrt_size, float* w1,
                                                float* w2, int B, int T, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && dots1 && dots2 && w1 && w2 && np1 >= 1 && np2 >= 1,
                "merge_weights: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (np1 == 1 && np2 == 1 && lens2 == nullptr) {
    TAVSR_CUDA_OK(launch_kernel(merge_weights_plain_kernel, dim3((B + 3) / 4), dim3(128), 0, s, 0,
                                reinterpret_cast<const float2*>(dots1),
                                reinterpret_cast<const float2*>(dots2), lens1, pool_b1, pool_b2,
                                wproj_b1, wproj_b2, inv_sqrt_size, w1, w2, B, T,
                                static_cast<const float*>(nullptr),
                                static_cast<const int32_t*>(nullptr)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  TAVSR_CUDA_OK(launch_kernel(merge_weights_kernel, dim3((B + 3) / 4), dim3(128), 0, s, 0,
                              reinterpret_cast<const float2*>(dots1), np1,
                              reinterpret_cast<const float2*>(dots2), np2, lens1, lens2, pool_b1,
                              pool_b2, wproj_b1, wproj_b2, inv_sqrt_size, w1, w2, B, T));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_merge_scores(const void* a1, long long ld1, int K1, const void* a2, long long ld2,
                                  int K2, const float* va1, const float* vb1, const float* va2,
                                  const float* vb2, const int32_t* lens, float pool_b1, float pool_b2,
                                  float wproj_b1, float wproj_b2, float inv_sqrt_size, float* w1,
                                  float* w2, int B, int T, int dtype, void* stream) {
  const bool bf16 = (dtype & TAVSR_DT_MASK) == TAVSR_DT_BF16;
  const int al = bf16 ? 8 : 4;
  TAVSR_REQUIRE(B > 0 && T > 0 && T <= kMsCtas * kMsMaxRows && a1 && a2 && va1 && vb1 && va2 && vb2 &&
                    w1 && w2,
                "merge_scores: bad arguments (B=%d T=%d)", B, T);
  TAVSR_REQUIRE(K1 == 256 && K2 == 1024 && ld1 % al == 0 && ld2 % al == 0,
                "merge_scores: built for a 256-wide attention context and a 1024-wide gated cgMLP "
                "hidden (K1=%d K2=%d)", K1, K2);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bf16)
    TAVSR_CUDA_OK(launch_kernel(merge_scores_kernel<true>, dim3(B * kMsCtas), dim3(256), 0, s, kMsCtas,
                                a1, ld1, a2, ld2, va1, vb1, va2, vb2, lens, pool_b1, pool_b2, wproj_b1,
                                wproj_b2, inv_sqrt_size, w1, w2, T));
  else
    TAVSR_CUDA_OK(launch_kernel(merge_scores_kernel<false>, dim3(B * kMsCtas), dim3(256), 0, s, kMsCtas,
                                a1, ld1, a2, ld2, va1, vb1, va2, vb2, lens, pool_b1, pool_b2, wproj_b1,
                                wproj_b2, inv_sqrt_size, w1, w2, T));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_merge_learned_ave_weights_dev(const float* dots1, const float* dots2,
                                                   const int32_t* lens, const int32_t* lens2,
                                                   const float* scal, float inv_sqrt_size, float* w1,
                                                   float* w2, int B, int T, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && dots1 && dots2 && scal && w1 && w2, "merge_weights: bad arguments");
  TAVSR_CUDA_OK(launch_kernel(merge_weights_plain_kernel, dim3((B + 3) / 4), dim3(128), 0,
                              static_cast<cudaStream_t>(stream), 0,
                              reinterpret_cast<const float2*>(dots1),
                              reinterpret_cast<const float2*>(dots2), lens, 0.f, 0.f, 0.f, 0.f,
                              inv_sqrt_size, w1, w2, B, T, scal, lens2));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_merge_learned_ave_weights(const float* dots1, const float* dots2,
                                               const int32_t* lens, float pool_b1, float pool_b2,
                                               float wproj_b1, float wproj_b2, float inv_sqrt_size,
                                               float* w1, float* w2, int B, int T, void* stream) {
  return tavsr_merge_learned_ave_weights2(dots1, 1, dots2, 1, lens, nullptr, pool_b1, pool_b2,
                                          wproj_b1, wproj_b2, inv_sqrt_size, w1, w2, B, T, stream);
}

extern "C" int tavsr_scale_add_rows(const float* a, long long lda, const float* b, long long ldb,
                                    const float* w1, const float* w2, int rows_per_seg, void* out,
                                    long long ldo, int M, int D, int dtype, void* stream) {
  TAVSR_REQUIRE(M > 0 && D > 0 && D % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldo % 4 == 0 &&
                    rows_per_seg > 0 && a && b && w1 && w2 && out,
                "scale_add_rows: bad arguments (M=%d D=%d)", M, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(M) * (D / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
  if (dtype & TAVSR_DT_OUT_BF16)
    TAVSR_CUDA_OK(launch_kernel(scale_add_rows_kernel<true>, dim3(static_cast<unsigned>(blocks)),
                                dim3(256), 0, s, 0, a, lda, b, ldb, w1, w2, rows_per_seg, out, ldo, M,
                                D / 4));
  else
    TAVSR_CUDA_OK(launch_kernel(scale_add_rows_kernel<false>, dim3(static_cast<unsigned>(blocks)),
                                dim3(256), 0, s, 0, a, lda, b, ldb, w1, w2, rows_per_seg, out, ldo, M,
                                D / 4));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_split_tf32(const float* in, long long ld, float* out, long long ldo, int M,
                                int K, int pattern, void* stream) {
  TAVSR_REQUIRE(M > 0 && K > 0 && K % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && in && out &&
                    (pattern == 0 || pattern == 1),
                "split_tf32: bad arguments (M=%d K=%d pattern=%d)", M, K, pattern);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(M) * (K / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
  TAVSR_CUDA_OK(launch_kernel(split_tf32_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, 0,
                              in, ld, out, ldo, M, K / 4, pattern));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_conv2d_sub_im2col(const float* x, int B, int Tin, int F, const float* w1,
                                       const float* b1, int C, void* A, int dtype, void* stream) {
  TAVSR_REQUIRE(B > 0 && Tin >= 7 && F >= 7 && C > 0 && x && w1 && b1 && A,
                "conv2d_sub: bad arguments (B=%d Tin=%d F=%d C=%d)", B, Tin, F, C);
  const int T2 = ((Tin - 1) / 2 - 1) / 2, F2 = ((F - 1) / 2 - 1) / 2;
  TAVSR_REQUIRE(T2 >= 1 && F2 >= 1 && 7 * F * 4 <= 48 * 1024, "conv2d_sub: unsupported shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype & TAVSR_DT_OUT_BF16)
    TAVSR_CUDA_OK(launch_kernel(conv2d_sub_im2col_kernel<true>, dim3(B * T2), dim3(256),
                                static_cast<size_t>(7 * F * 4), s, 0, x, Tin, F, w1, b1, C, T2, F2, A));
  else
    TAVSR_CUDA_OK(launch_kernel(conv2d_sub_im2col_kernel<false>, dim3(B * T2), dim3(256),
                                static_cast<size_t>(7 * F * 4), s, 0, x, Tin, F, w1, b1, C, T2, F2, A));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
