// Backward building blocks for the encoder.
//
// These kernels transcribe the autograd-free formulas of oracle/bwd_formulas.py (verified against
// torch.autograd on the CPU, tests/test_bwd_formulas_cpu.py) and are checked against them on the
// GPU (tests/test_backward_gpu.py).  No module of the product path calls them yet: the attention
// and merge backward kernels and the training-forward orchestration are still missing.
//
// All of them are memory-bound fp32 row kernels (coalesced along the channel axis, warp-shuffle or
// two-stage reductions, no atomics: results are bit-reproducible).  The GEMM-shaped parts of the
// backward (dgrad dX = dY.W, wgrad dW = dY^T.X) reuse the tcgen05 GEMM: tavsr_gemm_bias_act with a
// transposed weight copy, resp. with both operands transposed by tavsr_transpose_2d.
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;

namespace bwd {

// ------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c]  (32 x 32 tiles through shared memory, both sides coalesced)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out,
                 long long ld_out, int R, int C, int Rpad) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? ld_act(in + static_cast<long long>(r) * ld_in + c) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    // columns [R, Rpad) of the output are written as zeros: a wgrad product reduces over them
    if (c < C && r < Rpad) out[static_cast<long long>(c) * ld_out + r] = tile[tx][j];
  }
}

// ------------------------------------------------------------------------------------------------
// Column sums, two stages:  part[blk][c] = sum over the CTA's rows of a[r][c] (* b[r][c]),
// then out[c] = sum_blk part[blk][c].  (bias gradients; LayerNorm d gamma with b = x_hat.)
// ------------------------------------------------------------------------------------------------
constexpr int kColRows = 64;  // rows per CTA of the partial stage (250 x C/128 CTAs at M = 8000)

__global__ void __launch_bounds__(128)
col_sums_partial_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                        long long ldb, float* __restrict__ part, int R, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * kColRows;
  if (c >= C) return;
  const int r1 = r0 + kColRows < R ? r0 + kColRows : R;
  // eight independent row streams per thread: the loads of one pass are all in flight together
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ld_act(a + static_cast<long long>(r + k) * lda + c);
    if (b != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= ld_act(b + static_cast<long long>(r + k) * ldb + c);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
  }
  for (; r < r1; ++r) {
    float v = ld_act(a + static_cast<long long>(r) * lda + c);
    if (b != nullptr) v *= ld_act(b + static_cast<long long>(r) * ldb + c);
    s[0] += v;
  }
  part[static_cast<long long>(blockIdx.y) * C + c] =
      ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// out[c] = sum_blk part[blk][c]: CTA = 32 columns x 32 row groups (a fixed summation tree, so the
// result is bit-reproducible); the former one-thread-per-column loop took 93 us over 1000 partial
// rows on four CTAs.
__global__ void __launch_bounds__(1024)
col_sums_reduce_kernel(const float* __restrict__ part, int nblk, float* __restrict__ out, int C) {
  __shared__ float red[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    int b = ty;
    for (; b + 96 < nblk; b += 128) {
      s0 += ld_act(part + static_cast<long long>(b) * C + c);
      s1 += ld_act(part + static_cast<long long>(b + 32) * C + c);
      s2 += ld_act(part + static_cast<long long>(b + 64) * C + c);
      s3 += ld_act(part + static_cast<long long>(b + 96) * C + c);
    }
    for (; b < nblk; b += 32) s0 += ld_act(part + static_cast<long long>(b) * C + c);
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += red[k][tx];
    out[c] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// dz = dh * act'(z)   (Swish: s (1 + z (1 - s)); exact-erf GELU: Phi(z) + z phi(z); ReLU: z > 0)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case 1: {  // swish
      const float s = 1.0f / (1.0f + __expf(-z));
      return s * (1.0f + z * (1.0f - s));
    }
    case 2: {  // gelu
      const float phi = 0.3989422804014327f * __expf(-0.5f * z * z);
      const float Phi = 0.5f * (1.0f + erff(z * 0.7071067811865476f));
      return Phi + z * phi;
    }
    case 3: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

__device__ __forceinline__ float act_val(float z, int act) {
  switch (act) {
    case 1: return z / (1.0f + __expf(-z));
    case 2: return 0.5f * z * (1.0f + erff(z * 0.7071067811865476f));
    case 3: return fmaxf(z, 0.f);
    default: return z;
  }
}

// h = act(z) [* mask]: the training forward keeps the pre-activation z (the backward needs it) and
// re-evaluates h where a weight gradient reads it; `mask` (optional, same shape) carries a dropout
// keep-mask already scaled by 1 / (1 - p)
__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ mask,
               long long ldm, float* __restrict__ h, long long ldh, int M, int C4, int act) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / C4), q = static_cast<int>(i % C4);
    const float4 zz = ld_act4(reinterpret_cast<const float4*>(z + m * ldz) + q);
    float4 r = make_float4(act_val(zz.x, act), act_val(zz.y, act), act_val(zz.z, act), act_val(zz.w, act));
    if (mask != nullptr) {
      const float4 k = ld_act4(reinterpret_cast<const float4*>(mask + m * ldm) + q);
      r.x *= k.x; r.y *= k.y; r.z *= k.z; r.w *= k.w;
    }
    reinterpret_cast<float4*>(h + m * ldh)[q] = r;
  }
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ dh,
               long long ldh, float* __restrict__ dz, long long ldd, int M, int C4, int act) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(M) * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / C4), q = static_cast<int>(i % C4);
    const float4 zz = ld_act4(reinterpret_cast<const float4*>(z + m * ldz) + q);
    const float4 g = ld_act4(reinterpret_cast<const float4*>(dh + m * ldh) + q);
    reinterpret_cast<float4*>(dz + m * ldd)[q] =
        make_float4(g.x * act_grad(zz.x, act), g.y * act_grad(zz.y, act),
                    g.z * act_grad(zz.z, act), g.w * act_grad(zz.w, act));
  }
}

// ------------------------------------------------------------------------------------------------
// Elementwise op with a TRANSPOSED (and optionally also a row-major) output, 64 x 64 tiles, float4 on
// both sides:   v = a                         (kOp 0: plain transpose)
//               v = act(a) [* mask]           (kOp 1: the FFN hidden h recomputed from z, needed by
//                                              the backward only as the wgrad operand h^T)
//               v = b * act'(a)               (kOp 2: dz, needed row-major by dgrad and transposed
//                                              by wgrad: one read of z / dh, two writes)
// outT[c][r] = v[r][c] with the columns [R, Rpad) of outT zero-filled (a wgrad product reduces over
// them); out (optional) gets v row-major.  Replaces an elementwise pass + a separate transpose
// (2 x 131 MB at the FFN hidden's shape).
// ------------------------------------------------------------------------------------------------
template <int kOp>
__global__ void __launch_bounds__(256)
ew_transpose_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                    long long ldb, const float* __restrict__ mask, long long ldm,
                    float* __restrict__ out, long long ldo, float* __restrict__ outT, long long ldt,
                    int R, int C, int Rpad, int act) {
  __shared__ float tile[64][65];
  pdl_launch_dependents();
  pdl_wait();
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tq = threadIdx.x & 15, tr = threadIdx.x >> 4;   // 16 float4 per row x 16 rows per pass
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int rr = tr + 16 * pass;
    const int r = r0 + rr, c = c0 + 4 * tq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && c < C) {   // C % 4 == 0: a float4 is inside or outside as a whole
      v = ld_act4(reinterpret_cast<const float4*>(a + static_cast<long long>(r) * lda + c));
      if (kOp == 1) {
        v = make_float4(act_val(v.x, act), act_val(v.y, act), act_val(v.z, act), act_val(v.w, act));
        if (mask != nullptr) {
          const float4 k = ld_act4(reinterpret_cast<const float4*>(mask + static_cast<long long>(r) * ldm + c));
          v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
        }
      } else if (kOp == 2) {
        const float4 g = ld_act4(reinterpret_cast<const float4*>(b + static_cast<long long>(r) * ldb + c));
        v = make_float4(g.x * act_grad(v.x, act), g.y * act_grad(v.y, act), g.z * act_grad(v.z, act),
                        g.w * act_grad(v.w, act));
      }
      if (out != nullptr)
        *reinterpret_cast<float4*>(out + static_cast<long long>(r) * ldo + c) = v;
    }
    tile[rr][4 * tq] = v.x; tile[rr][4 * tq + 1] = v.y; tile[rr][4 * tq + 2] = v.z; tile[rr][4 * tq + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int cc = tr + 16 * pass;            // output row = input column
    const int c = c0 + cc, r = r0 + 4 * tq;   // output columns r .. r + 3
    if (c < C && r < Rpad)                    // Rpad % 4 == 0; rows >= R of the tile hold zeros
      *reinterpret_cast<float4*>(outT + static_cast<long long>(c) * ldt + r) =
          make_float4(tile[4 * tq][cc], tile[4 * tq + 1][cc], tile[4 * tq + 2][cc], tile[4 * tq + 3][cc]);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (D = 128 * kVec), rows of a CTA = 8 warps x kLnRowsPerWarp:
//   x_hat = (x - mu) rstd;  g = dy * gamma;  dx = rstd (g - mean(g) - x_hat mean(g x_hat)) [+ dres]
//   per-warp partials of d gamma = sum dy x_hat and d beta = sum dy -> part[warp slot][2][D]
// (reduced by col_sums_reduce_kernel over 2 D columns).  Statistics are recomputed two-pass.
// ------------------------------------------------------------------------------------------------
constexpr int kLnRowsPerWarp = 8;

template <int kVec>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ dy, long long ldy, const float* __restrict__ dres,
                     long long ldr, float* __restrict__ dx, long long ldd,
                     float* __restrict__ part, int M, float eps) {
  constexpr int D = 128 * kVec;
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  float4 gam[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) gam[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
  pdl_wait();
  float4 ag[kVec], ab[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int row_beg = (blockIdx.x * 8 + warp) * kLnRowsPerWarp;
  if (row_beg >= M) return;  // whole warp idle: its partial slot does not exist
  for (int rr = 0; rr < kLnRowsPerWarp; ++rr) {
    const int row = row_beg + rr;
    if (row >= M) break;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
    const float4* yr = reinterpret_cast<const float4*>(dy + static_cast<long long>(row) * ldy);
    float4 xv[kVec], gv[kVec];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      xv[i] = ld_act4(xr + lane + 32 * i);
      gv[i] = ld_act4(yr + lane + 32 * i);
      sum += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    const float inv_d = 1.0f / static_cast<float>(D);
    const float mean = warp_sum(sum) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      // x_hat in xv, dy in gv -> accumulate parameter gradients, then g = dy * gamma in gv
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;
      ag[i].x += gv[i].x * xv[i].x; ag[i].y += gv[i].y * xv[i].y;
      ag[i].z += gv[i].z * xv[i].z; ag[i].w += gv[i].w * xv[i].w;
      ab[i].x += gv[i].x; ab[i].y += gv[i].y; ab[i].z += gv[i].z; ab[i].w += gv[i].w;
      gv[i].x *= gam[i].x; gv[i].y *= gam[i].y; gv[i].z *= gam[i].z; gv[i].w *= gam[i].w;
      sg += gv[i].x + gv[i].y + gv[i].z + gv[i].w;
      sgx += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
    }
    const float mg = warp_sum(sg) * inv_d, mgx = warp_sum(sgx) * inv_d;
    float4* dr = reinterpret_cast<float4*>(dx + static_cast<long long>(row) * ldd);
    const float4* rs = dres ? reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * ldr)
                            : nullptr;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float4 o;
      o.x = rstd * (gv[i].x - mg - xv[i].x * mgx);
      o.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
      o.z = rstd * (gv[i].z - mg - xv[i].z * mgx);
      o.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
      if (rs != nullptr) {
        const float4 r4 = ld_act4(rs + lane + 32 * i);
        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
      }
      dr[lane + 32 * i] = o;
    }
  }
  // per-WARP partials of d gamma | d beta: part[(cta * 8 + warp)][2][D], summed by the reduce kernel
  float4* pg = reinterpret_cast<float4*>(part + static_cast<long long>(blockIdx.x * 8 + warp) * 2 * D);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    pg[lane + 32 * i] = ag[i];
    pg[D / 4 + lane + 32 * i] = ab[i];
  }
}

// ------------------------------------------------------------------------------------------------
// CSGU backward, convolution part.  Forward: h = [r | g], n = LN(g), c = dwconv_k(n) + cb, u = r c.
// Given du:   dr = du c;   dc = du r;   dn[t] = sum_j w[j] dc[t - j + half];
//             d w[j] = sum_t dc[t] n[t + j - half];   d cb = sum_t dc[t].
// (The LayerNorm part, dg = LN_bwd(g; dn), is layernorm_bwd_kernel on the gate half of h.)
// CTA = 128 channels x kSeg frames of one utterance; the n tile and the dc tile (kSeg + 2 half
// rows each) live in shared memory; thread = channel.  dr goes to dh[:, :C], dn to its own buffer;
// per-CTA partials of d w (k taps) and d cb go to part[cta][C-slab][32] (slot 31 = d cb).
// ------------------------------------------------------------------------------------------------
constexpr int kTaps = 31, kHalo = 15, kSeg = 64, kCh = 128, kRows = kSeg + 2 * kHalo;

__global__ void __launch_bounds__(kCh)
csgu_conv_bwd_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ norm_g,
                     const float* __restrict__ norm_b, const float* __restrict__ conv_w,
                     const float* __restrict__ conv_b, const float2* __restrict__ stats,
                     const float* __restrict__ du, long long ldu, float* __restrict__ dh,
                     long long lddh, float* __restrict__ dn, long long lddn,
                     float* __restrict__ part, int T, int Ch) {
  extern __shared__ float s_mem[];
  float* s_n = s_mem;                    // [kRows][kCh]  LN(g), exactly 0 outside [0, T)
  float* s_dc = s_mem + kRows * kCh;     // [kRows][kCh]  du * r, 0 outside [0, T)
  pdl_launch_dependents();
  const int c = blockIdx.x * kCh + threadIdx.x;
  float w[kTaps];
#pragma unroll
  for (int k = 0; k < kTaps; ++k) w[k] = __ldg(conv_w + static_cast<long long>(c) * kTaps + k);
  const float gam = __ldg(norm_g + c), bet = __ldg(norm_b + c), cb = __ldg(conv_b + c);
  pdl_wait();
  const int t0 = blockIdx.y * kSeg;
  const int b = blockIdx.z;
  const long long row0 = static_cast<long long>(b) * T;
  // tile load, eight rows per pass, software-pipelined: the 32 loads of pass p + 1 are issued
  // before pass p's values are consumed, so the kernel waits for one L2 round trip per tile instead
  // of one per pass (one row at a time it spent most of its 280 us on dependent round trips)
  struct RowBatch { float2 st[8]; float gq[8], uq[8], rq[8]; };
  auto fetch = [&](RowBatch& v, int rb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 - kHalo + rb + i;
      const bool in = rb + i < kRows && t >= 0 && t < T;
      v.st[i] = in ? stats[row0 + t] : make_float2(0.f, 0.f);
      v.gq[i] = in ? ld_act(h + (row0 + t) * ldh + Ch + c) : 0.f;
      v.uq[i] = in ? ld_act(du + (row0 + t) * ldu + c) : 0.f;
      v.rq[i] = in ? ld_act(h + (row0 + t) * ldh + c) : 0.f;
    }
  };
  auto stash = [&](const RowBatch& v, int rb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 - kHalo + rb + i;
      if (rb + i < kRows) {
        const bool in = t >= 0 && t < T;
        s_n[(rb + i) * kCh + threadIdx.x] = in ? (v.gq[i] - v.st[i].x) * v.st[i].y * gam + bet : 0.f;
        s_dc[(rb + i) * kCh + threadIdx.x] = v.uq[i] * v.rq[i];
      }
    }
  };
  {
    RowBatch ba, bb;
    fetch(ba, 0);
    for (int rb = 0; rb < kRows; rb += 16) {
      fetch(bb, rb + 8);
      stash(ba, rb);
      fetch(ba, rb + 16);
      stash(bb, rb + 8);
    }
  }
  // every thread only reads the column it wrote: no barrier needed
  float dw[kTaps];
#pragma unroll
  for (int k = 0; k < kTaps; ++k) dw[k] = 0.f;
  float dcb = 0.f;
  // 16 output frames per pass out of a REGISTER window of the 46 tile rows they touch (the
  // one-frame-at-a-time form issued 62 shared-memory loads per 93 FMAs and ran at 286 us; here the
  // 1488 FMAs of a pass read 92 values once)
  constexpr int kOut = 16, kWin = kOut + 2 * kHalo;
  static_assert(kSeg % kOut == 0, "segment is a whole number of passes");
  for (int o0 = 0; o0 < kSeg; o0 += kOut) {
    if (t0 + o0 >= T) break;
    float nw[kWin], dcw[kWin];
#pragma unroll
    for (int i = 0; i < kWin; ++i) {
      nw[i] = s_n[(o0 + i) * kCh + threadIdx.x];      // tile row o0 + i <-> frame t0 + o0 - kHalo + i
      dcw[i] = s_dc[(o0 + i) * kCh + threadIdx.x];
    }
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
      const int t = t0 + o0 + j;
      float conv = cb, dnv = 0.f;
      const float dct = dcw[j + kHalo];               // 0 for t >= T: no contribution below
#pragma unroll
      for (int k = 0; k < kTaps; ++k) {
        const float nk = nw[j + k];
        conv = fmaf(w[k], nk, conv);                  // c[t]
        dnv = fmaf(w[k], dcw[j + 2 * kHalo - k], dnv);  // dn[t]
        dw[k] = fmaf(dct, nk, dw[k]);                 // d w[k]
      }
      dcb += dct;
      if (t < T) {
        dh[(row0 + t) * lddh + c] = ld_act(du + (row0 + t) * ldu + c) * conv;  // dr
        dn[(row0 + t) * lddn + c] = dnv;
      }
    }
  }
  float* pp = part + ((static_cast<long long>(b) * gridDim.y + blockIdx.y) * Ch + c) * 32;
#pragma unroll
  for (int k = 0; k < kTaps; ++k) pp[k] = dw[k];
  pp[31] = dcb;
}

// sums the per-CTA partials: out_w[c][k] (k < 31), out_b[c]
__global__ void __launch_bounds__(256)
csgu_conv_bwd_reduce_kernel(const float* __restrict__ part, int nblk, int Ch,
                            float* __restrict__ dconv_w, float* __restrict__ dconv_b) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;  // (c, slot)
  if (i >= Ch * 32) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += ld_act(part + static_cast<long long>(b) * Ch * 32 + i);
  const int c = i >> 5, k = i & 31;
  if (k < kTaps) dconv_w[c * kTaps + k] = s; else dconv_b[c] = s;
}


// ------------------------------------------------------------------------------------------------
// learned_ave merge backward (oracle/bwd_formulas.py::learned_ave_merge_bwd), one CTA of 1024 threads
// per utterance, D == 256 (tests/test_backward_gpu.py::test_merge_learned_ave_bwd).
//   forward:  score_i[t] = (x_i[t].a_i + c_i)/sqrt(D) (t < len), s_i = softmax_t, pooled_i = sum_t
//             s_i[t] x_i[t], omega_i = pooled_i.b_i + e_i, w = softmax(omega), m = w1 x1 + w2 x2
//   given dm: dw_i = sum_{t,d} dm x_i (ALL T frames: the sum m is dense), domega = w (dw - w.dw),
//             dpool_i = domega_i b_i, ds_i[t] = x_i[t].dpool_i, dsc_i = s_i (ds_i - s_i.ds_i)/sqrt(D)
//             dx_i[t] = w_i dm[t] + s_i[t] dpool_i + dsc_i[t] a_i
//             da_i = sum_t dsc_i[t] x_i[t], dc_i = sum_t dsc_i[t], db_i = domega_i pooled_i, de_i = domega_i
// Row dots run warp-per-row, column accumulations thread-per-column (coalesced either way).
// Per-utterance parameter-gradient partials: part[b][4][D] = (da1, db1, da2, db2), part_s[b][4] =
// (dc1, de1, dc2, de2); the caller reduces them over b with tavsr_col_sums' reduce kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kMergeT = 2048;  // frames per utterance held in shared memory

constexpr int kMergeThreads = 1024;  // 32 warps: only B CTAs exist, so a CTA takes a whole SM
constexpr int kMergeWarps = kMergeThreads / 32;
constexpr int kMergeGroups = kMergeThreads / 256;   // row groups of the thread-per-column phases

__device__ __forceinline__ float block_sum_merge(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kMergeWarps; ++i) r += s_red[i];
  return r;
}
__device__ __forceinline__ float block_max_merge(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
#pragma unroll
  for (int i = 1; i < kMergeWarps; ++i) r = fmaxf(r, s_red[i]);
  return r;
}

struct MergeBwdParams {
  const float* x1; const float* x2; const float* dm;
  long long ld1, ld2, ldm;
  const int32_t* lens;
  const int32_t* lens2;   // optional: branch 2 masked by its own lengths (audio-visual fusion)
  const float* a1; const float* b1; const float* a2; const float* b2;  // [D] vectors
  const float* scal;   // device [4]: c1, e1, c2, e2 (the Linear(D,1) biases; no host read-back)
  float* dx1; float* dx2;
  long long ldd1, ldd2;
  float* part;    // [B][4][D]
  float* part_s;  // [B][4]
  int T;
};

__global__ void __launch_bounds__(kMergeThreads)
merge_learned_ave_bwd_kernel(const MergeBwdParams p) {
  constexpr int D = 256;
  extern __shared__ float s_mem[];
  float* s_s = s_mem;                  // [2][T]  scores -> softmax weights s_i[t]
  float* s_d = s_mem + 2 * p.T;        // [2][T]  ds_i[t] -> dsc_i[t]
  __shared__ float s_vec[4][D];        // a1, a2 -> later dpool1, dpool2 in rows 2, 3
  __shared__ float s_grp[2][kMergeGroups][D];   // per row-group partials of the column phases
  __shared__ float s_red[kMergeWarps];
  pdl_launch_dependents();
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int col = tid & (D - 1), grp = tid >> 8;
  const uint32_t lane = lane_id();
  const int T = p.T;
  if (tid < D) {
    s_vec[0][tid] = __ldg(p.a1 + tid);
    s_vec[1][tid] = __ldg(p.a2 + tid);
  }
  const float bb1 = __ldg(p.b1 + col), bb2 = __ldg(p.b2 + col);
  pdl_wait();
  const float pc1 = ld_act(p.scal), pe1 = ld_act(p.scal + 1), pc2 = ld_act(p.scal + 2), pe2 = ld_act(p.scal + 3);
  int len1 = p.lens ? p.lens[b] : T;
  len1 = len1 < 0 ? 0 : (len1 > T ? T : len1);
  int len2 = p.lens2 ? p.lens2[b] : len1;
  len2 = len2 < 0 ? 0 : (len2 > T ? T : len2);
  // softmax weights, ds and dsc of a branch are exactly 0 at and beyond that branch's length, so the
  // phases that walk both branches together run to the longer one
  const int len = len1 > len2 ? len1 : len2;
  const long long row0 = static_cast<long long>(b) * T;
  const float rs = rsqrtf(static_cast<float>(D));
  __syncthreads();
  // ---- phase 1 (warp per row): scores and the dense dw_i = sum dm . x_i ----
  float dw1 = 0.f, dw2 = 0.f;
  for (int t = warp; t < T; t += kMergeWarps) {
    const float4* r1 = reinterpret_cast<const float4*>(p.x1 + (row0 + t) * p.ld1);
    const float4* r2 = reinterpret_cast<const float4*>(p.x2 + (row0 + t) * p.ld2);
    const float4* rm = reinterpret_cast<const float4*>(p.dm + (row0 + t) * p.ldm);
    float sa1 = 0.f, sa2 = 0.f, sm1 = 0.f, sm2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 v1 = ld_act4(r1 + lane + 32 * i), v2 = ld_act4(r2 + lane + 32 * i);
      const float4 g = ld_act4(rm + lane + 32 * i);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_vec[0][4 * (lane + 32 * i)]);
      const float4 w2 = *reinterpret_cast<const float4*>(&s_vec[1][4 * (lane + 32 * i)]);
      sa1 += v1.x * w1.x + v1.y * w1.y + v1.z * w1.z + v1.w * w1.w;
      sa2 += v2.x * w2.x + v2.y * w2.y + v2.z * w2.z + v2.w * w2.w;
      sm1 += v1.x * g.x + v1.y * g.y + v1.z * g.z + v1.w * g.w;
      sm2 += v2.x * g.x + v2.y * g.y + v2.z * g.z + v2.w * g.w;
    }
    sa1 = warp_sum(sa1); sa2 = warp_sum(sa2);
    dw1 += warp_sum(sm1); dw2 += warp_sum(sm2);   // identical in every lane of the warp
    if (lane == 0) {
      s_s[t] = (sa1 + pc1) * rs;
      s_s[T + t] = (sa2 + pc2) * rs;
    }
  }
  // one lane per warp carries the warp's dw partial into the block sums
  dw1 = block_sum_merge(lane == 0 ? dw1 : 0.f, s_red);
  dw2 = block_sum_merge(lane == 0 ? dw2 : 0.f, s_red);
  // ---- phase 2: masked softmax over t < len for both branches ----
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    float* sc = s_s + br * T;
    const int lb = br == 0 ? len1 : len2;
    float mx = -INFINITY;
    for (int t = tid; t < lb; t += kMergeThreads) mx = fmaxf(mx, sc[t]);
    mx = block_max_merge(mx, s_red);
    float sum = 0.f;
    for (int t = tid; t < T; t += kMergeThreads) {
      const float e = t < lb ? expf(sc[t] - mx) : 0.f;
      sc[t] = e;
      sum += e;
    }
    const float se = block_sum_merge(sum, s_red);
    const float inv = lb > 0 ? 1.0f / se : 0.f;
    for (int t = tid; t < T; t += kMergeThreads) sc[t] *= inv;
  }
  __syncthreads();
  // ---- phase 3 (thread per column, rows split over the groups): pooled_i[d], omega, w, domega ----
  float pool1 = 0.f, pool2 = 0.f;
  for (int t = grp; t < len; t += kMergeGroups) {
    pool1 = fmaf(s_s[t], ld_act(p.x1 + (row0 + t) * p.ld1 + col), pool1);
    pool2 = fmaf(s_s[T + t], ld_act(p.x2 + (row0 + t) * p.ld2 + col), pool2);
  }
  s_grp[0][grp][col] = pool1;
  s_grp[1][grp][col] = pool2;
  __syncthreads();
  pool1 = 0.f; pool2 = 0.f;
#pragma unroll
  for (int k = 0; k < kMergeGroups; ++k) { pool1 += s_grp[0][k][col]; pool2 += s_grp[1][k][col]; }
  const float om1 = block_sum_merge(grp == 0 ? pool1 * bb1 : 0.f, s_red) + pe1;
  const float om2 = block_sum_merge(grp == 0 ? pool2 * bb2 : 0.f, s_red) + pe2;
  const float mo = fmaxf(om1, om2);
  const float e1 = expf(om1 - mo), e2 = expf(om2 - mo);
  const float w1 = e1 / (e1 + e2), w2 = e2 / (e1 + e2);
  const float wd = w1 * dw1 + w2 * dw2;
  const float dom1 = w1 * (dw1 - wd), dom2 = w2 * (dw2 - wd);
  const float dp1 = dom1 * bb1, dp2 = dom2 * bb2;   // dpool_i[d]
  float* pp = p.part + static_cast<long long>(b) * 4 * D;
  if (grp == 0) {
    s_vec[2][col] = dp1;
    s_vec[3][col] = dp2;
    pp[1 * D + col] = dom1 * pool1;  // db1
    pp[3 * D + col] = dom2 * pool2;  // db2
  }
  __syncthreads();
  // ---- phase 4 (warp per row): ds_i[t] = x_i[t] . dpool_i ----
  for (int t = warp; t < len; t += kMergeWarps) {
    const float4* r1 = reinterpret_cast<const float4*>(p.x1 + (row0 + t) * p.ld1);
    const float4* r2 = reinterpret_cast<const float4*>(p.x2 + (row0 + t) * p.ld2);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 v1 = ld_act4(r1 + lane + 32 * i), v2 = ld_act4(r2 + lane + 32 * i);
      const float4 q1 = *reinterpret_cast<const float4*>(&s_vec[2][4 * (lane + 32 * i)]);
      const float4 q2 = *reinterpret_cast<const float4*>(&s_vec[3][4 * (lane + 32 * i)]);
      s1 += v1.x * q1.x + v1.y * q1.y + v1.z * q1.z + v1.w * q1.w;
      s2 += v2.x * q2.x + v2.y * q2.y + v2.z * q2.z + v2.w * q2.w;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { s_d[t] = s1; s_d[T + t] = s2; }
  }
  __syncthreads();
  // ---- phase 5: dsc_i[t] = s_i[t] (ds_i[t] - sum_t s_i ds_i) / sqrt(D); dc_i = sum_t dsc_i[t] ----
  float dcs[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const int lb = br == 0 ? len1 : len2;
    float acc = 0.f;
    for (int t = tid; t < lb; t += kMergeThreads) acc += s_s[br * T + t] * s_d[br * T + t];
    const float sds = block_sum_merge(acc, s_red);
    float dc = 0.f;
    for (int t = tid; t < T; t += kMergeThreads) {
      const float v = t < lb ? s_s[br * T + t] * (s_d[br * T + t] - sds) * rs : 0.f;
      s_d[br * T + t] = v;
      dc += v;
    }
    dcs[br] = block_sum_merge(dc, s_red);
  }
  __syncthreads();
  // ---- phase 6 (thread per column, rows split over the groups): dx_i and da_i ----
  const float av1 = s_vec[0][col], av2 = s_vec[1][col];
  float da1 = 0.f, da2 = 0.f;
  for (int t = grp; t < T; t += kMergeGroups) {
    const float g = ld_act(p.dm + (row0 + t) * p.ldm + col);
    float o1 = w1 * g, o2 = w2 * g;
    if (t < len) {
      const float v1 = ld_act(p.x1 + (row0 + t) * p.ld1 + col);
      const float v2 = ld_act(p.x2 + (row0 + t) * p.ld2 + col);
      const float k1 = s_d[t], k2 = s_d[T + t];
      o1 += s_s[t] * dp1 + k1 * av1;
      o2 += s_s[T + t] * dp2 + k2 * av2;
      da1 = fmaf(k1, v1, da1);
      da2 = fmaf(k2, v2, da2);
    }
    p.dx1[(row0 + t) * p.ldd1 + col] = o1;
    p.dx2[(row0 + t) * p.ldd2 + col] = o2;
  }
  s_grp[0][grp][col] = da1;
  s_grp[1][grp][col] = da2;
  __syncthreads();
  if (grp == 0) {
    da1 = 0.f; da2 = 0.f;
#pragma unroll
    for (int k = 0; k < kMergeGroups; ++k) { da1 += s_grp[0][k][col]; da2 += s_grp[1][k][col]; }
    pp[0 * D + col] = da1;
    pp[2 * D + col] = da2;
  }
  if (tid == 0) {
    float* ps = p.part_s + b * 4;
    ps[0] = dcs[0]; ps[1] = dom1; ps[2] = dcs[1]; ps[3] = dom2;
  }
}

// ------------------------------------------------------------------------------------------------
// Conv2dSubsampling front end, backward of its first half (espnet Conv2dSubsampling:
// Conv2d(1, C, 3, 2) + ReLU + Conv2d(C, C, 3, 2) + ReLU; encoder.py:149-155).  The forward writes the
// im2col operand A[(b, t2, f2)][(kt, kf, c)] = relu(conv1(x))[b, c, 2 t2 + kt, 2 f2 + kf] of the second
// convolution (conv2d_sub_im2col_kernel); given dA (from the conv2 GEMM's dgrad) this kernel does the
// col2im gather  dh1[b, c, t1, f1] = sum over the <= 4 windows (kt, kf) with t1 - kt, f1 - kf even,
// re-evaluates conv1 (9 FMAs) for the ReLU mask, and accumulates d w1[c][3][3], d b1[c].
// CTA = one (b, t1), thread = channel c (coalesced dA reads); per-CTA partials part[cta][c * 10 + k]
// (k = 9: bias) are summed by col_sums_reduce_kernel.  The input features get no gradient.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv2d_sub_bwd_kernel(const float* __restrict__ x, int Tin, int F, const float* __restrict__ w1,
                      const float* __restrict__ b1, int C, int T1, int F1, int T2, int F2,
                      const float* __restrict__ dA, float* __restrict__ part) {
  extern __shared__ float s_x[];  // 3 input rows x F
  pdl_launch_dependents();
  const int b = blockIdx.x / T1, t1 = blockIdx.x % T1;
  pdl_wait();
  const float* xb = x + (static_cast<long long>(b) * Tin + 2 * t1) * F;
  for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) s_x[i] = ld_act(xb + i);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w[9], dw[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { w[k] = __ldg(w1 + c * 9 + k); dw[k] = 0.f; }
    const float bias = __ldg(b1 + c);
    float db = 0.f;
    for (int f1 = 0; f1 < F1; ++f1) {
      float dh = 0.f;
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int tt = t1 - kt;
        if (tt < 0 || (tt & 1) || (tt >> 1) >= T2) continue;
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          const int ff = f1 - kf;
          if (ff < 0 || (ff & 1) || (ff >> 1) >= F2) continue;
          const long long row = (static_cast<long long>(b) * T2 + (tt >> 1)) * F2 + (ff >> 1);
          dh += ld_act(dA + row * (9ll * C) + (kt * 3 + kf) * C + c);
        }
      }
      const float* xp = s_x + 2 * f1;
      float z = bias;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) z = fmaf(w[i * 3 + j], xp[i * F + j], z);
      const float dz = z > 0.f ? dh : 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dw[i * 3 + j] = fmaf(dz, xp[i * F + j], dw[i * 3 + j]);
      db += dz;
    }
    float* pp = part + (static_cast<long long>(blockIdx.x) * C + c) * 10;
#pragma unroll
    for (int k = 0; k < 9; ++k) pp[k] = dw[k];
    pp[9] = db;
  }
}

}  // namespace bwd
}  // namespace tavsr

using namespace tavsr;

static int launch_ew_transpose(int op, const float* a, long long lda, const float* b, long long ldb,
                               const float* mask, long long ldm, float* out, long long ldo,
                               float* outT, long long ldt, int R, int C, int act, cudaStream_t s) {
  // the tail of every transposed row up to the next multiple of 4 (within the pitch) is zero-filled
  const int Rpad = static_cast<int>(ldt < ((R + 3) / 4) * 4 ? ldt : ((R + 3) / 4) * 4);
  const dim3 grid((C + 63) / 64, (R + 63) / 64);
  if (op == 0)
    TAVSR_CUDA_OK(launch_kernel(bwd::ew_transpose_kernel<0>, grid, dim3(256), 0, s, 0, a, lda, b, ldb, mask,
                                ldm, out, ldo, outT, ldt, R, C, Rpad, act));
  else if (op == 1)
    TAVSR_CUDA_OK(launch_kernel(bwd::ew_transpose_kernel<1>, grid, dim3(256), 0, s, 0, a, lda, b, ldb, mask,
                                ldm, out, ldo, outT, ldt, R, C, Rpad, act));
  else
    TAVSR_CUDA_OK(launch_kernel(bwd::ew_transpose_kernel<2>, grid, dim3(256), 0, s, 0, a, lda, b, ldb, mask,
                                ldm, out, ldo, outT, ldt, R, C, Rpad, act));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_transpose_2d(const float* in, long long ld_in, float* out, long long ld_out,
                                  int R, int C, void* stream) {
  TAVSR_REQUIRE(R > 0 && C > 0 && in && out && ld_in >= C && ld_out >= R, "transpose: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ld_out >= ((R + 3) / 4) * 4)
    return launch_ew_transpose(0, in, ld_in, nullptr, 0, nullptr, 0, nullptr, 0, out, ld_out, R, C, 0, s);
  // the tail of every output row up to the next multiple of 4 (within the pitch) is zero-filled
  const int Rpad = static_cast<int>(ld_out < ((R + 3) / 4) * 4 ? ld_out : ((R + 3) / 4) * 4);
  TAVSR_CUDA_OK(launch_kernel(bwd::transpose_kernel, dim3((C + 31) / 32, (R + 31) / 32), dim3(256), 0,
                              s, 0, in, ld_in, out, ld_out, R, C, Rpad));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// h^T = (act(z) [* mask])^T: the recomputed FFN hidden in the operand form of its weight gradient.
extern "C" int tavsr_act_fwd_t(const float* z, long long ldz, const float* mask, long long ldm,
                               float* hT, long long ldt, int M, int C, int act, void* stream) {
  TAVSR_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && ldz % 4 == 0 && ldt % 4 == 0 && ldt >= ((M + 3) / 4) * 4 &&
                    (!mask || ldm % 4 == 0) && z && hT && act >= 0 && act <= 3,
                "act_fwd_t: bad arguments (M=%d C=%d act=%d)", M, C, act);
  return launch_ew_transpose(1, z, ldz, nullptr, 0, mask, ldm, nullptr, 0, hT, ldt, M, C, act,
                             static_cast<cudaStream_t>(stream));
}

// dz = dh * act'(z), written row-major (dgrad operand) AND transposed (wgrad operand) in one pass.
extern "C" int tavsr_act_bwd_t(const float* z, long long ldz, const float* dh, long long ldh, float* dz,
                               long long ldd, float* dzT, long long ldt, int M, int C, int act,
                               void* stream) {
  TAVSR_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && ldz % 4 == 0 && ldh % 4 == 0 && ldd % 4 == 0 &&
                    ldt % 4 == 0 && ldt >= ((M + 3) / 4) * 4 && z && dh && dz && dzT && act >= 0 && act <= 3,
                "act_bwd_t: bad arguments (M=%d C=%d act=%d)", M, C, act);
  return launch_ew_transpose(2, z, ldz, dh, ldh, nullptr, 0, dz, ldd, dzT, ldt, M, C, act,
                             static_cast<cudaStream_t>(stream));
}

extern "C" size_t tavsr_col_sums_workspace_bytes(int R, int C) {
  return static_cast<size_t>((R + bwd::kColRows - 1) / bwd::kColRows) * C * sizeof(float);
}

extern "C" int tavsr_col_sums(const float* a, long long lda, const float* b, long long ldb,
                              float* out, void* workspace, long long workspace_bytes, int R, int C,
                              void* stream) {
  TAVSR_REQUIRE(R > 0 && C > 0 && a && out && workspace, "col_sums: bad arguments");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_col_sums_workspace_bytes(R, C),
                "col_sums: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nblk = (R + bwd::kColRows - 1) / bwd::kColRows;
  float* part = static_cast<float*>(workspace);
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_partial_kernel, dim3((C + 127) / 128, nblk), dim3(128), 0,
                              s, 0, a, lda, b, ldb, part, R, C));
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_reduce_kernel, dim3((C + 31) / 32), dim3(1024), 0, s, 0,
                              static_cast<const float*>(part), nblk, out, C));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_act_fwd(const float* z, long long ldz, const float* mask, long long ldm, float* h,
                             long long ldh, int M, int C, int act, void* stream) {
  TAVSR_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && ldz % 4 == 0 && ldh % 4 == 0 && z && h && act >= 0 &&
                    act <= 3 && (!mask || ldm % 4 == 0),
                "act_fwd: bad arguments (M=%d C=%d act=%d)", M, C, act);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(M) * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
  TAVSR_CUDA_OK(launch_kernel(bwd::act_fwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s,
                              0, z, ldz, mask, ldm, h, ldh, M, C / 4, act));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_act_bwd(const float* z, long long ldz, const float* dh, long long ldh,
                             float* dz, long long ldd, int M, int C, int act, void* stream) {
  TAVSR_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && ldz % 4 == 0 && ldh % 4 == 0 && ldd % 4 == 0 && z &&
                    dh && dz && act >= 0 && act <= 3,
                "act_bwd: bad arguments (M=%d C=%d act=%d)", M, C, act);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(M) * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
  TAVSR_CUDA_OK(launch_kernel(bwd::act_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s,
                              0, z, ldz, dh, ldh, dz, ldd, M, C / 4, act));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" size_t tavsr_layernorm_bwd_workspace_bytes(int M, int D) {
  const size_t slots = (static_cast<size_t>(M) + bwd::kLnRowsPerWarp - 1) / bwd::kLnRowsPerWarp;
  return slots * 2 * D * sizeof(float);
}

extern "C" int tavsr_layernorm_bwd(const float* x, long long ldx, const float* gamma,
                                   const float* dy, long long ldy, const float* dres, long long ldr,
                                   float* dx, long long ldd, float* dgamma, float* dbeta,
                                   void* workspace, long long workspace_bytes, int M, int D,
                                   float eps, void* stream) {
  TAVSR_REQUIRE(M > 0 && (D == 256 || D == 512 || D == 1024) && x && gamma && dy && dx && dgamma &&
                    dbeta && workspace,
                "layernorm_bwd: built for D in {256, 512, 1024} (M=%d D=%d)", M, D);
  TAVSR_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldd % 4 == 0 && (!dres || ldr % 4 == 0) &&
                    dgamma + D == dbeta,
                "layernorm_bwd: pitches must be multiples of 4 and dgamma / dbeta one [2, D] buffer");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_layernorm_bwd_workspace_bytes(M, D),
                "layernorm_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_cta = 8 * bwd::kLnRowsPerWarp;
  const int nblk = (M + rows_per_cta - 1) / rows_per_cta;
  float* part = static_cast<float*>(workspace);
#define TAVSR_LNB(V)                                                                                 \
  TAVSR_CUDA_OK(launch_kernel(bwd::layernorm_bwd_kernel<V>, dim3(nblk), dim3(256), 0, s, 0, x, ldx,  \
                              gamma, dy, ldy, dres, ldr, dx, ldd, part, M, eps))
  if (D == 256) TAVSR_LNB(2); else if (D == 512) TAVSR_LNB(4); else TAVSR_LNB(8);
#undef TAVSR_LNB
  // [nblk][2 D] partials -> dgamma | dbeta (contiguous [2, D])
  const int slots = (M + bwd::kLnRowsPerWarp - 1) / bwd::kLnRowsPerWarp;
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_reduce_kernel, dim3((2 * D + 31) / 32), dim3(1024), 0, s, 0,
                              static_cast<const float*>(part), slots, dgamma, 2 * D));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" size_t tavsr_csgu_bwd_workspace_bytes(int B, int T, int Ch) {
  const size_t nseg = (static_cast<size_t>(T) + bwd::kSeg - 1) / bwd::kSeg;
  return static_cast<size_t>(B) * nseg * Ch * 32 * sizeof(float);
}

extern "C" int tavsr_csgu_conv_bwd(const float* h, long long ldh, const float* norm_g,
                                   const float* norm_b, const float* conv_w, const float* conv_b,
                                   const float* stats, const float* du, long long ldu, float* dh,
                                   long long lddh, float* dn, long long lddn, float* dconv_w,
                                   float* dconv_b, void* workspace, long long workspace_bytes, int B,
                                   int T, int Ch, int ksize, void* stream) {
  TAVSR_REQUIRE(ksize == bwd::kTaps, "csgu_bwd: only kernel size 31 is built (got %d)", ksize);
  TAVSR_REQUIRE(B > 0 && T > 0 && Ch > 0 && Ch % 128 == 0 && h && norm_g && norm_b && conv_w &&
                    conv_b && stats && du && dh && dn && dconv_w && dconv_b && workspace,
                "csgu_bwd: bad arguments");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_csgu_bwd_workspace_bytes(B, T, Ch),
                "csgu_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int smem = 2 * bwd::kRows * bwd::kCh * 4;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    TAVSR_CUDA_OK(cudaFuncSetAttribute(bwd::csgu_conv_bwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int nseg = (T + bwd::kSeg - 1) / bwd::kSeg;
  float* part = static_cast<float*>(workspace);
  TAVSR_CUDA_OK(launch_kernel(bwd::csgu_conv_bwd_kernel, dim3(Ch / bwd::kCh, nseg, B), dim3(bwd::kCh),
                              static_cast<size_t>(smem), s, 0, h, ldh, norm_g, norm_b, conv_w, conv_b,
                              reinterpret_cast<const float2*>(stats), du, ldu, dh, lddh, dn, lddn, part,
                              T, Ch));
  TAVSR_CUDA_OK(launch_kernel(bwd::csgu_conv_bwd_reduce_kernel, dim3((Ch * 32 + 255) / 256), dim3(256), 0,
                              s, 0, static_cast<const float*>(part), B * nseg, Ch, dconv_w, dconv_b));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" size_t tavsr_merge_learned_ave_bwd_workspace_bytes(int B) {
  return static_cast<size_t>(B) * (4 * 256 + 4) * sizeof(float);
}

// grads: [4][256] = (da1, db1, da2, db2) then [4] = (dc1, de1, dc2, de2), i.e. 1028 floats.
extern "C" int tavsr_merge_learned_ave_bwd(const float* x1, long long ld1, const float* x2,
                                           long long ld2, const float* dm, long long ldm,
                                           const int32_t* lens, const int32_t* lens2,
                                           const float* a1, const float* b1, const float* a2,
                                           const float* b2, const float* scal, float* dx1,
                                           long long ldd1,
                                           float* dx2, long long ldd2, float* grads, void* workspace,
                                           long long workspace_bytes, int B, int T, int D,
                                           void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && T <= bwd::kMergeT && D == 256,
                "merge_bwd: built for D == 256, T <= %d (B=%d T=%d D=%d)", bwd::kMergeT, B, T, D);
  TAVSR_REQUIRE(x1 && x2 && dm && a1 && b1 && a2 && b2 && scal && dx1 && dx2 && grads && workspace,
                "merge_bwd: null pointer");
  TAVSR_REQUIRE(ld1 % 4 == 0 && ld2 % 4 == 0 && ldm % 4 == 0, "merge_bwd: pitches must be multiples of 4");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_merge_learned_ave_bwd_workspace_bytes(B),
                "merge_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bwd::MergeBwdParams p;
  p.x1 = x1; p.x2 = x2; p.dm = dm; p.ld1 = ld1; p.ld2 = ld2; p.ldm = ldm; p.lens = lens;
  p.lens2 = lens2;
  p.a1 = a1; p.b1 = b1; p.a2 = a2; p.b2 = b2; p.scal = scal;
  p.dx1 = dx1; p.dx2 = dx2; p.ldd1 = ldd1; p.ldd2 = ldd2;
  p.part = static_cast<float*>(workspace);
  p.part_s = p.part + static_cast<size_t>(B) * 4 * 256;
  p.T = T;
  const int smem = 4 * T * 4;
  static PerDeviceMax configured;
  if (configured.raise(smem)) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(bwd::merge_learned_ave_bwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  TAVSR_CUDA_OK(launch_kernel(bwd::merge_learned_ave_bwd_kernel, dim3(B), dim3(bwd::kMergeThreads),
                              static_cast<size_t>(smem), s, 0, p));
  // reduce the per-utterance partials over b: [B][1024] -> grads[0:1024], [B][4] -> grads[1024:1028]
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_reduce_kernel, dim3(32), dim3(1024), 0, s, 0,
                              static_cast<const float*>(p.part), B, grads, 1024));
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_reduce_kernel, dim3(1), dim3(1024), 0, s, 0,
                              static_cast<const float*>(p.part_s), B, grads + 1024, 4));
  g_launches.fetch_add(3, std::memory_order_relaxed);
  return 0;
}

// d conv1.weight [C, 9] | d conv1.bias [C] of the Conv2dSubsampling front end from dA (see
// conv2d_sub_bwd_kernel); grads = [C][10] (nine taps, then the bias), workspace B * T1 * C * 10 floats.
extern "C" size_t tavsr_conv2d_sub_bwd_workspace_bytes(int B, int Tin, int C) {
  const size_t T1 = static_cast<size_t>((Tin - 1) / 2);
  return static_cast<size_t>(B) * T1 * C * 10 * sizeof(float);
}

extern "C" int tavsr_conv2d_sub_bwd(const float* x, int B, int Tin, int F, const float* w1,
                                    const float* b1, int C, const float* dA, float* grads,
                                    void* workspace, long long workspace_bytes, void* stream) {
  TAVSR_REQUIRE(B > 0 && Tin >= 7 && F >= 7 && C > 0 && x && w1 && b1 && dA && grads && workspace,
                "conv2d_sub_bwd: bad arguments (B=%d Tin=%d F=%d C=%d)", B, Tin, F, C);
  const int T1 = (Tin - 1) / 2, F1 = (F - 1) / 2;
  const int T2 = (T1 - 1) / 2, F2 = (F1 - 1) / 2;
  TAVSR_REQUIRE(T2 >= 1 && F2 >= 1 && 3 * F * 4 <= 48 * 1024, "conv2d_sub_bwd: unsupported shape");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_conv2d_sub_bwd_workspace_bytes(B, Tin, C),
                "conv2d_sub_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  TAVSR_CUDA_OK(launch_kernel(bwd::conv2d_sub_bwd_kernel, dim3(B * T1), dim3(256),
                              static_cast<size_t>(3 * F * 4), s, 0, x, Tin, F, w1, b1, C, T1, F1, T2, F2,
                              dA, part));
  TAVSR_CUDA_OK(launch_kernel(bwd::col_sums_reduce_kernel, dim3((C * 10 + 31) / 32), dim3(1024), 0, s, 0,
                              static_cast<const float*>(part), B * T1, grads, C * 10));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}
