// CTC scorer kernels: fp32 head (logits + log-softmax / softmax / argmax), warp-per-utterance
// log-domain forward-backward loss with gradient w.r.t. the logits, greedy decode, prefix scoring.
// Reference: src/ctc/ctc.py:58-69,133-188 (torch.nn.CTCLoss(reduction="none", zero_infinity)),
// greedy call sites src/models/espnet_model.py:590-592, prefix scoring via espnet
// CTCPrefixScoreTH (src/inference/asr_inference.py:142).
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;

namespace ctc {

constexpr int kVPad = 64;  // V <= 64 (EN 41, ES 37)

// ------------------------------------------------------------------------------------------------
// Head: logits in plain fp32 FMA (argmax must not flip on near ties, so no TF32 here).
// Block = 8 warps; W^T staged in shared memory as [D][64]; each warp handles 4 frames per pass,
// lane l produces vocabulary entries l and l+32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ctc_head_kernel(const float* __restrict__ hs, long long ldh, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ logits,
                float* __restrict__ logp, float* __restrict__ prob, int64_t* __restrict__ amax, int M,
                int D, int V) {
  pdl_launch_dependents();
  extern __shared__ float sm[];
  float* Wt = sm;                       // [D][64]
  float* Hs = sm + D * kVPad;           // [8 warps][4 frames][D]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < D * kVPad; i += 256) {
    const int k = i / kVPad, v = i % kVPad;
    Wt[i] = v < V ? __ldg(w + static_cast<long long>(v) * D + k) : 0.f;
  }
  __syncthreads();
  pdl_wait();  // W^T staging above only read weights
  const float b0 = lane < V ? __ldg(bias + lane) : 0.f;
  const float b1 = lane + 32 < V ? __ldg(bias + lane + 32) : 0.f;
  float* hw = Hs + warp * 4 * D;
  const int groups = (M + 3) / 4;
  for (int grp = blockIdx.x * 8 + warp; grp < groups; grp += gridDim.x * 8) {
    const int m0 = grp * 4;
    __syncwarp();
    for (int i = lane * 4; i < 4 * D; i += 128) {
      const int f = i / D, k = i % D;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + f < M)
        x = ld_act4(reinterpret_cast<const float4*>(hs + static_cast<long long>(m0 + f) * ldh + k));
      *reinterpret_cast<float4*>(hw + i) = x;
    }
    __syncwarp();
    float acc[4][2];
#pragma unroll
    for (int f = 0; f < 4; ++f) { acc[f][0] = b0; acc[f][1] = b1; }
    for (int k = 0; k < D; k += 4) {
      float4 x[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) x[f] = *reinterpret_cast<const float4*>(hw + f * D + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float w0 = Wt[(k + kk) * kVPad + lane];
        const float w1 = Wt[(k + kk) * kVPad + lane + 32];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const float xv = kk == 0 ? x[f].x : (kk == 1 ? x[f].y : (kk == 2 ? x[f].z : x[f].w));
          acc[f][0] = fmaf(xv, w0, acc[f][0]);
          acc[f][1] = fmaf(xv, w1, acc[f][1]);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int m = m0 + f;
      if (m >= M) break;  // warp-uniform
      const float x0 = lane < V ? acc[f][0] : -INFINITY;
      const float x1 = lane + 32 < V ? acc[f][1] : -INFINITY;
      // argmax with lowest-index tie-break (torch.argmax returns the first maximal index)
      float bv = x0;
      int bi = lane;
      if (x1 > bv) { bv = x1; bi = lane + 32; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      const float e0 = lane < V ? expf(x0 - bv) : 0.f;
      const float e1 = lane + 32 < V ? expf(x1 - bv) : 0.f;
      const float se = warp_sum(e0 + e1);
      const float lse = bv + logf(se);
      if (logits) {
        if (lane < V) logits[static_cast<long long>(m) * V + lane] = x0;
        if (lane + 32 < V) logits[static_cast<long long>(m) * V + lane + 32] = x1;
      }
      if (logp) {
        if (lane < V) logp[static_cast<long long>(m) * V + lane] = x0 - lse;
        if (lane + 32 < V) logp[static_cast<long long>(m) * V + lane + 32] = x1 - lse;
      }
      if (prob) {
        const float inv = 1.0f / se;
        if (lane < V) prob[static_cast<long long>(m) * V + lane] = e0 * inv;
        if (lane + 32 < V) prob[static_cast<long long>(m) * V + lane + 32] = e1 * inv;
      }
      if (amax && lane == 0) amax[m] = bi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Head for larger vocabularies (64 < V <= 256: the 256-token SentencePiece alternative of
// configs/ASR/branchformer_transformer+ctc_english.yaml:110-112).  CTA = 8 frames (rows in shared
// memory) x 256 threads; thread v owns vocabulary entry v for all 8 frames and streams its W row
// from L2 (float4, the 256 KB matrix is shared by every CTA), plain fp32 FMA in k order like the
// small-V kernel, so the argmax is just as stable; per frame a block-wide (max, first index) and
// sum-exp reduction.
// ------------------------------------------------------------------------------------------------
constexpr int kBigV = 256;
constexpr int kBigFrames = 8;

__global__ void __launch_bounds__(256)
ctc_head_big_kernel(const float* __restrict__ hs, long long ldh, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ logits,
                    float* __restrict__ logp, float* __restrict__ prob, int64_t* __restrict__ amax,
                    int M, int D, int V) {
  extern __shared__ float sm[];
  float* Hs = sm;                                   // [8][D]
  __shared__ float s_val[kBigFrames][8];
  __shared__ int s_idx[kBigFrames][8];
  __shared__ float s_sum[kBigFrames][8];
  pdl_launch_dependents();
  pdl_wait();
  const int v = threadIdx.x;
  const int warp = v >> 5, lane = v & 31;
  const int m0 = blockIdx.x * kBigFrames;
  for (int i = threadIdx.x * 4; i < kBigFrames * D; i += 1024) {
    const int f = i / D, k = i % D;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + f < M) x = ld_act4(reinterpret_cast<const float4*>(hs + static_cast<long long>(m0 + f) * ldh + k));
    *reinterpret_cast<float4*>(Hs + i) = x;
  }
  __syncthreads();
  float acc[kBigFrames];
  const float bv = v < V ? __ldg(bias + v) : 0.f;
#pragma unroll
  for (int f = 0; f < kBigFrames; ++f) acc[f] = bv;
  if (v < V) {
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<long long>(v) * D);
    for (int k4 = 0; k4 < D / 4; ++k4) {
      const float4 ww = __ldg(wr + k4);
#pragma unroll
      for (int f = 0; f < kBigFrames; ++f) {
        const float4 x = *reinterpret_cast<const float4*>(Hs + f * D + 4 * k4);
        acc[f] = fmaf(x.x, ww.x, acc[f]);
        acc[f] = fmaf(x.y, ww.y, acc[f]);
        acc[f] = fmaf(x.z, ww.z, acc[f]);
        acc[f] = fmaf(x.w, ww.w, acc[f]);
      }
    }
  }
  // ---- per frame: argmax (first maximal index), log-sum-exp ----
#pragma unroll
  for (int f = 0; f < kBigFrames; ++f) {
    float bvv = v < V ? acc[f] : -INFINITY;
    int bi = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bvv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bvv || (ov == bvv && oi < bi)) { bvv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[f][warp] = bvv; s_idx[f][warp] = bi; }
  }
  __syncthreads();
  float mx[kBigFrames];
  int am[kBigFrames];
#pragma unroll
  for (int f = 0; f < kBigFrames; ++f) {
    float bvv = s_val[f][0];
    int bi = s_idx[f][0];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float ov = s_val[f][q];
      const int oi = s_idx[f][q];
      if (ov > bvv || (ov == bvv && oi < bi)) { bvv = ov; bi = oi; }
    }
    mx[f] = bvv;
    am[f] = bi;
    const float e = v < V ? expf(acc[f] - bvv) : 0.f;
    const float se = warp_sum(e);
    if (lane == 0) s_sum[f][warp] = se;
  }
  __syncthreads();
#pragma unroll
  for (int f = 0; f < kBigFrames; ++f) {
    const int m = m0 + f;
    if (m >= M) break;
    float se = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) se += s_sum[f][q];
    if (v < V) {
      const long long o = static_cast<long long>(m) * V + v;
      if (logits) logits[o] = acc[f];
      if (logp) logp[o] = acc[f] - (mx[f] + logf(se));
      if (prob) prob[o] = expf(acc[f] - mx[f]) / se;
    }
    if (amax && v == 0) amax[m] = am[f];
  }
}

// ------------------------------------------------------------------------------------------------
// Vocabulary residual: out = x + p . W^T + b with p (M,V) posteriors and W (D,V) — the InterCTC
// self-conditioning update (encoder.py:393-399) and the second half of InterCTCResidualModule
// (interctc_residual_module.py:14).  Optionally also the LayerNorm of the updated row (the next
// block's norm_ff_macaron), so the conditioned stream needs no extra pass.  One warp per frame,
// lane owns channels lane + 32 c; W^T is staged once per CTA.
// ------------------------------------------------------------------------------------------------
// kBig: 64 < V <= 256 - the posterior row takes up to 8 values per lane and W^T (V x D floats, up
// to 512 KB) is read through L1 / L2 from a pre-transposed global copy instead of shared memory.
template <int kC, bool kBig = false>
__global__ void __launch_bounds__(256)
vocab_residual_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ p,
                      const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, long long ldo, const float* __restrict__ ln_g,
                      const float* __restrict__ ln_b, float eps, void* __restrict__ xn,
                      long long ldn, int M, int V, int xn_bf16) {
  pdl_launch_dependents();
  constexpr int D = kC * 32;
  constexpr int kPS = kBig ? 8 : 2;   // posterior values per lane
  extern __shared__ float sm[];
  const float* Wt = sm;  // [V][D]
  if constexpr (kBig) {
    Wt = w;   // the caller passes W^T (V, D) row-major
  } else {
    for (int i = threadIdx.x; i < V * D; i += 256) {
      const int v = i / D, d = i % D;
      sm[i] = __ldg(w + static_cast<long long>(d) * V + v);
    }
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float bb[kC], gg[kC], be[kC];
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    bb[c] = __ldg(bias + lane + 32 * c);
    gg[c] = ln_g ? __ldg(ln_g + lane + 32 * c) : 1.f;
    be[c] = ln_b ? __ldg(ln_b + lane + 32 * c) : 0.f;
  }
  __syncthreads();
  pdl_wait();
  for (int m = blockIdx.x * 8 + warp; m < M; m += gridDim.x * 8) {
    float pr[kPS];
#pragma unroll
    for (int q = 0; q < kPS; ++q)
      pr[q] = lane + 32 * q < V ? ld_act(p + static_cast<long long>(m) * V + lane + 32 * q) : 0.f;
    float acc[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) acc[c] = bb[c];
#pragma unroll
    for (int q = 0; q < kPS; ++q) {
      const int vend = min(32, V - 32 * q);
      for (int vv = 0; vv < vend; ++vv) {
        const int v = 32 * q + vv;
        const float pv = __shfl_sync(0xffffffffu, pr[q], vv);
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          const float wv = kBig ? __ldg(Wt + static_cast<long long>(v) * D + lane + 32 * c)
                                : Wt[v * D + lane + 32 * c];
          acc[c] = fmaf(pv, wv, acc[c]);
        }
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      acc[c] += ld_act(x + static_cast<long long>(m) * ldx + lane + 32 * c);
      out[static_cast<long long>(m) * ldo + lane + 32 * c] = acc[c];
      sum += acc[c];
    }
    if (xn) {
      const float mean = warp_sum(sum) * (1.0f / D);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float d = acc[c] - mean;
        var = fmaf(d, d, var);
      }
      const float rstd = rsqrtf(warp_sum(var) * (1.0f / D) + eps);
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float y = (acc[c] - mean) * rstd * gg[c] + be[c];
        const long long o = static_cast<long long>(m) * ldn + lane + 32 * c;
        if (xn_bf16) static_cast<uint16_t*>(xn)[o] = static_cast<uint16_t>(pack_bf16x2(y, 0.f) & 0xFFFFu);
        else static_cast<float*>(xn)[o] = y;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Loss: one warp per utterance; the S = 2L+1 lattice states are striped over the lanes in
// contiguous chunks of kC states, so s-1 / s-2 neighbours are mostly in-thread and the chunk
// boundary is crossed with two warp shuffles per time step.
// ------------------------------------------------------------------------------------------------
// The recursions run in the BASE-2 log domain (alpha2 = alpha / ln 2): a log-sum-exp is then
// 3 FADD + 3 ex2 + 2 FADD + lg2 + FADD with no multiplies, and "log zero" is the finite kNeg
// instead of -inf so no branch or NaN guard is needed (kNeg + anything finite stays ~kNeg, and
// ex2(kNeg - m) == 0).  The single warp that owns an utterance is issue-bound, so instruction
// count per lattice state is what sets the speed of this kernel.
constexpr float kNeg = -1e30f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lse2_b2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + lg2f(ex2f(a - m) + ex2f(b - m));
}
__device__ __forceinline__ float lse3_b2(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  return m + lg2f(ex2f(a - m) + ex2f(b - m) + ex2f(c - m));
}
// natural-log versions used by the prefix scorer (logzero = -1e10 there, espnet's constant)
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + __logf(__expf(a - m) + __expf(b - m));
}

// kVS = vocabulary values per lane: 2 (V <= 64) or 8 (V <= 256)
template <int kC, int kVS = 2>
__global__ void __launch_bounds__(128)
ctc_loss_kernel(const float* __restrict__ logp, const int64_t* __restrict__ targets,
                long long ld_targets, const int32_t* __restrict__ hlens,
                const int32_t* __restrict__ tlens, float* __restrict__ nll_out,
                float* __restrict__ grad, float gscale, float* __restrict__ alpha_ws, int B, int T,
                int V, int Lmax, int zero_infinity) {
  constexpr int kVP = 32 * kVS;
  __shared__ float s_row[4][2][kVP];   // per warp, double-buffered log2-prob row
  __shared__ float s_occ[4][kVP];
  __shared__ float s_fin[4][2];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  int L = tlens[b];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  int Tb = hlens[b];
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  const int S = 2 * L + 1;
  const int Sws = 2 * Lmax + 1;
  const float* lp = logp + static_cast<long long>(b) * T * V;
  const int64_t* tg = targets + static_cast<long long>(b) * ld_targets;
  float* aws = alpha_ws ? alpha_ws + static_cast<long long>(b) * T * Sws : nullptr;
  float* gb = grad ? grad + static_cast<long long>(b) * T * V : nullptr;

  // per-state constants
  int lab[kC];
  bool skip_in[kC];   // transition s-2 -> s allowed
  bool skip_out[kC];  // transition s -> s+2 allowed
  bool live[kC];      // s < S
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    const int s = lane * kC + i;
    int l = 0;
    bool si = false, so = false;
    if (s < S && (s & 1)) {
      l = static_cast<int>(tg[(s - 1) >> 1]);
      if (s >= 3) si = static_cast<int>(tg[(s - 3) >> 1]) != l;
      if (s + 2 < S) so = static_cast<int>(tg[(s + 1) >> 1]) != l;
    }
    lab[i] = (l >= 0 && l < V) ? l : 0;
    skip_in[i] = si;
    skip_out[i] = so;
    live[i] = s < S;
  }

  // log-prob rows are prefetched one time step ahead into registers (fetch) and published to the
  // warp through shared memory (commit, converted to base 2), so the global-load latency is off
  // the serial chain.
  float pre[kVS];
#pragma unroll
  for (int q = 0; q < kVS; ++q) pre[q] = 0.f;
  auto fetch = [&](int t) {
    if (t >= 0 && t < Tb) {
#pragma unroll
      for (int q = 0; q < kVS; ++q)
        if (lane + 32 * q < V) pre[q] = ld_act(lp + static_cast<long long>(t) * V + lane + 32 * q);
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int q = 0; q < kVS; ++q)
      if (lane + 32 * q < V) s_row[warp][buf][lane + 32 * q] = pre[q] * kLog2e;
  };

  float nll = INFINITY;
  float a[kC];
  if (Tb > 0) {
    // ---------------- alpha ----------------
    fetch(0);
    commit(0);
    fetch(1);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int s = lane * kC + i;
      a[i] = (s < S && s < 2) ? s_row[warp][0][lab[i]] : kNeg;
      if (aws && s < S) aws[s] = a[i];
    }
    for (int t = 1; t < Tb; ++t) {
      const int buf = t & 1;
      commit(buf);
      fetch(t + 1);
      // neighbours from the previous lane
      float pm1 = __shfl_up_sync(0xffffffffu, a[kC - 1], 1);
      float pm2 = __shfl_up_sync(0xffffffffu, a[kC - 2], 1);
      if (lane == 0) { pm1 = kNeg; pm2 = kNeg; }
      __syncwarp();
      float n[kC];
#pragma unroll
      for (int i = 0; i < kC; ++i) {
        const float x1 = i >= 1 ? a[i - 1] : pm1;
        // s-2 neighbour: in-thread for i >= 2, else the previous lane's last / second-to-last state
        const float x2 = i >= 2 ? a[i - 2] : (i == 1 ? pm1 : pm2);
        const float two = skip_in[i] ? x2 : kNeg;
        const float v = lse3_b2(a[i], x1, two) + s_row[warp][buf][lab[i]];
        n[i] = live[i] ? v : kNeg;
      }
#pragma unroll
      for (int i = 0; i < kC; ++i) {
        a[i] = n[i];
        const int s = lane * kC + i;
        if (aws && s < S) aws[static_cast<long long>(t) * Sws + s] = a[i];
      }
      __syncwarp();
    }
    // final: logaddexp(alpha[S-1], alpha[S-2])
    if (lane == 0) { s_fin[warp][0] = kNeg; s_fin[warp][1] = kNeg; }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int s = lane * kC + i;
      if (s == S - 1) s_fin[warp][0] = a[i];
      if (S >= 2 && s == S - 2) s_fin[warp][1] = a[i];
    }
    __syncwarp();
    const float fin = lse2_b2(s_fin[warp][0], s_fin[warp][1]);
    nll = fin < -1e29f ? INFINITY : -fin * kLn2;
  } else {
    nll = (L == 0) ? 0.f : INFINITY;
  }
  const bool infeasible = !(nll < INFINITY);  // inf or nan
  if (lane == 0) nll_out[b] = (infeasible && zero_infinity) ? 0.f : nll;
  if (!gb) return;

  // ---------------- beta + gradient ----------------
  if (infeasible || Tb == 0) {
    for (int i = lane; i < T * V; i += 32) gb[i] = zero_infinity || Tb == 0 ? 0.f : NAN;
    return;
  }
  for (int i = Tb * V + lane; i < T * V; i += 32) gb[i] = 0.f;
  const float nll2 = nll * kLog2e;
  float bt[kC];
  float apre[kC];  // alpha_t prefetched one step ahead of the beta recursion
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    const int s = lane * kC + i;
    apre[i] = s < S ? aws[static_cast<long long>(Tb - 1) * Sws + s] : kNeg;
    bt[i] = kNeg;
  }
  fetch(Tb - 1);
  for (int t = Tb - 1; t >= 0; --t) {
    const int buf = t & 1;
    commit(buf);
    fetch(t - 1);
#pragma unroll
    for (int q = 0; q < kVS; ++q) s_occ[warp][lane + 32 * q] = 0.f;
    float np1 = __shfl_down_sync(0xffffffffu, bt[0], 1);
    float np2 = __shfl_down_sync(0xffffffffu, bt[1], 1);
    if (lane == 31) { np1 = kNeg; np2 = kNeg; }
    __syncwarp();
    float n[kC];
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int s = lane * kC + i;
      float v;
      if (t == Tb - 1) {
        v = (s < S && s >= S - 2) ? 0.f : kNeg;
      } else {
        const float x1 = i + 1 < kC ? bt[i + 1] : np1;
        const float x2 = i + 2 < kC ? bt[i + 2] : (i + 2 == kC ? np1 : np2);
        v = lse3_b2(bt[i], x1, skip_out[i] ? x2 : kNeg);
      }
      n[i] = live[i] ? v + s_row[warp][buf][lab[i]] : kNeg;
    }
    // occupancies: the L + 1 blank states (even s, label 0 = blank) all add into one slot - summed
    // in registers and reduced by shuffles instead of ~100 same-address shared-memory atomics per
    // time step, which made the gradient pass 10x the forward pass; label states keep the atomics
    // (a label repeats a few times per utterance at most)
    float occ_blank = 0.f;
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      bt[i] = n[i];
      const int s = lane * kC + i;
      if (s < S) {
        const float e = ex2f(apre[i] + bt[i] + nll2 - s_row[warp][buf][lab[i]]);
        if ((s & 1) == 0) occ_blank += e;
        else if (e > 0.f) atomicAdd(&s_occ[warp][lab[i]], e);
        if (t > 0) apre[i] = aws[static_cast<long long>(t - 1) * Sws + s];
      }
    }
    occ_blank = warp_sum(occ_blank);
    if (lane == 0 && occ_blank > 0.f) atomicAdd(&s_occ[warp][0], occ_blank);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kVS; ++q)
      if (lane + 32 * q < V)
        gb[static_cast<long long>(t) * V + lane + 32 * q] =
            gscale * (ex2f(s_row[warp][buf][lane + 32 * q]) - s_occ[warp][lane + 32 * q]);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Softmax backward over the vocabulary, one warp per row:  dl = p * (dp - sum_v p dp).
// (InterCTC self-conditioning on the training path: the gradient that comes back through
// conditioning_layer(ctc.softmax(tap)), encoder.py:393-401, on its way to ctc_lo and the tap.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ dl,
                   int M, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* pr = p + static_cast<long long>(row) * V;
  const float* dr = dp + static_cast<long long>(row) * V;
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s = fmaf(ld_act(pr + v), ld_act(dr + v), s);
  s = warp_sum(s);
  for (int v = lane; v < V; v += 32)
    dl[static_cast<long long>(row) * V + v] = ld_act(pr + v) * (ld_act(dr + v) - s);
}

// ------------------------------------------------------------------------------------------------
// Greedy decode: collapse repeats, drop blank.  One warp per utterance, ballot compaction.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
ctc_greedy_kernel(const int64_t* __restrict__ amax, const int32_t* __restrict__ lens,
                  int64_t* __restrict__ tokens, int32_t* __restrict__ ntok, int B, int T,
                  int blank) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int n = lens ? lens[b] : T;
  n = n < 0 ? 0 : (n > T ? T : n);
  const int64_t* a = amax + static_cast<long long>(b) * T;
  int64_t* out = tokens + static_cast<long long>(b) * T;
  int count = 0;
  for (int t0 = 0; t0 < n; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int64_t tok = blank;
    if (t < n) {
      tok = a[t];
      keep = tok != blank && (t == 0 || a[t - 1] != tok);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = tok;
    count += __popc(m);
  }
  for (int t = count + lane; t < T; t += 32) out[t] = -1;
  if (lane == 0) ntok[b] = count;
}

// ------------------------------------------------------------------------------------------------
// Prefix scoring step: one thread per (hypothesis, candidate token), serial in T.
// r layouts: r_prev [nhyp][T][2], r_new [nhyp][T][V][2].
// ------------------------------------------------------------------------------------------------
constexpr float kLogZero = -1e10f;

__global__ void __launch_bounds__(128)
ctc_prefix_kernel(const float* __restrict__ logp, const float* __restrict__ r_prev,
                  const int32_t* __restrict__ last, const int32_t* __restrict__ plen,
                  const float* __restrict__ psi_prev, float* __restrict__ r_new,
                  float* __restrict__ score, int T, int Tvalid, int V, int nhyp, int blank,
                  int eos) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nhyp * V) return;
  const int hy = idx / V, c = idx % V;
  const float* rp = r_prev + static_cast<long long>(hy) * T * 2;
  float2* rn = reinterpret_cast<float2*>(r_new) + static_cast<long long>(hy) * T * V + c;
  const int pl = plen[hy];
  const int lastl = last[hy];
  const int start = pl > 1 ? pl : 1;
  auto lpx = [&](int t, int v) -> float {
    if (t < Tvalid) return logp[static_cast<long long>(t) * V + v];
    return v == blank ? 0.f : kLogZero;
  };
  auto phi = [&](int t) -> float {
    const float rn_ = rp[2 * t], rb_ = rp[2 * t + 1];
    return c == lastl ? rb_ : lse2(rn_, rb_);
  };
  float psi;
  if (c == blank) {
    for (int t = 0; t < T; ++t) rn[static_cast<long long>(t) * V] = make_float2(kLogZero, kLogZero);
    score[idx] = kLogZero - psi_prev[hy];
    return;
  }
  float rnn = kLogZero, rnb = kLogZero;
  for (int t = 0; t < start - 1; ++t) rn[static_cast<long long>(t) * V] = make_float2(kLogZero, kLogZero);
  if (pl == 0) rnn = lpx(0, c);
  rn[static_cast<long long>(start - 1) * V] = make_float2(rnn, rnb);
  psi = rnn;
  for (int t = start; t < T; ++t) {
    const float ph = phi(t - 1);
    const float xc = lpx(t, c);
    const float nn = lse2(rnn, ph) + xc;
    const float nb = lse2(rnn, rnb) + lpx(t, blank);
    psi = lse2(psi, ph + xc);
    rnn = nn;
    rnb = nb;
    rn[static_cast<long long>(t) * V] = make_float2(rnn, rnb);
  }
  if (c == eos) {
    const int te = (Tvalid > 0 ? Tvalid : 1) - 1;
    psi = lse2(rp[2 * te], rp[2 * te + 1]);
  }
  score[idx] = psi - psi_prev[hy];
}


// ------------------------------------------------------------------------------------------------
// Backward of the CTC head  logits = hs . W^T + b  (ctc.py:143: ctc_lo) given d loss / d logits
// from ctc_loss_kernel:   d hs = g . W,   d W = g^T . hs,   d b = sum_m g   with g = dlogits scaled
// per utterance by the upstream gradient (d total / d nll_b).  fp32 FMA like the forward head.
// CTA = 64 rows x all D = 256 columns (thread = column): W's column and the V partial sums of d W
// live in registers, the scaled g rows are broadcast from shared memory; per-CTA partials of d W /
// d b go to a workspace that ctc_head_bwd_reduce_kernel sums (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
constexpr int kBwdRows = 64;

__global__ void __launch_bounds__(256)
ctc_head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ row_scale,
                    int rows_per_seg, const float* __restrict__ hs, long long ldh,
                    const float* __restrict__ w, float* __restrict__ dhs, long long ldd,
                    float* __restrict__ part_w, float* __restrict__ part_b, int M, int V) {
  __shared__ float s_g[kBwdRows][kVPad];
  pdl_launch_dependents();
  const int t = threadIdx.x;  // column of hs / W
  constexpr int D = 256;
  float wcol[kVPad], acc[kVPad];
#pragma unroll
  for (int v = 0; v < kVPad; ++v) {
    wcol[v] = v < V ? __ldg(w + static_cast<long long>(v) * D + t) : 0.f;
    acc[v] = 0.f;
  }
  pdl_wait();
  const int m0 = blockIdx.x * kBwdRows;
  for (int i = t; i < kBwdRows * kVPad; i += 256) {
    const int r = i / kVPad, v = i % kVPad;
    const int m = m0 + r;
    float g = 0.f;
    if (m < M && v < V) {
      g = ld_act(dlogits + static_cast<long long>(m) * V + v);
      if (row_scale != nullptr) g *= ld_act(row_scale + m / rows_per_seg);
    }
    s_g[r][v] = g;
  }
  __syncthreads();
  for (int r = 0; r < kBwdRows; ++r) {
    const int m = m0 + r;
    if (m >= M) break;
    const float h = ld_act(hs + static_cast<long long>(m) * ldh + t);
    float dh = 0.f;
#pragma unroll
    for (int v = 0; v < kVPad; ++v) {
      if (v < V) {
        const float g = s_g[r][v];
        acc[v] = fmaf(g, h, acc[v]);
        dh = fmaf(g, wcol[v], dh);
      }
    }
    dhs[static_cast<long long>(m) * ldd + t] = dh;
  }
  float* pw = part_w + static_cast<long long>(blockIdx.x) * kVPad * D;
#pragma unroll
  for (int v = 0; v < kVPad; ++v)
    if (v < V) pw[v * D + t] = acc[v];
  if (t < kVPad) {
    float sb = 0.f;
    for (int r = 0; r < kBwdRows; ++r) sb += s_g[r][t];
    part_b[blockIdx.x * kVPad + t] = sb;
  }
}

__global__ void __launch_bounds__(256)
ctc_head_bwd_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b,
                           int nblk, float* __restrict__ dw, float* __restrict__ db, int V) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = 256;
  const int v = blockIdx.x, t = threadIdx.x;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += ld_act(part_w + (static_cast<long long>(b) * kVPad + v) * D + t);
  dw[v * D + t] = s;
  if (t == 0) {
    float sb = 0.f;
    for (int b = 0; b < nblk; ++b) sb += ld_act(part_b + b * kVPad + v);
    db[v] = sb;
  }
}

}  // namespace ctc
}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_ctc_head(const float* hs, long long ldh, const float* w, const float* b,
                              float* logits, float* logp, float* prob, int64_t* amax, int M,
                              int D, int V, void* stream) {
  TAVSR_REQUIRE(M > 0 && D > 0 && D % 4 == 0 && D <= 512, "ctc_head: bad D=%d", D);
  TAVSR_REQUIRE(V > 0 && V <= ctc::kBigV, "ctc_head: V=%d > 256 not built", V);
  TAVSR_REQUIRE(hs && w && b && ldh % 4 == 0, "ctc_head: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (V > ctc::kVPad) {
    TAVSR_CUDA_OK(launch_kernel(ctc::ctc_head_big_kernel, dim3((M + ctc::kBigFrames - 1) / ctc::kBigFrames),
                                dim3(256), static_cast<size_t>(ctc::kBigFrames * D * 4), s, 0, hs, ldh, w, b,
                                logits, logp, prob, amax, M, D, V));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  const int smem = (D * ctc::kVPad + 8 * 4 * D) * 4;
  static PerDeviceMax configured;
  if (configured.raise(smem)) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(ctc::ctc_head_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int groups = (M + 3) / 4;
  int grid = (groups + 7) / 8;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  TAVSR_CUDA_OK(launch_kernel(ctc::ctc_head_kernel, dim3(grid), dim3(256), smem, s, 0, hs, ldh, w, b, logits,
                              logp, prob, amax, M, D, V));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_vocab_residual(const float* x, long long ldx, const float* p, const float* w,
                                    const float* b, float* out, long long ldo, const float* ln_g,
                                    const float* ln_b, float eps, void* xn, long long ldn, int M,
                                    int D, int V, int dtype, void* stream) {
  const int xn_bf16 = (dtype & TAVSR_DT_LNA_BF16) ? 1 : 0;
  TAVSR_REQUIRE(M > 0 && (D == 128 || D == 256 || D == 512), "vocab_residual: D=%d not built", D);
  TAVSR_REQUIRE(V > 0 && V <= ctc::kBigV, "vocab_residual: V=%d > 256 not built", V);
  TAVSR_REQUIRE(x && p && w && b && out, "vocab_residual: null pointer");
  TAVSR_REQUIRE(!xn || (ln_g && ln_b), "vocab_residual: LayerNorm output needs gamma and beta");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool big = V > ctc::kVPad;   // w is then W^T (V, D): see include/tavsr.h
  const int smem = big ? 0 : V * D * 4;
  int grid = (M + 7) / 8;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
#define TAVSR_VR_CASE(C)                                                                        \
  do {                                                                                          \
    static PerDeviceMax configured;                                                                  \
    if (configured.raise(smem)) {                                                                    \
      TAVSR_CUDA_OK(cudaFuncSetAttribute(ctc::vocab_residual_kernel<C>,                         \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
    }                                                                                           \
    if (big)                                                                                    \
      TAVSR_CUDA_OK(launch_kernel(ctc::vocab_residual_kernel<C, true>, dim3(grid), dim3(256), 0, s, 0, \
                                  x, ldx, p, w, b, out, ldo, ln_g, ln_b, eps, xn, ldn, M, V, xn_bf16)); \
    else                                                                                        \
      TAVSR_CUDA_OK(launch_kernel(ctc::vocab_residual_kernel<C>, dim3(grid), dim3(256), smem, s, 0, x, \
                                  ldx, p, w, b, out, ldo, ln_g, ln_b, eps, xn, ldn, M, V, xn_bf16)); \
  } while (0)
  if (D == 128) TAVSR_VR_CASE(4);
  else if (D == 256) TAVSR_VR_CASE(8);
  else TAVSR_VR_CASE(16);
#undef TAVSR_VR_CASE
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" size_t tavsr_ctc_workspace_bytes(int B, int T, int Lmax) {
  return static_cast<size_t>(B) * T * (2 * static_cast<size_t>(Lmax) + 1) * sizeof(float);
}

extern "C" int tavsr_ctc_loss(const float* logp, const int64_t* targets, long long ld_targets,
                              const int32_t* hlens, const int32_t* tlens, float* nll, float* grad,
                              float gscale, void* alpha_ws, int B, int T, int V, int Lmax,
                              int zero_infinity, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && V > 0 && V <= ctc::kBigV && Lmax >= 0, "ctc_loss: bad shape (V <= 256)");
  TAVSR_REQUIRE(logp && targets && hlens && tlens && nll, "ctc_loss: null pointer");
  TAVSR_REQUIRE(!grad || alpha_ws, "ctc_loss: gradient needs the alpha workspace");
  const int S = 2 * Lmax + 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (B + 3) / 4;
  float* aws = static_cast<float*>(alpha_ws);
  if (!grad) aws = nullptr;
#define TAVSR_CTC_CASE(C)                                                                       \
  do {                                                                                          \
    if (V > ctc::kVPad)                                                                         \
      TAVSR_CUDA_OK(launch_kernel(ctc::ctc_loss_kernel<C, 8>, dim3(grid), dim3(128), 0, s, 0, logp, \
                                  targets, ld_targets, hlens, tlens, nll, grad, gscale, aws, B, T, V, \
                                  Lmax, zero_infinity));                                        \
    else                                                                                        \
      TAVSR_CUDA_OK(launch_kernel(ctc::ctc_loss_kernel<C>, dim3(grid), dim3(128), 0, s, 0, logp,    \
                                  targets, ld_targets, hlens, tlens, nll, grad, gscale, aws, B, T, V, \
                                  Lmax, zero_infinity));                                        \
  } while (0)
  if (S <= 64) TAVSR_CTC_CASE(2);
  else if (S <= 128) TAVSR_CTC_CASE(4);
  else if (S <= 256) TAVSR_CTC_CASE(8);
  else if (S <= 512) TAVSR_CTC_CASE(16);
  else if (S <= 1024) TAVSR_CTC_CASE(32);
  else return set_error(TAVSR_ERR_UNSUPPORTED, "ctc_loss: target length %d > 511 not built", Lmax);
#undef TAVSR_CTC_CASE
  TAVSR_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_ctc_greedy(const int64_t* amax, const int32_t* lens, int64_t* tokens,
                                int32_t* ntok, int B, int T, int blank, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && amax && tokens && ntok, "ctc_greedy: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TAVSR_CUDA_OK(launch_kernel(ctc::ctc_greedy_kernel, dim3((B + 3) / 4), dim3(128), 0, s, 0, amax, lens,
                              tokens, ntok, B, T, blank));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_ctc_prefix_score(const float* logp, const float* r_prev, const int32_t* last,
                                      const int32_t* plen, const float* psi_prev, float* r_new,
                                      float* score, int T, int Tvalid, int V, int nhyp, int blank,
                                      int eos, void* stream) {
  TAVSR_REQUIRE(T > 0 && V > 0 && nhyp > 0 && Tvalid >= 0 && Tvalid <= T,
                "ctc_prefix_score: bad shape");
  TAVSR_REQUIRE(logp && r_prev && last && plen && psi_prev && r_new && score,
                "ctc_prefix_score: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = nhyp * V;
  TAVSR_CUDA_OK(launch_kernel(ctc::ctc_prefix_kernel, dim3((n + 127) / 128), dim3(128), 0, s, 0, logp,
                              r_prev, last, plen, psi_prev, r_new, score, T, Tvalid, V, nhyp, blank,
                              eos));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" size_t tavsr_ctc_head_bwd_workspace_bytes(int M) {
  const size_t nblk = (static_cast<size_t>(M) + ctc::kBwdRows - 1) / ctc::kBwdRows;
  return nblk * ctc::kVPad * (256 + 1) * sizeof(float);
}

extern "C" int tavsr_ctc_head_bwd(const float* dlogits, const float* row_scale, int rows_per_seg,
                                  const float* hs, long long ldh, const float* w, float* dhs,
                                  long long ldd, float* dw, float* db, void* workspace,
                                  long long workspace_bytes, int M, int D, int V, void* stream) {
  TAVSR_REQUIRE(M > 0 && D == 256 && V > 0 && V <= ctc::kVPad,
                "ctc_head_bwd: built for D == 256, V <= 64 (M=%d D=%d V=%d)", M, D, V);
  TAVSR_REQUIRE(dlogits && hs && w && dhs && dw && db && workspace, "ctc_head_bwd: null pointer");
  TAVSR_REQUIRE(!row_scale || rows_per_seg > 0, "ctc_head_bwd: row_scale needs rows_per_seg");
  TAVSR_REQUIRE(static_cast<size_t>(workspace_bytes) >= tavsr_ctc_head_bwd_workspace_bytes(M),
                "ctc_head_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nblk = (M + ctc::kBwdRows - 1) / ctc::kBwdRows;
  float* part_w = static_cast<float*>(workspace);
  float* part_b = part_w + static_cast<size_t>(nblk) * ctc::kVPad * 256;
  TAVSR_CUDA_OK(launch_kernel(ctc::ctc_head_bwd_kernel, dim3(nblk), dim3(256), 0, s, 0, dlogits,
                              row_scale, rows_per_seg, hs, ldh, w, dhs, ldd, part_w, part_b, M, V));
  TAVSR_CUDA_OK(launch_kernel(ctc::ctc_head_bwd_reduce_kernel, dim3(V), dim3(256), 0, s, 0,
                              static_cast<const float*>(part_w), static_cast<const float*>(part_b),
                              nblk, dw, db, V));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

extern "C" int tavsr_softmax_bwd(const float* p, const float* dp, float* dlogits, int M, int V,
                                 void* stream) {
  TAVSR_REQUIRE(M > 0 && V > 0 && p && dp && dlogits, "softmax_bwd: bad arguments (M=%d V=%d)", M, V);
  TAVSR_CUDA_OK(launch_kernel(ctc::softmax_bwd_kernel, dim3((M + 7) / 8), dim3(256), 0,
                              static_cast<cudaStream_t>(stream), 0, p, dp, dlogits, M, V));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
