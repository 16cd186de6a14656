// Fused relative-position multi-head self-attention, flash style (scores never reach HBM).
//
// Reference semantics (espnet RelPositionMultiHeadedAttention.forward, called from
// src/encoder/branchformer/encoder_layer.py:208 and tailored/encoder_layer.py:192,239):
//   ac[i,j] = (q_i + u) . k_j
//   bd[i,j] = (q_i + v) . p_{T-1-i+j}            (rel_shift of (q+v) P^T, P = linear_pos(pos_emb))
//   attn    = softmax_j((ac + bd) / sqrt(d_k)) with keys j >= len[b] masked to probability 0
//   ctx_i   = sum_j attn[i,j] v_j
//
// Round-1 implementation: one CTA = 64 query rows of one (utterance, head); 4 warps x 16 rows;
// K / V / P-band tiles staged in shared memory with cp.async; tensor-core products with
// mma.sync.m16n8k8 TF32 (fp32 accumulate); online softmax in registers.  The rel-shift is done
// on chip: each warp computes the 16 x 80 band (q+v) . P_band^T it needs, parks it in shared
// memory and re-reads it along the diagonals.  (A tcgen05/TMEM version is the planned successor;
// see DESIGN.md.)
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;
int relpos_attn_tc_launch(const float* qkv, long long ld_qkv, const float* pos, long long ld_pos,
                          const float* u, const float* v, const int32_t* lens, float* ctx,
                          long long ld_ctx, int B, int T, int H, int round_out, const float* dva,
                          const float* dvb, float* dots_out, cudaStream_t s);

namespace attn {

constexpr int kQT = 64;       // query rows per CTA
constexpr int kKT = 64;       // keys per tile
constexpr int kD = 64;        // head dim
constexpr int kLd = 72;       // smem row pitch (floats) of K / V / P tiles: 8 banks per row shift ->
                              // the 64-bit fragment loads of a half-warp hit 32 distinct banks
constexpr int kLdV = 68;      // V tile pitch: its fragments are row-strided 32-bit loads (4 banks / row)
constexpr int kPRows = 128;   // P band rows (127 used)
constexpr int kRLd = 84;      // per-warp R band pitch
constexpr int kSmemFloats = kKT * kLd + kKT * kLdV + kPRows * kLd + 4 * 16 * kRLd;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(128, 2)
relpos_attn_kernel(const float* __restrict__ qkv, long long ld_qkv, const float* __restrict__ pos,
                   long long ld_pos, const float* __restrict__ bias_u,
                   const float* __restrict__ bias_v, const int32_t* __restrict__ lens,
                   float* __restrict__ ctx, long long ld_ctx, int T, int H, int round_out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float smem[];
  float* Ks = smem;
  float* Vs = Ks + kKT * kLd;
  float* Ps = Vs + kKT * kLdV;
  float* Rs = Ps + kPRows * kLd;

  const int i0 = blockIdx.x * kQT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  int len = lens ? lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  const long long row_base = static_cast<long long>(b) * T;
  const int HD = H * kD;
  const float* qbase = qkv + h * kD;
  const float* kbase = qkv + HD + h * kD;
  const float* vbase = qkv + 2 * HD + h * kD;
  const float* pbase = pos + h * kD;
  float* Rw = Rs + warp * 16 * kRLd;

  // ---- query fragments (q + u) and (q + v), rounded to tf32 once ----
  uint32_t qu[8][4], qv[8][4];
  {
    const int r0 = i0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // contraction index permuted inside every 8-block: MMA k = t <-> d = 2t, k = t+4 <-> d = 2t+1,
      // so the matching B values (k, k+4) sit next to each other in the K-major smem tiles
      const int c0 = ks * 8 + 2 * t, c1 = c0 + 1;
      const float u0 = __ldg(bias_u + h * kD + c0), u1 = __ldg(bias_u + h * kD + c1);
      const float v0 = __ldg(bias_v + h * kD + c0), v1 = __ldg(bias_v + h * kD + c1);
      const float q00 = r0 < T ? ld_act(qbase + (row_base + r0) * ld_qkv + c0) : 0.f;
      const float q10 = r1 < T ? ld_act(qbase + (row_base + r1) * ld_qkv + c0) : 0.f;
      const float q01 = r0 < T ? ld_act(qbase + (row_base + r0) * ld_qkv + c1) : 0.f;
      const float q11 = r1 < T ? ld_act(qbase + (row_base + r1) * ld_qkv + c1) : 0.f;
      qu[ks][0] = tf32_bits(q00 + u0); qu[ks][1] = tf32_bits(q10 + u0);
      qu[ks][2] = tf32_bits(q01 + u1); qu[ks][3] = tf32_bits(q11 + u1);
      qv[ks][0] = tf32_bits(q00 + v0); qv[ks][1] = tf32_bits(q10 + v0);
      qv[ks][2] = tf32_bits(q01 + v1); qv[ks][3] = tf32_bits(q11 + v1);
    }
  }

  float O[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY};
  float lrow[2] = {0.f, 0.f};
  const float kScaleLog2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

  const int nkt = (len + kKT - 1) / kKT;
  for (int kt = 0; kt < nkt; ++kt) {
    const int j0 = kt * kKT;
    const int rb0 = T - 1 - (i0 + kQT - 1) + j0;  // P row held in Ps[0]
    __syncthreads();                              // previous tile fully consumed
    for (int idx = threadIdx.x; idx < kKT * 16; idx += 128) {
      const int r = idx >> 4, ch = idx & 15;
      const int j = j0 + r;
      const bool ok = j < T;
      const long long grow = row_base + (ok ? j : 0);
      cp_async16(Ks + r * kLd + ch * 4, kbase + grow * ld_qkv + ch * 4, ok);
      cp_async16(Vs + r * kLdV + ch * 4, vbase + grow * ld_qkv + ch * 4, ok);
    }
    for (int idx = threadIdx.x; idx < (kPRows - 1) * 16; idx += 128) {
      const int x = idx >> 4, ch = idx & 15;
      const int r = rb0 + x;
      const bool ok = r >= 0 && r <= 2 * T - 2;
      cp_async16(Ps + x * kLd + ch * 4, pbase + static_cast<long long>(ok ? r : 0) * ld_pos + ch * 4,
                 ok);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- bd band: R[a][c] = (q_a + v) . P[rb0 + (48 - 16*warp) + c],  c = 15 - a + j_local ----
    {
      const float* Pw = Ps + (48 - 16 * warp) * kLd;
#pragma unroll
      for (int n = 0; n < 10; ++n) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* prow = Pw + (8 * n + g) * kLd + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const float2 bb = *reinterpret_cast<const float2*>(prow + ks * 8);
          mma_tf32(acc, qv[ks], __float_as_uint(bb.x), __float_as_uint(bb.y));
        }
        float* rw = Rw + g * kRLd + 8 * n + 2 * t;
        rw[0] = acc[0]; rw[1] = acc[1];
        rw[8 * kRLd] = acc[2]; rw[8 * kRLd + 1] = acc[3];
      }
    }
    __syncwarp();

    // ---- ac + bd, scale, mask ----
    float S[8][4];
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* krow = Ks + (8 * n + g) * kLd + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const float2 bb = *reinterpret_cast<const float2*>(krow + ks * 8);
        mma_tf32(acc, qu[ks], __float_as_uint(bb.x), __float_as_uint(bb.y));
      }
      const int jl = 8 * n + 2 * t;
      const float* r_lo = Rw + g * kRLd + (15 - g) + jl;
      const float* r_hi = Rw + (g + 8) * kRLd + (7 - g) + jl;
      acc[0] = (acc[0] + r_lo[0]) * kScaleLog2;
      acc[1] = (acc[1] + r_lo[1]) * kScaleLog2;
      acc[2] = (acc[2] + r_hi[0]) * kScaleLog2;
      acc[3] = (acc[3] + r_hi[1]) * kScaleLog2;
      if (j0 + jl >= len) { acc[0] = -INFINITY; acc[2] = -INFINITY; }
      if (j0 + jl + 1 >= len) { acc[1] = -INFINITY; acc[3] = -INFINITY; }
      S[n][0] = acc[0]; S[n][1] = acc[1]; S[n][2] = acc[2]; S[n][3] = acc[3];
      tmax[0] = fmaxf(tmax[0], fmaxf(acc[0], acc[1]));
      tmax[1] = fmaxf(tmax[1], fmaxf(acc[2], acc[3]));
    }
    __syncwarp();  // all lanes done reading Rw before the next tile overwrites it

    // ---- online softmax (base-2 domain) ----
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float m = tmax[r];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const float mnew = fmaxf(mrow[r], m);  // finite: every tile has >= 1 unmasked key
      corr[r] = exp2f(mrow[r] - mnew);
      mrow[r] = mnew;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      S[n][0] = exp2f(S[n][0] - mrow[0]);
      S[n][1] = exp2f(S[n][1] - mrow[0]);
      S[n][2] = exp2f(S[n][2] - mrow[1]);
      S[n][3] = exp2f(S[n][3] - mrow[1]);
      psum[0] += S[n][0] + S[n][1];
      psum[1] += S[n][2] + S[n][3];
    }
    lrow[0] = lrow[0] * corr[0] + psum[0];
    lrow[1] = lrow[1] * corr[1] + psum[1];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      O[n][0] *= corr[0]; O[n][1] *= corr[0];
      O[n][2] *= corr[1]; O[n][3] *= corr[1];
    }

    // ---- O += P V.  The k index of the A fragment is mapped to keys (2t, 2t+1) of each 8-key
    //      block so the score fragment can be reused as the A operand without shuffles. ----
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t a[4];
      a[0] = tf32_bits(S[kk][0]);
      a[1] = tf32_bits(S[kk][2]);
      a[2] = tf32_bits(S[kk][1]);
      a[3] = tf32_bits(S[kk][3]);
      const float* v0 = Vs + (8 * kk + 2 * t) * kLdV + g;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        const uint32_t b0 = __float_as_uint(v0[8 * nd]);
        const uint32_t b1 = __float_as_uint(v0[kLdV + 8 * nd]);
        mma_tf32(O[nd], a, b0, b1);
      }
    }
  }

  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = lrow[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    lrow[r] = l > 0.f ? 1.0f / l : 0.f;  // len == 0 -> all probabilities are zero -> ctx = 0
  }
  const int r0 = i0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int c = h * kD + 8 * nd + 2 * t;
    float2 lo = make_float2(O[nd][0] * lrow[0], O[nd][1] * lrow[0]);
    float2 hi = make_float2(O[nd][2] * lrow[1], O[nd][3] * lrow[1]);
    if (round_out) {
      lo.x = round_tf32(lo.x); lo.y = round_tf32(lo.y);
      hi.x = round_tf32(hi.x); hi.y = round_tf32(hi.y);
    }
    if (r0 < T) *reinterpret_cast<float2*>(ctx + (row_base + r0) * ld_ctx + c) = lo;
    if (r1 < T) *reinterpret_cast<float2*>(ctx + (row_base + r1) * ld_ctx + c) = hi;
  }
}

}  // namespace attn
}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_relpos_attn_fwd(const float* qkv, long long ld_qkv, const float* pos,
                                     long long ld_pos, const float* u, const float* v,
                                     const int32_t* lens, float* ctx, long long ld_ctx, int B,
                                     int T, int H, int round_out, void* stream) {
  return tavsr_relpos_attn_fwd_dots(qkv, ld_qkv, pos, ld_pos, u, v, lens, ctx, ld_ctx, B, T, H,
                                    round_out, nullptr, nullptr, nullptr, stream);
}

extern "C" int tavsr_relpos_attn_fwd_dots(const float* qkv, long long ld_qkv, const float* pos,
                                          long long ld_pos, const float* u, const float* v,
                                          const int32_t* lens, float* ctx, long long ld_ctx, int B,
                                          int T, int H, int round_out, const float* dva,
                                          const float* dvb, float* dots_out, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && H > 0, "attn: bad shape B=%d T=%d H=%d", B, T, H);
  TAVSR_REQUIRE(!dots_out || (dva && dvb && (reinterpret_cast<uintptr_t>(dva) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(dvb) & 15) == 0),
                "attn: dots_out needs 16-byte aligned dva / dvb");
  TAVSR_REQUIRE(qkv && pos && u && v && ctx, "attn: null pointer");
  TAVSR_REQUIRE(ld_qkv % 4 == 0 && ld_pos % 4 == 0 && ld_ctx % 2 == 0,
                "attn: pitches must be multiples of 4 (qkv, pos) / 2 (ctx)");
  TAVSR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(ctx) & 7) == 0,
                "attn: misaligned pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tcgen05 / TMEM kernel (attention_sm100.cu) whenever TMA can address the operands;
  // g_debug[8] = 1 forces the mma.sync kernel below.
  if (g_debug[8] != 1 && ld_ctx % 4 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0) {
    const int rc = relpos_attn_tc_launch(qkv, ld_qkv, pos, ld_pos, u, v, lens, ctx, ld_ctx, B, T, H,
                                         round_out, dva, dvb, dots_out, s);
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc;
  }
  TAVSR_REQUIRE(dots_out == nullptr, "attn: the fused row dots need the tcgen05 kernel "
                                     "(16-byte aligned ctx with a pitch multiple of 4)");
  const int smem = attn::kSmemFloats * 4;
  static bool configured = false;
  if (!configured) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(attn::relpos_attn_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((T + attn::kQT - 1) / attn::kQT, H, B);
  TAVSR_CUDA_OK(launch_kernel(attn::relpos_attn_kernel, grid, dim3(128), smem, s, 0, qkv, ld_qkv, pos,
                              ld_pos, u, v, lens, ctx, ld_ctx, T, H, round_out));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
