// Backward of the fused relative-position multi-head self-attention (training path).
//
// Forward (attention_sm100.cu; espnet RelPositionMultiHeadedAttention.forward called from
// src/encoder/branchformer/encoder_layer.py:208), per utterance b and head h, raw scores
//   a_ij = (q_i + u) . k_j + (q_i + v) . p_{r(i,j)},   r(i,j) = T-1-i+j      (rel_shift)
//   P_ij = softmax_j(a_ij / sqrt(d_k)) over keys j < len[b] (0 elsewhere),   o_i = sum_j P_ij v_j
// Backward, given do (oracle/bwd_formulas.py::relpos_attn_core_bwd_tiled, verified against autograd):
//   dP_ij = do_i . v_j,  D_i = do_i . o_i,  g_ij = P_ij (dP_ij - D_i) / sqrt(d_k)     (= dL / d a_ij)
//   dv_j = sum_i P_ij do_i                 dk_j = sum_i g_ij (q_i + u)
//   dq_i = sum_j g_ij k_j  +  sum_j g_ij p_{r(i,j)}          (the "ac" and "bd" parts)
//   dp_r = sum_{(i,j): r(i,j) = r} g_ij (q_i + v)          (summed over utterances too)
//   du   = sum_i (ac part of dq_i),  dv_bias = sum_i (bd part of dq_i)     (column sums, host side)
// P is recomputed from the forward's per-row log-sum-exp (flash style): nothing of size T x T is
// kept between forward and backward.
//
// One CTA = 64 keys of one (utterance, head); it loops over the 64-query tiles and owns its dk / dv
// rows (plain stores); dq parts and dp are accumulated with fp32 atomics (4 key tiles per query row
// at T = 250; dp additionally across utterances).  All products run as 64 x 64 x 64 register-tiled
// FMA mini-GEMMs out of shared memory (fp32 CUDA cores: the first training version; the tensor-core
// tiling is fixed by the same index algebra).
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;

namespace attn_bwd {

constexpr int kT = 64;     // tile edge (queries and keys)
constexpr int kD = 64;     // head dim
constexpr int kLD = 68;    // row pitch of the 64-wide tiles (floats): 16-byte aligned, 4-bank skew
constexpr int kBand = 128; // band columns per tile pair (127 used)
constexpr int kLDR = 132;  // row pitch of the band-score tile
constexpr int kThreads = 256;

constexpr int kOffK = 0;
constexpr int kOffV = kOffK + kT * kLD;
constexpr int kOffQu = kOffV + kT * kLD;
constexpr int kOffQv = kOffQu + kT * kLD;
constexpr int kOffdO = kOffQv + kT * kLD;
constexpr int kOffS = kOffdO + kT * kLD;
constexpr int kOffdS = kOffS + kT * kLD;
constexpr int kOffPb = kOffdS + kT * kLD;
constexpr int kOffR = kOffPb + kBand * kLD;
constexpr int kOffLse = kOffR + kT * kLDR;
constexpr int kOffDi = kOffLse + kT;
constexpr int kSmemFloats = kOffDi + kT;
constexpr int kSmemBytes = kSmemFloats * 4;

struct Params {
  const float* qkv; long long ld_qkv;   // [B*T, 3*H*64]
  const float* pos; long long ld_pos;   // [2T-1, H*64]
  const float* u; const float* v;       // [H*64]
  const int32_t* lens;
  const float* o; long long ld_o;       // forward context [B*T, H*64]
  const float* dout; long long ld_do;   // d loss / d context
  const float* lse;                     // [B, H, T] base-2 log-sum-exp of the scaled scores
  float* dqkv; long long ld_dqkv;       // [B*T, 3*H*64]: k and v blocks written here
  float* dq_ac; float* dq_bd;           // [B*T, H*64] each, zero-initialised, atomically accumulated
  float* dpos;                          // [2T-1, H*64] zero-initialised, atomically accumulated
  // attention-probability dropout of the forward (optional): keep[b][h][i][j] bytes, row pitch
  // ld_drop, kept probabilities scaled by drop_scale.  With m_ij = keep_ij * drop_scale:
  //   o_i = sum_j m_ij P_ij v_j,  dP_ij = m_ij (do_i . v_j),  dv_j = sum_i m_ij P_ij do_i,
  // and D_i = do_i . o_i = sum_j P_ij dP_ij still holds, so g keeps its form.
  const uint8_t* drop_keep; long long ld_drop; float drop_scale;
  int T, H;
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
relpos_attn_bwd_kernel(const Params p) {
  extern __shared__ __align__(16) float sm[];
  float* sK = sm + kOffK;
  float* sV = sm + kOffV;
  float* sQu = sm + kOffQu;
  float* sQv = sm + kOffQv;
  float* sdO = sm + kOffdO;
  float* sS = sm + kOffS;
  float* sdS = sm + kOffdS;
  float* sPb = sm + kOffPb;
  float* sR = sm + kOffR;
  float* s_lse = sm + kOffLse;
  float* s_Di = sm + kOffDi;

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;   // 16 x 16 thread grid
  const int T = p.T, H = p.H;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * kT;
  const int HD = H * kD;
  const int hcol = h * kD;
  const long long row0 = static_cast<long long>(b) * T;
  int len = p.lens ? p.lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);

  float* dk_out = p.dqkv + HD + hcol;
  float* dv_out = p.dqkv + 2 * HD + hcol;
  if (j0 >= len) {
    // every key of this tile is masked: zero gradients for its k / v rows
    for (int idx = tid; idx < kT * (kD / 4); idx += kThreads) {
      const int jj = idx / (kD / 4), q4 = idx % (kD / 4);
      if (j0 + jj < T) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dk_out + (row0 + j0 + jj) * p.ld_dqkv + 4 * q4) = z;
        *reinterpret_cast<float4*>(dv_out + (row0 + j0 + jj) * p.ld_dqkv + 4 * q4) = z;
      }
    }
    return;
  }

  // ---- the key tile: K and V rows of this head (rows >= T are zero) ----
  for (int idx = tid; idx < kT * (kD / 4); idx += kThreads) {
    const int jj = idx / (kD / 4), q4 = idx % (kD / 4);
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if (j0 + jj < T) {
      const float* base = p.qkv + (row0 + j0 + jj) * p.ld_qkv + hcol + 4 * q4;
      kk = ld_act4(reinterpret_cast<const float4*>(base + HD));
      vv = ld_act4(reinterpret_cast<const float4*>(base + 2 * HD));
    }
    *reinterpret_cast<float4*>(sK + jj * kLD + 4 * q4) = kk;
    *reinterpret_cast<float4*>(sV + jj * kLD + 4 * q4) = vv;
  }

  const float scale = 0.125f;                       // 1 / sqrt(d_k)
  const float scale2 = 0.125f * 1.4426950408889634f; // the forward's exp2-domain scale
  // dk / dv accumulators: rows (keys) 4 ti .. +3, columns 4 tj .. +3
  float acc_dk[4][4], acc_dv[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc_dk[a][c] = 0.f; acc_dv[a][c] = 0.f; }

  const int n_qt = (T + kT - 1) / kT;
  for (int qt = 0; qt < n_qt; ++qt) {
    const int i0 = qt * kT;
    const int rbase = T - 1 - i0 + j0 - (kT - 1);   // band column c <-> relative-position row rbase + c
    __syncthreads();  // the previous pair's readers of sQu / sQv / sdO / sPb / sR are done
    // ---- query-tile operands: q + u, q + v, do; D_i = do_i . o_i; lse ----
    for (int idx = tid; idx < kT * (kD / 4); idx += kThreads) {
      const int ii = idx / (kD / 4), q4 = idx % (kD / 4);
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f), d4 = q;
      if (i0 + ii < T) {
        q = ld_act4(reinterpret_cast<const float4*>(p.qkv + (row0 + i0 + ii) * p.ld_qkv + hcol + 4 * q4));
        d4 = ld_act4(reinterpret_cast<const float4*>(p.dout + (row0 + i0 + ii) * p.ld_do + hcol + 4 * q4));
      }
      const float4 uu = __ldg(reinterpret_cast<const float4*>(p.u + hcol + 4 * q4));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(p.v + hcol + 4 * q4));
      *reinterpret_cast<float4*>(sQu + ii * kLD + 4 * q4) = make_float4(q.x + uu.x, q.y + uu.y, q.z + uu.z, q.w + uu.w);
      *reinterpret_cast<float4*>(sQv + ii * kLD + 4 * q4) = make_float4(q.x + vb.x, q.y + vb.y, q.z + vb.z, q.w + vb.w);
      *reinterpret_cast<float4*>(sdO + ii * kLD + 4 * q4) = d4;
    }
    {
      // D_i: four threads per query row, 16 columns each
      const int ii = tid >> 2, part = tid & 3;
      float dsum = 0.f;
      if (i0 + ii < T) {
        const float* op = p.o + (row0 + i0 + ii) * p.ld_o + hcol + 16 * part;
        const float* dp = p.dout + (row0 + i0 + ii) * p.ld_do + hcol + 16 * part;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 a = ld_act4(reinterpret_cast<const float4*>(op) + q4);
          const float4 c = ld_act4(reinterpret_cast<const float4*>(dp) + q4);
          dsum += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
      if (part == 0) {
        s_Di[ii] = dsum;
        s_lse[ii] = i0 + ii < T ? ld_act(p.lse + (static_cast<long long>(b) * H + h) * T + i0 + ii) : 0.f;
      }
    }
    // ---- the band of relative-position rows this tile pair touches ----
    for (int idx = tid; idx < kBand * (kD / 4); idx += kThreads) {
      const int c = idx / (kD / 4), q4 = idx % (kD / 4);
      const int r = rbase + c;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0 && r <= 2 * T - 2 && c < kBand - 1)
        x = ld_act4(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(r) * p.ld_pos + hcol + 4 * q4));
      *reinterpret_cast<float4*>(sPb + c * kLD + 4 * q4) = x;
    }
    __syncthreads();

    // ---- S = (q+u) K^T, dP = do V^T   (rows i = ti + 16 a, columns j = tj + 16 c) ----
    float cs[4][4], ce[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) { cs[a][c] = 0.f; ce[a][c] = 0.f; }
#pragma unroll 4
    for (int d4 = 0; d4 < kD / 4; ++d4) {
      float4 qa[4], da[4], kb[4], vb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        qa[a] = lds4(sQu + (ti + 16 * a) * kLD + 4 * d4);
        da[a] = lds4(sdO + (ti + 16 * a) * kLD + 4 * d4);
        kb[a] = lds4(sK + (tj + 16 * a) * kLD + 4 * d4);
        vb[a] = lds4(sV + (tj + 16 * a) * kLD + 4 * d4);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          cs[a][c] += qa[a].x * kb[c].x + qa[a].y * kb[c].y + qa[a].z * kb[c].z + qa[a].w * kb[c].w;
          ce[a][c] += da[a].x * vb[c].x + da[a].y * vb[c].y + da[a].z * vb[c].z + da[a].w * vb[c].w;
        }
    }
    // ---- R = (q+v) Pband^T   (rows ti + 16 a, band columns tj + 16 c, c < 8) ----
    {
      float cr[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) cr[a][c] = 0.f;
#pragma unroll 2
      for (int d4 = 0; d4 < kD / 4; ++d4) {
        float4 qa[4], pb[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) qa[a] = lds4(sQv + (ti + 16 * a) * kLD + 4 * d4);
#pragma unroll
        for (int c = 0; c < 8; ++c) pb[c] = lds4(sPb + (tj + 16 * c) * kLD + 4 * d4);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            cr[a][c] += qa[a].x * pb[c].x + qa[a].y * pb[c].y + qa[a].z * pb[c].z + qa[a].w * pb[c].w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) sR[(ti + 16 * a) * kLDR + tj + 16 * c] = cr[a][c];
    }
    __syncthreads();
    // ---- P and g = dL / d a  for this thread's 16 (i, j) ----
    float g[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ii = ti + 16 * a;
      const float lse = s_lse[ii], Di = s_Di[ii];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = tj + 16 * c;
        const float r = sR[ii * kLDR + jj - ii + (kT - 1)];
        const bool valid = (i0 + ii < T) && (j0 + jj < len);
        const float pr = valid ? exp2f((cs[a][c] + r) * scale2 - lse) : 0.f;
        float mk = 1.0f;
        if (p.drop_keep != nullptr && valid)
          mk = p.drop_keep[((static_cast<long long>(b) * H + h) * T + i0 + ii) * p.ld_drop + j0 + jj] != 0
                   ? p.drop_scale : 0.f;
        g[a][c] = pr * (ce[a][c] * mk - Di) * scale;
        sS[ii * kLD + jj] = pr * mk;
        sdS[ii * kLD + jj] = g[a][c];
      }
    }
    __syncthreads();  // every band score has been read
    for (int idx = tid; idx < kT * kLDR / 4; idx += kThreads)
      reinterpret_cast<float4*>(sR)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ii = ti + 16 * a, jj = tj + 16 * c;
        sR[ii * kLDR + jj - ii + (kT - 1)] = g[a][c];   // g in band layout: column = jj - ii + 63
      }
    __syncthreads();

    // ---- dv[j][d] += sum_i P[i][j] do[i][d],  dk[j][d] += sum_i g[i][j] (q+u)[i][d]
    //      (rows j = 4 ti .. +3, columns d = 4 tj .. +3) ----
#pragma unroll 4
    for (int i = 0; i < kT; ++i) {
      const float4 pr = lds4(sS + i * kLD + 4 * ti);
      const float4 gg = lds4(sdS + i * kLD + 4 * ti);
      const float4 od = lds4(sdO + i * kLD + 4 * tj);
      const float4 qu = lds4(sQu + i * kLD + 4 * tj);
      const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
      const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
      const float ov[4] = {od.x, od.y, od.z, od.w};
      const float qv[4] = {qu.x, qu.y, qu.z, qu.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc_dv[a][c] = fmaf(pv[a], ov[c], acc_dv[a][c]);
          acc_dk[a][c] = fmaf(gv[a], qv[c], acc_dk[a][c]);
        }
    }
    // ---- dq_ac[i][d] = sum_j g[i][j] K[j][d],  dq_bd[i][d] = sum_c gband[i][c] Pband[c][d]
    //      (rows i = 4 ti .. +3, columns d = 4 tj .. +3) ----
    {
      float qa[4][4], qb[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) { qa[a][c] = 0.f; qb[a][c] = 0.f; }
#pragma unroll 2
      for (int j4 = 0; j4 < kT / 4; ++j4) {
        float4 gs[4], kr[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gs[a] = lds4(sdS + (4 * ti + a) * kLD + 4 * j4);
          kr[a] = lds4(sK + (4 * j4 + a) * kLD + 4 * tj);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float ga[4] = {gs[a].x, gs[a].y, gs[a].z, gs[a].w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            qa[a][0] = fmaf(ga[jj], kr[jj].x, qa[a][0]);
            qa[a][1] = fmaf(ga[jj], kr[jj].y, qa[a][1]);
            qa[a][2] = fmaf(ga[jj], kr[jj].z, qa[a][2]);
            qa[a][3] = fmaf(ga[jj], kr[jj].w, qa[a][3]);
          }
        }
      }
#pragma unroll 2
      for (int c4 = 0; c4 < kBand / 4; ++c4) {
        float4 gs[4], pr[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gs[a] = lds4(sR + (4 * ti + a) * kLDR + 4 * c4);
          pr[a] = lds4(sPb + (4 * c4 + a) * kLD + 4 * tj);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float ga[4] = {gs[a].x, gs[a].y, gs[a].z, gs[a].w};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            qb[a][0] = fmaf(ga[cc], pr[cc].x, qb[a][0]);
            qb[a][1] = fmaf(ga[cc], pr[cc].y, qb[a][1]);
            qb[a][2] = fmaf(ga[cc], pr[cc].z, qb[a][2]);
            qb[a][3] = fmaf(ga[cc], pr[cc].w, qb[a][3]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + 4 * ti + a;
        if (i < T) {
          float* pa = p.dq_ac + (row0 + i) * HD + hcol + 4 * tj;
          float* pb = p.dq_bd + (row0 + i) * HD + hcol + 4 * tj;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            red_add(pa + c, qa[a][c]);
            red_add(pb + c, qb[a][c]);
          }
        }
      }
    }
    // ---- dp[rbase + c][d] += sum_i gband[i][c] (q+v)[i][d]   (band rows 8 ti .. +7, columns 4 tj .. +3)
    {
      float pp[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) pp[a][c] = 0.f;
#pragma unroll 4
      for (int i = 0; i < kT; ++i) {
        const float4 g0 = lds4(sR + i * kLDR + 8 * ti);
        const float4 g1 = lds4(sR + i * kLDR + 8 * ti + 4);
        const float4 qv = lds4(sQv + i * kLD + 4 * tj);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          pp[a][0] = fmaf(gv[a], qv.x, pp[a][0]);
          pp[a][1] = fmaf(gv[a], qv.y, pp[a][1]);
          pp[a][2] = fmaf(gv[a], qv.z, pp[a][2]);
          pp[a][3] = fmaf(gv[a], qv.w, pp[a][3]);
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int r = rbase + 8 * ti + a;
        if (r >= 0 && r <= 2 * T - 2 && 8 * ti + a < kBand - 1) {
          float* dst = p.dpos + static_cast<long long>(r) * HD + hcol + 4 * tj;
#pragma unroll
          for (int c = 0; c < 4; ++c) red_add(dst + c, pp[a][c]);
        }
      }
    }
  }
  // ---- this CTA's dk / dv rows ----
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + 4 * ti + a;
    if (j < T) {
      *reinterpret_cast<float4*>(dk_out + (row0 + j) * p.ld_dqkv + 4 * tj) =
          make_float4(acc_dk[a][0], acc_dk[a][1], acc_dk[a][2], acc_dk[a][3]);
      *reinterpret_cast<float4*>(dv_out + (row0 + j) * p.ld_dqkv + 4 * tj) =
          make_float4(acc_dv[a][0], acc_dv[a][1], acc_dv[a][2], acc_dv[a][3]);
    }
  }
}

}  // namespace attn_bwd
}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_relpos_attn_bwd(const float* qkv, long long ld_qkv, const float* pos,
                                     long long ld_pos, const float* u, const float* v,
                                     const int32_t* lens, const float* ctx, long long ld_ctx,
                                     const float* dctx, long long ld_dctx, const float* lse,
                                     float* dqkv, long long ld_dqkv, float* dq_ac, float* dq_bd,
                                     float* dpos, const uint8_t* drop_keep, long long ld_drop,
                                     float drop_scale, int B, int T, int H, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && H > 0, "attn_bwd: bad shape B=%d T=%d H=%d", B, T, H);
  TAVSR_REQUIRE(qkv && pos && u && v && ctx && dctx && lse && dqkv && dq_ac && dq_bd && dpos,
                "attn_bwd: null pointer");
  TAVSR_REQUIRE(ld_qkv % 4 == 0 && ld_pos % 4 == 0 && ld_ctx % 4 == 0 && ld_dctx % 4 == 0 &&
                    ld_dqkv % 4 == 0,
                "attn_bwd: pitches must be multiples of 4 floats");
  attn_bwd::Params p;
  p.qkv = qkv; p.ld_qkv = ld_qkv; p.pos = pos; p.ld_pos = ld_pos; p.u = u; p.v = v; p.lens = lens;
  p.o = ctx; p.ld_o = ld_ctx; p.dout = dctx; p.ld_do = ld_dctx; p.lse = lse;
  p.dqkv = dqkv; p.ld_dqkv = ld_dqkv; p.dq_ac = dq_ac; p.dq_bd = dq_bd; p.dpos = dpos;
  TAVSR_REQUIRE(drop_keep == nullptr || (ld_drop >= T && drop_scale >= 1.0f),
                "attn_bwd: keep mask pitch %lld < T or drop_scale < 1", ld_drop);
  p.drop_keep = drop_keep; p.ld_drop = ld_drop; p.drop_scale = drop_scale;
  p.T = T; p.H = H;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    TAVSR_CUDA_OK(cudaFuncSetAttribute(attn_bwd::relpos_attn_bwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       attn_bwd::kSmemBytes));
  dim3 grid((T + attn_bwd::kT - 1) / attn_bwd::kT, H, B);
  TAVSR_CUDA_OK(launch_kernel(attn_bwd::relpos_attn_bwd_kernel, grid, dim3(attn_bwd::kThreads),
                              attn_bwd::kSmemBytes, static_cast<cudaStream_t>(stream), 0, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
