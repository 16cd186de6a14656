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
// at T = 250; dp additionally across utterances).  All products run as warp-level TF32 tensor-core
// MMAs (mma.sync.m16n8k8, fp32 accumulate) out of shared memory: the eight products of a tile pair
// need their operands in both orientations (P^T do, g^T (q+u), g K, ...), which the register
// fragments of mma.sync read from ONE row-major copy - strides (1, 68) / (68, 1) are both bank-
// conflict free - where tcgen05 would need K-major transposed copies of five tiles.  Operands are
// rounded to TF32 (nearest) when they enter shared memory, like the forward's TMA loads.  The fp32
// FMA version of this kernel took 506 us per launch at the C2 shape.
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
  float* dqkv; long long ld_dqkv;       // [B*T, 3*H*64]: k and v blocks written; the q block is
                                        // zero-initialised by the caller and atomically accumulated
  float* du; float* dvb;                // [H*64] each, zero-initialised: d pos_bias_u / d pos_bias_v
  float* dpos;                          // [B][2T-1][H*64] zero-initialised: PER-UTTERANCE slabs (the
                                        // caller sums them): no cross-utterance atomic contention
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

// Accumulate one mma C fragment (rows g / g + 8, column pairs 2 t4) with 16-byte reductions: lanes
// t4 and t4 ^ 1 swap halves so that the even lane owns four consecutive columns of row g and the
// odd lane four of row g + 8 (a quarter of the scalar red count: the kernel is bound by its
// reductions into dq / dpos once the products run on tensor cores).
// `base` points at (row 0, column 8 nt) of the fragment; `row_ok(r)` guards the fragment row r.
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}
template <typename RowOk>
__device__ __forceinline__ void red_fragment(float* base, long long ld, const float (&c)[4], int g, int t4,
                                             RowOk row_ok) {
  const bool odd = (t4 & 1) != 0;
  const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];
  const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  const float4 v = odd ? make_float4(r0, r1, c[2], c[3]) : make_float4(c[0], c[1], r0, r1);
  const int row = odd ? g + 8 : g;
  if (row_ok(row)) red_add_v4(base + row * ld + 2 * (t4 & ~1), v);
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
               "{%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// One warp: acc[16 x 8 NT] += A[16 x 8 KS] . B[8 KS x 8 NT] out of shared memory.
// A(m, k) = a[m * a_sm + k * a_sk], B(k, n) = b[k * b_sk + n * b_sn]; g = lane >> 2, t = lane & 3.
// Fragment layout of mma.m16n8k8: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (t, g)
// b1 (t+4, g); c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
template <int NT, int KS>
__device__ __forceinline__ void warp_mma(float (&acc)[NT][4], const float* a, int a_sm, int a_sk,
                                         const float* b, int b_sk, int b_sn, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = 8 * ks + t;
    const uint32_t a0 = __float_as_uint(a[g * a_sm + k0 * a_sk]);
    const uint32_t a1 = __float_as_uint(a[(g + 8) * a_sm + k0 * a_sk]);
    const uint32_t a2 = __float_as_uint(a[g * a_sm + (k0 + 4) * a_sk]);
    const uint32_t a3 = __float_as_uint(a[(g + 8) * a_sm + (k0 + 4) * a_sk]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint32_t b0 = __float_as_uint(b[k0 * b_sk + (8 * nt + g) * b_sn]);
      const uint32_t b1 = __float_as_uint(b[(k0 + 4) * b_sk + (8 * nt + g) * b_sn]);
      mma_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
    }
  }
}

__device__ __forceinline__ float4 tf32x4(float4 v) {
  return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
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
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;        // mma fragment coordinates
  const int wr = 16 * (warp & 3);                // this warp's 16 rows of a 64-row output tile
  const int wc = 32 * (warp >> 2);               // ... and its 32 columns
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
    *reinterpret_cast<float4*>(sK + jj * kLD + 4 * q4) = tf32x4(kk);
    *reinterpret_cast<float4*>(sV + jj * kLD + 4 * q4) = tf32x4(vv);
  }

  const float scale = 0.125f;                       // 1 / sqrt(d_k)
  const float scale2 = 0.125f * 1.4426950408889634f; // the forward's exp2-domain scale
  // dk / dv accumulators in mma layout: keys wr + g (+8), columns wc + 8 nt + 2 t4 (+1)
  float acc_dk[4][4], acc_dv[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc_dk[a][c] = 0.f; acc_dv[a][c] = 0.f; }

  float su[4][2], sv[4][2];   // per-thread column sums of the two dq parts (d pos_bias_u / _v)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { su[nt][0] = su[nt][1] = 0.f; sv[nt][0] = sv[nt][1] = 0.f; }
  float pp[2][4][4];          // dpos accumulator carried across two consecutive query tiles
  float* dpos_b = p.dpos + static_cast<long long>(b) * (2 * T - 1) * HD;
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
      *reinterpret_cast<float4*>(sQu + ii * kLD + 4 * q4) = tf32x4(make_float4(q.x + uu.x, q.y + uu.y, q.z + uu.z, q.w + uu.w));
      *reinterpret_cast<float4*>(sQv + ii * kLD + 4 * q4) = tf32x4(make_float4(q.x + vb.x, q.y + vb.y, q.z + vb.z, q.w + vb.w));
      *reinterpret_cast<float4*>(sdO + ii * kLD + 4 * q4) = tf32x4(d4);
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
      *reinterpret_cast<float4*>(sPb + c * kLD + 4 * q4) = tf32x4(x);
    }
    __syncthreads();

    // ---- R = (q+v) Pband^T: 64 x 128, this warp rows wr .., band columns 64 (warp >> 2) .. ----
    {
      float cr[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) cr[nt][e] = 0.f;
      const int wcb = 64 * (warp >> 2);
      warp_mma<8, 8>(cr, sQv + wr * kLD, kLD, 1, sPb + wcb * kLD, 1, kLD, g, t4);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = wcb + 8 * nt + 2 * t4;
        *reinterpret_cast<float2*>(sR + (wr + g) * kLDR + col) = make_float2(cr[nt][0], cr[nt][1]);
        *reinterpret_cast<float2*>(sR + (wr + g + 8) * kLDR + col) = make_float2(cr[nt][2], cr[nt][3]);
      }
    }
    // ---- S = (q+u) K^T, dP = do V^T: this warp rows wr .., key columns wc .. ----
    float cs[4][4], ce[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { cs[nt][e] = 0.f; ce[nt][e] = 0.f; }
    warp_mma<4, 8>(cs, sQu + wr * kLD, kLD, 1, sK + wc * kLD, 1, kLD, g, t4);
    warp_mma<4, 8>(ce, sdO + wr * kLD, kLD, 1, sV + wc * kLD, 1, kLD, g, t4);
    __syncthreads();   // every band score is in sR
    // ---- P and g = dL / d a  for this thread's 16 (i, j) ----
    float gk[4][4];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int ii = wr + g + 8 * h2;
      const float lse = s_lse[ii], Di = s_Di[ii];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int e = 2 * h2 + e2;
          const int jj = wc + 8 * nt + 2 * t4 + e2;
          const float r = sR[ii * kLDR + jj - ii + (kT - 1)];
          const bool valid = (i0 + ii < T) && (j0 + jj < len);
          const float pr = valid ? exp2f((cs[nt][e] + r) * scale2 - lse) : 0.f;
          float mk = 1.0f;
          if (p.drop_keep != nullptr && valid)
            mk = p.drop_keep[((static_cast<long long>(b) * H + h) * T + i0 + ii) * p.ld_drop + j0 + jj] != 0
                     ? p.drop_scale : 0.f;
          gk[nt][e] = round_tf32(pr * (ce[nt][e] * mk - Di) * scale);
          sS[ii * kLD + jj] = round_tf32(pr * mk);
          sdS[ii * kLD + jj] = gk[nt][e];
        }
    }
    __syncthreads();  // every band score has been read
    for (int idx = tid; idx < kT * kLDR / 4; idx += kThreads)
      reinterpret_cast<float4*>(sR)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ii = wr + g + 8 * (e >> 1), jj = wc + 8 * nt + 2 * t4 + (e & 1);
        sR[ii * kLDR + jj - ii + (kT - 1)] = gk[nt][e];   // g in band layout: column = jj - ii + 63
      }
    __syncthreads();

    // ---- dv[j][d] += sum_i P[i][j] do[i][d],  dk[j][d] += sum_i g[i][j] (q+u)[i][d]
    //      (this warp: keys wr .., columns wc ..; A read transposed out of the row-major tiles) ----
    warp_mma<4, 8>(acc_dv, sS + wr, 1, kLD, sdO + wc, kLD, 1, g, t4);
    warp_mma<4, 8>(acc_dk, sdS + wr, 1, kLD, sQu + wc, kLD, 1, g, t4);
    // ---- dq[i][d] = sum_j g[i][j] K[j][d]  +  sum_c gband[i][c] Pband[c][d]; the column sums of
    //      the two parts are d pos_bias_u / d pos_bias_v (kept per thread, reduced at the end) ----
    {
      float qa[4][4], qb[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { qa[nt][e] = 0.f; qb[nt][e] = 0.f; }
      warp_mma<4, 8>(qa, sdS + wr * kLD, kLD, 1, sK + wc, kLD, 1, g, t4);
      warp_mma<4, 16>(qb, sR + wr * kLDR, kLDR, 1, sPb + wc, kLD, 1, g, t4);
      float* pq = p.dqkv + (row0 + i0 + wr) * p.ld_dqkv + hcol + wc;
      auto q_ok = [&](int r) { return i0 + wr + r < T; };
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        // rows >= T have g == 0, so the unguarded column sums are exact
        su[nt][0] += qa[nt][0] + qa[nt][2]; su[nt][1] += qa[nt][1] + qa[nt][3];
        sv[nt][0] += qb[nt][0] + qb[nt][2]; sv[nt][1] += qb[nt][1] + qb[nt][3];
        const float qs[4] = {qa[nt][0] + qb[nt][0], qa[nt][1] + qb[nt][1], qa[nt][2] + qb[nt][2],
                             qa[nt][3] + qb[nt][3]};
        red_fragment(pq + 8 * nt, p.ld_dqkv, qs, g, t4, q_ok);
      }
    }
    // ---- dp[rbase + c][d] += sum_i gband[i][c] (q+v)[i][d]: 128 band rows x 64.  Consecutive query
    //      tiles shift the band by 64 rows, so the lower half of this pair's band is the upper half
    //      of the next pair's: a warp takes band chunk ((warp & 3) + 2 qt) mod 4 (32 rows), i.e. it
    //      alternates lower half -> upper half over the SAME absolute rows, keeps the accumulator in
    //      registers in between and emits once per two pairs (half the reductions) ----
    {
      const int chunk = ((warp & 3) + 2 * qt) & 3;
      const bool lower = chunk < 2;
      if (lower || qt == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) pp[mt][nt][e] = 0.f;
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        warp_mma<4, 8>(pp[mt], sR + 32 * chunk + 16 * mt, 1, kLDR, sQv + wc, kLD, 1, g, t4);
      if (!lower || qt == n_qt - 1) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int c0 = 32 * chunk + 16 * mt;
          // (rbase + c0 may be negative: the row guard keeps every dereferenced address in range;
          // band column 127 only ever carries the previous pair's row 63 - its own operands are 0)
          float* dst = dpos_b + (static_cast<long long>(rbase) + c0) * HD + hcol + wc;
          auto r_ok = [&](int r) {
            const int rr = rbase + c0 + r;
            return rr >= 0 && rr <= 2 * T - 2;
          };
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) red_fragment(dst + 8 * nt, HD, pp[mt][nt], g, t4, r_ok);
        }
      }
    }
  }
  // ---- d pos_bias_u / d pos_bias_v: column sums over this CTA's query rows (all tiles) ----
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = su[nt][e], c = sv[nt][e];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        c += __shfl_xor_sync(0xffffffffu, c, off);
      }
      if (g == 0) {
        red_add(p.du + hcol + wc + 8 * nt + 2 * t4 + e, a);
        red_add(p.dvb + hcol + wc + 8 * nt + 2 * t4 + e, c);
      }
    }
  // ---- this CTA's dk / dv rows ----
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int j = j0 + wr + g + 8 * h2;
    if (j < T) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int d = wc + 8 * nt + 2 * t4;
        *reinterpret_cast<float2*>(dk_out + (row0 + j) * p.ld_dqkv + d) =
            make_float2(acc_dk[nt][2 * h2], acc_dk[nt][2 * h2 + 1]);
        *reinterpret_cast<float2*>(dv_out + (row0 + j) * p.ld_dqkv + d) =
            make_float2(acc_dv[nt][2 * h2], acc_dv[nt][2 * h2 + 1]);
      }
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
                                     float* dqkv, long long ld_dqkv, float* du, float* dvb,
                                     float* dpos, const uint8_t* drop_keep, long long ld_drop,
                                     float drop_scale, int B, int T, int H, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && H > 0, "attn_bwd: bad shape B=%d T=%d H=%d", B, T, H);
  TAVSR_REQUIRE(qkv && pos && u && v && ctx && dctx && lse && dqkv && du && dvb && dpos,
                "attn_bwd: null pointer");
  TAVSR_REQUIRE(ld_qkv % 4 == 0 && ld_pos % 4 == 0 && ld_ctx % 4 == 0 && ld_dctx % 4 == 0 &&
                    ld_dqkv % 4 == 0,
                "attn_bwd: pitches must be multiples of 4 floats");
  attn_bwd::Params p;
  p.qkv = qkv; p.ld_qkv = ld_qkv; p.pos = pos; p.ld_pos = ld_pos; p.u = u; p.v = v; p.lens = lens;
  p.o = ctx; p.ld_o = ld_ctx; p.dout = dctx; p.ld_do = ld_dctx; p.lse = lse;
  p.dqkv = dqkv; p.ld_dqkv = ld_dqkv; p.du = du; p.dvb = dvb; p.dpos = dpos;
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
