// Fused relative-position multi-head self-attention on tcgen05 / TMEM / TMA (sm_100a).
//
// Reference semantics (espnet RelPositionMultiHeadedAttention.forward, called from
// src/encoder/branchformer/encoder_layer.py:208 and tailored/encoder_layer.py:192,239):
//   ac[i,j] = (q_i + u) . k_j
//   bd[i,j] = (q_i + v) . p_{T-1-i+j}            (rel_shift of (q+v) P^T, P = linear_pos(pos_emb))
//   attn    = softmax_j((ac + bd) / sqrt(d_k)) with keys j >= len[b] masked to probability 0
//   ctx_i   = sum_j attn[i,j] v_j
//
// One CTA = 128 query rows of one (utterance, head), looping over 128-key tiles (flash style, the
// (B,h,T,T) scores and the (B,h,T,2T-1) pre-shift tensor never exist).  Per key tile:
//   S  = Q K^T            128 x 128, 8 tcgen05.mma (K = 64), D in TMEM columns [0,128)
//   R  = Q Pband^T        128 x 256: the 255 relative positions this tile pair can touch,
//                         TMEM columns [128,384)
//   the bias terms are rank-1:  (q+u).k = q.k + u.k,  (q+v).p = q.p + v.p; the softmax warps
//   compute u.k_j and v.p_r from the smem tiles while the MMAs run, so Q is used as TMA wrote it
//   softmax warps (thread == query row == TMEM lane): rel-shift = every row reads its R row at a
//   lane-dependent offset (bounced through a private smem strip that aliases the consumed
//   P-band tile), online softmax in the exp2
//   domain, probabilities written back IN PLACE over S as TF32
//   O += P V              A operand read from TMEM; the V tile ([key][d] as TMA wrote it) is
//                         transposed by the softmax warps into a K-major [d][key] tile (tf32
//                         MMAs take MN-major operands only in the 32B-atom swizzle, which a
//                         {32 x rows} TMA box does not produce), D in TMEM columns [384,448)
// Warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 softmax (two per
// TMEM lane quadrant, splitting the keys of a tile).
#include <atomic>

#include "host.h"
#include "ptx.cuh"

namespace tavsr {
extern std::atomic<long long> g_launches;
extern void* g_debug_ptr;

namespace attn_tc {

constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 128;   // keys per tile
constexpr int kD = 64;     // head dim
constexpr int kBand = 256; // relative positions per (query tile, key tile): 255 used
constexpr int kThreads = 320;  // TMA warp, MMA warp, 8 softmax warps
constexpr int kBouncePitch = 36;  // floats; 16-byte aligned rows, skewed per-lane reads conflict-free

constexpr int kOffQ = 0;                       // 2 atoms x 128 rows x 128 B
constexpr int kOffK = kOffQ + 32768;
constexpr int kOffV = kOffK + 32768;
constexpr int kOffP = kOffV + 32768;           // 2 atoms x 256 rows x 128 B
constexpr int kOffVt = kOffP + 65536;          // V^T: 4 key atoms x 64 d-rows x 128 B
constexpr int kOffXch = kOffVt + 32768;        // row max / sum exchange: 2 parities x 2 halves x 128
constexpr int kOffCu = kOffXch + 2 * 256 * 4;
constexpr int kOffCv = kOffCu + 2 * kKT * 4;      // cu / cv are double-buffered by tile parity
constexpr int kOffU = kOffCv + 2 * kBand * 4;
constexpr int kOffVb = kOffU + kD * 4;
constexpr int kOffBar = kOffVb + kD * 4;
constexpr int kSmemBytes = 1024 + kOffBar + 128;

constexpr uint32_t kColS = 0, kColR = 128, kColO = 384, kTmemCols = 512;

struct Params {
  CUtensorMap tmQKV;  // [B*T, 3*H*64], box {32, 128}
  CUtensorMap tmPos;  // [2T-1, H*64],  box {32, 256}
  const float* u;
  const float* v;
  const int32_t* lens;
  float* ctx;
  long long ld_ctx;
  int T, H, round_out;
  // optional: partial row dots of the stored context with two (H*64)-vectors, one (a, b) pair per
  // (head, 32-column half): dots_out[(row * 2H + 2h + half)] - the learned_ave pooling scores
  const float* dva;
  const float* dvb;
  float2* dots_out;
  long long* dbg;  // optional phase timestamps (16 per CTA), tools/time_attn.py
};

__device__ __forceinline__ long long attn_globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ATTN_STAMP(slot)                                                                       \
  do {                                                                                         \
    if (p.dbg != nullptr && threadIdx.x == 64)                                                 \
      p.dbg[((static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x +       \
             blockIdx.x) * 16 + (slot)] = attn_globaltimer_ns();                                \
  } while (0)

// dot of a 64-vector `w` (smem, plain) with row `row` of a K-major SW128 tile of `rows` rows
__device__ __forceinline__ float dot_row_sw128(const uint8_t* tile, int rows, int row,
                                               const float* w) {
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const uint8_t* r = tile + a * rows * 128 + row * 128;
#pragma unroll
    for (int lc = 0; lc < 8; ++lc) {
      const float4 x = *reinterpret_cast<const float4*>(r + ((lc ^ (row & 7)) << 4));
      const float4 y = *reinterpret_cast<const float4*>(w + a * 32 + lc * 4);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
      acc = fmaf(x.z, y.z, acc);
      acc = fmaf(x.w, y.w, acc);
    }
  }
  return acc;
}

__global__ void __launch_bounds__(kThreads, 1)
relpos_attn_tc_kernel(const __grid_constant__ Params p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align_smem_1024(smem_raw);
  uint8_t* sQ = sm + kOffQ;
  uint8_t* sK = sm + kOffK;
  uint8_t* sV = sm + kOffV;
  uint8_t* sP = sm + kOffP;
  uint8_t* sVt = sm + kOffVt;
  float* s_xch = reinterpret_cast<float*>(sm + kOffXch);
  float* s_cu_all = reinterpret_cast<float*>(sm + kOffCu);
  float* s_cv_all = reinterpret_cast<float*>(sm + kOffCv);
  float* s_u = reinterpret_cast<float*>(sm + kOffU);
  float* s_v = reinterpret_cast<float*>(sm + kOffVb);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* s_done = bars + 4;
  uint64_t* kp_free = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int T = p.T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = blockIdx.x * kQT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmPos);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(p_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_done, 1);
    mbar_init(kp_free, 8);
    mbar_init(p_ready, 8);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 2) {  // pos_bias_u / pos_bias_v of this head (weights: legal before the PDL wait)
    const int t = threadIdx.x - 64;
    if (t < kD) s_u[t] = __ldg(p.u + h * kD + t);
    else if (t < 2 * kD) s_v[t - kD] = __ldg(p.v + h * kD + t - kD);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  int len = p.lens ? p.lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  const int n_kv = (len + kKT - 1) / kKT;
  const int hcol = h * kD;
  const int row0 = b * T;

  if (warp == 0) {
    // ======================================= TMA producer ======================================
    if (lane == 0 && n_kv > 0) {
      mbar_arrive_expect_tx(q_full, 32768);
      tma_load_2d(sQ, &p.tmQKV, q_full, hcol, row0 + i0);
      tma_load_2d(sQ + 16384, &p.tmQKV, q_full, hcol + 32, row0 + i0);
      for (int t = 0; t < n_kv; ++t) {
        const int j0 = t * kKT;
        if (t > 0) mbar_wait(kp_free, (t - 1) & 1);
        mbar_arrive_expect_tx(k_full, 32768);
        tma_load_2d(sK, &p.tmQKV, k_full, p.H * kD + hcol, row0 + j0);
        tma_load_2d(sK + 16384, &p.tmQKV, k_full, p.H * kD + hcol + 32, row0 + j0);
        const int rbase = T - kQT - i0 + j0;  // band column c <-> relative-position row rbase + c
        mbar_arrive_expect_tx(p_full, 65536);
        tma_load_2d(sP, &p.tmPos, p_full, hcol, rbase);
        tma_load_2d(sP + 32768, &p.tmPos, p_full, hcol + 32, rbase);
        if (t > 0) mbar_wait(p_ready, (t - 1) & 1);  // the V tile has been transposed out of sV
        mbar_arrive_expect_tx(v_full, 32768);
        tma_load_2d(sV, &p.tmQKV, v_full, 2 * p.H * kD + hcol, row0 + j0);
        tma_load_2d(sV + 16384, &p.tmQKV, v_full, 2 * p.H * kD + hcol + 32, row0 + j0);
      }
    }
  } else if (warp == 1) {
    // ======================================== MMA issuer =======================================
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idescS = umma_idesc(UMMA_FMT_TF32, 128, 128);
      constexpr uint32_t idescR = umma_idesc(UMMA_FMT_TF32, 128, 256);
      constexpr uint32_t idescO = umma_idesc(UMMA_FMT_TF32, 128, 64);
      const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK));
      const uint64_t dP = umma_desc_kmajor_sw128(smem_u32(sP));
      const uint64_t dV = umma_desc_kmajor_sw128(smem_u32(sVt));
      mbar_wait(q_full, 0);
      for (int t = 0; t < n_kv; ++t) {
        mbar_wait(k_full, t & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss<true>(tmem_base + kColS, dQ + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3),
                        dK + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3), idescS, ks ? 1u : 0u);
        mbar_wait(p_full, t & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss<true>(tmem_base + kColR, dQ + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3),
                        dP + (ks >> 2) * (32768 >> 4) + 2 * (ks & 3), idescR, ks ? 1u : 0u);
        umma_commit(s_done);
        mbar_wait(p_ready, t & 1);  // P in TMEM, V^T in smem
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          umma_ts_tf32(tmem_base + kColO, tmem_base + kColS + 8 * ks,
                       dV + (ks >> 2) * (8192 >> 4) + 2 * (ks & 3), idescO, (t | ks) ? 1u : 0u);
        umma_commit(o_done);
      }
    }
  } else {
    // ====================================== softmax warps ======================================
    // 8 warps: TMEM quadrant q = warp % 4 (rows 32q..32q+31), key half hf (keys 64hf..64hf+63 of
    // the tile, output columns 32hf..32hf+31).  The two threads of a row exchange their maxima
    // through smem every tile and their sums once at the end.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int row = q * 32 + static_cast<int>(lane);
    const int tid = threadIdx.x - 64;           // 0..255
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* br = reinterpret_cast<float*>(sP) + tid * kBouncePitch;  // skew strip, aliases the P band
    const float scale = 0.125f * 1.4426950408889634f;  // 1/sqrt(d_k) in the exp2 domain
    float m_run = -INFINITY, l_run = 0.f;
    ATTN_STAMP(0);
    for (int t = 0; t < n_kv; ++t) {
      const int j0 = t * kKT;
      // a fast warp may start tile t+1 while a slow one still reads tile t's bias terms / maxima
      float* s_cu = s_cu_all + (t & 1) * kKT;
      float* s_cv = s_cv_all + (t & 1) * kBand;
      float* s_mx = s_xch + (t & 1) * 256;
      // rank-1 bias terms from the tiles TMA just delivered (overlaps the S / R MMAs)
      mbar_wait(k_full, t & 1);
      if (t < 2) ATTN_STAMP(1 + 6 * t);
      if (tid < kKT) s_cu[tid] = dot_row_sw128(sK, kKT, tid, s_u);
      mbar_wait(p_full, t & 1);
      s_cv[tid] = dot_row_sw128(sP, kBand, tid, s_v);
      named_bar_sync(1, 256);
      if (t < 2) ATTN_STAMP(2 + 6 * t);
      mbar_wait(s_done, t & 1);
      tc_fence_after_sync();
      if (t < 2) ATTN_STAMP(3 + 6 * t);
      // ---- pass 1: s = (S + u.k + rel_shift(R + v.p)) * scale, masked; row max; s back to TMEM
      float tile_max = -INFINITY;
      // R chunk m, read back circularly: w[e] = R[row][32 m + ((31 - lane + e) & 31)] + v.p
      auto skew_chunk = [&](int m, float (&w)[32]) {
        uint32_t ra[32];
        tmem_ld32(trow + kColR + 32 * m, ra);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 ca = *reinterpret_cast<const float4*>(s_cv + 32 * m + 4 * i);
          *reinterpret_cast<float4*>(br + 4 * i) =
              make_float4(__uint_as_float(ra[4 * i]) + ca.x, __uint_as_float(ra[4 * i + 1]) + ca.y,
                          __uint_as_float(ra[4 * i + 2]) + ca.z, __uint_as_float(ra[4 * i + 3]) + ca.w);
        }
        // lane l wrote its own strip and reads it back rotated by 31 - l: no cross-lane hazard
#pragma unroll
        for (int e = 0; e < 32; ++e) w[e] = br[(31 - static_cast<int>(lane) + e) & 31];
      };
      float wa[32], wb[32];
      skew_chunk(2 * hf - q + 3, wa);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * hf + cc;
        // key 32c+e of this row sits at band column 32 (c - q + 3) + (31 - lane + e): the first
        // lane+1 keys come from chunk c-q+3, the rest from the next chunk
        uint32_t rs[32];
        tmem_ld32(trow + kColS + 32 * c, rs);
        skew_chunk(c - q + 4, wb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 cu4 = *reinterpret_cast<const float4*>(s_cu + 32 * c + 4 * i);
          const float cu[4] = {cu4.x, cu4.y, cu4.z, cu4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = 4 * i + k;
            const int j = j0 + 32 * c + e;
            const float r = e <= static_cast<int>(lane) ? wa[e] : wb[e];
            float sc = (__uint_as_float(rs[e]) + cu[k] + r) * scale;
            sc = j < len ? sc : -INFINITY;
            tile_max = fmaxf(tile_max, sc);
            rs[e] = __float_as_uint(sc);
            wa[e] = wb[e];
          }
        }
        tmem_st32(trow + kColS + 32 * c, rs);
      }
      s_mx[hf * 128 + row] = tile_max;
      tmem_st_wait();
      named_bar_sync(1, 256);                 // maxima exchanged; every strip read of sP is done
      if (lane == 0) mbar_arrive(kp_free);    // K / Pband smem: MMAs retired and our reads done
      if (t < 2) ATTN_STAMP(4 + 6 * t);
      // ---- pass 2: p = exp2(s - m), row sum, TF32 probabilities in place
      const float m_new = fmaxf(m_run, fmaxf(tile_max, s_mx[(hf ^ 1) * 128 + row]));
      const float alpha = exp2f(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      float sum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * hf + cc;
        uint32_t rs[32];
        tmem_ld32(trow + kColS + 32 * c, rs);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float pr = exp2f(__uint_as_float(rs[e]) - m_new);
          sum += pr;
          rs[e] = __float_as_uint(round_tf32(pr));
        }
        tmem_st32(trow + kColS + 32 * c, rs);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      if (t > 0) {  // rescale the running output once the previous P.V has retired
        mbar_wait(o_done, (t - 1) & 1);
        tc_fence_after_sync();
        uint32_t ro[32];
        tmem_ld32(trow + kColO + 32 * hf, ro);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) ro[e] = __float_as_uint(__uint_as_float(ro[e]) * alpha);
        tmem_st32(trow + kColO + 32 * hf, ro);
      }
      if (t < 2) ATTN_STAMP(5 + 6 * t);
      // ---- V tile [key][d] -> V^T [d][key], K-major SW128; sVt is free: the previous P.V
      //      retired (o_done above).  Thread = (key, d half).
      mbar_wait(v_full, t & 1);
      {
        const int key = tid & 127;
        const int a = tid >> 7;
        uint8_t* dst = sVt + (key >> 5) * 8192 + (key & 3) * 4;
        const int lc = (key & 31) >> 2;
        const uint8_t* src = sV + a * 16384 + key * 128;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 x = *reinterpret_cast<const float4*>(src + ((c4 ^ (key & 7)) << 4));
          const int d = a * 32 + c4 * 4;
          *reinterpret_cast<float*>(dst + (d + 0) * 128 + ((lc ^ ((d + 0) & 7)) << 4)) = x.x;
          *reinterpret_cast<float*>(dst + (d + 1) * 128 + ((lc ^ ((d + 1) & 7)) << 4)) = x.y;
          *reinterpret_cast<float*>(dst + (d + 2) * 128 + ((lc ^ ((d + 2) & 7)) << 4)) = x.z;
          *reinterpret_cast<float*>(dst + (d + 3) * 128 + ((lc ^ ((d + 3) & 7)) << 4)) = x.w;
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes of V^T -> visible to the MMA (async proxy)
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      if (t < 2) ATTN_STAMP(6 + 6 * t);
    }
    // ---- epilogue: ctx = O / l   (this thread: output columns 32hf..32hf+31 of its row)
    const int i = i0 + row;
    float* out = p.ctx + static_cast<long long>(row0 + i) * p.ld_ctx + hcol + 32 * hf;
    if (n_kv > 0) {
      float* s_sum = s_xch + (n_kv & 1) * 256;  // slot not in use by the last tile's maxima
      s_sum[hf * 128 + row] = l_run;
      named_bar_sync(1, 256);
      const float inv = 1.0f / (l_run + s_sum[(hf ^ 1) * 128 + row]);
      mbar_wait(o_done, (n_kv - 1) & 1);
      tc_fence_after_sync();
      ATTN_STAMP(13);
      const int dbg = p.round_out >> 8;  // 1: dump the last tile's probabilities of keys 0..63
      uint32_t ro[32];
      tmem_ld32(trow + (dbg == 1 ? kColS : kColO) + 32 * hf, ro);
      tmem_ld_wait();
      if (i < T) {
        float da = 0.f, db = 0.f;
        const float4* va4 = reinterpret_cast<const float4*>(p.dva + hcol + 32 * hf);
        const float4* vb4 = reinterpret_cast<const float4*>(p.dvb + hcol + 32 * hf);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float4 o = make_float4(__uint_as_float(ro[4 * e]) * inv, __uint_as_float(ro[4 * e + 1]) * inv,
                                 __uint_as_float(ro[4 * e + 2]) * inv, __uint_as_float(ro[4 * e + 3]) * inv);
          if (p.round_out & 1) {
            o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
          }
          *reinterpret_cast<float4*>(out + 4 * e) = o;
          if (p.dots_out != nullptr) {
            const float4 a4 = __ldg(va4 + e), b4 = __ldg(vb4 + e);
            da += o.x * a4.x + o.y * a4.y + o.z * a4.z + o.w * a4.w;
            db += o.x * b4.x + o.y * b4.y + o.z * b4.z + o.w * b4.w;
          }
        }
        if (p.dots_out != nullptr)
          p.dots_out[static_cast<long long>(row0 + i) * (2 * p.H) + 2 * h + hf] = make_float2(da, db);
      }
    } else if (i < T) {
      for (int e = 0; e < 8; ++e)
        *reinterpret_cast<float4*>(out + 4 * e) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.dots_out != nullptr)
        p.dots_out[static_cast<long long>(row0 + i) * (2 * p.H) + 2 * h + hf] = make_float2(0.f, 0.f);
    }
  }

  ATTN_STAMP(14);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace attn_tc

int relpos_attn_tc_launch(const float* qkv, long long ld_qkv, const float* pos, long long ld_pos,
                          const float* u, const float* v, const int32_t* lens, float* ctx,
                          long long ld_ctx, int B, int T, int H, int round_out, const float* dva,
                          const float* dvb, float* dots_out, cudaStream_t s) {
  attn_tc::Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap_2d(&p.tmQKV, qkv, 4, false, static_cast<uint64_t>(B) * T, 3ull * H * 64,
                         ld_qkv, 128, 32)))
    return rc;
  if ((rc = make_tmap_2d(&p.tmPos, pos, 4, false, 2ull * T - 1, static_cast<uint64_t>(H) * 64, ld_pos,
                         256, 32)))
    return rc;
  p.u = u;
  p.v = v;
  p.lens = lens;
  p.ctx = ctx;
  p.ld_ctx = ld_ctx;
  p.T = T;
  p.H = H;
  p.dva = dva;
  p.dvb = dvb;
  p.dots_out = reinterpret_cast<float2*>(dots_out);
  p.dbg = reinterpret_cast<long long*>(g_debug_ptr);
  p.round_out = (round_out ? 1 : 0) | (g_debug[9] << 8);
  static bool configured = false;
  if (!configured) {
    TAVSR_CUDA_OK(cudaFuncSetAttribute(attn_tc::relpos_attn_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       attn_tc::kSmemBytes));
    configured = true;
  }
  dim3 grid((T + attn_tc::kQT - 1) / attn_tc::kQT, H, B);
  TAVSR_CUDA_OK(launch_kernel(attn_tc::relpos_attn_tc_kernel, grid, dim3(attn_tc::kThreads),
                              attn_tc::kSmemBytes, s, 0, p));
  return 0;
}

}  // namespace tavsr
