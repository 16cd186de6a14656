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
//
// bf16 mode (kind::f16): Q / K / V rows are exactly one 128-byte swizzle row (64 bf16), so every
// tile is half the bytes and half the MMAs (K = 16 per instruction); the probabilities go back to
// TMEM as PACKED bf16 pairs (the A-from-TMEM layout) and the V tile is consumed as TMA wrote it, as
// an MN-major B operand - the transpose pass of the tf32 kernel disappears; the freed shared memory
// double-buffers the K / P-band / V tiles so the loads of key tile t+1 overlap tile t.
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

template <bool kBf16>
struct Cfg {
  static constexpr int kTile = kBf16 ? 16384 : 32768;       // a 128-row Q / K / V tile
  static constexpr int kPBytes = 2 * kTile;                  // the 256-row P band
  static constexpr int kStages = kBf16 ? 2 : 1;              // K / P / V buffers
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kTile;
  static constexpr int kOffV = kOffK + kStages * kTile;
  static constexpr int kOffP = kOffV + kStages * kTile;
  // tf32: V^T (4 key atoms x 64 d-rows x 128 B) and the skew strip aliasing the consumed P band;
  // bf16: a dedicated skew strip (256 threads x 36 floats)
  static constexpr int kOffVt = kOffP + kStages * kPBytes;
  static constexpr int kOffXch = kOffVt + (kBf16 ? 256 * kBouncePitch * 4 : 32768);
  static constexpr int kOffCu = kOffXch + 2 * 256 * 4;      // row max / sum exchange
  static constexpr int kOffCv = kOffCu + 2 * kKT * 4;       // cu / cv are double-buffered by tile parity
  static constexpr int kOffU = kOffCv + 2 * kBand * 4;
  static constexpr int kOffVb = kOffU + kD * 4;
  static constexpr int kOffBar = kOffVb + kD * 4;
  static constexpr int kSmemBytes = 1024 + kOffBar + 256;
};

constexpr uint32_t kColS = 0, kColR = 128, kColO = 384, kTmemCols = 512;

struct Params {
  CUtensorMap tmQKV;  // [B*T, 3*H*64], box {32, 128}
  CUtensorMap tmPos;  // [2T-1, H*64],  box {32, 256}
  const float* u;
  const float* v;
  const int32_t* lens;
  void* ctx;       // fp32 (tf32 mode) or bf16
  long long ld_ctx;
  int T, H, round_out;
  float* lse;      // optional [B, H, T]: base-2 log-sum-exp of the scaled scores (training forward)
  // training only (tf32 kernel): attention-probability dropout, espnet attention.py
  // `matmul(self.dropout(self.attn), value)`.  keep[b][h][i][j] != 0 keeps P_ij; the kept
  // probabilities are scaled by drop_scale = 1 / (1 - p) (folded into the final 1 / l); the row
  // sum and the log-sum-exp stay those of the undropped softmax.
  const uint8_t* drop_keep;
  long long ld_drop;  // row pitch in bytes, multiple of 16
  float drop_scale;
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

// same for a bf16 tile: a row is ONE 128-byte swizzle row of 64 values
__device__ __forceinline__ float dot_row_sw128_bf16(const uint8_t* tile, int row, const float* w) {
  float acc = 0.f;
  const uint8_t* r = tile + row * 128;
#pragma unroll
  for (int lc = 0; lc < 8; ++lc) {
    const uint4 x = *reinterpret_cast<const uint4*>(r + ((lc ^ (row & 7)) << 4));
    const float4 y0 = *reinterpret_cast<const float4*>(w + lc * 8);
    const float4 y1 = *reinterpret_cast<const float4*>(w + lc * 8 + 4);
    acc = fmaf(bf16_lo(x.x), y0.x, acc);
    acc = fmaf(bf16_hi(x.x), y0.y, acc);
    acc = fmaf(bf16_lo(x.y), y0.z, acc);
    acc = fmaf(bf16_hi(x.y), y0.w, acc);
    acc = fmaf(bf16_lo(x.z), y1.x, acc);
    acc = fmaf(bf16_hi(x.z), y1.y, acc);
    acc = fmaf(bf16_lo(x.w), y1.z, acc);
    acc = fmaf(bf16_hi(x.w), y1.w, acc);
  }
  return acc;
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
relpos_attn_tc_kernel(const __grid_constant__ Params p) {
  using C = Cfg<kBf16>;
  constexpr int kTile = C::kTile;
  constexpr int kStages = C::kStages;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align_smem_1024(smem_raw);
  uint8_t* sQ = sm + C::kOffQ;
  uint8_t* sK0 = sm + C::kOffK;
  uint8_t* sV0 = sm + C::kOffV;
  uint8_t* sP0 = sm + C::kOffP;
  uint8_t* sVt = sm + C::kOffVt;   // tf32: V^T tile; bf16: the skew strip
  float* s_xch = reinterpret_cast<float*>(sm + C::kOffXch);
  float* s_cu_all = reinterpret_cast<float*>(sm + C::kOffCu);
  float* s_cv_all = reinterpret_cast<float*>(sm + C::kOffCv);
  float* s_u = reinterpret_cast<float*>(sm + C::kOffU);
  float* s_v = reinterpret_cast<float*>(sm + C::kOffVb);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2] by stage
  uint64_t* p_full = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* kp_free = bars + 7;   // [2] by stage: K / P-band buffer consumed
  uint64_t* v_free = bars + 9;    // [2] by stage (bf16): P.V of the tile in this V buffer retired
  uint64_t* s_done = bars + 11;
  uint64_t* p_ready = bars + 12;
  uint64_t* o_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int T = p.T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = blockIdx.x * kQT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmPos);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&p_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kp_free[i], 8);
      mbar_init(&v_free[i], 1);
    }
    mbar_init(s_done, 1);
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
    // Stage st = t % kStages, use u = t / kStages.  A K / P-band buffer is free once the S / R MMAs
    // of the tile that used it retired and the softmax warps finished reading it (kp_free[st]); a V
    // buffer once the P.V MMAs of its tile retired (bf16: v_free[st]) or the tile was transposed out
    // of it (tf32: p_ready).  Every release barrier belongs to ONE stage, so the next phase of a
    // barrier this thread waits on cannot complete before this thread has refilled that stage (a
    // parity wait is only sound while the barrier is at most one phase ahead of the waiter).  With
    // two stages the loads of tile t+1 are in flight during tile t.
    if (lane == 0 && n_kv > 0) {
      constexpr int kHalf = kBf16 ? 64 : 32;   // elements per 128-byte box row
      constexpr int kBoxes = kD / kHalf;       // boxes per 64-wide operand row
      mbar_arrive_expect_tx(q_full, kTile);
      for (int a = 0; a < kBoxes; ++a)
        tma_load_2d(sQ + a * 16384, &p.tmQKV, q_full, hcol + a * kHalf, row0 + i0);
      for (int t = 0; t < n_kv; ++t) {
        const int j0 = t * kKT;
        const int st = t % kStages;
        const int u = t / kStages;           // use count of this stage
        if (u > 0) mbar_wait(&kp_free[st], (u - 1) & 1);
        uint8_t* sK = sK0 + st * kTile;
        uint8_t* sP = sP0 + st * C::kPBytes;
        uint8_t* sV = sV0 + st * kTile;
        mbar_arrive_expect_tx(&k_full[st], kTile);
        for (int a = 0; a < kBoxes; ++a)
          tma_load_2d(sK + a * 16384, &p.tmQKV, &k_full[st], p.H * kD + hcol + a * kHalf, row0 + j0);
        const int rbase = T - kQT - i0 + j0;  // band column c <-> relative-position row rbase + c
        mbar_arrive_expect_tx(&p_full[st], C::kPBytes);
        for (int a = 0; a < kBoxes; ++a)
          tma_load_2d(sP + a * 32768, &p.tmPos, &p_full[st], hcol + a * kHalf, rbase);
        if (u > 0) {
          if (kBf16) mbar_wait(&v_free[st], (u - 1) & 1);    // P.V of that tile retired
          else mbar_wait(p_ready, (t - 1) & 1);              // the V tile has been transposed out
        }
        mbar_arrive_expect_tx(&v_full[st], kTile);
        for (int a = 0; a < kBoxes; ++a)
          tma_load_2d(sV + a * 16384, &p.tmQKV, &v_full[st], 2 * p.H * kD + hcol + a * kHalf, row0 + j0);
      }
    }
  } else if (warp == 1) {
    // ======================================== MMA issuer =======================================
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t fmt = kBf16 ? UMMA_FMT_BF16 : UMMA_FMT_TF32;
      constexpr uint32_t idescS = umma_idesc(fmt, 128, 128);
      constexpr uint32_t idescR = umma_idesc(fmt, 128, 256);
      constexpr uint32_t idescO = umma_idesc(fmt, 128, 64) | (kBf16 ? kUmmaBMajorMN : 0u);
      const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      for (int t = 0; t < n_kv; ++t) {
        const int st = t % kStages;
        const uint32_t ph = static_cast<uint32_t>(t / kStages) & 1u;
        const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK0 + st * kTile));
        const uint64_t dP = umma_desc_kmajor_sw128(smem_u32(sP0 + st * C::kPBytes));
        mbar_wait(&k_full[st], ph);
        tc_fence_after_sync();
        if constexpr (kBf16) {
          // (the S columns still hold tile t-1's packed probabilities: MMAs of one thread execute
          // in issue order, so this overwrite runs after P.V(t-1))
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss<false>(tmem_base + kColS, dQ + 2 * ks, dK + 2 * ks, idescS, ks ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss<true>(tmem_base + kColS, dQ + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3),
                          dK + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3), idescS, ks ? 1u : 0u);
        }
        mbar_wait(&p_full[st], ph);
        tc_fence_after_sync();
        if constexpr (kBf16) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss<false>(tmem_base + kColR, dQ + 2 * ks, dP + 2 * ks, idescR, ks ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss<true>(tmem_base + kColR, dQ + (ks >> 2) * (16384 >> 4) + 2 * (ks & 3),
                          dP + (ks >> 2) * (32768 >> 4) + 2 * (ks & 3), idescR, ks ? 1u : 0u);
        }
        umma_commit(s_done);
        mbar_wait(p_ready, t & 1);  // P in TMEM (tf32: V^T in smem)
        tc_fence_after_sync();
        if constexpr (kBf16) {
          mbar_wait(&v_full[st], ph);
          tc_fence_after_sync();
          // V tile [key][d] as TMA wrote it == MN-major B operand: 8 keys per 1024-byte atom, one
          // K = 16 step = two atoms; the packed probabilities of keys [64 h, 64 h + 64) sit in S
          // columns [64 h, 64 h + 32)
          const uint32_t v_addr = smem_u32(sV0 + st * kTile);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ts_f16(tmem_base + kColO, tmem_base + kColS + (ks >> 2) * 64 + (ks & 3) * 8,
                        umma_desc_mnmajor_sw128_b16(v_addr + ks * 2048, 1024), idescO,
                        (t | ks) ? 1u : 0u);
          umma_commit(&v_free[st]);
        } else {
          const uint64_t dV = umma_desc_kmajor_sw128(smem_u32(sVt));
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_ts_tf32(tmem_base + kColO, tmem_base + kColS + 8 * ks,
                         dV + (ks >> 2) * (8192 >> 4) + 2 * (ks & 3), idescO, (t | ks) ? 1u : 0u);
        }
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
    const float scale = 0.125f * 1.4426950408889634f;  // 1/sqrt(d_k) in the exp2 domain
    float m_run = -INFINITY, l_run = 0.f;
    ATTN_STAMP(0);
    for (int t = 0; t < n_kv; ++t) {
      const int j0 = t * kKT;
      const int st = t % kStages;
      const uint32_t ph = static_cast<uint32_t>(t / kStages) & 1u;
      const uint8_t* sK = sK0 + st * kTile;
      uint8_t* sP = sP0 + st * C::kPBytes;
      const uint8_t* sV = sV0 + st * kTile;
      // skew strip: tf32 aliases the consumed P band, bf16 owns a dedicated region
      float* br = reinterpret_cast<float*>(kBf16 ? sVt : sP) + tid * kBouncePitch;
      // a fast warp may start tile t+1 while a slow one still reads tile t's bias terms / maxima
      float* s_cu = s_cu_all + (t & 1) * kKT;
      float* s_cv = s_cv_all + (t & 1) * kBand;
      float* s_mx = s_xch + (t & 1) * 256;
      // rank-1 bias terms from the tiles TMA just delivered (overlaps the S / R MMAs)
      mbar_wait(&k_full[st], ph);
      if (t < 2) ATTN_STAMP(1 + 6 * t);
      if (tid < kKT) s_cu[tid] = kBf16 ? dot_row_sw128_bf16(sK, tid, s_u) : dot_row_sw128(sK, kKT, tid, s_u);
      mbar_wait(&p_full[st], ph);
      s_cv[tid] = kBf16 ? dot_row_sw128_bf16(sP, tid, s_v) : dot_row_sw128(sP, kBand, tid, s_v);
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
      if (lane == 0) mbar_arrive(&kp_free[st]);  // K / Pband smem: MMAs retired and our reads done
      if (t < 2) ATTN_STAMP(4 + 6 * t);
      // ---- pass 2: p = exp2(s - m), row sum, probabilities in place (TF32 values / packed bf16)
      const float m_new = fmaxf(m_run, fmaxf(tile_max, s_mx[(hf ^ 1) * 128 + row]));
      const float alpha = exp2f(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      float sum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * hf + cc;
        uint32_t rs[32];
        tmem_ld32(trow + kColS + 32 * c, rs);
        tmem_ld_wait();
        if constexpr (kBf16) {
          // keys [32 c, 32 c + 32) -> 16 packed columns at [64 hf + 16 cc, +16): inside the column
          // range this thread alone reads, behind the chunk just consumed.  The row sum takes the
          // ROUNDED probabilities, i.e. exactly the weights the P.V product applies.
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = exp2f(__uint_as_float(rs[2 * e]) - m_new);
            const float p1 = exp2f(__uint_as_float(rs[2 * e + 1]) - m_new);
            w[e] = pack_bf16x2(p0, p1);
            sum += bf16_lo(w[e]) + bf16_hi(w[e]);
          }
          tmem_st16(trow + kColS + 64 * hf + 16 * cc, w);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float pr = exp2f(__uint_as_float(rs[e]) - m_new);
            sum += pr;
            rs[e] = __float_as_uint(round_tf32(pr));
          }
          if (p.drop_keep != nullptr && i0 + row < T) {
            // keys kv0 + 32 c .. + 31 of this query row: 32 keep bytes (the pitch covers the
            // last tile's overhang; masked keys are already zero)
            const uint8_t* kp = p.drop_keep +
                ((static_cast<long long>(b) * p.H + h) * T + i0 + row) * p.ld_drop + t * kKT + 32 * c;
            const uint4 k0 = ld_act_u4(reinterpret_cast<const uint4*>(kp));
            const uint4 k1 = ld_act_u4(reinterpret_cast<const uint4*>(kp) + 1);
            const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (((kw[e >> 2] >> (8 * (e & 3))) & 0xffu) == 0u) rs[e] = 0u;
          }
          tmem_st32(trow + kColS + 32 * c, rs);
        }
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
      if constexpr (!kBf16) {
        // ---- V tile [key][d] -> V^T [d][key], K-major SW128; sVt is free: the previous P.V
        //      retired (o_done above).  Thread = (key, d half).
        mbar_wait(&v_full[st], ph);
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
        fence_proxy_async_smem();  // generic-proxy writes of V^T -> visible to the MMA (async proxy)
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      if (t < 2) ATTN_STAMP(6 + 6 * t);
    }
    // ---- epilogue: ctx = O / l   (this thread: output columns 32hf..32hf+31 of its row)
    const int i = i0 + row;
    const long long off = static_cast<long long>(row0 + i) * p.ld_ctx + hcol + 32 * hf;
    uint32_t ro[32];
    float inv = 0.f;
    if (n_kv > 0) {
      float* s_sum = s_xch + (n_kv & 1) * 256;  // slot not in use by the last tile's maxima
      s_sum[hf * 128 + row] = l_run;
      named_bar_sync(1, 256);
      const float l_tot = l_run + s_sum[(hf ^ 1) * 128 + row];
      inv = (p.drop_keep != nullptr ? p.drop_scale : 1.0f) / l_tot;
      if (p.lse != nullptr && hf == 0 && i < T)
        p.lse[(static_cast<long long>(b) * p.H + h) * T + i] = m_run + log2f(l_tot);
      mbar_wait(o_done, (n_kv - 1) & 1);
      tc_fence_after_sync();
      ATTN_STAMP(13);
      tmem_ld32(trow + kColO + 32 * hf, ro);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) ro[e] = 0u;   // empty utterance: zero context
      if (p.lse != nullptr && hf == 0 && i < T) p.lse[(static_cast<long long>(b) * p.H + h) * T + i] = 0.f;
    }
    if (i < T) {
      if constexpr (kBf16) {
        uint4* out = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.ctx) + off);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ro[8 * e]) * inv, __uint_as_float(ro[8 * e + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ro[8 * e + 2]) * inv, __uint_as_float(ro[8 * e + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ro[8 * e + 4]) * inv, __uint_as_float(ro[8 * e + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ro[8 * e + 6]) * inv, __uint_as_float(ro[8 * e + 7]) * inv);
          out[e] = w;
        }
      } else {
        float* out = static_cast<float*>(p.ctx) + off;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float4 o = make_float4(__uint_as_float(ro[4 * e]) * inv, __uint_as_float(ro[4 * e + 1]) * inv,
                                 __uint_as_float(ro[4 * e + 2]) * inv, __uint_as_float(ro[4 * e + 3]) * inv);
          if (p.round_out & 1) {
            o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
          }
          *reinterpret_cast<float4*>(out + 4 * e) = o;
        }
      }
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

template <bool kBf16>
static int relpos_attn_launch(const void* qkv, long long ld_qkv, const void* pos, long long ld_pos,
                              const float* u, const float* v, const int32_t* lens, void* ctx,
                              long long ld_ctx, int B, int T, int H, int round_out, float* lse,
                              cudaStream_t s, const uint8_t* drop_keep = nullptr,
                              long long ld_drop = 0, float drop_scale = 1.0f) {
  using C = attn_tc::Cfg<kBf16>;
  attn_tc::Params p;
  memset(&p, 0, sizeof(p));
  constexpr int eb = kBf16 ? 2 : 4;
  int rc;
  if ((rc = make_tmap_2d(&p.tmQKV, qkv, eb, kBf16, static_cast<uint64_t>(B) * T, 3ull * H * 64,
                         ld_qkv, 128, 128 / eb)))
    return rc;
  if ((rc = make_tmap_2d(&p.tmPos, pos, eb, kBf16, 2ull * T - 1, static_cast<uint64_t>(H) * 64, ld_pos,
                         256, 128 / eb)))
    return rc;
  p.u = u;
  p.v = v;
  p.lens = lens;
  p.ctx = ctx;
  p.ld_ctx = ld_ctx;
  p.T = T;
  p.H = H;
  p.lse = lse;
  p.drop_keep = drop_keep;
  p.ld_drop = ld_drop;
  p.drop_scale = drop_scale;
  p.dbg = reinterpret_cast<long long*>(g_debug_ptr);
  p.round_out = round_out ? 1 : 0;
  auto kern = attn_tc::relpos_attn_tc_kernel<kBf16>;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    TAVSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::kSmemBytes));
  dim3 grid((T + attn_tc::kQT - 1) / attn_tc::kQT, H, B);
  TAVSR_CUDA_OK(launch_kernel(kern, grid, dim3(attn_tc::kThreads), C::kSmemBytes, s, 0, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace tavsr

using namespace tavsr;

extern "C" int tavsr_relpos_attn_fwd(const void* qkv, long long ld_qkv, const void* pos,
                                     long long ld_pos, const float* u, const float* v,
                                     const int32_t* lens, void* ctx, long long ld_ctx, int B,
                                     int T, int H, int round_out, int dtype, float* lse,
                                     void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && H > 0, "attn: bad shape B=%d T=%d H=%d", B, T, H);
  TAVSR_REQUIRE(qkv && pos && u && v && ctx, "attn: null pointer");
  const int op = dtype & TAVSR_DT_MASK;
  TAVSR_REQUIRE(op == TAVSR_DT_TF32 || op == (TAVSR_DT_BF16) , "attn: dtype must be tf32 or bf16");
  const bool bf16 = op == TAVSR_DT_BF16;
  TAVSR_REQUIRE(!bf16 || (dtype & TAVSR_DT_OUT_BF16), "attn: the bf16 kernel stores a bf16 context");
  const int al = bf16 ? 8 : 4;  // elements per 16 bytes
  TAVSR_REQUIRE(ld_qkv % al == 0 && ld_pos % al == 0 && ld_ctx % al == 0,
                "attn: pitches must be multiples of %d elements (16 bytes)", al);
  TAVSR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
                "attn: operands must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bf16 ? relpos_attn_launch<true>(qkv, ld_qkv, pos, ld_pos, u, v, lens, ctx, ld_ctx, B, T, H,
                                         round_out, lse, s)
              : relpos_attn_launch<false>(qkv, ld_qkv, pos, ld_pos, u, v, lens, ctx, ld_ctx, B, T, H,
                                          round_out, lse, s);
}

// Training forward with attention-probability dropout (tf32 storage): `drop_keep` is a (B, H, T)
// x ld_drop byte matrix, keep[b][h][i][j] != 0 keeps P_ij, ld_drop a multiple of 128 >= T (a key
// tile never crosses a row); kept probabilities are scaled by drop_scale = 1 / (1 - p).  `lse`
// stays the log-sum-exp of the undropped scores, which is what tavsr_relpos_attn_bwd wants.
extern "C" int tavsr_relpos_attn_fwd_dropout(const float* qkv, long long ld_qkv, const float* pos,
                                             long long ld_pos, const float* u, const float* v,
                                             const int32_t* lens, float* ctx, long long ld_ctx,
                                             int B, int T, int H, float* lse,
                                             const uint8_t* drop_keep, long long ld_drop,
                                             float drop_scale, void* stream) {
  TAVSR_REQUIRE(B > 0 && T > 0 && H > 0, "attn: bad shape B=%d T=%d H=%d", B, T, H);
  TAVSR_REQUIRE(qkv && pos && u && v && ctx && drop_keep, "attn: null pointer");
  TAVSR_REQUIRE(ld_qkv % 4 == 0 && ld_pos % 4 == 0 && ld_ctx % 4 == 0,
                "attn: pitches must be multiples of 4 elements (16 bytes)");
  TAVSR_REQUIRE(ld_drop % 128 == 0 && ld_drop >= T && (reinterpret_cast<uintptr_t>(drop_keep) & 15) == 0,
                "attn: the keep mask needs a 16-byte aligned base and a row pitch that is a multiple "
                "of 128 bytes >= T (got %lld)", ld_drop);
  TAVSR_REQUIRE(drop_scale >= 1.0f, "attn: drop_scale = 1 / (1 - p) must be >= 1");
  return relpos_attn_launch<false>(qkv, ld_qkv, pos, ld_pos, u, v, lens, ctx, ld_ctx, B, T, H, 0, lse,
                                   static_cast<cudaStream_t>(stream), drop_keep, ld_drop, drop_scale);
}
