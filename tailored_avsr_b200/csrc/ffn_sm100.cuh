// Fused position-wise feed-forward block for sm_100a:
//
//   y = x + alpha * ( act(LN(x) W1^T + b1) W2^T + b2 )   [+ the row-complete LayerNorm epilogue]
//
// (reference: espnet PositionwiseFeedForward called at src/encoder/branchformer/encoder_layer.py
//  :193-194 and :313-314, followed by the norms at :202,216,316).
//
// The 2048-wide hidden activation never leaves the SM.  Work decomposition:
//   * one 128-row tile of frames per 2-CTA cluster; CTA r of the cluster owns hidden units
//     [r*1024, (r+1)*1024) and produces a PARTIAL 128x256 output;
//   * per CTA the hidden half is processed in 8 chunks of 128 units:
//       GEMM1_j : H_j[128x128] = Xn[128x256] . W1[chunk j]^T       (A, B from smem; D in TMEM)
//       act_j   : 8 warps read H_j from TMEM, add b1, apply the activation and write the result back
//                 IN PLACE as the next GEMM's A operand (tcgen05.ld / tcgen05.st; TMEM lane == frame):
//                 TF32 values in tf32 mode, PACKED bf16 pairs in bf16 mode
//       GEMM2_j : D2[128x256] += H_j . W2[:, chunk j]^T             (A from TMEM, B from smem)
//     issue order GEMM1_0, GEMM1_1, GEMM2_0, GEMM1_2, GEMM2_1, ... so act_{j+1} overlaps GEMM2_j;
//   * Xn stays resident in smem (128 KB fp32 / 64 KB bf16), W1/W2 stream through a TMA ring of 32 KB
//     super-slots (3 in tf32 mode, 5 in bf16 mode);
//   * TMEM: D2 = columns [0,256), H double buffer = [256,384) and [384,512);
//   * the two partial outputs are exchanged over distributed shared memory (each CTA finishes 64
//     of the 128 rows): all eight warps then finish rows warp-per-row (residual, LayerNorms,
//     coalesced stores).
//
// bf16 mode (kind::f16): a 128-byte swizzle row holds 64 K elements instead of 32 and one MMA
// consumes K = 16, so the MMA count and the shared-memory operand bytes halve; the activation packs
// two bf16 per TMEM column (the A-from-TMEM layout pinned by tools/umma_probe.cu).
#pragma once
#include "gemm_sm100.cuh"

namespace tavsr {

struct alignas(64) FfnParams {
  CUtensorMap tmX;   // LN(x):  (256 inner, M rows),     box {128 B, 128}
  CUtensorMap tmW1;  // W1:     (256 inner, 2048 rows),  box {128 B, 128}
  CUtensorMap tmW2;  // W2:     (2048 inner, 256 rows),  box {128 B, 128}
  GemmParams ep;     // epilogue description: bias = b2, residual, alpha, ln0/lnA/lnB, raw outputs
  const float* b1;   // [2048]
  int act;
  long long* dbg;    // optional [gridDim.x][8] phase timestamps (globaltimer ns)
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FFN_STAMP(slot)                                                              \
  do {                                                                               \
    if (p.dbg != nullptr && threadIdx.x == 128)                                      \
      p.dbg[static_cast<long long>(blockIdx.x) * 8 + (slot)] = globaltimer_ns();     \
  } while (0)

__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// Warp roles (12 warps):
//   0: TMA producer            1: GEMM1 issuer (H_j = Xn . W1_j^T, 128x128 MMAs, A/B from smem)
//   2: GEMM2 issuer (D2 += H_j . W2_j^T, 128x256 MMAs, A from TMEM)       3: idle
//   4-11: activation / epilogue warps (TMEM lane quadrant = warp & 3, two warps per quadrant: the
//         activation sits between GEMM1_j and GEMM2_j on the tensor pipe's critical path, so its
//         latency is halved rather than its instruction count)
// Two issuing threads because one thread issues a tcgen05.mma at best every ~89 cycles (127 with a
// commit every 4; tools/mma_bench.cu) while a 128x128x8 TF32 MMA occupies the tensor pipe for
// ~70: the GEMM1 and GEMM2 streams are independent between handshakes, so issuing them from two
// warps keeps the pipe busy (measured 83 cycles per MMA for two issuers vs 127 for one).
// Weight ring: super-slots of 32 KB in one FIFO.  A W1 unit holds TWO k-blocks of one hidden chunk
// (8 MMAs per barrier wait + commit), a W2 unit one k-block of all 256 output rows (4 MMAs of
// N = 256).  The schedule
//   W1_0 W1_1 | W2_0 W1_2 | W2_1 W1_3 | ... | W2_6 | W2_7
// is walked identically by the producer and both issuers.
// ------------------------------------------------------------------------------------------------
template <bool kBf16>
struct FfnCfg {
  static constexpr int kD = 256;          // model width
  static constexpr int kHid = 2048;       // hidden width
  static constexpr int kHidCta = 1024;    // hidden units per CTA (cluster of 2)
  static constexpr int kChunk = 128;      // hidden units per chunk
  static constexpr int kNChunk = kHidCta / kChunk;
  static constexpr int kUnitBytes = 128 * 128;            // one 128-row k-block: 128 rows x 128 B
  static constexpr int kKB = kBf16 ? 64 : 32;             // elements per 128-byte k-block
  static constexpr int kXBlocks = kD / kKB;                // k-blocks of the resident LN(x) tile
  static constexpr int kXBytes = kXBlocks * kUnitBytes;
  static constexpr int kSuper = kBf16 ? 5 : 3;             // ring super-slots of 32 KB
  static constexpr int kW1Units = kXBlocks / 2;            // per hidden chunk
  static constexpr int kW2Units = kChunk / kKB;            // per hidden chunk
  static constexpr int kActWarps = 8;
  static constexpr int kThreads = 128 + 32 * kActWarps;
  static constexpr int kSmemBytes = 1024 + kXBytes + kSuper * 2 * kUnitBytes + 512;
  // epilogue scratch inside the (then idle) operand area: peer partial rows at [0, 64 KB) (the same
  // offset in both CTAs), own partial rows at [64 KB, 128 KB), parameter vectors at 160 KB
  static constexpr int kOwnOff = 64 * 1024;
  static constexpr int kParamOff = 160 * 1024;
  static_assert(kParamOff + 9 * 256 * 4 <= kXBytes + kSuper * 2 * kUnitBytes, "epilogue scratch");
  static constexpr uint32_t kColD2 = 0, kColH = 256;
};

namespace ffn {
constexpr int kHid = 2048;
template <int kSuper>
struct RingWalker {
  int slot = 0;
  uint32_t parity = 0;
  __device__ __forceinline__ uint32_t phase() const { return (parity >> slot) & 1u; }
  __device__ __forceinline__ void next() {
    parity ^= 1u << slot;
    if (++slot == kSuper) slot = 0;
  }
};
}  // namespace ffn

template <int kAct, bool kBf16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FfnCfg<kBf16>::kThreads, 1)
ffn_fused_kernel(const __grid_constant__ FfnParams p) {
  using C = FfnCfg<kBf16>;
  constexpr int kSuper = C::kSuper;
  constexpr int kUnitBytes = C::kUnitBytes;
  constexpr int kNChunk = C::kNChunk;
  constexpr int kChunk = C::kChunk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_x = smem;                       // resident LN(x) tile, later: peer partial rows (64 KB)
  uint8_t* s_ring = s_x + C::kXBytes;        // kSuper x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + kSuper * 2 * kUnitBytes);
  uint64_t* w_full = bars;                   // [kSuper]
  uint64_t* w_empty = bars + kSuper;         // [kSuper]
  uint64_t* w2_avail = bars + 2 * kSuper;    // [kSuper]  W2 unit landed (forwarded by the GEMM1 issuer)
  uint64_t* x_full = bars + 3 * kSuper;      // [1]
  uint64_t* h_full = x_full + 1;             // [2]  GEMM1 chunk complete
  uint64_t* h_ready = h_full + 2;            // [2]  activation written back
  uint64_t* h_free = h_ready + 2;            // [2]  GEMM2 finished reading the H buffer
  uint64_t* d_full = h_free + 2;             // [1]  all GEMM2 complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(d_full + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const int m0 = static_cast<int>(blockIdx.x >> 1) * 128;
  const int hid0 = static_cast<int>(rank) * C::kHidCta;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
    for (int i = 0; i < kSuper; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
      mbar_init(&w2_avail[i], 1);
    }
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&h_ready[i], C::kActWarps);
      mbar_init(&h_free[i], 1);
    }
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();
  FFN_STAMP(0);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, C::kXBytes);
      for (int kb = 0; kb < C::kXBlocks; ++kb)
        tma_load_2d(s_x + kb * kUnitBytes, &p.tmX, x_full, kb * C::kKB, m0);
      ffn::RingWalker<kSuper> rw;
      auto load_w1 = [&](int j) {  // units of 2 k-blocks of W1[hid0 + j*128 .. +128, :]
        for (int u = 0; u < C::kW1Units; ++u) {
          const int s = rw.slot;
          mbar_wait(&w_empty[s], rw.phase() ^ 1);
          mbar_arrive_expect_tx(&w_full[s], 2 * kUnitBytes);
          uint8_t* dst = s_ring + s * 2 * kUnitBytes;
          tma_load_2d(dst, &p.tmW1, &w_full[s], (2 * u) * C::kKB, hid0 + j * kChunk);
          tma_load_2d(dst + kUnitBytes, &p.tmW1, &w_full[s], (2 * u + 1) * C::kKB, hid0 + j * kChunk);
          rw.next();
        }
      };
      auto load_w2 = [&](int j) {  // units: one k-block of W2[:, hid0 + j*128 .. +128] each
        for (int kb = 0; kb < C::kW2Units; ++kb) {
          const int s = rw.slot;
          mbar_wait(&w_empty[s], rw.phase() ^ 1);
          mbar_arrive_expect_tx(&w_full[s], 2 * kUnitBytes);
          uint8_t* dst = s_ring + s * 2 * kUnitBytes;
          tma_load_2d(dst, &p.tmW2, &w_full[s], hid0 + j * kChunk + kb * C::kKB, 0);
          tma_load_2d(dst + kUnitBytes, &p.tmW2, &w_full[s], hid0 + j * kChunk + kb * C::kKB, 128);
          rw.next();
        }
      };
      load_w1(0);
      load_w1(1);
      for (int j = 0; j < kNChunk; ++j) {
        load_w2(j);
        if (j + 2 < kNChunk) load_w1(j + 2);
      }
    }
  } else if (warp == 1) {
    // ===================================== GEMM1 issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBf16 ? UMMA_FMT_BF16 : UMMA_FMT_TF32, 128, 128);
      ffn::RingWalker<kSuper> rw;
      const uint32_t x_addr = smem_u32(s_x);
      mbar_wait(x_full, 0);
      tc_fence_after_sync();
      auto gemm1 = [&](int j) {
        if (j >= 2) {  // H buffer j&1 was read by GEMM2 of chunk j-2
          mbar_wait(&h_free[j & 1], ((j - 2) >> 1) & 1);
          tc_fence_after_sync();
        }
        const uint32_t d = tmem_base + C::kColH + (j & 1) * kChunk;
        for (int u = 0; u < C::kW1Units; ++u) {
          const int s = rw.slot;
          mbar_wait(&w_full[s], rw.phase());
          tc_fence_after_sync();
          const uint32_t b_addr = smem_u32(s_ring + s * 2 * kUnitBytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kb = 2 * u + h;
            const uint64_t a_desc = umma_desc_kmajor_sw128(x_addr + kb * kUnitBytes);
            const uint64_t b_desc = umma_desc_kmajor_sw128(b_addr + h * kUnitBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<!kBf16>(d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&w_empty[s]);
          rw.next();
        }
        umma_commit(&h_full[j & 1]);
      };
      // This thread is the ONLY waiter on w_full: it observes every unit in FIFO order (an mbarrier
      // parity wait is only sound for a waiter that has seen every earlier phase) and forwards the
      // arrival of W2 units to the GEMM2 issuer through w2_avail.
      auto forward_w2 = [&]() {
        for (int kb = 0; kb < C::kW2Units; ++kb) {
          const int s = rw.slot;
          mbar_wait(&w_full[s], rw.phase());
          mbar_arrive(&w2_avail[s]);
          rw.next();
        }
      };
      gemm1(0);
      gemm1(1);
      for (int j = 0; j < kNChunk; ++j) {
        forward_w2();
        if (j + 2 < kNChunk) gemm1(j + 2);
      }
    }
  } else if (warp == 2) {
    // ===================================== GEMM2 issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBf16 ? UMMA_FMT_BF16 : UMMA_FMT_TF32, 128, 256);
      // slot sequence of the shared FIFO, but phase bits only for this thread's own barrier
      // (w2_avail): it waits on every phase of it, in order
      int slot = (2 * C::kW1Units) % kSuper;   // after the two leading W1 runs
      uint32_t parity2 = 0;
      for (int j = 0; j < kNChunk; ++j) {
        mbar_wait(&h_ready[j & 1], (j >> 1) & 1);  // activation of chunk j is back in TMEM
        tc_fence_after_sync();
        const uint32_t a0 = tmem_base + C::kColH + (j & 1) * kChunk;
        for (int kb = 0; kb < C::kW2Units; ++kb) {
          const int s = slot;
          mbar_wait(&w2_avail[s], (parity2 >> s) & 1u);
          tc_fence_after_sync();
          const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(s_ring + s * 2 * kUnitBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if constexpr (kBf16) {
              // packed pairs: hidden units [64 kb, 64 kb + 64) of the chunk sit in columns
              // [64 kb, 64 kb + 32) (written by the activation warps of column half kb)
              umma_ts_f16(tmem_base + C::kColD2, a0 + kb * 64 + k * 8, b_desc + 2 * k, idesc,
                          (j | kb | k) ? 1u : 0u);
            } else {
              umma_ts_tf32(tmem_base + C::kColD2, a0 + kb * 32 + k * 8, b_desc + 2 * k, idesc,
                           (j | kb | k) ? 1u : 0u);
            }
          }
          umma_commit(&w_empty[s]);
          parity2 ^= 1u << s;
          if (++slot == kSuper) slot = 0;
        }
        umma_commit(&h_free[j & 1]);
        if (j + 2 < kNChunk) slot = (slot + C::kW1Units) % kSuper;  // the W1_{j+2} run
      }
      umma_commit(d_full);
    }
  } else if (warp >= 4) {
    // =============================== activation warps =======================================
    // Quadrant q = warp & 3; the two warps of a quadrant split the chunk's four 32-column groups.
    const int q = warp & 3;  // TMEM lane quadrant
    const int part = (warp - 4) >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    for (int j = 0; j < kNChunk; ++j) {
      mbar_wait(&h_full[j & 1], (j >> 1) & 1);
      tc_fence_after_sync();
      if (j == 0) FFN_STAMP(1);
      if (j == 1) FFN_STAMP(6);
      const uint32_t th = tmem_base + lane_off + C::kColH + (j & 1) * kChunk;
      const float4* b1p = reinterpret_cast<const float4*>(p.b1 + hid0 + j * kChunk);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = part * 2 + cc;
        uint32_t r[32];
        tmem_ld32(th + c * 32, r);
        float bb[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // uniform 16-byte loads (every lane reads the same bias)
          const float4 f = p.b1 ? __ldg(b1p + c * 8 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          bb[4 * i] = f.x; bb[4 * i + 1] = f.y; bb[4 * i + 2] = f.z; bb[4 * i + 3] = f.w;
        }
        tmem_ld_wait();
        if constexpr (kBf16) {
          // group c (32 hidden units) -> 16 packed columns at [64 part + 16 cc, +16): inside the
          // column range this warp alone reads, behind the group just consumed
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0 = __uint_as_float(r[2 * i]) + bb[2 * i];
            float x1 = __uint_as_float(r[2 * i + 1]) + bb[2 * i + 1];
            if constexpr (kAct == ACT_SWISH) {   // one SFU op per value: the result is bf16
              x0 = swish_tanh(x0);
              x1 = swish_tanh(x1);
            } else {
              x0 = apply_act<kAct>(x0, p.act);
              x1 = apply_act<kAct>(x1, p.act);
            }
            w[i] = pack_bf16x2(x0, x1);
          }
          tmem_st16(th + part * 64 + cc * 16, w);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = apply_act<kAct>(__uint_as_float(r[i]) + bb[i], p.act);
            r[i] = __float_as_uint(round_tf32(x));
          }
          tmem_st32(th + c * 32, r);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[j & 1]);
    }
    mbar_wait(d_full, 0);  // every MMA of this CTA has retired: D2 final, s_x and s_ring free
    tc_fence_after_sync();
    FFN_STAMP(2);
  }

  // ---- epilogue parameters into the (now free) operand area; exchange partial outputs ----
  __syncthreads();     // reconverge the single-lane role loops before the aligned cluster barrier
  float* s_param = reinterpret_cast<float*>(smem + C::kParamOff);
  {
    const float* srcs[7] = {p.ep.bias, p.ep.ln0_g, p.ep.ln0_b, p.ep.lnA_g, p.ep.lnA_b,
                            p.ep.lnB_g, p.ep.lnB_b};
    for (int v = 0; v < 7; ++v)
      for (int i = threadIdx.x; i < 256; i += C::kThreads)
        s_param[v * 256 + i] = srcs[v] ? srcs[v][i] : 0.0f;
  }
  cluster_sync_all();  // both CTAs' MMAs are done -> both operand areas may be overwritten
  FFN_STAMP(3);
  // ---- warp-per-row epilogue: the partial accumulators leave TMEM once - rows this CTA finishes
  // into its own smem, the other 64 rows into the peer's - and then all eight warps finish rows with
  // coalesced global accesses and shuffle reductions (lane = 8 consecutive columns).
  uint8_t* s_own = smem + C::kOwnOff;  // [64 rows][1 KB], 16-byte chunks XOR-swizzled by row & 7
  if (warp >= 4) {
    // all eight activation warps: quadrant q, column half (warp - 4) >> 2
    const int q = warp & 3;
    const int chalf = (warp - 4) >> 2;
    const uint32_t td2 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + C::kColD2;
    const int row_local = (q & 1) * 32 + static_cast<int>(lane);
    const bool mine = static_cast<uint32_t>(q >> 1) == rank;
    const uint32_t base = mine ? smem_u32(s_own) + row_local * 1024
                               : mapa_cluster(smem_u32(s_x) + row_local * 1024, rank ^ 1u);
    for (int c = 4 * chalf; c < 4 * chalf + 4; ++c) {
      uint32_t r[32];
      tmem_ld32(td2 + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t dst = base + (((c * 8 + i) ^ (row_local & 7)) << 4);
        const float4 f = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                     __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        if (mine)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f.x), "f"(f.y),
                       "f"(f.z), "f"(f.w) : "memory");
        else
          st_cluster_v4(dst, f);
      }
    }
  }
  cluster_sync_all();  // release/acquire: own rows in s_own, the peer's partials of them in s_x
  FFN_STAMP(4);
  const GemmParams& e = p.ep;
  const bool has_ln0 = e.ln0_g != nullptr, has_lnA = e.lnA_g != nullptr, has_lnB = e.lnB_g != nullptr;
  const int col = 8 * static_cast<int>(lane);
  auto ld8 = [&](const float* src, float (&dst)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
    dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
  };
  float bias[8];
  ld8(s_param + col, bias);
  // the residual row of the NEXT iteration is requested before the current row is processed: its
  // L2 latency (the longest single wait of a row) hides behind the reductions and stores
  float4 pre0 = make_float4(0.f, 0.f, 0.f, 0.f), pre1 = pre0;
  auto fetch_residual = [&](int r) {
    const int mm = m0 + static_cast<int>(rank) * 64 + r;
    if (e.residual != nullptr && r < 64 && mm < e.M) {
      const float4* rp = reinterpret_cast<const float4*>(e.residual + static_cast<long long>(mm) * e.ldr + col);
      pre0 = ld_act4(rp);
      pre1 = ld_act4(rp + 1);
    }
  };
  fetch_residual(warp);
  for (int r = warp; r < 64; r += C::kThreads / 32) {
    const int m = m0 + static_cast<int>(rank) * 64 + r;
    if (m >= e.M) break;  // warp-uniform; rows only grow
    float v[8];
    {
      const uint32_t sw0 = static_cast<uint32_t>(((2 * lane) ^ (r & 7)) << 4);
      const uint32_t sw1 = static_cast<uint32_t>(((2 * lane + 1) ^ (r & 7)) << 4);
      const float4 o0 = *reinterpret_cast<const float4*>(s_own + r * 1024 + sw0);
      const float4 o1 = *reinterpret_cast<const float4*>(s_own + r * 1024 + sw1);
      const float4 q0 = *reinterpret_cast<const float4*>(s_x + r * 1024 + sw0);
      const float4 q1 = *reinterpret_cast<const float4*>(s_x + r * 1024 + sw1);
      const float4 r0 = pre0, r1 = pre1;
      fetch_residual(r + C::kThreads / 32);
      v[0] = r0.x + e.alpha * (o0.x + q0.x + bias[0]);
      v[1] = r0.y + e.alpha * (o0.y + q0.y + bias[1]);
      v[2] = r0.z + e.alpha * (o0.z + q0.z + bias[2]);
      v[3] = r0.w + e.alpha * (o0.w + q0.w + bias[3]);
      v[4] = r1.x + e.alpha * (o1.x + q1.x + bias[4]);
      v[5] = r1.y + e.alpha * (o1.y + q1.y + bias[5]);
      v[6] = r1.z + e.alpha * (o1.z + q1.z + bias[6]);
      v[7] = r1.w + e.alpha * (o1.w + q1.w + bias[7]);
    }
    auto stats = [&](float eps, float& mean, float& rstd) {  // two-pass, like torch's LayerNorm
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
      mean = warp_sum(s) * (1.0f / 256.0f);
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dk = v[k] - mean;
        ss = fmaf(dk, dk, ss);
      }
      rstd = rsqrtf(warp_sum(ss) * (1.0f / 256.0f) + eps);
    };
    auto emit = [&](void* out, long long ld, const float* g, const float* b, float mean, float rstd,
                    bool affine, bool rnd, bool as_bf16) {
      float y[8];
      if (affine) {
        float gg[8], bb[8];
        ld8(g + col, gg);
        ld8(b + col, bb);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = (v[k] - mean) * rstd * gg[k] + bb[k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = v[k];
      }
      if (kBf16 && as_bf16) {
        uint4 w;
        w.x = pack_bf16x2(y[0], y[1]); w.y = pack_bf16x2(y[2], y[3]);
        w.z = pack_bf16x2(y[4], y[5]); w.w = pack_bf16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(static_cast<uint16_t*>(out) + static_cast<long long>(m) * ld + col) = w;
        return;
      }
      if (rnd) {
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = round_tf32(y[k]);
      }
      float4* op = reinterpret_cast<float4*>(static_cast<float*>(out) + static_cast<long long>(m) * ld + col);
      op[0] = make_float4(y[0], y[1], y[2], y[3]);
      op[1] = make_float4(y[4], y[5], y[6], y[7]);
    };
    float mean = 0.f, rstd = 1.f;
    if (has_ln0) {  // v1 = LN0(v0) replaces v0 (norm_final)
      stats(e.eps0, mean, rstd);
      float gg[8], bb[8];
      ld8(s_param + 256 + col, gg);
      ld8(s_param + 512 + col, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * gg[k] + bb[k];
    }
    if (e.has_main) emit(e.out_main, e.ld_main, nullptr, nullptr, 0.f, 1.f, false, e.round_c != 0, e.c_bf16 != 0);
    if (has_lnA || has_lnB) {
      stats(e.eps, mean, rstd);
      if (has_lnA) emit(e.out_lnA, e.ld_lnA, s_param + 768, s_param + 1024, mean, rstd, true, e.round_lnA != 0, e.lnA_bf16 != 0);
      if (has_lnB) emit(e.out_lnB, e.ld_lnB, s_param + 1280, s_param + 1536, mean, rstd, true, e.round_lnB != 0, e.lnB_bf16 != 0);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  FFN_STAMP(5);
  tc_fence_after_sync();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tavsr
