// tcgen05 / TMEM / TMA GEMM family for sm_100a.
//
//   Y[M,N] = epilogue( X[M,K] . W[N,K]^T )          (both operands K-major == torch nn.Linear)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulators live in TMEM)
//   warps 2..   : epilogue      (tcgen05.ld -> registers -> fused math -> swizzled smem -> TMA store)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of
// tile i+1.
//
// Two epilogue families:
//   kModeTiled : out = act(acc + bias), arbitrary N (tiled by kBlockN), 8 epilogue warps.
//   kModeRowLN : kBlockN == N == 256, one thread owns one output row (TMEM lane == row), so
//                LayerNorm statistics, chained LayerNorms, residual adds and row dot-products are
//                thread-local.  Optionally two A operands / two accumulators combined with
//                per-utterance scalars (the Branchformer learned_ave merge,
//                reference src/encoder/branchformer/encoder_layer.py:291-293).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx.cuh"

namespace tavsr {

enum : int { ACT_NONE = 0, ACT_SWISH = 1, ACT_GELU = 2, ACT_RELU = 3 };
enum : int { kModeTiled = 0, kModeRowLN = 1 };

struct alignas(64) GemmParams {
  CUtensorMap tmA;    // activations  (K inner, M rows), box {128 B, 128}
  CUtensorMap tmA2;   // second activation operand (dual mode)
  CUtensorMap tmB;    // weights      (K inner, N rows), box {128 B, kBlockN}
  CUtensorMap tmC;    // main output  (N inner, M rows), box {128 B, 32}
  CUtensorMap tmLnA;  // LayerNorm output A
  CUtensorMap tmLnB;  // LayerNorm output B
  int M, N, K;
  int num_m_tiles, num_n_tiles;
  // ---- tiled split-K (weight gradients: few output tiles, K = all frames): ksplit > 1 multiplies
  // the unit count; split s reduces k-blocks [s * split_kb, ...) and stores its partial tile into
  // slab s of the output map, i.e. at row s * slab_rows + m (slab_rows a multiple of the unit's
  // rows); the slabs are summed by sum_slabs_kernel.  Bias is added by split 0 only. ----
  int ksplit, split_kb, slab_rows;
  const float* bias;
  int act;
  int round_c;  // round main output to tf32 (output only feeds tensor-core GEMMs)
  // ---- RowLN mode ----
  const float* residual;  // [M, ldr] fp32 or null
  long long ldr;
  float alpha;                // v0 = residual + alpha * (combine(acc) + bias)
  const float* rowscale1;     // dual: per-segment scalars, segment = row / rows_per_seg
  const float* rowscale2;
  int rows_per_seg;
  const float* ln0_g;  // optional first LayerNorm applied to v0 (-> v1); main output is v1
  const float* ln0_b;
  const float* lnA_g;  // optional LayerNorm outputs of v1
  const float* lnA_b;
  const float* lnB_g;
  const float* lnB_b;
  int has_main, round_lnA, round_lnB;
  const float* dot1;  // optional row dot-products of v1 with two 256-vectors
  const float* dot2;
  float* dots_out;    // [M,2]
  float eps;
  float eps0;  // eps of ln0
  // ---- RowLN split-K (num_n_tiles == 2): unit parity selects the K half; the odd unit dumps its
  // accumulator to `partial`, the even unit adds it in its epilogue ----
  float* partial;          // [M_padded, 256] fp32 scratch
  unsigned int* flags;     // [num_m_tiles * kCtas * 4], zero between launches
  // ---- sequential dual (kDual with seq_kb1 > 0): the reduction axis is the concatenation of the
  // two operands, A = [A1 | A2], W = [W1 | W2]; k-blocks < seq_kb1 accumulate A1.W1^T into the
  // first accumulator, the rest A2.W2^T into the second (merge with the branch output projections
  // folded in).  seg_bias: dot1 / dot2 hold per-branch bias vectors scaled like the accumulators.
  int seq_kb1;
  int seg_bias;
  // ---- raw output pointers (the fused FFN's warp-per-row epilogue stores directly, coalesced) ----
  void* out_main;
  long long ld_main;
  void* out_lnA;
  long long ld_lnA;
  void* out_lnB;
  long long ld_lnB;
  // ---- bf16 mode (kTf32 == false): which outputs are stored as bf16 (tensor-core operands of the
  // next kernel) instead of fp32 (residual stream, encoder output) ----
  int c_bf16, lnA_bf16, lnB_bf16;
  // ---- grouped tiled mode (kAct2 != kNoGroup): a SECOND problem Y2 = act2(X2 . W2^T + b2) with the
  // same M and K shares the launch; per row block the n tiles of problem 1 come first, then the
  // num_n_tiles2 tiles of problem 2 (tmA2 = X2) ----
  CUtensorMap tmB2;
  CUtensorMap tmC2;
  const float* bias2;
  int N2, num_n_tiles2;
};

constexpr int kNoGroup = -2;

// kWideEpi (tiled mode): 16 epilogue warps (four per TMEM lane quadrant, a quarter of the tile's
// columns each) instead of 8.  The K = 256 projections are bound by their epilogue, and ncu shows it
// latency-bound rather than issue-bound (8 epilogue warps: 39 % issue-active, stalls on the first
// use of each TMEM / bias load and on fixed-latency dependencies), so more warps per scheduler.
template <bool kTf32, int kBlockN, int kMode, bool kDual, int kCtas, bool kSeq = false,
          bool kWideEpi = false>
struct GemmCfg {
  static constexpr int kBlockM = 128;               // rows per CTA (a CTA pair covers 256)
  static constexpr int kElemBytes = kTf32 ? 4 : 2;
  static constexpr int kBlockK = 128 / kElemBytes;  // one 128B swizzle row
  static constexpr int kUmmaK = 32 / kElemBytes;
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBRows = kBlockN / kCtas;    // B rows held by this CTA (half in a pair)
  static constexpr int kBBytes = kBRows * 128;
  // a stage holds two A tiles only in the parallel dual mode; the sequential dual mode (kSeq)
  // streams ONE A tile per k-block, so its stages are as small - and its ring as deep - as the
  // single-operand kernel's
  static constexpr int kATiles = (kDual && !kSeq) ? 2 : 1;
  static constexpr int kStageBytes = kABytes * kATiles + kBBytes;
  static constexpr int kAccCols = kBlockN * (kDual ? 2 : 1);
  static constexpr int kAccStages = (512 / kAccCols) >= 2 ? 2 : 1;
  static constexpr int kEpiWarps = kMode == kModeTiled ? (kWideEpi ? 16 : 8) : 4;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kStageBufs = kWideEpi ? 1 : 2;  // 16 warps: single-buffered store staging
  static_assert(kMode == kModeTiled || !kWideEpi, "the wide epilogue is a tiled-mode variant");
  static constexpr int kStagingBytes = kEpiWarps * kStageBufs * 4096;
  static constexpr int kEpiParts = kEpiWarps / 4;      // column slices of a tile (one per warp of a quadrant)
  // params staged in smem: bias[2][kBlockN] (tiled) or 9 x 256 floats (rowln)
  static constexpr int kParamFloats = kMode == kModeTiled ? 2 * kBlockN : 9 * 256;
  static constexpr int kFixedBytes = 1024 /*align slack*/ + kStagingBytes + kParamFloats * 4 + 256;
  static constexpr int kStagesFit = (227 * 1024 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
};

// Swish with ONE special-function op: x sigmoid(x) = h + h tanh(h), h = x / 2 (tanh.approx.f32,
// relative error ~2^-11).  The exact form below costs two (ex2 + rcp) and the SFU pipe (16 per
// clock per SM) is what bounds the fused FFN's activation stage: 128 x 128 values per chunk are
// 2048 cycles of SFU time, as much as the two GEMMs of the chunk.  Used where the result is rounded
// to bf16 (2^-9) anyway.
__device__ __forceinline__ float swish_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

template <int kAct>
__device__ __forceinline__ float apply_act(float x, int act_rt) {
  const int act = kAct >= 0 ? kAct : act_rt;
  switch (act) {
    case ACT_SWISH: return __fdividef(x, 1.0f + __expf(-x));
    case ACT_GELU: {
      // exact-erf GELU  x * Phi(x)  with the upper tail  Q(t) = 0.5 erfc(t / sqrt 2) = 2^R(t),
      // R a degree-6 minimax fit of log2 Q on [0, 9] weighted by t Q(t) (tools/fit_gelu.py):
      // |gelu - exact| <= 1.6e-7 max(1, |x|), i.e. the fp32 rounding level of the reference's erff,
      // in 12 issue slots (1 SFU) instead of ~30 - the epilogue warps are the bottleneck of the
      // channel_proj1 GEMM.  No cancellation in the negative tail: Phi(x) = Q(|x|) there.
      const float t = fminf(fabsf(x), 9.0f);
      float r = fmaf(2.904336725e-05f, t, -7.323304308e-04f);
      r = fmaf(r, t, 7.953807712e-03f);
      r = fmaf(r, t, -5.320513994e-02f);
      r = fmaf(r, t, -4.589348137e-01f);
      r = fmaf(r, t, -1.151144981e+00f);
      r = fmaf(r, t, -9.999991059e-01f);
      float qv;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(qv) : "f"(r));
      return x * (x >= 0.f ? 1.0f - qv : qv);
    }
    case ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}

// Write one 32-float row chunk into a per-warp staging box (32 rows x 128 B, 128B-swizzled so that
// it matches a TMA store with CU_TENSOR_MAP_SWIZZLE_128B) and issue the TMA store.
struct WarpStager {
  uint8_t* base;  // nbuf x 4096 B, 1024-aligned
  int buf;
  int nbuf = 2;
  __device__ __forceinline__ void store(const CUtensorMap* tm, const float (&v)[32], int col0,
                                        int row0) {
    const uint32_t lane = lane_id();
    if (lane == 0) {
      if (nbuf == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
    }
    __syncwarp();
    // explicit st.shared: the pointer travels through this struct, where the compiler may lose
    // the address space and fall back to generic stores
    const uint32_t dst = smem_u32(base) + buf * 4096 + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ (lane & 7)) << 4)),
                   "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                   : "memory");
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm, base + buf * 4096, col0, row0);
      tma_store_commit();
    }
    if (nbuf == 2) buf ^= 1;
  }
  // ---- explicit-buffer forms (bf16 mode) ----
  // fp32 box into buffer b
  __device__ __forceinline__ void store_at(int b, const CUtensorMap* tm, const float (&v)[32],
                                           int col0, int row0) {
    const uint32_t lane = lane_id();
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    const uint32_t dst = smem_u32(base) + b * 4096 + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ (lane & 7)) << 4)),
                   "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                   : "memory");
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm, base + b * 4096, col0, row0);
      tma_store_commit();
    }
  }
  // A bf16 box is 32 rows x 64 columns (128 B rows): 32 fp32 values per lane fill HALF a row
  // (16-byte chunks 4*half .. 4*half+3, XOR-swizzled like the TMA map expects).  `pending` = how
  // many earlier bulk groups may still be reading OTHER buffers when half 0 claims buffer b.
  template <int kPending>
  __device__ __forceinline__ void put_bf16(int b, const float (&v)[32], int half) {
    const uint32_t lane = lane_id();
    if (half == 0) {
      if (lane == 0) tma_store_wait_read<kPending>();
      __syncwarp();
    }
    const uint32_t dst = smem_u32(base) + b * 4096 + lane * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                       dst + (((4 * half + j) ^ (lane & 7)) << 4)),
                   "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                   "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                   : "memory");
  }
  __device__ __forceinline__ void flush_bf16(int b, const CUtensorMap* tm, int col0, int row0) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane_id() == 0) {
      tma_store_2d(tm, base + b * 4096, col0, row0);
      tma_store_commit();
    }
  }
  __device__ __forceinline__ void drain() {
    if (lane_id() == 0) tma_store_wait_all<0>();
    __syncwarp();
  }
};

// One 32-column chunk `c` (0..7) of an N = 256 row-complete output: fp32 boxes go out per chunk,
// bf16 boxes per chunk PAIR (64 columns).  `b` = staging buffer (0 / 1) reserved for this output.
__device__ __forceinline__ void rowln_emit(WarpStager& stager, int b, const CUtensorMap* tm,
                                           bool as_bf16, const float (&v)[32], int c, int row0) {
  if (as_bf16) {
    stager.put_bf16<0>(b, v, c & 1);
    if (c & 1) stager.flush_bf16(b, tm, (c - 1) * 32, row0);
  } else {
    stager.store_at(b, tm, v, c * 32, row0);
  }
}

// ------------------------------------------------------------------------------------------------
// Row-complete epilogue math, shared by the RowLN GEMM and the fused FFN kernel.  The calling warp
// owns TMEM lanes [lane_off, +32) == output rows row0 .. row0+31 (one row per thread) and columns
// [tacc, tacc+256) hold the finished fp32 accumulator (tacc2: second accumulator in dual mode).
//   v0 = residual + alpha * (w1*acc + w2*acc2 + partial + bias);  v1 = LN0(v0) (optional)
//   main = v1;  lnA / lnB = LayerNorm(v1);  dots = (v1.dot1, v1.dot2)
// Values round-trip TMEM between passes (tcgen05.st), statistics are two-pass (mean, then centred
// second moment) like torch's LayerNorm.
// ------------------------------------------------------------------------------------------------
template <bool kDual, bool kBf16 = false>
__device__ __forceinline__ void rowln_finish(const GemmParams& p, const float* s_param,
                                             uint32_t tacc, int row0, uint32_t lane,
                                             WarpStager& stager, const float* part_row) {
  const float* s_bias = s_param;
  const float* s_g0 = s_param + 256;
  const float* s_b0 = s_param + 512;
  const float* s_gA = s_param + 768;
  const float* s_bA = s_param + 1024;
  const float* s_gB = s_param + 1280;
  const float* s_bB = s_param + 1536;
  const float* s_d1 = s_param + 1792;
  const float* s_d2 = s_param + 2048;
  const int m = row0 + static_cast<int>(lane);
  const bool valid = m < p.M;
  const bool has_ln0 = p.ln0_g != nullptr;
  const bool has_lnA = p.lnA_g != nullptr;
  const bool has_lnB = p.lnB_g != nullptr;
  const bool any_ln = has_ln0 || has_lnA || has_lnB;
  const bool has_dots = p.dots_out != nullptr;
  const bool splitk = part_row != nullptr;
  float w1 = 1.0f, w2 = 0.0f;
  if (kDual) {
    const int seg = (valid ? m : p.M - 1) / p.rows_per_seg;
    w1 = p.rowscale1[seg];
    w2 = p.rowscale2[seg];
  }
    // PASS A: v0 = residual + alpha * (combine(acc) + bias)
    // The residual row is read straight from global memory in this thread-per-row layout (each
    // lane a different 1 KB row): the loads of chunk c+1 are issued before chunk c is processed
    // so their L2 latency overlaps the TMEM traffic and the math instead of being exposed 8 times.
    float sum = 0.0f, dd1 = 0.0f, dd2 = 0.0f;
    const bool has_res = p.residual != nullptr && valid;
    const float4* rp0 = reinterpret_cast<const float4*>(
        p.residual + (has_res ? static_cast<long long>(m) * p.ldr : 0));
    float resn[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 f = has_res ? ld_act4(rp0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      resn[4 * j] = f.x; resn[4 * j + 1] = f.y; resn[4 * j + 2] = f.z; resn[4 * j + 3] = f.w;
    }
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
      uint32_t r2[kDual ? 32 : 1];
      tmem_ld32(tacc + c * 32, r);
      if constexpr (kDual) tmem_ld32(tacc + 256 + c * 32, r2);  // both loads in flight, one wait
      float res[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) res[j] = resn[j];
      if (c + 1 < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 f = has_res ? ld_act4(rp0 + (c + 1) * 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          resn[4 * j] = f.x; resn[4 * j + 1] = f.y; resn[4 * j + 2] = f.z; resn[4 * j + 3] = f.w;
        }
      }
      tmem_ld_wait();
      float v[32];
      if constexpr (kDual) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = w1 * __uint_as_float(r[j]) + w2 * __uint_as_float(r2[j]);
        if (p.seg_bias) {  // per-branch biases, scaled like the branches (dot slots reused)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w1 * s_d1[c * 32 + j] + w2 * s_d2[c * 32 + j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      }
      if (splitk) {
        const float4* pp = reinterpret_cast<const float4*>(part_row + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 f = __ldcg(pp + j);
          v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = res[j] + p.alpha * (v[j] + s_bias[c * 32 + j]);
        sum += v[j];
      }
      if (any_ln) {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v[j]);
        tmem_st32(tacc + c * 32, w);
      }
      if (!has_ln0) {
        if (has_dots) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            dd1 += v[j] * s_d1[c * 32 + j];
            dd2 += v[j] * s_d2[c * 32 + j];
          }
        }
        if (p.has_main) {
          if (p.round_c) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
          }
          if constexpr (kBf16) rowln_emit(stager, 0, &p.tmC, p.c_bf16 != 0, v, c, row0);
          else stager.store(&p.tmC, v, c * 32, row0);
        }
      }
    }
    if (any_ln) {
      tmem_st_wait();
      float mean = sum * (1.0f / 256.0f);
      // PASS B: centred second moment of v0
      float ss = 0.0f;
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld32(tacc + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(r[j]) - mean;
          ss += d * d;
        }
      }
      float rstd = rsqrtf(ss * (1.0f / 256.0f) + (has_ln0 ? p.eps0 : p.eps));
      if (has_ln0) {
        // PASS C: v1 = LN0(v0) -> TMEM, main output, dots, new sum
        float sum1 = 0.0f;
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld32(tacc + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = (__uint_as_float(r[j]) - mean) * rstd * s_g0[c * 32 + j] + s_b0[c * 32 + j];
            sum1 += v[j];
            r[j] = __float_as_uint(v[j]);
          }
          tmem_st32(tacc + c * 32, r);
          if (has_dots) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              dd1 += v[j] * s_d1[c * 32 + j];
              dd2 += v[j] * s_d2[c * 32 + j];
            }
          }
          if (p.has_main) {
            if (p.round_c) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
            }
            if constexpr (kBf16) rowln_emit(stager, 0, &p.tmC, p.c_bf16 != 0, v, c, row0);
          else stager.store(&p.tmC, v, c * 32, row0);
          }
        }
        tmem_st_wait();
        mean = sum1 * (1.0f / 256.0f);
        // PASS D: centred second moment of v1
        ss = 0.0f;
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld32(tacc + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(r[j]) - mean;
            ss += d * d;
          }
        }
        rstd = rsqrtf(ss * (1.0f / 256.0f) + p.eps);
      }
      // PASS E: LayerNorm outputs of v1
      if (has_lnA || has_lnB) {
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld32(tacc + c * 32, r);
          tmem_ld_wait();
          float v[32];
          if (has_lnA) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float y =
                  (__uint_as_float(r[j]) - mean) * rstd * s_gA[c * 32 + j] + s_bA[c * 32 + j];
              v[j] = p.round_lnA ? round_tf32(y) : y;
            }
            if constexpr (kBf16) rowln_emit(stager, 0, &p.tmLnA, p.lnA_bf16 != 0, v, c, row0);
            else stager.store(&p.tmLnA, v, c * 32, row0);
          }
          if (has_lnB) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float y =
                  (__uint_as_float(r[j]) - mean) * rstd * s_gB[c * 32 + j] + s_bB[c * 32 + j];
              v[j] = p.round_lnB ? round_tf32(y) : y;
            }
            if constexpr (kBf16) rowln_emit(stager, 1, &p.tmLnB, p.lnB_bf16 != 0, v, c, row0);
            else stager.store(&p.tmLnB, v, c * 32, row0);
          }
        }
      }
    }
    if (has_dots && valid) {
      reinterpret_cast<float2*>(p.dots_out)[m] = make_float2(dd1, dd2);
    }
}

template <bool kTf32, int kBlockN, int kMode, bool kDual, int kCtas, int kAct, bool kSeq = false,
          bool kWideEpi = false, int kAct2 = kNoGroup>
__global__ void __launch_bounds__(GemmCfg<kTf32, kBlockN, kMode, kDual, kCtas, kSeq, kWideEpi>::kThreads, 1)
gemm_sm100_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<kTf32, kBlockN, kMode, kDual, kCtas, kSeq, kWideEpi>;
  constexpr bool kGroup = kAct2 != kNoGroup;
  static_assert(!kGroup || (kMode == kModeTiled && !kDual), "grouping is a tiled-mode feature");
  static_assert(!kSeq || kDual, "the sequential mode is a dual mode");
  constexpr bool kPair = kCtas == 2;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_stage = smem;                                        // kStages * kStageBytes
  uint8_t* s_staging = s_stage + kStages * Cfg::kStageBytes;      // kEpiWarps * 8192
  float* s_param = reinterpret_cast<float*>(s_staging + Cfg::kStagingBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_param + Cfg::kParamFloats);
  uint64_t* full_bar = s_bar;                  // [kStages]
  uint64_t* empty_bar = s_bar + kStages;       // [kStages]
  uint64_t* tfull_bar = s_bar + 2 * kStages;   // [kAccStages]
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  // Work units: a unit is one CTA tile (kCtas == 1) or one CTA-pair tile of 256 rows (kCtas == 2).
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const int unit0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int n_per_m = p.num_n_tiles + (kGroup ? p.num_n_tiles2 : 0);   // tiles per row block
  const int tiles_mn = p.num_m_tiles * n_per_m;  // in units
  const bool ksplit_on = kMode == kModeTiled && !kGroup && p.ksplit > 1;
  const int num_tiles = ksplit_on ? tiles_mn * p.ksplit : tiles_mn;
  const int num_kb = (p.K + Cfg::kBlockK - 1) / Cfg::kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (kDual || kGroup) tma_prefetch_desc(&p.tmA2);
    if (kGroup) tma_prefetch_desc(&p.tmB2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], Cfg::kEpiWarps * kCtas);  // both CTAs' epilogues free the leader
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_2sm(s_tmem, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(s_tmem, 512);
      tmem_relinquish();
    }
  }
  if (kMode == kModeRowLN && warp >= 2) {
    // stage the per-column parameter vectors once (N == 256)
    const float* srcs[9] = {p.bias, p.ln0_g, p.ln0_b, p.lnA_g, p.lnA_b,
                            p.lnB_g, p.lnB_b, p.dot1, p.dot2};
    for (int v = 0; v < 9; ++v) {
      for (int i = threadIdx.x - 64; i < 256; i += 32 * Cfg::kEpiWarps)
        s_param[v * 256 + i] = srcs[v] ? srcs[v][i] : 0.0f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();  // everything above only touched weights / on-chip state

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int split = ksplit_on ? tile / tiles_mn : 0;
        const int t_mn = ksplit_on ? tile - split * tiles_mn : tile;
        const int m_blk = (t_mn / n_per_m) * kCtas + static_cast<int>(cta_rank);
        int n_blk = t_mn % n_per_m;
        const bool prob2 = kGroup && n_blk >= p.num_n_tiles;
        if (prob2) n_blk -= p.num_n_tiles;
        const CUtensorMap* tma_b = prob2 ? &p.tmB2 : &p.tmB;
        // RowLN: N is one tile, the unit's n index selects the K split instead
        const int b_row0 = (kMode == kModeRowLN ? 0 : n_blk * kBlockN) +
                           static_cast<int>(cta_rank) * Cfg::kBRows;
        int kb_cnt = kMode == kModeRowLN ? num_kb / p.num_n_tiles : num_kb;
        int kb_beg = kMode == kModeRowLN ? n_blk * kb_cnt : 0;
        if (ksplit_on) {
          kb_beg = split * p.split_kb;
          kb_cnt = num_kb - kb_beg < p.split_kb ? num_kb - kb_beg : p.split_kb;
        }
        for (int kb = kb_beg; kb < kb_beg + kb_cnt; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = s_stage + s * Cfg::kStageBytes;
          uint8_t* sb = st + Cfg::kABytes * Cfg::kATiles;
          constexpr bool seq = kSeq;   // one A tile per stage, picked by k-block
          const bool second = seq && kb >= p.seq_kb1;
          const CUtensorMap* tma_a = (second || prob2) ? &p.tmA2 : &p.tmA;
          const int ka = (second ? kb - p.seq_kb1 : kb) * Cfg::kBlockK;
          const uint32_t bytes = seq ? Cfg::kABytes + Cfg::kBBytes : Cfg::kStageBytes;
          if (kPair) {
            // the leader's barrier collects the bytes of both CTAs' loads
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], bytes * 2);
            tma_load_2d_2sm(st, tma_a, &full_bar[s], ka, m_blk * Cfg::kBlockM);
            if (kDual && !seq)
              tma_load_2d_2sm(st + Cfg::kABytes, &p.tmA2, &full_bar[s], kb * Cfg::kBlockK,
                              m_blk * Cfg::kBlockM);
            tma_load_2d_2sm(sb, tma_b, &full_bar[s], kb * Cfg::kBlockK, b_row0);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], bytes);
            tma_load_2d(st, tma_a, &full_bar[s], ka, m_blk * Cfg::kBlockM);
            if (kDual && !seq)
              tma_load_2d(st + Cfg::kABytes, &p.tmA2, &full_bar[s], kb * Cfg::kBlockK,
                          m_blk * Cfg::kBlockM);
            tma_load_2d(sb, tma_b, &full_bar[s], kb * Cfg::kBlockK, b_row0);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc(kTf32 ? UMMA_FMT_TF32 : UMMA_FMT_BF16,
                                            Cfg::kBlockM * kCtas, kBlockN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after_sync();
        const uint32_t d0 = tmem_base + as * Cfg::kAccCols;
        int kb_cnt = kMode == kModeRowLN ? num_kb / p.num_n_tiles : num_kb;
        if (ksplit_on) {
          const int kb_beg = (tile / tiles_mn) * p.split_kb;
          kb_cnt = num_kb - kb_beg < p.split_kb ? num_kb - kb_beg : p.split_kb;
        }
        for (int kb = 0; kb < kb_cnt; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(s_stage + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes * Cfg::kATiles;
          const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
          const uint64_t b_desc = umma_desc_kmajor_sw128(b_addr);
          const uint64_t a2_desc = umma_desc_kmajor_sw128(a_addr + Cfg::kABytes);
          constexpr bool seq = kSeq;
          const bool second = seq && kb >= p.seq_kb1;
          const uint32_t dd = second ? d0 + kBlockN : d0;
          const int kb_rel = second ? kb - p.seq_kb1 : kb;
#pragma unroll
          for (int k = 0; k < Cfg::kBlockK / Cfg::kUmmaK; ++k) {
            const uint32_t acc = (kb_rel | k) ? 1u : 0u;
            // advancing K inside the 128B swizzle row: +32 bytes == +2 in the (addr>>4) field
            if (kPair) {
              umma_ss_2sm<kTf32>(dd, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
              if (kDual && !seq)
                umma_ss_2sm<kTf32>(d0 + kBlockN, a2_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            } else {
              umma_ss<kTf32>(dd, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
              if (kDual && !seq)
                umma_ss<kTf32>(d0 + kBlockN, a2_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            }
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (kPair) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (kPair) umma_commit_2sm(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
        if (++as == kAccStages) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int ew = warp - 2;          // 0 .. kEpiWarps-1
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    WarpStager stager{s_staging + ew * Cfg::kStageBufs * 4096, 0, Cfg::kStageBufs};
    int as = 0;
    uint32_t aph = 0;
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
      const int split = ksplit_on ? tile / tiles_mn : 0;
      const int t_mn = ksplit_on ? tile - split * tiles_mn : tile;
      const int m_blk = (t_mn / n_per_m) * kCtas + static_cast<int>(cta_rank);
      int n_blk = t_mn % n_per_m;
      const bool prob2 = kGroup && n_blk >= p.num_n_tiles;   // warp-uniform
      if (prob2) n_blk -= p.num_n_tiles;
      const int m0 = m_blk * Cfg::kBlockM + (ksplit_on ? split * p.slab_rows : 0);
      const int n0 = n_blk * kBlockN;
      const float* e_bias = prob2 ? p.bias2 : (split == 0 ? p.bias : nullptr);
      const int e_N = prob2 ? p.N2 : p.N;
      const CUtensorMap* e_tmC = prob2 ? &p.tmC2 : &p.tmC;
      const uint32_t tacc = tmem_base + lane_off + as * Cfg::kAccCols;

      if constexpr (kMode == kModeTiled) {
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after_sync();
        const int half = ew >> 2;              // column slice of this warp (0 .. kEpiParts-1)
        constexpr int kChunks = kBlockN / 32 / Cfg::kEpiParts;  // 32-col chunks per warp
        const int col0 = half * (kBlockN / Cfg::kEpiParts);
        uint32_t r[2][32];
        tmem_ld32(tacc + col0, r[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col = col0 + c * 32;
          // warp-uniform; a bf16 box covers a 64-column pair, stored when its first half is live
          // (columns past N are clipped by the TMA store)
          const bool live = (!kTf32 && p.c_bf16) ? (n0 + (col & ~63) < e_N) : (n0 + col < e_N);
          // bias: uniform (same address in every lane) 16-byte loads served by L1
          float bv[32];
          if (e_bias != nullptr && n0 + col + 32 <= e_N) {
            const float4* bp = reinterpret_cast<const float4*>(e_bias + n0 + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 f = __ldg(bp + j);
              bv[4 * j] = f.x; bv[4 * j + 1] = f.y; bv[4 * j + 2] = f.z; bv[4 * j + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              bv[j] = (e_bias != nullptr && n0 + col + j < e_N) ? __ldg(e_bias + n0 + col + j) : 0.f;
          }
          tmem_ld_wait();
          if (c + 1 < kChunks) {
            tmem_ld32(tacc + col + 32, r[(c + 1) & 1]);  // overlaps the math of chunk c
          } else {
            // every accumulator value of this warp is in registers: release TMEM to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
              if (kPair) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
            }
          }
          if (live) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(r[c & 1][j]) + bv[j];
              if (kGroup && prob2) x = apply_act<kGroup ? kAct2 : kAct>(x, p.act);
              else x = apply_act<kAct>(x, p.act);
              v[j] = p.round_c ? round_tf32(x) : x;
            }
            if (!kTf32 && p.c_bf16) {
              static_assert(kTf32 || kChunks % 2 == 0, "bf16 boxes need an even number of chunks per warp");
              stager.template put_bf16<Cfg::kStageBufs - 1>(stager.buf, v, c & 1);
              if (c & 1) {
                stager.flush_bf16(stager.buf, e_tmC, n0 + col - 32, m0 + q * 32);
                if (Cfg::kStageBufs == 2) stager.buf ^= 1;
              }
            } else {
              stager.store(e_tmC, v, n0 + col, m0 + q * 32);
            }
          }
        }
      } else {
        // ---------------------------- row-complete epilogue ----------------------------------
        const int m = m0 + q * 32 + static_cast<int>(lane);
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after_sync();

        const bool splitk = p.num_n_tiles == 2;
        unsigned int* flag =
            splitk ? p.flags + ((static_cast<long long>(m_blk) * 4 + q)) : nullptr;
        if (splitk && n_blk == 1) {
          // ---- split-K writer: dump the raw partial accumulator, publish, next tile ----
          float* prow = p.partial + static_cast<long long>(m) * 256;
          for (int c = 0; c < 8; ++c) {
            uint32_t r[32];
            tmem_ld32(tacc + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(reinterpret_cast<float4*>(prow + c * 32) + j,
                     make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
          }
          tc_fence_before_sync();
          __threadfence();
          __syncwarp();
          if (lane == 0) {
            if (kPair) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
            atomicExch(flag, 1u);
          }
          if (++as == kAccStages) { as = 0; aph ^= 1; }
          continue;
        }
        if (splitk) {
          // ---- split-K reader: wait for the other half of K (bounded spin) ----
          if (lane == 0) {
            const long long t0 = clock64();
            while (atomicAdd(flag, 0u) == 0u) {
              if (clock64() - t0 > (1ll << 31)) {
                printf("tavsr: split-K flag timeout (block %d)\n", blockIdx.x);
                __trap();
              }
            }
            *flag = 0u;  // re-armed for the next launch (stream order)
            __threadfence();
          }
          __syncwarp();
        }
        const float* part_row = splitk ? p.partial + static_cast<long long>(m) * 256 : nullptr;
        rowln_finish<kDual, !kTf32>(p, s_param, tacc, m0 + q * 32, lane, stager, part_row);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
      }
      if (++as == kAccStages) { as = 0; aph ^= 1; }
    }
    stager.drain();
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the peer may still be reading this CTA's smem / TMEM
  tc_fence_after_sync();
  if (warp == 1) {
    __syncwarp();
    if (kPair) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tavsr
