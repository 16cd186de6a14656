/*
 * tavsr.h — C ABI of libtavsr_sm100.so: the B200 (sm_100a) kernels behind the Branchformer encoder
 * stack and the CTC scorer of david-gimeno/tailored-avsr.
 *
 * The reference has no FFI for this path: its plug-in surface is the Python module API
 * (src/encoder/branchformer/encoder.py:324, src/encoder/branchformer/encoder_layer.py:153,
 * src/encoder/audiovisual/tailored/encoder_layer.py:118, src/ctc/ctc.py:133-188).  The drop-in
 * Python modules in tailored_avsr_b200/ keep those signatures and bind the entry points below
 * through ctypes; each entry point names the reference op sequence it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return value 0 = ok, negative = error (tavsr_last_error() gives a thread-local message);
 *   - no exceptions cross the boundary, no hidden allocations on the data path: scratch memory is
 *     passed in by the caller (tavsr_*_workspace_bytes says how much);
 *   - "tf32 mode" = fp32 storage, operands rounded to TF32 for the tensor cores, fp32 accumulate.
 *   - row-major activations: (rows = utterance-major frames, b*T + t), leading dimension in
 *     ELEMENTS.
 */
#ifndef TAVSR_H_
#define TAVSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAVSR_VERSION 100

#define TAVSR_OK 0
#define TAVSR_ERR_INVALID (-1)
#define TAVSR_ERR_CUDA (-2)
#define TAVSR_ERR_UNSUPPORTED (-3)

/* activation codes (espnet get_activation / cgMLP GELU) */
#define TAVSR_ACT_NONE 0
#define TAVSR_ACT_SWISH 1
#define TAVSR_ACT_GELU 2
#define TAVSR_ACT_RELU 3

/* `dtype` arguments: operand storage of the tensor-core products in the low byte, OR-ed with flags
 * saying which OUTPUTS are stored as bf16 (only with bf16 operands: outputs that feed the next
 * tensor-core product are bf16, the residual stream / encoder output stay fp32).
 * "tf32x3" (fp32-class accuracy on the TF32 pipe) is composed by the caller: tavsr_split_tf32
 * writes [hi | hi | lo] / [hi | lo | hi] operand triples and the product runs as a TF32 GEMM over
 * the tripled reduction axis (a_hi.b_hi + a_hi.b_lo + a_lo.b_hi). */
#define TAVSR_DT_TF32 0 /* fp32 storage, kind::tf32 */
#define TAVSR_DT_BF16 1 /* bf16 storage, kind::f16  */
#define TAVSR_DT_MASK 0xff
#define TAVSR_DT_OUT_BF16 0x100 /* main output (y / out_main / ctx / u) stored as bf16 */
#define TAVSR_DT_LNA_BF16 0x200 /* out_lnA stored as bf16 */
#define TAVSR_DT_LNB_BF16 0x400 /* out_lnB stored as bf16 */

int tavsr_version(void);
const char* tavsr_last_error(void);
/* debug knobs (descriptor variants etc.); not part of the stable surface */
int tavsr_debug_set(int key, int value);
int tavsr_debug_set_ptr(void* device_buffer); /* kernel phase timestamps (tools/ only) */
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long tavsr_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Y[M,N] = act(X[M,K] . W[N,K]^T + bias)          tcgen05 / TMEM / TMA GEMM
 * Replaces torch.nn.Linear + activation: FFN w_1 + Swish (espnet PositionwiseFeedForward, called at
 * encoder_layer.py:194,314), cgMLP channel_proj1 + GELU (encoder_layer.py:220), fused
 * linear_q|k|v (encoder_layer.py:208), linear_pos.
 * round_out != 0 rounds Y to TF32 (legal when Y only feeds further tensor-core products).
 * dtype: TAVSR_DT_TF32 (fp32 x / w / y) or TAVSR_DT_BF16 (bf16 x / w; y fp32, or bf16 with
 * TAVSR_DT_OUT_BF16).  Requirements: 16-byte row pitches and bases (K % 4 / N % 4 elements in
 * fp32, % 8 in bf16).
 * ---------------------------------------------------------------------------------------------- */
int tavsr_gemm_bias_act(const void* x, long long ldx, const void* w, long long ldw,
                        const float* bias, void* y, long long ldy, int M, int N, int K, int act,
                        int round_out, int dtype, void* stream);

/* Two projections of the same rows in ONE launch: Y1 = X1 . W1^T + b1 (no activation: the fused
 * linear_q|k|v, encoder_layer.py:208) and Y2 = gelu(X2 . W2^T + b2) (cgMLP channel_proj1,
 * encoder_layer.py:220); X1 / X2 are the two LayerNorm outputs of the same block input.  Same M and
 * K; the persistent tile scheduler walks both problems' 256-wide tiles.  dtype as above. */
int tavsr_gemm_group2(const void* x1, long long ldx1, const void* w1, long long ldw1,
                      const float* bias1, void* y1, long long ldy1, int N1, const void* x2,
                      long long ldx2, const void* w2, long long ldw2, const float* bias2, void* y2,
                      long long ldy2, int N2, int M, int K, int dtype, void* stream);

/* Weight-gradient product (training): out[M, N] = A[M, K] . B[N, K]^T, fp32 storage, TF32 operands,
 * for a long reduction axis K (all frames of the batch) and few output tiles: the library splits K
 * over CTA pairs (partial tiles in `workspace`, tavsr_gemm_wgrad_workspace_bytes) and sums the
 * partials in a fixed order (bit-reproducible).  dW = dY^T X is the call
 * (A = dY^T [N_out, frames], B = X^T [N_in, frames], both from tavsr_transpose_2d). */
size_t tavsr_gemm_wgrad_workspace_bytes(int M, int N, int K);
int tavsr_gemm_wgrad(const float* a, long long lda, const float* b, long long ldb, float* out,
                     long long ldo, int M, int N, int K, void* workspace, long long workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-complete GEMM, N == 256 (the model width): one thread owns one output row, so everything that
 * follows the projection in the reference layer is fused into the epilogue:
 *
 *   acc  = X . W^T                      (or  rowscale1[seg]*X.W^T + rowscale2[seg]*X2.W^T,
 *                                        seg = row / rows_per_seg — the learned_ave merge,
 *                                        encoder_layer.py:291-293)
 *   v0   = residual + alpha * (acc + bias)                      (encoder_layer.py:194,291,314)
 *   v1   = ln0 ? LayerNorm(v0; ln0, eps0) : v0                  (norm_final, encoder_layer.py:316)
 *   out_main = v1            (optional, fp32)
 *   out_lnA  = LayerNorm(v1; lnA, eps)   out_lnB = LayerNorm(v1; lnB, eps)   (next block's norms)
 *   dots_out[row] = (v1 . dot1, v1 . dot2)                      (pooling_proj / weight_proj,
 *                                                                encoder_layer.py:243,258)
 * All pointer members may be NULL to disable that stage.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tavsr_rowln_args {
  int struct_size; /* sizeof(tavsr_rowln_args), for forward compatibility */
  int M, K;        /* N is 256 */
  int dtype;       /* TAVSR_DT_* operand type | TAVSR_DT_{OUT,LNA,LNB}_BF16 output flags */
  const void* x;
  long long ldx;
  const void* x2; /* second operand (merge) or NULL */
  long long ldx2;
  const void* w; /* [256, K] */
  long long ldw;
  const float* bias; /* [256] or NULL */
  const float* residual;
  long long ldr;
  float alpha;
  const float* rowscale1; /* [ceil(M / rows_per_seg)] */
  const float* rowscale2;
  int rows_per_seg;
  const float* ln0_g;
  const float* ln0_b;
  float eps0;
  void* out_main; /* fp32, or bf16 with TAVSR_DT_OUT_BF16 */
  long long ld_main;
  int round_main;
  const float* lnA_g;
  const float* lnA_b;
  void* out_lnA; /* fp32, or bf16 with TAVSR_DT_LNA_BF16 */
  long long ld_lnA;
  int round_lnA;
  const float* lnB_g;
  const float* lnB_b;
  void* out_lnB; /* fp32, or bf16 with TAVSR_DT_LNB_BF16 */
  long long ld_lnB;
  int round_lnB;
  float eps;
  const float* dot1; /* [256] */
  const float* dot2;
  float* dots_out; /* [M,2] */
  /* Optional split-K scratch (tavsr_rowln_workspace_bytes(M) bytes, ZERO-initialised once, owned
   * by the caller, one per concurrently used stream).  When present and the problem leaves more
   * than half of the SMs idle (K >= 1024, few row tiles) the reduction is split in two across CTA
   * pairs and recombined in the epilogue.  The kernel leaves the flag words zeroed again. */
  void* workspace;
  long long workspace_bytes;
  /* Sequential dual mode (x2 != NULL and k1 > 0): the reduction axis is the concatenation of the
   * two operands — x is [M,k1], x2 is [M,K-k1], w is [256,K] = [W1 | W2] — and
   *   acc = rowscale1[seg]*(X.W1^T + segbias1) + rowscale2[seg]*(X2.W2^T + segbias2).
   * This is the learned_ave / fixed_ave merge with the branch output projections folded into
   * merge_proj (W1 = Wm.Wo, W2 = Wm.W_proj2; encoder_layer.py:208-209,220,291-293), so the
   * attention context and the gated cgMLP activations feed the merge GEMM directly.
   * k1 and K-k1 must be multiples of 32 (64 in bf16).  segbias1/2: [256] or NULL; not combinable
   * with dots.  The bf16 kernel has the sequential dual mode only. */
  int k1;
  const float* segbias1;
  const float* segbias2;
} tavsr_rowln_args;

int tavsr_gemm_rowln(const tavsr_rowln_args* args, void* stream);
size_t tavsr_rowln_workspace_bytes(int M);

/* ------------------------------------------------------------------------------------------------
 * Fused position-wise feed-forward block (espnet PositionwiseFeedForward + the residual / LayerNorm
 * lines around it, encoder_layer.py:193-194, 202, 216, 313-316):
 *
 *   acc = act(xn . W1^T + b1) . W2^T            hidden = 2048 stays in TMEM / shared memory
 *   then exactly the tavsr_gemm_rowln epilogue on acc (ep.bias = b2, ep.residual, ep.alpha, ep.ln0,
 *   ep.out_main, ep.lnA/lnB ...).  ep.x / ep.w / ep.K / ep.x2 / ep.workspace are ignored.
 * Built for model width 256 and hidden width 2048 (every shipped config).  ep.dtype selects tf32
 * (fp32 xn / w1 / w2, TF32 hidden in TMEM) or bf16 (bf16 xn / w1 / w2, packed bf16 hidden in TMEM);
 * ep.dots_out is not supported here.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tavsr_ffn_args {
  int struct_size; /* sizeof(tavsr_ffn_args) */
  int hidden;      /* 2048 */
  int act;         /* TAVSR_ACT_* */
  int reserved;
  const void* xn;  /* [M, 256] LayerNorm-ed input */
  long long ldxn;
  const void* w1;  /* [hidden, 256] */
  long long ldw1;
  const float* b1; /* [hidden] */
  const void* w2;  /* [256, hidden] */
  long long ldw2;
  tavsr_rowln_args ep;
} tavsr_ffn_args;

int tavsr_ffn_fused(const tavsr_ffn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone LayerNorm over the last dim D (D % 128 == 0, D <= 2048) producing up to two affine
 * variants of the same normalised row (espnet LayerNorm, eps 1e-12; torch LayerNorm of the
 * `linear` input layer, encoder.py:126, eps 1e-5).  scale multiplies the result (x * sqrt(d) of
 * RelPositionalEncoding).
 * ---------------------------------------------------------------------------------------------- */
int tavsr_layernorm(const float* x, long long ldx, int M, int D, float eps, const float* gA,
                    const float* bA, void* outA, long long ldA, int roundA, const float* gB,
                    const float* bB, void* outB, long long ldB, int roundB, float scale,
                    int dtype /* TAVSR_DT_LNA_BF16 / TAVSR_DT_LNB_BF16: bf16 outputs */,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused relative-position multi-head self-attention (espnet RelPositionMultiHeadedAttention.forward
 * as called at encoder_layer.py:208 / tailored encoder_layer.py:192,239), flash style: the
 * (B,h,T,T) score tensors and the (B,h,T,2T-1) pre-shift tensor never reach HBM.
 *   qkv   [B*T, ld_qkv]   q | k | v, each H*64 wide (output of the fused projection)
 *   pos   [2T-1, ld_pos]  linear_pos(pos_emb) for this layer, row k <-> relative position T-1-k
 *   u, v  [H*64]          pos_bias_u / pos_bias_v
 *   lens  [B] int32       valid keys per utterance (mask.sum); keys >= lens[b] get probability 0
 *   ctx   [B*T, ld_ctx]   softmax(((q+u)k^T + rel_shift((q+v)p^T)) / 8) v, heads concatenated
 * d_k is fixed to 64.  dtype: TAVSR_DT_TF32 (fp32 qkv / pos / ctx) or
 * TAVSR_DT_BF16 | TAVSR_DT_OUT_BF16 (bf16 qkv / pos / ctx; scores, softmax and the output
 * accumulation stay fp32).  All pitches and bases 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int tavsr_relpos_attn_fwd(const void* qkv, long long ld_qkv, const void* pos, long long ld_pos,
                          const float* u, const float* v, const int32_t* lens, void* ctx,
                          long long ld_ctx, int B, int T, int H, int round_out, int dtype,
                          float* lse /* optional [B,H,T]: base-2 log-sum-exp of the scaled scores,
                                        kept by the training forward for tavsr_relpos_attn_bwd */,
                          void* stream);

/* Training forward with attention-probability dropout (espnet attention.py:
 * `matmul(self.dropout(self.attn), value)`, active when attention_dropout_rate > 0 in train mode).
 * fp32 / TF32 only.  drop_keep [(B*H*T), ld_drop] bytes: keep[b][h][i][j] != 0 keeps P_ij; ld_drop a
 * multiple of 128 >= T; kept probabilities are scaled by drop_scale = 1 / (1 - p).  lse stays the
 * log-sum-exp of the undropped scores. */
int tavsr_relpos_attn_fwd_dropout(const float* qkv, long long ld_qkv, const float* pos,
                                  long long ld_pos, const float* u, const float* v,
                                  const int32_t* lens, float* ctx, long long ld_ctx, int B, int T,
                                  int H, float* lse, const uint8_t* drop_keep, long long ld_drop,
                                  float drop_scale, void* stream);

/* Backward of the attention core (training).  Given dctx = d loss / d ctx, the forward's qkv, pos,
 * u, v, lens, ctx and lse (fp32), writes the k and v blocks of dqkv [B*T, 3*H*64] and ACCUMULATES
 * (fp32 atomics; zero them first): d q into the q block of dqkv, d pos_bias_u / d pos_bias_v into
 * du / dvb [H*64], and d linear_pos(pos_emb) into PER-UTTERANCE slabs dpos [B][2T-1][H*64] (the
 * caller sums the slabs: tavsr_col_sums over a [B, (2T-1)*H*64] view).  Probabilities are
 * recomputed from lse: nothing of size T x T is stored.  The products run as TF32 mma.sync
 * tensor-core MMAs with fp32 accumulation.  Arithmetic: oracle/bwd_formulas.py::relpos_attn_core_bwd
 * (verified against autograd). */
int tavsr_relpos_attn_bwd(const float* qkv, long long ld_qkv, const float* pos, long long ld_pos,
                          const float* u, const float* v, const int32_t* lens, const float* ctx,
                          long long ld_ctx, const float* dctx, long long ld_dctx, const float* lse,
                          float* dqkv, long long ld_dqkv, float* du, float* dvb, float* dpos,
                          const uint8_t* drop_keep /* NULL, or the forward's keep mask */,
                          long long ld_drop, float drop_scale, int B, int T, int H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutional spatial gating unit (espnet ConvolutionalSpatialGatingUnit.forward, reached through
 * self.cgmlp at encoder_layer.py:220):  h = [r | g] (C = 2*Ch wide, post-GELU);
 *   out = r * (dwconv_k(LayerNorm_Ch(g)) + conv_bias),  zero padding (k-1)/2 on the normalised
 * sequence at t<0 and t>=T, the key-padding mask is NOT applied (espnet ignores it).
 *   stats [B*T,2] scratch (mean, rstd).  Kernel size must be 31.
 * dtype: TAVSR_DT_TF32 (fp32 h / out) or TAVSR_DT_BF16 | TAVSR_DT_OUT_BF16 (bf16 h / out;
 * LayerNorm statistics, the convolution and the gate product stay fp32).
 * ---------------------------------------------------------------------------------------------- */
int tavsr_csgu_fwd(const void* h, long long ldh, const float* norm_g, const float* norm_b,
                   const float* conv_w /* [Ch,31] */, const float* conv_b, void* out,
                   long long ldo, float* stats, int B, int T, int Ch, int ksize, float eps,
                   int round_out, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * learned_ave merge weights (encoder_layer.py:241-289).  dotsK[row] = (x_K . pooling_projK.weight,
 * x_K . weight_projK.weight) come from tavsr_gemm_rowln.  Per utterance b:
 *   s_K = softmax_{t < lens[b]}((dotsK[.,0] + pool_bK) / sqrt(size));  omega_K = sum_t s_K dotsK[.,1] + wproj_bK
 *   (w1, w2) = softmax(omega_1, omega_2)  -> weight_global / weight_local
 * ---------------------------------------------------------------------------------------------- */
/* Row dots  out[m] = (a[m,:] . va, a[m,:] . vb)  for one or two (a2 may be NULL) activation
 * matrices in one launch: the pooling_proj / weight_proj scores of the learned_ave merge
 * (encoder_layer.py:243,258) taken on the attention context and the gated cgMLP activations with
 * the branch output projections folded into the vectors (va = Wo^T a, ...).  out: [M,2]. */
int tavsr_row_dots(const void* a1, long long ld1, int K1, const float* va1, const float* vb1,
                   float* out1, const void* a2, long long ld2, int K2, const float* va2,
                   const float* vb2, float* out2, int M, int dtype /* a1 / a2 fp32 or bf16 */,
                   void* stream);

int tavsr_merge_learned_ave_weights(const float* dots1, const float* dots2, const int32_t* lens,
                                    float pool_b1, float pool_b2, float wproj_b1, float wproj_b2,
                                    float inv_sqrt_size, float* w1, float* w2, int B, int T,
                                    void* stream);
/* The two steps above in ONE launch for the folded two-branch block (a1 = attention context, 256
 * wide; a2 = gated cgMLP hidden, 1024 wide; fp32 or bf16 by `dtype`): a cluster of 4 CTAs per
 * utterance computes the row dots, the masked softmax pooling over time and the 2-way softmax, the
 * partials crossing the cluster over distributed shared memory.  T <= 2048. */
int tavsr_merge_scores(const void* a1, long long ld1, int K1, const void* a2, long long ld2, int K2,
                       const float* va1, const float* vb1, const float* va2, const float* vb2,
                       const int32_t* lens, float pool_b1, float pool_b2, float wproj_b1,
                       float wproj_b2, float inv_sqrt_size, float* w1, float* w2, int B, int T,
                       int dtype, void* stream);
/* Training form: the four biases (pool_b1, pool_b2, wproj_b1, wproj_b2) are read from the DEVICE
 * array `scal` instead of being passed by value (no host read-back of parameters per step). */
int tavsr_merge_learned_ave_weights_dev(const float* dots1, const float* dots2, const int32_t* lens,
                                        const int32_t* lens2 /* NULL, or branch 2's own lengths */,
                                        const float* scal, float inv_sqrt_size, float* w1, float* w2,
                                        int B, int T, void* stream);
/* General form: dotsK holds npK partial pairs per frame ([B*T, npK, 2], summed inside) and each
 * branch has its own length array (lens2 == NULL: lens1) - the two modalities of
 * AdaptiveAudioVisualFusion carry separate masks (adaptive_audiovisual_fusion.py:150-183). */
int tavsr_merge_learned_ave_weights2(const float* dots1, int np1, const float* dots2, int np2,
                                     const int32_t* lens1, const int32_t* lens2, float pool_b1,
                                     float pool_b2, float wproj_b1, float wproj_b2,
                                     float inv_sqrt_size, float* w1, float* w2, int B, int T,
                                     void* stream);
/* out[m,:] = w1[m / rows_per_seg] * a[m,:] + w2[m / rows_per_seg] * b[m,:]: the weighted modality
 * average in front of the fusion FFN (adaptive_audiovisual_fusion.py:187-194).  D % 4 == 0. */
int tavsr_scale_add_rows(const float* a, long long lda, const float* b, long long ldb,
                         const float* w1, const float* w2, int rows_per_seg, void* out,
                         long long ldo, int M, int D,
                         int dtype /* TAVSR_DT_OUT_BF16: out stored as bf16 (w = (1, 0): a cast) */,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3xTF32 operand split (the "tf32x3" compute mode): out is [M, 3K],
 *   pattern 0 (activations): [hi | hi | lo],   pattern 1 (weights): [hi | lo | hi],
 * hi = tf32(x), lo = x - hi.  tavsr_gemm_bias_act / tavsr_gemm_rowln over the tripled reduction
 * axis then compute a_hi.b_hi + a_hi.b_lo + a_lo.b_hi with fp32 accumulation.
 * ---------------------------------------------------------------------------------------------- */
int tavsr_split_tf32(const float* in, long long ld, float* out, long long ldo, int M, int K,
                     int pattern, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2dSubsampling front end (espnet Conv2dSubsampling inside the encoder module,
 * encoder.py:149-155): writes the im2col operand of the second 3x3 stride-2 convolution with the
 * first convolution + ReLU evaluated on the fly,
 *   A[(b, t2, f2), (i*3+j)*C + c] = relu(conv1)[b, c, 2 t2 + i, 2 f2 + j],
 * T2 = ((Tin-1)/2-1)/2, F2 = ((F-1)/2-1)/2, A is [B*T2*F2, 9*C] row-major.  The second convolution
 * is tavsr_gemm_bias_act(A, W2r, b2, RELU) with W2r[c2, (i*3+j)*C + c1] = w2[c2, c1, i, j].
 *   x [B, Tin, F] contiguous;  w1 [C, 9] (= conv.0.weight flattened);  b1 [C].
 * ---------------------------------------------------------------------------------------------- */
int tavsr_conv2d_sub_im2col(const float* x, int B, int Tin, int F, const float* w1,
                            const float* b1, int C, void* A,
                            int dtype /* TAVSR_DT_OUT_BF16: A stored as bf16 */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CTC head: logits = hs . W^T + b in fp32 FMA (argmax must be bit-stable), then log-softmax /
 * softmax / argmax over V <= 256 (ctc.py:143,160-188; V <= 64 - the char vocabularies - on the
 * warp-per-frame kernel with W^T in shared memory, 64 < V <= 256 - the 256-token SentencePiece
 * alternative - on a thread-per-token kernel).  Any of logits / logp / prob / amax may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int tavsr_ctc_head(const float* hs, long long ldh, const float* w /* [V,D] */, const float* b,
                   float* logits, float* logp, float* prob, int64_t* amax, int M, int D, int V,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vocabulary residual  out = x + p . W^T + b,  p (M,V) posteriors, W (D,V):
 *   - InterCTC self-conditioning  x + conditioning_layer(ctc.softmax(after_norm(x)))
 *     (src/encoder/branchformer/encoder.py:393-399);
 *   - InterCTCResidualModule      x + proj_2(softmax(proj_1(x)))
 *     (src/ctc/interctc_residual_module.py:11-16).
 * When xn != NULL it also receives LayerNorm(out; ln_g, ln_b, eps) — the next block's
 * norm_ff_macaron, so the conditioned stream needs no extra pass.  D in {128,256,512}, V <= 256;
 * for V > 64 `w` must be the TRANSPOSED matrix W^T (V,D) (rows streamed from L2 instead of staged
 * in shared memory).
 * ---------------------------------------------------------------------------------------------- */
int tavsr_vocab_residual(const float* x, long long ldx, const float* p, const float* w /* [D,V] */,
                         const float* b, float* out, long long ldo, const float* ln_g,
                         const float* ln_b, float eps, void* xn, long long ldn, int M, int D, int V,
                         int dtype /* TAVSR_DT_LNA_BF16: xn stored as bf16 */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CTC loss, log domain, blank = 0 (torch.nn.CTCLoss(reduction="none", zero_infinity) at
 * ctc.py:41,60-61): one warp per utterance, V <= 256.
 *   logp    [B, T, V]      log-softmax (batch-major; the reference's (T,B,V) transpose is a view)
 *   targets [B, Lmax]      int64, padded (ys_pad); tlens [B] int64/int32 as int32 here
 *   hlens   [B] int32      valid frames
 *   nll     [B]            per-utterance negative log likelihood (inf -> 0 when zero_infinity)
 *   grad    [B, T, V]      optional: d(sum_b gscale*nll_b)/d(logits) = gscale*(softmax - occupancy)
 *                          for t < hlens[b], 0 elsewhere (what ATen's backward folds through
 *                          log_softmax); needs `alpha_ws` of tavsr_ctc_workspace_bytes.
 * ---------------------------------------------------------------------------------------------- */
size_t tavsr_ctc_workspace_bytes(int B, int T, int Lmax);
int tavsr_ctc_loss(const float* logp, const int64_t* targets, long long ld_targets,
                   const int32_t* hlens, const int32_t* tlens, float* nll, float* grad,
                   float gscale, void* alpha_ws, int B, int T, int V, int Lmax, int zero_infinity,
                   void* stream);

/* Backward of the CTC head logits = hs . W^T + b (ctc.py:143) given dlogits [M,V] (the `grad` of
 * tavsr_ctc_loss), scaled per utterance by row_scale[m / rows_per_seg] (the upstream d total / d
 * nll_b; NULL = 1):  dhs [M,D] = g . W,  dw [V,D] = g^T . hs,  db [V] = sum_m g.  D == 256, V <= 64.
 * `workspace` (tavsr_ctc_head_bwd_workspace_bytes) holds per-CTA partials: no atomics, results are
 * bit-reproducible.  Outputs are overwritten, not accumulated. */
size_t tavsr_ctc_head_bwd_workspace_bytes(int M);
int tavsr_ctc_head_bwd(const float* dlogits, const float* row_scale, int rows_per_seg,
                       const float* hs, long long ldh, const float* w, float* dhs, long long ldd,
                       float* dw, float* db, void* workspace, long long workspace_bytes, int M,
                       int D, int V, void* stream);

/* Softmax backward over the vocabulary (training of InterCTC self-conditioning, encoder.py:393-401):
 * dlogits[m][v] = p[m][v] * (dp[m][v] - sum_v p[m][v] dp[m][v]); p, dp, dlogits contiguous [M, V]. */
int tavsr_softmax_bwd(const float* p, const float* dp, float* dlogits, int M, int V, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward kernels of the encoder (training): transcriptions of
 * oracle/bwd_formulas.py (verified against autograd on the CPU), checked on the GPU in
 * tests/test_backward_gpu.py and composed into autograd nodes by tailored_avsr_b200/training.py.
 * All fp32, two-stage reductions through caller-provided workspaces (atomics only in
 * tavsr_relpos_attn_bwd).
 *   tavsr_transpose_2d    out[c][r] = in[r][c]           (wgrad operands, W^T for dgrad)
 *   tavsr_col_sums        out[c] = sum_r a[r][c] (* b[r][c] when b != NULL)     (bias gradients)
 *   tavsr_act_bwd         dz = dh * act'(z), act = TAVSR_ACT_*   (Swish / exact GELU / ReLU)
 *   tavsr_layernorm_bwd   dx = LN_bwd(x; gamma, dy) (+ dres), dgamma | dbeta into one [2, D] buffer
 *                         (dbeta == dgamma + D); D in {256, 512, 1024}; statistics recomputed
 *   tavsr_csgu_conv_bwd   CSGU backward up to the LayerNorm: given h = [r | g], the forward's
 *                         (mean, rstd) `stats` and du, writes dr into dh[:, :Ch], dn = d LN(g) into
 *                         `dn`, and dconv_w [Ch,31], dconv_b [Ch]; the gate half of dh is then
 *                         tavsr_layernorm_bwd(x = h + Ch, dy = dn, dx = dh + Ch, D = Ch)
 * ---------------------------------------------------------------------------------------------- */
int tavsr_transpose_2d(const float* in, long long ld_in, float* out, long long ld_out, int R, int C,
                       void* stream);
/* Backward of the first half of the Conv2dSubsampling front end (training): given dA, the gradient
 * of the im2col operand tavsr_conv2d_sub_im2col wrote (same [(b,t2,f2)][(kt,kf,c)] layout, fp32),
 * gathers it back onto conv1's output grid, applies conv1's ReLU mask (conv1 is re-evaluated from x)
 * and reduces d conv1.weight / d conv1.bias: grads [C][10] = nine taps then the bias per channel.
 * The input features get no gradient. */
size_t tavsr_conv2d_sub_bwd_workspace_bytes(int B, int Tin, int C);
int tavsr_conv2d_sub_bwd(const float* x, int B, int Tin, int F, const float* w1, const float* b1,
                         int C, const float* dA, float* grads, void* workspace,
                         long long workspace_bytes, void* stream);

/* Fused elementwise + transpose forms (one pass instead of an elementwise pass and a transpose):
 *   tavsr_act_fwd_t   hT[c][m] = act(z[m][c]) (* mask[m][c]); hT rows zero-padded to a multiple of 4
 *   tavsr_act_bwd_t   dz = dh * act'(z) row-major and dzT = dz^T                                   */
int tavsr_act_fwd_t(const float* z, long long ldz, const float* mask, long long ldm, float* hT,
                    long long ldt, int M, int C, int act, void* stream);
int tavsr_act_bwd_t(const float* z, long long ldz, const float* dh, long long ldh, float* dz,
                    long long ldd, float* dzT, long long ldt, int M, int C, int act, void* stream);
size_t tavsr_col_sums_workspace_bytes(int R, int C);
int tavsr_col_sums(const float* a, long long lda, const float* b, long long ldb, float* out,
                   void* workspace, long long workspace_bytes, int R, int C, void* stream);
/* h = act(z) (* mask when mask != NULL: a dropout keep-mask pre-scaled by 1/(1-p)) */
int tavsr_act_fwd(const float* z, long long ldz, const float* mask, long long ldm, float* h,
                  long long ldh, int M, int C, int act, void* stream);
int tavsr_act_bwd(const float* z, long long ldz, const float* dh, long long ldh, float* dz,
                  long long ldd, int M, int C, int act, void* stream);
size_t tavsr_layernorm_bwd_workspace_bytes(int M, int D);
int tavsr_layernorm_bwd(const float* x, long long ldx, const float* gamma, const float* dy,
                        long long ldy, const float* dres, long long ldr, float* dx, long long ldd,
                        float* dgamma, float* dbeta, void* workspace, long long workspace_bytes,
                        int M, int D, float eps, void* stream);
size_t tavsr_csgu_bwd_workspace_bytes(int B, int T, int Ch);
int tavsr_csgu_conv_bwd(const float* h, long long ldh, const float* norm_g, const float* norm_b,
                        const float* conv_w, const float* conv_b, const float* stats,
                        const float* du, long long ldu, float* dh, long long lddh, float* dn,
                        long long lddn, float* dconv_w, float* dconv_b, void* workspace,
                        long long workspace_bytes, int B, int T, int Ch, int ksize, void* stream);

/* learned_ave merge backward (encoder_layer.py:241-291 up to m = w1 x1 + w2 x2), D == 256,
 * T <= 2048.  Given dm = d loss / d m: dx1, dx2 and grads = [da1 | db1 | da2 | db2] (4 x 256:
 * pooling_projK.weight, weight_projK.weight) followed by [dc1, de1, dc2, de2] (their biases), 1028
 * floats.  aK / bK are the pooling_projK / weight_projK weight vectors; scal is a DEVICE array
 * [c1, e1, c2, e2] of their biases (read on the device: no host synchronisation inside a training
 * step).  Arithmetic: oracle/bwd_formulas.py::learned_ave_merge_bwd. */
size_t tavsr_merge_learned_ave_bwd_workspace_bytes(int B);
int tavsr_merge_learned_ave_bwd(const float* x1, long long ld1, const float* x2, long long ld2,
                                const float* dm, long long ldm, const int32_t* lens,
                                const int32_t* lens2 /* NULL: branch 2 uses lens too; else its own
                                                        lengths (audio-visual fusion) */,
                                const float* a1, const float* b1, const float* a2, const float* b2,
                                const float* scal, float* dx1, long long ldd1, float* dx2,
                                long long ldd2, float* grads, void* workspace,
                                long long workspace_bytes, int B, int T, int D, void* stream);

/* Greedy CTC decode: collapse repeats of `amax` and drop blank (espnet_model.py:590-592,
 * maskctc_model.py:287-291).  lens == NULL collapses over all T frames (what _calc_ctc_loss does).
 * tokens [B,T] padded with -1, ntok [B]. */
int tavsr_ctc_greedy(const int64_t* amax, const int32_t* lens, int64_t* tokens, int32_t* ntok,
                     int B, int T, int blank, void* stream);

/* CTC prefix scoring step (espnet CTCPrefixScoreTH.__call__ as driven by CTCPrefixScorer,
 * src/inference/asr_inference.py:142).  One utterance, `nhyp` hypotheses, all V candidates.
 *   logp   [T,V]         log-softmax of the utterance
 *   r_prev [nhyp,T,2]    (log r^n, log r^b) of each hypothesis
 *   last   [nhyp] int32  last label of each hypothesis (<0: empty prefix)
 *   plen   [nhyp] int32  prefix length |h|
 *   psi_prev [nhyp]      previous prefix score
 *   r_new  [nhyp,V,T,2]  new forward variables per candidate
 *   score  [nhyp,V]      log psi(h.c) - psi_prev; eos gets the full-sequence score, blank logzero */
int tavsr_ctc_prefix_score(const float* logp, const float* r_prev, const int32_t* last,
                           const int32_t* plen, const float* psi_prev, float* r_new, float* score,
                           int T, int Tvalid, int V, int nhyp, int blank, int eos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAVSR_H_ */
