#!/bin/bash
# compute-sanitizer memcheck / racecheck over the CI-shape kernel and parity tests.
# Usage (on the GPU box): bash tools/sanitize.sh [outdir]; summaries go to <outdir>/sanitizer_*.log
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL_KERNELS='test_gemm_bias_act or test_gemm_rowln_residual_two_ln or test_gemm_rowln_dual_merge or test_gemm_rowln_sequential or test_relpos_attention or test_csgu or test_ffn_fused or test_ctc_head or test_ctc_loss_and_grad or test_ctc_greedy or test_row_dots or test_merge_weights or test_vocab_residual or test_layernorm or test_conv2d'
export TAVSR_SANITIZER=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "($SEL_KERNELS) and not 8000 and not 4864 and not 1992" \
  > "$OUT/sanitizer_memcheck_kernels.log" 2>&1
echo "memcheck kernels exit $?" | tee -a "$OUT/sanitizer_memcheck_kernels.log"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "asr_small or av_tailored_small or asr_tailored_small" \
  > "$OUT/sanitizer_memcheck_parity.log" 2>&1
echo "memcheck parity exit $?" | tee -a "$OUT/sanitizer_memcheck_parity.log"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "(test_csgu or test_ctc_loss_and_grad or test_ctc_head or test_row_dots or test_layernorm or test_merge_weights) and not 8000 and not 1992" \
  > "$OUT/sanitizer_racecheck_rowops.log" 2>&1
echo "racecheck rowops exit $?" | tee -a "$OUT/sanitizer_racecheck_rowops.log"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "(test_gemm_bias_act or test_ffn_fused or test_relpos_attention or test_gemm_rowln_dual_merge) and not 8000 and not 1992 and not 4864" \
  > "$OUT/sanitizer_racecheck_tc.log" 2>&1
echo "racecheck tc exit $?" | tee -a "$OUT/sanitizer_racecheck_tc.log"
# round-2 additions: bf16 kernels and the training (backward) kernels at CI shapes
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_bf16_gpu.py -q -x -m gpu -k "not 8000 and not 4864 and not 1992" \
  > "$OUT/sanitizer_memcheck_bf16.log" 2>&1
echo "memcheck bf16 exit $?" | tee -a "$OUT/sanitizer_memcheck_bf16.log"
SEL_BWD='test_transpose_and_col_sums or test_act_bwd or test_layernorm_bwd or test_csgu_bwd or test_merge_learned_ave_bwd or test_relpos_attention_bwd or test_relpos_attention_dropout_fwd_bwd or test_fused_elementwise_transpose or test_gemm_wgrad_split_k or test_linear_bwd or test_act_fwd'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_backward_gpu.py -q -x -m gpu -k "($SEL_BWD) and not 8000" \
  > "$OUT/sanitizer_memcheck_backward.log" 2>&1
echo "memcheck backward exit $?" | tee -a "$OUT/sanitizer_memcheck_backward.log"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_backward_gpu.py -q -x -m gpu -k "test_encoder_training_with_dropout_matches_reference and vsr_small" \
  > "$OUT/sanitizer_memcheck_training.log" 2>&1
echo "memcheck training exit $?" | tee -a "$OUT/sanitizer_memcheck_training.log"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_backward_gpu.py -q -x -m gpu -k "(test_relpos_attention_bwd or test_merge_learned_ave_bwd or test_csgu_bwd or test_fused_elementwise_transpose or test_layernorm_bwd or test_transpose_and_col_sums) and not 8000 and not 250" \
  > "$OUT/sanitizer_racecheck_backward.log" 2>&1
echo "racecheck backward exit $?" | tee -a "$OUT/sanitizer_racecheck_backward.log"
