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
