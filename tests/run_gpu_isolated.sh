#!/bin/bash
# Runs every GPU kernel test function in its own process so that one trapped kernel (sticky CUDA
# error) cannot mask the others.  Output: gpurun_out/kernels_<name>.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_gemm_plain test_gemm_conv test_gemm_gru_epilogues test_pack_unpack test_corr_lookup \
         test_upsample_and_small_kernels test_corr_build test_attn_lse_pv_finalize test_encoder_norm_kernels test_fused_encoder test_corr_lookup0; do
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$t" --timeout=120 -x 2>&1 | tail -25 > gpurun_out/kernels_$t.log
  echo "== $t: $(tail -1 gpurun_out/kernels_$t.log)"
done
