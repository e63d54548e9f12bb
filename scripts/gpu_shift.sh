#!/bin/bash
# shifted-window kernel: kernel tests (forced on for every eligible geometry), per-launch ncu table, A/B against the per-class kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "size_gated or many_tiles or tensor_core" 2>&1 | tail -15
bash scripts/gpu_ncu_kernel.sh tc_conv_shift shift 12 2>&1 | tail -4
bash scripts/gpu_ab.sh VARSEP_DISABLE_SHIFT=1 "$@"
