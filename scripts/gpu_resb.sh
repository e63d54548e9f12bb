#!/bin/bash
# resident-weight pair kernel: kernel tests (forced on for every eligible geometry), then A/B against the ring-only kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "resident or tensor_core" 2>&1 | tail -5
bash scripts/gpu_ab.sh VARSEP_RESIDENT_OC=1 VARSEP_RESIDENT_OC=2 VARSEP_RESIDENT_OC=3
