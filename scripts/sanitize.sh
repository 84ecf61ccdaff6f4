#!/bin/bash
# compute-sanitizer jobs for the hand-written kernels. Run on a GPU box:
#   gpurun -- 'bash scripts/sanitize.sh memcheck'      (or racecheck / synccheck / initcheck)
# The selection keeps the instrumented run short; the full suite is `pytest tests/test_gpu_kernels.py`.
TOOL=${1:-memcheck}
SEL=${2:-"test_gemm_kmajor and 256-512-512 or test_conv_nhwc and 3-4-16-8 or test_lean_epilogue and 4096 or test_strided and 4-64-32-128 or test_fused_optimizer or test_herding or test_fused_augmentation and default-dtype0 or test_comm_single_rank"}
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --error-exitcode 9 --log-file "gpurun_out/sanitizer_${TOOL}.log" \
  python -m pytest tests/test_gpu_kernels.py -q -x --timeout 1200 -p no:cacheprovider -k "$SEL" 2>&1 | tail -4
echo "sanitizer exit: $?"
tail -5 "gpurun_out/sanitizer_${TOOL}.log"
