#!/bin/bash
# compute-sanitizer over the tiny-model GPU tests, ONE tool per gpurun call (B200_PROFILING.md): usage tools/gpu_sanitizer.sh memcheck|racecheck
cd "$(dirname "$0")/.."
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL="act_quant_codes_bit_exact or int8_conv or conv1x1_f32 or groupnorm_silu_quant_fused or fused_rowprog_equals_layerwise or ddim_loop or attention_core or percentile or group_wise or calibration_vs_reference"
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?" | tee -a gpurun_out/sanitizer_$TOOL.log
grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" gpurun_out/sanitizer_$TOOL.log | tail -12
