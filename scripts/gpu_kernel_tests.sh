#!/bin/bash
# Runs each kernel-level GPU test group in its own process under a timeout, so that one trapped/hung launch does
# not hide the results of the others. Logs to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/kt_gpu.txt 2>&1
for t in test_layout_converters test_linear test_conv_s2_fprop_stats test_conv_s2_dgrad test_conv_s2_wgrad test_convT_fprop_dgrad_wgrad test_edge_in test_edge_out test_batchnorm_relu test_losses test_optimizers_match_torch; do
  echo "=== $t" | tee -a gpurun_out/kt_summary.txt
  timeout 300 python -m pytest -q tests/test_kernels_gpu.py -m gpu -k "$t" -x --no-header -p no:cacheprovider > gpurun_out/kt_$t.log 2>&1
  echo "exit $?" | tee -a gpurun_out/kt_summary.txt
  tail -n 25 gpurun_out/kt_$t.log | grep -E "passed|failed|error|Error|assert|FmriError|timeout" | tail -n 8 | tee -a gpurun_out/kt_summary.txt
done
