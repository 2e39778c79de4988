#!/bin/sh
# compute-sanitizer over the kernels that mix asynchronous-proxy (bulk / TMA) traffic with generic loads and stores:
# K4c (gridnet.cu: bulk zero fill, then generic stores over the same rows), K3 / K0 (gather.cu), K5 (gridnet_sample.cu).
# Logs -> gpurun_out/sanitizer/{memcheck,racecheck,synccheck}.log (copied to profiles/r02/sanitizer/).
#   sh benchmarks/sanitize.sh            (on a GPU box: gpurun -- 'sh benchmarks/sanitize.sh')
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
WIDE='test_gridnet_fwd_bwd or test_gridnet_action_dtypes or test_fused_microrts or test_fused_lux_multi_head or test_fused_lux_after_scaling or test_fused_standardize or test_fused_with_masked or test_gridnet_bf16 or test_fused_bf16 or test_gather_rows_bit_exact or test_rollout_store_step or test_gridnet_vs_reference or test_categorical_fwd_bwd or test_fused_categorical or test_fused_gaussian or test_gae_scalar_gamma_bit_exact or test_segmented_gae or test_running or test_reward'
NARROW='test_gridnet_fwd_bwd or test_fused_microrts or test_fused_lux_multi_head or test_fused_with_masked or test_gather_rows_bit_exact or test_rollout_store_step'
for tool in memcheck racecheck synccheck; do
  sel="$NARROW"; [ "$tool" = memcheck ] && sel="$WIDE"
  echo "== compute-sanitizer --tool $tool : pytest -m gpu -k \"$sel\"" > "$OUT/$tool.log"
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --launch-timeout 0 \
    python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$sel" >> "$OUT/$tool.log" 2>&1
  echo "== exit code $?" >> "$OUT/$tool.log"
  tail -4 "$OUT/$tool.log"
done
