#!/bin/bash
# Decompose the fused decoder-GEMM + log-sum-exp kernel (K4) with the B200VAE_TC_DBG probe bits:
#   0 real kernel | 1 no operand loads | 2 no epilogue work | 4 no MMAs     (results are garbage for != 0)
# 6 = loads only (TMA / L2 / HBM delivery), 3 = MMAs only, 5 = epilogue only, 7 = empty pipeline (fixed cost)
cd "$(dirname "$0")/.."
for dbg in 0 1 2 4 6 3 5 7; do
  echo "== B200VAE_TC_DBG=$dbg"
  B200VAE_TC_DBG=$dbg timeout 300 python scripts/k4_sweep.py --batches "${BATCHES:-125,250,500}" "$@" 2>&1 | grep -E "^B="
done
