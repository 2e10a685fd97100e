#!/bin/bash
# ablation matrix for the channels-last SYRK main kernel (profiling aid)
L=a576,a1152,a2304,a4608,g1024,g256,g64,g256s
for dbg in 0 1 2; do
  echo "== DBG=$dbg (1: no TMA, 2: no MMA)"
  CURVATURE_B200_DBG=$dbg python scripts/profile_layer.py --prec bf16 $L
done
for kb in 32 48; do
  echo "== STAGE_KB=$kb"
  CURVATURE_B200_STAGE_KB=$kb python scripts/profile_layer.py --prec bf16 $L
done
