#!/bin/bash
# developer: which engine component makes config 5's TF32 gradients leave the truncated-operand oracle
for env in "X=1" "SB_NO_BN_POOL=1" "SB_BN_TWO_PASS=1" "SB_TC_HALO_SLICES=0" "SB_TC_HALO_WGRAD=0" "SB_TC_HALO=0"; do
  echo "--- $env"
  env $env timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "config5_tf32_gradients" 2>&1 | grep -E "^E +(8/weights|4/weights|10/gamma|6/gamma|0/weights)|passed"
done
