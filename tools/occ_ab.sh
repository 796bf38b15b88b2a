#!/bin/bash
# A/B of the register kernels' occupancy targets for n = 3, 4 (built on the GPU box).
set -e
for cfg in "2 1" "3 1" "2 3" "3 2"; do
  set -- $cfg
  touch bayesian_dlms_b200/csrc/kf_small.cu
  BDLM_NVCC_EXTRA="-DBDLM_OCC3=$1 -DBDLM_OCC4=$2" python -m bayesian_dlms_b200.build > /dev/null
  echo "== OCC3=$1 OCC4=$2"
  python tools/kernel_time.py --n 3 --waves 3 | sed 's/^lib=default //'
  python tools/kernel_time.py --n 4 --waves 3 | sed 's/^lib=default //'
done
