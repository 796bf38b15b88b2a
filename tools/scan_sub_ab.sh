#!/bin/bash
# A/B of the scan's rows-per-thread (level-1 granularity), built on the GPU box.
for sub in ${SUBS:-128 64}; do
  touch bayesian_dlms_b200/csrc/scan.cu
  BDLM_NVCC_EXTRA="-DBDLM_SCAN_SUB=$sub" python -m bayesian_dlms_b200.build > /dev/null || { echo "build failed for $sub"; continue; }
  echo "== kSub=$sub"
  python tools/scan_time.py --cases 2:24,1:24,2:20 | grep "^scan"
done
touch bayesian_dlms_b200/csrc/scan.cu
python -m bayesian_dlms_b200.build > /dev/null
python tools/scan_time.py --cases 2:24 --reps 2 > gpurun_out/plain_scan.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r2_scan_launches.csv python tools/scan_time.py --cases 2:24 --reps 2 > /dev/null 2>&1
