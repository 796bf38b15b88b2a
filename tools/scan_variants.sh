#!/bin/bash
# Time prebuilt kSub variants of the library (build_variants/libbdlm_<sub>.so, built on the dev box).
cp bayesian_dlms_b200/libbdlm.so /tmp/libbdlm_keep.so
for sub in ${SUBS:-64 32 16}; do
  cp build_variants/libbdlm_$sub.so bayesian_dlms_b200/libbdlm.so
  echo "== kSub=$sub"
  python tools/scan_time.py --cases ${CASES:-2:24,2:21,2:20,1:21} | grep "^scan"
done
cp /tmp/libbdlm_keep.so bayesian_dlms_b200/libbdlm.so
