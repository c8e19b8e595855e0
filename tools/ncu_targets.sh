#!/bin/bash
# Targeted `ncu --set full` captures of the kernels that carry the step (run under gpurun, one GPU).
# Reports stay on the box (/tmp); only text exports (details page, raw CSV, per-instruction source CSV) come back.
set -u
CMD="python tools/profile_step.py --clips 128 --plan 2"
OUT=gpurun_out
$CMD > $OUT/r2_ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/r2_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $OUT/r2_ncu_$1_details.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/r2_ncu_$1_raw.csv 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/r2_ncu_$1_source.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap pw "tap_gemm_tc_kernel<\(int\)256, \(int\)1," 10 2
cap k3 "tap_gemm_tc_kernel<\(int\)256, \(int\)3, \(bool\)0, \(int\)2" 24 1
cap n64 "tap_gemm_tc_kernel<\(int\)64, \(int\)3," 0 3
cap n128 "tap_gemm_tc_kernel<\(int\)128, \(int\)3," 0 1
if [ "${1:-all}" = "all" ]; then
cap rb0 "resblock0_fused_kernel" 0 1
cap dw "dwconv_ln_kernel" 2 1
fi
ls -la $OUT | grep r2_ncu | head -40
