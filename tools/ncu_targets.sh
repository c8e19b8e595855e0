#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): bash tools/ncu_targets.sh [all|gemm]
#  1. launch list of one full step (256 clips): every launch with its device time (cold-cache, serialised: compare SHARES)
#  2. targeted `ncu --set full` captures of the kernels that carry the step (128-clip step = one decoder chunk).
# Reports stay on the box (/tmp); only text exports (details page, raw CSV, per-instruction stall CSV) come back.
set -u
OUT=gpurun_out
P=${NCU_PREFIX:-r2f}
python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_ps256.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${P}_launches_256.csv python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_launches.log 2>&1
CMD="python tools/profile_step.py --clips 128 --plan 2"
$CMD > $OUT/${P}_ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/${P}_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $OUT/${P}_ncu_$1_details.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${P}_ncu_$1_raw.csv 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/${P}_ncu_$1_source.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap gemm1 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)2" 5 1      # ConvNeXt GEMM-1 (GELU-only epilogue), CTA pairs
cap gemm2 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)0" 5 1      # ConvNeXt GEMM-2 (layer scale + residual)
cap k3 "tap_gemm_tc_kernel<\(int\)256, \(int\)3, \(int\)0, \(int\)2" 12 2   # decoder k3 convs 768 -> 768 (3-pass, CTA pairs)
if [ "${1:-all}" = "all" ]; then
cap n64 "tap_gemm_tc_kernel<\(int\)64, \(int\)3," 0 3      # encoder: level-0 strided, level-1 tail, level-2 k3
cap n128 "tap_gemm_tc_kernel<\(int\)128, \(int\)3," 0 1    # encoder: level-1 strided
cap rb0 "resblock0_fused_kernel" 0 1
cap dw "dwconv_ln_kernel" 2 1
cap gn "groupnorm_kernel" 2 1
cap lstm "lstm_persistent_kernel" 1 1
fi
ls -la $OUT | grep ${P}_ | head -60
