#!/bin/bash
# End-of-round-2 profiling recipe (run under gpurun, one GPU): bash tools/ncu_round2_final.sh
#  1. launch list of one full step (256 clips): every launch with its device time (cold-cache, serialised: compare SHARES)
#  2. `ncu --set full` captures of the kernels that changed this round (64-clip encoder chunk / 128-clip step)
# Only text exports come back (details page, raw CSV, per-instruction stall CSV).
set -u
OUT=gpurun_out
P=${NCU_PREFIX:-r2g}
python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_ps256.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${P}_launches_256.csv python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_launches.log 2>&1
CMD="python tools/profile_step.py --clips 128 --plan 2"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/${P}_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $OUT/${P}_ncu_$1_details.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${P}_ncu_$1_raw.csv 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/${P}_ncu_$1_source.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap l0tc "enc_l0_tc_kernel" 0 1
cap l1f "enc_l1_fused_kernel" 0 1
cap lstm "lstm_persistent_kernel" 1 1
cap n128 "tap_gemm_tc_kernel<\(int\)128, \(int\)3," 0 1    # encoder: level-1 strided conv
cap gemm1 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)2" 5 1
cap gemm2 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)0" 5 1
cap k3 "tap_gemm_tc_kernel<\(int\)256, \(int\)3, \(int\)0, \(int\)2" 12 1
python tools/ncu_table.py $OUT/${P}_ncu_*_raw.csv > $OUT/${P}_ncu_table.txt 2>&1
for k in l0tc l1f lstm gemm1; do echo "#### $k"; python tools/ncu_stalls.py $OUT/${P}_ncu_${k}_source.csv 12; done > $OUT/${P}_ncu_stalls.txt 2>&1
ls $OUT | grep ${P}_ | head -40
