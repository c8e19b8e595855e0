#!/bin/bash
# dram traffic per launch of the modal launch shapes after the pass-size change (128-clip encoder passes, one 256-clip
# decoder pass): one `ncu --set full` capture each, text exports only. Run under gpurun: bash tools/ncu_round2_traffic.sh
set -u
OUT=gpurun_out
P=${NCU_PREFIX:-r2i}
CMD="python tools/profile_step.py --clips 256 --plan 2"
cap() {  # name regex skip count
  ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/${P}_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $OUT/${P}_ncu_$1_details.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${P}_ncu_$1_raw.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap k3 "tap_gemm_tc_kernel<\(int\)256, \(int\)3, \(int\)0, \(int\)2" 14 1   # decoder k3 conv 768 -> 768, 256 clips
cap gemm1 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)2" 5 1          # ConvNeXt GEMM-1
cap gemm2 "tap_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)0" 5 1          # ConvNeXt GEMM-2
cap n128 "tap_gemm_tc_kernel<\(int\)128, \(int\)3," 0 1                      # encoder level-1 strided conv, 128 clips
cap l0tc "enc_l0_tc_kernel" 0 1
cap l1f "enc_l1_fused_kernel" 0 1
python tools/ncu_table.py $OUT/${P}_ncu_*_raw.csv > $OUT/${P}_ncu_table.txt 2>&1
cat $OUT/${P}_ncu_table.txt
