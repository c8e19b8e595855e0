#!/bin/bash
# Round-2 (last session) profiling recipe, run under gpurun on one GPU: bash tools/ncu_round2_mem.sh
#  1. launch list of one full step (256 clips, 128-clip encoder passes, one 256-clip decoder pass)
#  2. `ncu --set full` captures of the memory-bound kernels that changed in this session
# Only text exports come back (details page, raw CSV).
set -u
OUT=gpurun_out
P=${NCU_PREFIX:-r2h}
python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_ps256.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${P}_launches_256.csv python tools/profile_step.py --clips 256 --plan 2 > $OUT/${P}_launches.log 2>&1
CMD="python tools/profile_step.py --clips 256 --plan 2"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/${P}_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > $OUT/${P}_ncu_$1_details.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${P}_ncu_$1_raw.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap dw "dwconv_ln_hi2_kernel" 2 1
cap gn "groupnorm_kernel" 2 1
cap spec "spectral_rows_kernel" 0 1
cap ola "overlap_add4_kernel" 0 1
python tools/ncu_table.py $OUT/${P}_ncu_*_raw.csv > $OUT/${P}_ncu_table.txt 2>&1
ls $OUT | grep ${P}_ | head -40
