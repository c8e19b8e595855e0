#!/bin/bash
# One targeted `ncu --set full` capture (run under gpurun, one GPU): bash tools/ncu_one.sh <name> <kernel regex> <skip> <count>
# Text exports only (details page, raw CSV, per-instruction stall CSV) come back under gpurun_out/.
set -u
OUT=gpurun_out
P=${NCU_PREFIX:-r3}
CMD="python tools/profile_step.py --clips ${NCU_CLIPS:-128} --plan 2"
$CMD > $OUT/${P}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${P}_ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:$2" -s $3 -c $4 -f -o /tmp/$1 $CMD > $OUT/${P}_ncu_$1.log 2>&1
ncu -i /tmp/$1.ncu-rep --page details > $OUT/${P}_ncu_$1_details.txt 2>&1
ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${P}_ncu_$1_raw.csv 2>&1
ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/${P}_ncu_$1_source.csv 2>&1
rm -f /tmp/$1.ncu-rep
python tools/ncu_stalls.py $OUT/${P}_ncu_$1_source.csv 2>&1 | head -60
