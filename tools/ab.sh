# Same-box A/B helper for gpurun sessions:  source tools/ab.sh; run <label> ENV=VALUE ...
# Runs `bench.py --quick` (10 steps) with the given environment, keeps the JSON line under gpurun_out/ab_<label>.json and
# prints step time, LSTM / encoder category times and the median SM clock. Variants are compared on ONE box in one call
# because boxes of the pool differ by a few per cent (sw_power_cap, 1.65 - 1.95 GHz).
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), "lstm", d["breakdown"]["lstm"]["ms"], "enc", d["breakdown"]["enc_conv"]["ms"], d["clocks"]["sm_mhz"])
except Exception as e: print("$label ERR", e, open("gpurun_out/ab_$label.err").read()[-600:])
PY
}
run2() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$label.json").read().strip().splitlines()[-1])
    print("$label", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "enc", d["breakdown"]["enc_conv"]["ms"], d["clocks"]["sm_mhz"])
except Exception as e: print("$label ERR", e, open("gpurun_out/ab_$label.err").read()[-600:])
PY
}
