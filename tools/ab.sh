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
