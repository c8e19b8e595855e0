source tools/ab.sh
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k=d['kernels']
print(sys.argv[1], {n:k[n]['ms'] for n in ('dwconv_ln_kernel','groupnorm_kernel') if n in k}, 'mem', d['breakdown']['memory_bound']['ms'], 'pw', d['breakdown']['pwconv']['ms'])
PY
}
timeout 300 python -m pytest tests/test_fallback_paths.py -m gpu -x -q -k memory 2>&1 | tail -8
run2 n1; show n1
run2 o1 WT_MEM_V1=1; show o1
run2 n2; show n2
