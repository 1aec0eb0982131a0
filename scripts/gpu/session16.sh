#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s16; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/env.txt
tail -4 $O/pytest_gpu.log
run() { tag=$1; cfg=$2; shift 2; ( for e in "$@"; do export $e; done; timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_${cfg}_$tag.json 2> $O/bench_${cfg}_$tag.err ); }
run sparse c3
run dense c3 QHBM_NO_SPARSE_INIT=1
run sparse c4
run dense c4 QHBM_NO_SPARSE_INIT=1
run sparse c3q
run sparse c3l7
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  p=d.get("parity") or {}
  print({k:d[k] for k in ("value","ms_per_step") if k in d}, d.get("config",{}).get("ms_per_4096_bitstrings"), "parity max_rel_err", p.get("max_rel_err"), "pass", p.get("pass"))
except Exception as e:
  print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
