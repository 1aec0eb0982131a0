#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s8; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/env.txt
tail -4 $O/pytest_gpu.log
timeout 300 python scripts/debug_shard_draw.py > $O/shard_draw.txt 2>&1; tail -30 $O/shard_draw.txt
run() { tag=$1; cfg=$2; shift 2; ( for e in "$@"; do export $e; done; timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_${cfg}_$tag.json 2> $O/bench_${cfg}_$tag.err ); }
run lean c3
run nolean c3 QHBM_NO_LEAN=1
run lean c4
run nolean c4 QHBM_NO_LEAN=1
run lean c3q
run lean c2
run lean c3l7
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
cp qhbm-library_b200/libqhbm_b200.so $O/libqhbm_b200.so
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
