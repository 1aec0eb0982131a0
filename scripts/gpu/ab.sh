#!/bin/bash
# A/B timing of library variants on config 3 (and c3q): usage ab.sh <outdir> [lib.so[:ENV=1,ENV2=1] ...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/$1; shift; mkdir -p $O
nvidia-smi -L > $O/env.txt
for spec in "$@"; do
  lib=${spec%%:*}; envs=""; [[ "$spec" == *:* ]] && envs=${spec#*:}
  tag=$(basename $lib .so)_$(echo "$envs" | tr ',=' '__')
  for C in c3 c3q; do
    ( export QHBM_B200_LIB=$PWD/$lib; for e in ${envs//,/ }; do export $e; done
      timeout 600 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_${C}_$tag.json 2> $O/bench_${C}_$tag.err )
  done
done
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
