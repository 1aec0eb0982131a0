#!/bin/bash
# Round-2 GPU session 5: diagonal-gate gradients through marginal vectors + reduction tasks + descriptors.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-s5}; mkdir -p $O
nvidia-smi -L > $O/env.txt; nproc >> $O/env.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/env.txt
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_c3.json 2> $O/bench_c3.err
for C in c2 c3q c3l7; do
  timeout 900 python bench.py --config $C --steps 3 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_$C.json 2> $O/bench_$C.err
done
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
