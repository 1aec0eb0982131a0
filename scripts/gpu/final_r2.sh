#!/bin/bash
# Round-2 final single-GPU measurements: GPU suite, bench lines of every config, reference arm, ncu capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-final}; mkdir -p $O
nvidia-smi -L > $O/env.txt; nproc >> $O/env.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/env.txt
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for C in c1 c2 c3q c3l7 c4; do
  timeout 900 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$C.json 2> $O/bench_$C.err
done
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python scripts/bench_ebm.py > $O/bench_ebm.txt 2>&1
timeout 300 python scripts/bench_api.py 16 500 > $O/bench_api_500.txt 2>&1
timeout 300 python scripts/bench_api.py 16 100000 > $O/bench_api_100k.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/launches_c3.log 2>&1
cp qhbm-library_b200/libqhbm_b200.so $O/libqhbm_b200.so
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
  p=d.get("parity") or {}
  print({k:d[k] for k in ("value","ms_per_step","impl") if k in d}, d.get("config",{}).get("ms_per_4096_bitstrings"), "parity", p.get("max_rel_err"), p.get("pass"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
  print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
grep -v Warn $O/bench_api_500.txt | tail -1; grep -v Warn $O/bench_api_100k.txt | tail -1
