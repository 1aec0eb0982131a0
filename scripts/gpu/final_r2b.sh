#!/bin/bash
# Last session of round 2: GPU suite + headline bench + rows timing + ncu capture / launch list of c3 on the final build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-f25}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 120 python scripts/bench_rows.py > $O/bench_rows.json 2> $O/bench_rows.err; tail -1 $O/bench_rows.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/launches_c3.log 2>&1
cp qhbm-library_b200/libqhbm_b200.so $O/libqhbm_b200.so
python -c "
import json
d=json.loads([l for l in open('$O/bench_c3.json').read().strip().splitlines() if l.startswith('{')][-1]); print('c3', d['value'], d['config'].get('ms_per_4096_bitstrings'), d['parity']['pass'], d['e2e']['value'])"
