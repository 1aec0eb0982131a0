#!/bin/bash
# Reduced final session: the configs whose plans changed (c3, c3l7) + the ncu capture / launch list of c3.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-final_short}; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --config c3l7 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3l7.json 2> $O/bench_c3l7.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/launches_c3.log 2>&1
cp qhbm-library_b200/libqhbm_b200.so $O/libqhbm_b200.so
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "
import json
for c in ('c3','c3l7'):
  d=json.loads([l for l in open('$O/bench_'+c+'.json').read().strip().splitlines() if l.startswith('{')][-1]); print(c, d['value'], d['config']['ms_per_4096_bitstrings'], d['parity']['pass'])"
