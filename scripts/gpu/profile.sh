#!/bin/bash
# ncu --set full capture of one bench step (config 3) + launch list; usage: profile.sh <tag>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-prof}; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/launches_c3.log 2>&1
cp qhbm-library_b200/libqhbm_b200.so $O/libqhbm_b200.so
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
tail -c 600 $O/bench_c3.json
