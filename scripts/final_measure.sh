#!/bin/bash
# Round-end measurement set on one B200 (outputs under gpurun_out/; scripts/make_profiles.py summarises).
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for c in c1 c2 c3l7 c3q c4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
python scripts/bench_ebm.py > gpurun_out/bench_ebm.log 2>&1
(python scripts/bench_api.py 16 500 | tail -2; python scripts/bench_api.py 16 100000 | tail -2) > gpurun_out/bench_api.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/bench_launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 24 -c 4 -f -o gpurun_out/prof_r1_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_final.log 2>&1
timeout 900 bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
tail -n 3 gpurun_out/bench_c3.json gpurun_out/bench_ref.json gpurun_out/bench_c4.json gpurun_out/bench_api.log
