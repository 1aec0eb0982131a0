#!/bin/bash
# Round-2 GPU session 1: packed-f32x2 kernels vs the round-1 build, FFMA2 microbenchmark, parity numbers.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s1; mkdir -p $O
nvidia-smi -L > $O/env.txt; nproc >> $O/env.txt; lscpu | grep -E "Model name|Flags" | cut -c1-600 >> $O/env.txt
(cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench ffma2_bench.cu && /tmp/ffma2_bench) > $O/ffma2.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/env.txt
R1=$PWD/qhbm-library_b200/libqhbm_b200_r1.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-parity-fail > $O/bench_c3_new.json 2> $O/bench_c3_new.err
QHBM_B200_LIB=$R1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_c3_r1.json 2> $O/bench_c3_r1.err
for T in 11 13; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-fail --tile-qubits $T > $O/bench_c3_T$T.json 2> $O/bench_c3_T$T.err
done
for C in c1 c2 c3q c4 c5; do
  timeout 900 python bench.py --config $C --steps 3 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_$C.json 2> $O/bench_$C.err
done
QHBM_B200_LIB=$R1 timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_c4_r1.json 2> $O/bench_c4_r1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 96 -c 4 -f -o $O/prof_c3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/prof_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/launches_c3.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print({k:d[k] for k in ("value","ms_per_step") if k in d}, d.get("config",{}).get("ms_per_4096_bitstrings"), json.dumps(d.get("parity"))[:900])
except Exception as e:
  print("ERR", e)
PY
done
cat $O/ffma2.txt
