#!/bin/bash
# Multi-GPU session: usage multi.sh <N> <tag> [list of rank counts, default "1 2 4 8"].  Runs the 2-rank NCCL
# tests (N >= 2) and the strong-scaling bench lines of configs 3, 4 and 5 at the listed rank counts <= N.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$1; O=gpurun_out/${2:-multi$N}; mkdir -p $O
nvidia-smi -L > $O/env.txt; nproc >> $O/env.txt
if [ "$N" -ge 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -rA > $O/pytest_gpu_multi.log 2>&1; echo "multi pytest rc=$?" >> $O/env.txt
  tail -15 $O/pytest_gpu_multi.log
fi
P=29500
for G in ${3:-1 2 4 8}; do
  [ "$G" -gt "$N" ] && continue
  for C in c3 c4 c5; do
    P=$((P+1))
    if [ "$G" -eq 1 ]; then
      timeout 900 python bench.py --config $C --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${C}_n$G.json 2> $O/bench_${C}_n$G.err
    else
      NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P \
        bench.py --config $C --gpus $G --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${C}_n$G.json 2> $O/bench_${C}_n$G.err
    fi
    echo "$C n=$G rc=$?" >> $O/env.txt
    grep -h "NCCL INFO" $O/bench_${C}_n$G.json 2>/dev/null | grep -i "nvls\|nranks" | head -6 > $O/nccl_${C}_n$G.txt
  done
done
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
  d=[l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1]
  d=json.loads(d)
  p=d.get("parity") or {}
  print({k:d[k] for k in ("value","ms_per_step","n_gpus") if k in d}, "parity max_rel_err", p.get("max_rel_err"), "pass", p.get("pass"))
except Exception as e:
  print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
