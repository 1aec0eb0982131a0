#!/bin/bash
# ab.sh over three configs: usage ab3.sh <outdir> lib1.so lib2.so
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/$1; shift; mkdir -p $O
for lib in "$@"; do
  tag=$(basename $lib .so)
  for C in c3 c3q c3l7; do
    QHBM_B200_LIB=$PWD/$lib timeout 600 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline --no-parity-fail > $O/bench_${C}_$tag.json 2> $O/bench_${C}_$tag.err
    python -c "
import json; d=json.loads(open('$O/bench_${C}_$tag.json').read().strip().splitlines()[-1]); print('$C', '$tag', round(d['config']['ms_per_4096_bitstrings'],3), d['parity']['pass'])"
  done
done
