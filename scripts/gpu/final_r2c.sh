#!/bin/bash
# After the row-per-thread coefficient preparation: GPU suite, rows timing, headline line with current counters.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-f26}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 120 python scripts/bench_rows.py > $O/bench_rows.json 2> $O/bench_rows.err; tail -1 $O/bench_rows.json
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_c3.json').read().strip().splitlines() if l.startswith('{')][-1]); print('c3', d['value'], d['config'].get('ms_per_4096_bitstrings'), d['parity']['pass'], d['e2e']['value'], d['roofline']['traffic'])"
