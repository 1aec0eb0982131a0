#!/bin/bash
# Last call of round 2: GPU suite + the VQT step through the API after the host-side changes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${1:-f27}; mkdir -p $O
timeout 80 python -m pytest tests -m gpu -q -x --timeout 60 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
(timeout 25 python scripts/bench_api.py 16 500 | tail -1; timeout 25 python scripts/bench_api.py 16 100000 | tail -1) > $O/bench_api.txt 2>&1; cat $O/bench_api.txt
