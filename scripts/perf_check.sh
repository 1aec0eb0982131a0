#!/bin/bash
# Development loop on the GPU box: parity of the expectation engine, then the timings that matter.
python -m pytest tests/test_gpu_expectation.py tests/test_golden_fixtures.py -q -x -m gpu 2>&1 | tail -3
python scripts/profile_case.py 16 2 4096 12 4 1 xxz 5
python scripts/profile_case.py 16 2 4096 0 0 0 xxz 5
python scripts/profile_case.py 16 2 4096 0 0 0 tfim 5
python scripts/profile_case.py 20 2 2048 0 0 0 tfim 3
python scripts/profile_case.py 20 2 2048 0 0 0 xxz 3
QHBM_ONE_STAGE=1 python scripts/profile_case.py 20 2 2048 0 0 0 tfim 3
QHBM_NO_HPASS=1 python scripts/profile_case.py 20 2 2048 0 0 0 tfim 3
