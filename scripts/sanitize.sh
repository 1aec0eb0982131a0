#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family (GPU box).
set -o pipefail
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "== $tool: expectation engine"
  $CS --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_expectation.py -q -x -m gpu \
      -k "hea_against_oracle or tfq_fd_mode or walsh or errors or empty or host_buffer or single_observable or observable_passes" 2>&1 | tail -4
  echo "== $tool: ebm + sampled"
  $CS --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_ebm.py tests/test_gpu_sampled.py -q -x -m gpu \
      -k "not uneven and not large_state and not moments and not full_size and not 2p" 2>&1 | tail -4
done
