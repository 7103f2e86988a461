#!/usr/bin/env bash
# compute-sanitizer over the small GPU tests (race / memory checks of the shared-memory
# histograms, the union-find labels and the mbarrier pipelines).  Run on a GPU box:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# The matcher spins on mbarriers with a clock watchdog; under the sanitizer's slowdown only the small
# cases are run, each tool under its own timeout.
set -u
cd "$(dirname "$0")/.."
TESTS="tests/test_gpu_hough_affine.py tests/test_gpu_postprocess.py tests/test_gpu_pipeline.py"
MATCH='tests/test_gpu_match.py -k "ties or extreme or key_exchange"'
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 0 \
    python -m pytest $TESTS -x -q -m gpu 2>&1 | tail -25
  echo "exit: $?"
done
echo "=== compute-sanitizer --tool memcheck (matcher, small cases)"
eval timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $MATCH -x -q -m gpu 2>&1 | tail -25
