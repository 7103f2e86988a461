#!/usr/bin/env bash
# compute-sanitizer over the small GPU tests (race / memory checks of the shared-memory
# histograms, the union-find labels, the peer-exchange flags and the mbarrier pipelines).
# ONE tool per GPU-box call (the profiling guide: running the four tools back to back in one call has left
# a B200 unusable):
#   gpurun --timeout 1200 -- 'bash tools/sanitize.sh memcheck  > gpurun_out/sanitize_memcheck.log 2>&1'
#   gpurun --timeout 1200 -- 'bash tools/sanitize.sh racecheck > gpurun_out/sanitize_racecheck.log 2>&1'
#   (synccheck, initcheck likewise)
# The matcher spins on mbarriers with a clock watchdog; under the sanitizer's slowdown only the small
# cases are run, each invocation under its own timeout.
set -u
cd "$(dirname "$0")/.."
tool="${1:-memcheck}"
case "$tool" in memcheck|racecheck|synccheck|initcheck) ;; *) echo "unknown tool $tool"; exit 2;; esac
# small cases only: the golden scenes through Hough / affine / clustering / the pipeline, and the matcher's
# ties / extreme-value / key-exchange / peer-exchange cases
SUITE=(tests/test_gpu_hough_affine.py tests/test_gpu_postprocess.py tests/test_gpu_pipeline.py tests/test_gpu_match.py)
SELECT='golden or rank_deficient or boundaries or residual or postprocess or pipeline_to_final or ties or extreme or key_exchange or peer_exchange'
echo "=== compute-sanitizer --tool $tool  ($(date -u +%FT%TZ))"
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
timeout 1000 compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 0 \
  python -m pytest "${SUITE[@]}" -x -q -m gpu -k "$SELECT" 2>&1 | tail -60
echo "exit: ${PIPESTATUS[0]}"
