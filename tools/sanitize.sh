#!/bin/bash
# compute-sanitizer passes over smoke() and one forced-lanes session test.
# Usage (on the GPU box): bash tools/sanitize.sh ; logs land in gpurun_out/sanitizer_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export EUT_NO_GRAPHS=1   # the sanitizer tracks kernels launched directly; graph replays are the same kernels
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke exit $?" | tee -a gpurun_out/sanitizer_summary.txt
done
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_diffusion_gpu.py -x -q -m gpu -k "session_lanes_match_oracle or two_lane_chunk" \
    > gpurun_out/sanitizer_${tool}_lanes.log 2>&1
  echo "$tool lanes exit $?" | tee -a gpurun_out/sanitizer_summary.txt
done
tail -n 6 gpurun_out/sanitizer_*.log
