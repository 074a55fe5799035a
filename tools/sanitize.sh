#!/bin/bash
# The reference's sanitizer sweep (vision-gpu/src/lib.rs:703-728: compute-sanitizer --tool {memcheck, racecheck,
# synccheck, initcheck} --report-api-errors all --error-exitcode 1 around its one smoke test) over the parity tests of
# this backend: the trait path on the golden synthetic cases and the two concurrent branches, the device batch and the
# host batch.  usage: tools/sanitize.sh <tag>   (run on the GPU box; summaries land in gpurun_out/<tag>_sanitize_*.txt)
tag=${1:-san}
mkdir -p gpurun_out
sel='test_golden_synthetic_cases or test_two_branches_concurrently or test_map_closed_and_threshold_boundary or test_debug_views or test_strip_seams or test_random_noise_frames or test_batch_with_closed_maps_and_rgb or test_host_batch or test_stream_over_a_ring or test_ui_map_as_its_l_plane'
status=0
for tool in memcheck racecheck synccheck initcheck; do
  log=gpurun_out/${tag}_sanitize_${tool}.txt
  timeout 1500 compute-sanitizer --tool $tool --report-api-errors all --error-exitcode 1 \
     python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_multi.py -m gpu -x -q -k "$sel" > $log 2>&1
  rc=$?
  echo "== $tool: exit code $rc" | tee -a $log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $log | tail -4
  [ $rc -ne 0 ] && status=1
done
exit $status
