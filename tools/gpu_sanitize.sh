# compute-sanitizer passes over the small end of the GPU test-suite (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards in the staged kernels; synccheck: barrier misuse). Summaries into gpurun_out/.
export PYTHONUNBUFFERED=1
SEL="tests/test_raster_gpu.py::test_edge_cases tests/test_raster_gpu.py::test_sh_degrees_and_background tests/test_raster_gpu.py::test_equal_depth_ties_follow_gaussian_index tests/test_mapoptim_gpu.py::test_map_step_matches_torch_activations_and_adam tests/test_mapsurgery_gpu.py"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -x -q > gpurun_out/sanitizer_$tool.pytest 2>&1
  echo "== $tool rc=$?"; tail -2 gpurun_out/sanitizer_$tool.pytest; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Misaligned" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -10
done
