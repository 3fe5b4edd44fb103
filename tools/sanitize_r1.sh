# compute-sanitizer legs (analogue of the reference CI's valgrind / ASAN legs, .github/workflows/build.yml:13,38-41)
T="python -m pytest tests/test_gpu_parity.py -q -x -k utest_small_or_golden_or_ragged_or_device_pointer_or_planar"
T="python -m pytest tests/test_gpu_parity.py -q -x -k"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke.log 2>&1; echo "$tool smoke exit $?"
  tail -4 gpurun_out/sanitizer_${tool}_smoke.log
done
echo "== memcheck on fused/overlap tests"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "device_pointer or ragged or planar or golden" > gpurun_out/sanitizer_memcheck_tests.log 2>&1; echo "memcheck tests exit $?"
tail -5 gpurun_out/sanitizer_memcheck_tests.log
echo "== racecheck on fused path"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "device_pointer" > gpurun_out/sanitizer_racecheck_tests.log 2>&1; echo "racecheck tests exit $?"
tail -5 gpurun_out/sanitizer_racecheck_tests.log
