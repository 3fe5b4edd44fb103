# compute-sanitizer legs for the code added late in round 1: the equalizer kernel (TMA-staged),
# the staged transforms (multi-frame calls), the early pending MAC, the radix-8 first pass.
echo "== memcheck: equalizer tests (small ranks)"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_equalizer.py -q -x -k "7-31 or 8-256 or 9-77 or 10-1024 or handover and not 14 or clear or caller_stream" > gpurun_out/sanitizer_memcheck_eq.log 2>&1; echo "memcheck eq exit $?"
tail -4 gpurun_out/sanitizer_memcheck_eq.log
echo "== racecheck: equalizer (shared-memory hazards of the fused middle / TMA slots)"
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_equalizer.py -q -x -k "8-256 or 10-1024 or 8-256-256" > gpurun_out/sanitizer_racecheck_eq.log 2>&1; echo "racecheck eq exit $?"
tail -4 gpurun_out/sanitizer_racecheck_eq.log
echo "== memcheck: multi-frame calls (k_fwd_staged / k_inv_staged) and eager / early pending MAC"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "multi_frame_calls_on_many or eager or (early_pending and 5-40000)" > gpurun_out/sanitizer_memcheck_multi.log 2>&1; echo "memcheck multi exit $?"
tail -4 gpurun_out/sanitizer_memcheck_multi.log
echo "== memcheck: primitives on large batches (staged transforms)"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_primitives.py -q -x -k "large_batches and (11-2000 or 8-6000)" > gpurun_out/sanitizer_memcheck_prim.log 2>&1; echo "memcheck prim exit $?"
tail -4 gpurun_out/sanitizer_memcheck_prim.log
echo "== racecheck: smoke (radix-8 first pass in k_frame / k_fwd / k_inv)"
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke2.log 2>&1; echo "racecheck smoke exit $?"
tail -3 gpurun_out/sanitizer_racecheck_smoke2.log
echo "== synccheck: equalizer"
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_equalizer.py -q -x -k "8-256 or 10-1024" > gpurun_out/sanitizer_synccheck_eq.log 2>&1; echo "synccheck eq exit $?"
tail -3 gpurun_out/sanitizer_synccheck_eq.log
