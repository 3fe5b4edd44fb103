# compute-sanitizer legs for what changed late in round 2: the spread direct-form answers and cp.async
# staging of the job-list launch, the row fold of the eager pending MAC, the radix-16 passes of the
# in-place transforms (ranks 14..16), k_mac with several partitions per stage on small bin tiles.
OUT=gpurun_out/r2_compute_sanitizer_late.txt
: > $OUT
run() {  # name, tool, pytest -k expression, file
    echo "== $2: $1" >> $OUT
    timeout 400 compute-sanitizer --tool $2 --error-exitcode 9 python -m pytest $4 -q -x -k "$3" > gpurun_out/_san.log 2>&1
    echo "exit $?" >> $OUT
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/_san.log | tail -4 >> $OUT
}
run "job-list launch: spread answers in place, cascade behind it" memcheck "spread_direct_form and (2-60000 or 8-30000 or 70-9000) or cascade_behind" tests/test_gpu_parity.py
run "job-list launch: shared-memory hazards (cp.async staging, answers under the inverse)" racecheck "spread_direct_form and (2-60000 or 8-30000) or utest_small" tests/test_gpu_parity.py
run "eager pending MAC with folded rows" memcheck "eager_pending_mac or early_pending_mac and 5-40000" tests/test_gpu_parity.py
run "eager pending MAC with folded rows, racecheck" racecheck "eager_pending_mac and False-1" tests/test_gpu_parity.py
run "radix-16 passes, ranks 14..16; small MAC bin tiles" memcheck "smaller_mac_bin_tiles or (matches_oracle_any_rank and (9000-14 or 40000-15 or 40000-16))" tests/test_gpu_parity.py
run "radix-16 passes, racecheck" racecheck "smaller_mac_bin_tiles" tests/test_gpu_parity.py
run "synccheck: job-list launch" synccheck "spread_direct_form and 2-60000" tests/test_gpu_parity.py
cat $OUT
