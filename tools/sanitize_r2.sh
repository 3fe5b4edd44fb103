# compute-sanitizer legs for the round-2 kernels: the job-list k_frame (general path), the tiled
# head term (partial_outputs / k_partial_tiles), the pipelined tails (slot-buffered rows), the batched
# IR ingest, the SpectralProcessor kernel, the chirp operator.  Small shapes: the tools slow kernels
# down 10-100 x.  Output: gpurun_out/r2_compute_sanitizer.txt
OUT=gpurun_out/r2_compute_sanitizer.txt
: > $OUT
run() {  # name, tool, pytest -k expression, file
    echo "== $2: $1" >> $OUT
    timeout 300 compute-sanitizer --tool $2 --error-exitcode 9 python -m pytest $4 -q -x -k "$3" > gpurun_out/_san.log 2>&1
    echo "exit $?" >> $OUT
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/_san.log | tail -4 >> $OUT
}
run "general path (k_frame_gen, k_partial_fused): utest shapes, random call sizes" memcheck "utest_small or utest_large or inplace_and_random" tests/test_gpu_parity.py
run "general path, shared-memory hazards of the tiled head term" racecheck "utest_small or inplace_and_random" tests/test_gpu_parity.py
run "dump / rank 16 tiny calls (k_partial_tiles), init_many, shared spectra" memcheck "dump_fields and (300000 or 129 or 31-9) or init_many" tests/test_gpu_parity.py
run "pipelined tails + early transform + cascades (small grids)" memcheck "cascaded and 16-9000 or early_input and 8-20000 or switching_between" tests/test_gpu_parity.py
run "pipelined tails, racecheck" racecheck "switching_between or (overlapped_launches and 7-50000)" tests/test_gpu_parity.py
run "SpectralProcessor kernel" memcheck "reference_utest_simple or 7-31 or 9-256 or 12-333 or rank_change" tests/test_gpu_spectral.py
run "SpectralProcessor kernel, racecheck" racecheck "reference_utest_simple or 9-256 or 14-8192" tests/test_gpu_spectral.py
run "chirp operator, shared IR, stream-ordered primitives" memcheck "chirp and (in_len0 or in_len3) or shared_impulse or parse_apply_restore and (8 or 13 or 16)" tests/test_gpu_primitives.py
run "synccheck: general path + spectral" synccheck "utest_small" tests/test_gpu_parity.py
cat $OUT
