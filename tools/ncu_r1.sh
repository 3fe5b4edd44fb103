# ncu captures for round 1 (run under gpurun, 1 GPU). Outputs land in gpurun_out/.
B="python bench.py --steps 1 --warmup 3 --frames-per-step 16 --e2e-frames 2 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 128 -c 150 --csv --log-file gpurun_out/launches_fused0.csv $B --fused 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 128 -c 60 --csv --log-file gpurun_out/launches_fused1.csv $B --fused 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 20 -c 2 -o gpurun_out/prof_frame_r1 -f $B --fused 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 20 -c 1 -o gpurun_out/prof_mac_r1 -f $B --fused 0 > /dev/null 2>&1
ls -la gpurun_out
