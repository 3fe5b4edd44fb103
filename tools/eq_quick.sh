#!/bin/bash
# quick check of the equalizer path: parity tests + throughput table
timeout 400 python -m pytest tests/test_gpu_equalizer.py -x -q 2>&1 | tail -3
timeout 200 python tools/bench_eq.py --no-cpu "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['instances'], d['fir_rank'], '%.1f G/s %.1f us/block share %.2f' % (d['device_samples_per_s']/1e9, d['device_us_per_block'], d['share_of_hbm_roofline']))
    else: print(l, end='')
"
