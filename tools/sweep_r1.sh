set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests5.log 2>&1; echo pytest-exit $?; tail -5 gpurun_out/gpu_tests5.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 20 "$@" 2>>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('RES', '$*', '| value %.1f M/s  frame %.2f us  kernel %.2f us  frac %.3f  e2e %.1f' % (d['value']/1e6, d['ms_per_step']*1000/469, r['launch_ms']*1000, r['frac'], d['e2e']['value']/1e6))"; }
run --fused 1 --bias 3
run --fused 1 --bias 0
run --fused 1 --bias 2
run --fused 1 --bias 5
run --fused 1 --bias 8
run --fused 1 --bias 3 --pdl 0
run --fused 1 --bias 3 --stages 4
run --fused 1 --bias 3 --splits 8
run --fused 1 --bias 3 --splits 7 --stages 4
run --fused 0
