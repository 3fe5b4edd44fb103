set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests8.log 2>&1; echo pytest-exit $?; tail -5 gpurun_out/gpu_tests8.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 20 "$@" 2>>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('RES', '$*', '| value %.1f M/s  frame %.2f us  iso-kernel %.2f us  frac %.3f  e2e %.1f' % (d['value']/1e6, d['ms_per_step']*1000/469, r['isolated_launch_ms']*1000, r['frac'], d['e2e']['value']/1e6))"; }
run --bias 3
run --bias 6 --stages 2
run --bias 6 --stages 2 --splits 8
run --bias 6 --stages 2 --splits 7
run --bias 6 --stages 2 --splits 6
run --bias 3 --stages 2 --splits 8
run --bias 10 --stages 2 --splits 8
