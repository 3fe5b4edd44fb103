run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-frames 20 "$@" 2>>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('RES', '$*', '| value %.1f M/s  frame %.2f us  iso-kernel %.2f us  frac %.3f  e2e %.1f' % (d['value']/1e6, d['ms_per_step']*1000/469, r['isolated_launch_ms']*1000, r['frac'], d['e2e']['value']/1e6))"; }
run
run --stages 3
run --splits 6
run --splits 7
run --splits 9
run --splits 10
run --bias 3
run --bias 10
run --stages 3 --splits 6
