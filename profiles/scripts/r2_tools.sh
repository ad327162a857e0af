# Device-resident numbers of the tools (round 2): bench.py --tool ... one line each.
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  timeout -k 5 300 python bench.py --tool $1 --shape $2 --lattices $3 --steps ${4:-5} --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/tb.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1 $2 $3', 'ms/step %.2f'%d['ms_per_step'], 'arcs/s %.3g'%d['value'], 'entries %d'%d['batch']['index_entries'], {k:round(v['ms_per_launch']*v['launches_per_step'],2) for k,v in d['roofline']['kernels'].items()})
except Exception as ex:
    print('$1 $2 $3 FAILED', ex)
"
  tail -2 gpurun_out/tb.err | cut -c1-300
done
