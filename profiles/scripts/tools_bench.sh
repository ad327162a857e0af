for cfg in "fwd_bwd c2 10000 300" "segment c2 3000 200" "prune_dyn_beam c2 3000 200" "best_path2 c2 2000 100" "utterance c2 300 30" "position c2 200 20" "segment c4 300 30"; do
  set -- $cfg
  timeout -k 5 240 python bench.py --tool $1 --shape $2 --lattices $3 --ref-lattices $4 --steps 5 --warmup 3 --e2e-steps 1 2>gpurun_out/tb.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    cb=d.get('cpu_baseline') or {}
    e=d.get('e2e') or {}
    print('$1 $2 $3', 'ms/step %.2f'%d['ms_per_step'], 'arcs/s %.3g'%d['value'], 'cpu %.3g (%s thr)'%(cb.get('value') or 0, cb.get('cores')), 'e2e %.3g'%(e.get('value') or 0), {k:round(v['ms_per_launch']*v['launches_per_step'],2) for k,v in d['roofline']['kernels'].items()})
except Exception as ex:
    print('$1 $2 $3 FAILED', ex)
"
  tail -1 gpurun_out/tb.err | cut -c1-200
done
