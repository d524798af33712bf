O=gpurun_out; mkdir -p $O
for mc in 8 32; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc"
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 100 python tools/step_timeline.py $O/step_timeline_r02_d_mc32.csv > $O/tl.log 2>&1; tail -1 $O/tl.log
