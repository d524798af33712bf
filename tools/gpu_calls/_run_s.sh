for v in 0 1 0 1; do
  GPTST_B200_DW_SLIM=$v timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('DW_SLIM=$v', round(d['ms_per_step'],4), round(d['e2e_ms_per_step'],4))"
done
