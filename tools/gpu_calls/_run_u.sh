for v in "GPTST_B200_PDL=0" "GPTST_B200_PDL=s" "GPTST_B200_PDL=0" "GPTST_B200_PDL=s"; do
  env $v timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['e2e_ms_per_step'],4), d['last_loss'])"
done
