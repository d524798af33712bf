O=gpurun_out; mkdir -p $O
timeout 60 ./tools/kbench 2>&1 | grep -E "score_head"
timeout 60 ./tools/cap_check 2>&1 | grep -E "chain \(us\)" | tail -1
GPTST_B200_PDL=0 timeout 60 ./tools/cap_check 2>&1 | grep -E "chain \(us\)" | tail -1
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_s.log
tail -2 $O/pytest_r02_s.log
for v in "GPTST_B200_PDL=0" "GPTST_B200_PDL=1" "GPTST_B200_PDL=0" "GPTST_B200_PDL=1"; do
  env $v timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['e2e_ms_per_step'],4), d['last_loss'])"
done
