O=gpurun_out; mkdir -p $O
timeout 60 ./tools/kbench 2>&1 | grep -E "cap_route" 
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_n.log
tail -2 $O/pytest_r02_n.log
timeout 120 python bench.py --no-cpu-baseline > $O/bench_r02_h.json 2> $O/bench_r02_h.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_h.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cap', d['roofline']['ms'], d['roofline']['frac'], 'loss', d['last_loss'])
PY
GPTST_B200_CAP_Z=recompute timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | cut -c1-120
