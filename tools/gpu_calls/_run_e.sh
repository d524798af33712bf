O=gpurun_out; mkdir -p $O
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_g.log
tail -4 $O/pytest_r02_g.log
timeout 150 python bench.py > $O/bench_r02_d.json 2> $O/bench_r02_d.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_d.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cap', d['roofline']['ms'], d['roofline']['frac'], 'loss', d['last_loss'])
PY
GPTST_B200_CAP=split timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | cut -c1-200
timeout 200 bash tools/ncu_cap_traffic.sh r02 2>&1 | tail -2
