O=gpurun_out; mkdir -p $O
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x -rs 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_o.log
tail -9 $O/pytest_r02_o.log
timeout 120 python bench.py --no-cpu-baseline > $O/bench_r02_i.json 2> $O/bench_r02_i.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_i.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), 'cap', d['roofline']['ms'], d['roofline']['frac'], 'loss', d['last_loss'])
PY
