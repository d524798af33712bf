O=gpurun_out; mkdir -p $O
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_p.log
tail -2 $O/pytest_r02_p.log
timeout 120 python bench.py --no-cpu-baseline > $O/bench_r02_j.json 2> $O/bench_r02_j.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_j.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), 'cap', d['roofline']['ms'], d['roofline']['frac'], 'loss', d['last_loss'])
PY
timeout 100 python tools/step_timeline.py $O/step_timeline_r02_f.csv > $O/tl.log 2>&1; tail -1 $O/tl.log
