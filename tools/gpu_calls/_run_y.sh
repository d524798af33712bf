O=gpurun_out; mkdir -p $O
timeout 60 ./tools/cap_check 2>&1 | grep -E "chain \(us\)|time \(us\)" | tail -2
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_t.log
tail -2 $O/pytest_r02_t.log
timeout 200 python bench.py --no-cpu-baseline > $O/bench_r02_k.json 2> $O/bench_r02_k.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_k.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), 'cap', d['roofline']['ms'], round(d['roofline']['frac'],4), 'htem_fwd', d['roofline_hypertem_fwd']['ms'], 'loss', d['last_loss'])
PY
