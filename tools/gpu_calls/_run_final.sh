O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -rs 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -12 > $O/pytest_gpu_r02_final.log
tail -4 $O/pytest_gpu_r02_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke_r02_final.log | cut -c1-160
timeout 400 python bench.py > $O/bench_r02_final.json 2> $O/bench_r02_final.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), 'cap', d['roofline']['ms'], round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'cpu', d['cpu_baseline'], 'launches', d['gpu_launches'])
PY
