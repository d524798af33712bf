O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_h.log
tail -3 $O/pytest_r02_h.log
timeout 120 python bench.py --no-cpu-baseline > $O/bench_r02_e.json 2> $O/bench_r02_e.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_e.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cap', d['roofline']['ms'], d['roofline']['frac'], 'loss', d['last_loss'])
PY
timeout 100 python tools/step_timeline.py $O/step_timeline_r02_c.csv > $O/tl.log 2>&1; tail -2 $O/tl.log
# memcheck of one small eager step + graph steps (smoke geometry): item 9 of the round-1 verdict
timeout 280 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1; tail -4 $O/sanitizer_memcheck_smoke.log
