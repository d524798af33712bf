#!/bin/bash
# One GPU-box call: parity tests, the bench line, A/B runs of the scheduling knobs, ncu evidence.  Outputs under gpurun_out/.
# Ordered by importance; every step has its own timeout so that a slow one cannot starve the rest.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
# stand-alone C++ checks first: no Python start-up, a few seconds each (build them here: see the header of each .cu)
[ -x tools/kbench ] && timeout 20 ./tools/kbench > $O/kbench.log 2>&1 && el "kbench: $(grep -c ' us ' $O/kbench.log) kernels timed"
[ -x tools/htem_check ] && timeout 20 ./tools/htem_check > $O/htem_check.log 2>&1 && el "htem_check done"
[ -x tools/cap_check ] && timeout 20 ./tools/cap_check > $O/cap_check.log 2>&1 && el "cap_check done"
timeout 80 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pt_b.log
el "pytest: $(tail -1 $O/pt_b.log)"
timeout 70 python bench.py > $O/bench_v.json 2> $O/bench_v.err
el "bench default"
timeout 50 python tools/ab_knobs.py > $O/ab_knobs.jsonl 2> $O/ab_knobs.err
el "ab knobs"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_graph.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rooflines --profiler-range > $O/ncu_list.log 2>&1
el "ncu list"
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
el "smoke: $(tail -1 $O/smoke.log | cut -c1-120)"
timeout 55 ncu --set full --clock-control none --import-source on -k regex:"gproj2_bwd|htem_bwd" -s 22 -c 6 -f -o $O/prof_bwd \
    python tools/prof_blocks.py all > $O/ncu_bwd.log 2>&1
el "ncu full"
python - <<'PY'
import json, glob
for f in ["gpurun_out/bench_v.json"]:
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l)
                e = d["e2e"]["value"] if isinstance(d["e2e"], dict) else d["e2e"]
                print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 4), round(e, 1), d["last_loss"], d.get("roofline", {}).get("ms"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat $O/ab_knobs.jsonl
tail -4 $O/pt_b.log
