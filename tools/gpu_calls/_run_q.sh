O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 240 bash tools/ncu_cap_traffic.sh r02b > $O/ncu_cap_traffic.log 2>&1; tail -1 $O/ncu_cap_traffic.log | cut -c1-300; el traffic
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_r02_b_graph_step.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rooflines --profiler-range > $O/ncu_list.log 2>&1; el "ncu list $(wc -l < $O/launches_r02_b_graph_step.csv) lines"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02.log 2>&1; el "smoke: $(tail -1 $O/smoke_r02.log | cut -c1-100)"
# racecheck on the stand-alone kernel harness (every heavy kernel of both blocks through the C ABI; shared-memory hazards)
timeout 280 compute-sanitizer --tool racecheck --print-limit 10 ./tools/kbench 2 170 2 > $O/sanitizer_racecheck_kbench_r02.log 2>&1; tail -3 $O/sanitizer_racecheck_kbench_r02.log; el racecheck
