O=gpurun_out; mkdir -p $O
timeout 20 ./tools/kbench > $O/kbench_r02_c.log 2>&1
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_e.log
tail -3 $O/pytest_r02_e.log
timeout 120 python bench.py > $O/bench_r02_c.json 2> $O/bench_r02_c.err
cat $O/bench_r02_c.json | cut -c1-400
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_r02_graph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rooflines --profiler-range > $O/ncu_list_r02.log 2>&1
wc -l $O/launches_r02_graph.csv
timeout 60 python tools/step_timeline.py $O/step_timeline_r02_b.csv 200 2>&1 | tail -1
