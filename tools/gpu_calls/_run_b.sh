O=gpurun_out; mkdir -p $O
timeout 150 python bench.py --workload metr_la --no-cpu-baseline > $O/bench_r02_metr_la.json 2> $O/bench_r02_metr_la.err
cut -c1-300 $O/bench_r02_metr_la.json
timeout 300 python bench.py --workload synthetic2048 --batch 16 --steps 10 --warmup 4 --no-cpu-baseline > $O/bench_r02_syn2048_b16.json 2> $O/bench_r02_syn2048_b16.err
cut -c1-300 $O/bench_r02_syn2048_b16.json; tail -3 $O/bench_r02_syn2048_b16.err
timeout 400 python bench.py --workload synthetic2048 --steps 6 --warmup 4 --no-cpu-baseline > $O/bench_r02_syn2048_b128.json 2> $O/bench_r02_syn2048_b128.err
cut -c1-300 $O/bench_r02_syn2048_b128.json; tail -3 $O/bench_r02_syn2048_b128.err
nvidia-smi --query-gpu=memory.used --format=csv
