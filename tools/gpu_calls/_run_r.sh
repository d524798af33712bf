O=gpurun_out; mkdir -p $O
for n in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline --no-rooflines 2>$O/bench_${n}gpu.err | grep '^{' | tee $O/bench_r02_e_${n}gpu.json | cut -c1-260
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | grep '^{' > $O/bench_r02_f_8gpu_full.json; cut -c1-400 $O/bench_r02_f_8gpu_full.json
