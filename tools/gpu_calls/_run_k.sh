O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_grad_check.py 2>&1 | grep -v "Warn\|warn" | tail -5 | tee $O/dp_grad_check_r02_b.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | tee $O/bench_r02_g_2gpu.json | cut -c1-300
timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | cut -c1-200
GPTST_B200_OPT_PREFETCH=1 timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | cut -c1-200
