O=gpurun_out; mkdir -p $O
timeout 60 ./tools/kbench 2>&1 | grep -E "gproj|linear_bwd|hypertem dW" 
timeout 250 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_r.log
tail -2 $O/pytest_r02_r.log
timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | cut -c1-130
