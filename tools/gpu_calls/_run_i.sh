O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_j.log
tail -2 $O/pytest_r02_j.log
timeout 200 python tools/exp_two_chains.py 2>&1 | grep -v Warn | tail -8 | tee $O/exp_two_chains.log
