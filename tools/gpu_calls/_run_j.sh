O=gpurun_out; mkdir -p $O
GPTST_B200_CAPTURE_DEBUG=1 timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -80 > $O/pytest_r02_k.log
grep -n "INVALIDATED\|ops.py:[0-9]*: in\|GPTST.py:[0-9]*: in\|train.py:[0-9]*: in\|optim.py:[0-9]*: in\|passed\|failed" $O/pytest_r02_k.log | head -30
