O=gpurun_out; mkdir -p $O
echo "== TPW=2"; timeout 60 ./tools/kbench 2>&1 | grep -E "cap_route_fwd|cap_dv_dcr|route_bwd" 
echo "== TPW=1"; GPTST_B200_ROUTE_TPW=1 timeout 60 ./tools/kbench 2>&1 | grep -E "cap_route_fwd"
timeout 60 ./tools/cap_check > $O/cap_check_r02_e.log 2>&1; grep -E "chain \(us\)|time \(us\)" $O/cap_check_r02_e.log | tail -4
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_m.log
tail -2 $O/pytest_r02_m.log
