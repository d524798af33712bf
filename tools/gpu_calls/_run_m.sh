O=gpurun_out; mkdir -p $O
GPTST_B200_ROUTE_TPW=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:cap_route2_fwd -s 3 -c 1 -f -o $O/prof_route_tpw1 ./tools/kbench > $O/ncu_m1.log 2>&1; tail -2 $O/ncu_m1.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cap_route2_fwd -s 3 -c 1 -f -o $O/prof_route_tpw2 ./tools/kbench > $O/ncu_m2.log 2>&1; tail -2 $O/ncu_m2.log
ls -la $O/*.ncu-rep
