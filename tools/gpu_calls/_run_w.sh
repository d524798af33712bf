O=gpurun_out; mkdir -p $O
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gproj2_bwd -s 6 -c 2 -f -o $O/prof_gproj2_bwd_r02 ./tools/kbench > $O/ncu_w.log 2>&1; tail -1 $O/ncu_w.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:htem_bwd -s 3 -c 1 -f -o $O/prof_htem_bwd_r02 ./tools/kbench > $O/ncu_w2.log 2>&1; tail -1 $O/ncu_w2.log
