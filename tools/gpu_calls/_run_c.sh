O=gpurun_out; mkdir -p $O
timeout 200 python tools/step_timeline.py $O/step_timeline_r02_syn2048_b16.csv 200 synthetic2048 16 2>&1 | tail -1
