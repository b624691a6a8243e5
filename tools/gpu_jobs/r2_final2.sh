mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gputests_r2_final.log; cat gpurun_out/gputests_r2_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches_ncu.csv python tools/train_one_step.py --steps 1 > gpurun_out/ncu_train_launches.log 2>&1
tail -1 gpurun_out/ncu_train_launches.log
python tools/train_launch_summary.py gpurun_out/r2_train_launches_ncu.csv > gpurun_out/r2_train_launch_summary.txt; head -12 gpurun_out/r2_train_launch_summary.txt
