mkdir -p gpurun_out
python tools/pass_timeline.py --quiet > gpurun_out/r2_timeline_c.txt 2>&1
