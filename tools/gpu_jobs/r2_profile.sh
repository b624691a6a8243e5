mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -2 gpurun_out/r2_bench_a.err
python tools/pass_timeline.py --quiet > gpurun_out/r2_timeline_a.txt 2>&1
python tools/pass_timeline.py --quiet --dense > gpurun_out/r2_timeline_a_dense.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_pass_ncu.csv python tools/one_pass.py --passes 1 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
# EDM step kernels: a short sampler run (eager) under ncu --set full, only the edm kernels
ncu --set full --clock-control none --import-source on -k regex:edm_ -c 8 -o gpurun_out/r2_edm python tools/gpu_jobs/edm_only.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
