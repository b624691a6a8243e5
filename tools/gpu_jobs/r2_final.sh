mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_pass_ncu.csv python tools/one_pass.py --passes 1 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
python tools/pass_timeline.py --quiet > gpurun_out/r2_timeline_final.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -1 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --config coco --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_coco_final.json 2>> gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
for n in ('final','coco_final'):
    d=json.load(open(f'gpurun_out/r2_bench_{n}.json')); e=d['e2e']
    print(n, 'value', round(d['value'],2), round(d['ms_per_pass'],3), d['clocks']['sm_mhz'], '| e2e', round(e['value'],2), round(e['ms_per_pass'],3), e['clocks']['sm_mhz'], 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), 'edm', round(d['roofline_edm_step']['frac'],3))
d=json.load(open('gpurun_out/r2_bench_reference_arm.json')); print('reference arm', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
