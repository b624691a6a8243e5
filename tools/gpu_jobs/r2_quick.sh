mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "loss_backward or final_step" 2>&1 | tail -3
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -2 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d['clocks'])
PY
for B in 8 64; do timeout 600 python bench.py --steps 2 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/r2_bench_b$B.json 2>> gpurun_out/r2_bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b$B.json')); print('B=$B value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks']['sm_mhz'])"; done
for B in 8 64; do DSG_NO_GRAPH=1 timeout 600 python bench.py --steps 2 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/r2_bench_b${B}_nograph.json 2>> gpurun_out/r2_bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b${B}_nograph.json')); print('B=$B no-graph value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])"; done
