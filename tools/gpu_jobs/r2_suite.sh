mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --config coco --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_coco.json 2> gpurun_out/r2_bench_coco.err; tail -1 gpurun_out/r2_bench_coco.err
timeout 900 python bench.py --config n64w16 --num-steps 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n64w16.json 2> gpurun_out/r2_bench_n64w16.err; tail -1 gpurun_out/r2_bench_n64w16.err
python - <<'PY'
import json
for n in ('coco','n64w16'):
    d=json.load(open(f'gpurun_out/r2_bench_{n}.json')); print(n, 'value', d['value'], 'e2e', d['e2e']['value'], d['ms_per_step'], d['ms_per_pass'], d['clocks']['sm_mhz'])
PY
