mkdir -p gpurun_out
python tools/pass_timeline.py --quiet > gpurun_out/r2_timeline_b.txt 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -2 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json')); print('value', d['value'], d['ms_per_step'], d['raw_denoiser_passes_per_step'], d['ms_per_pass'], d['clocks']); e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['raw_denoiser_passes_per_step'], e['ms_per_pass'], e['clocks']); print(d['cpu_baseline']['value'])
PY
