mkdir -p gpurun_out
for flag in "" "--e2e-last"; do
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $flag > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -2 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json')); print('value', d['value'], d['ms_per_step'], d['raw_denoiser_passes_per_step'], d['ms_per_pass'], d['clocks']); e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['raw_denoiser_passes_per_step'], e['ms_per_pass'], e['clocks'])
PY
sleep 20
done
