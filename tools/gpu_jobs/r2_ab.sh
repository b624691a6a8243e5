mkdir -p gpurun_out
for v in 0 1 0 1; do
DSG_NO_SKIP2=$v timeout 900 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json')); e=d['e2e']
print('NO_SKIP2=$v value', round(d['value'],2), round(d['ms_per_pass'],3), d['clocks']['sm_mhz'], '| e2e', round(e['value'],2), round(e['ms_per_pass'],3), e['clocks']['sm_mhz'])
PY
done
