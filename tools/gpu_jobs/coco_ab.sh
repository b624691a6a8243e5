mkdir -p gpurun_out
run() { # label, env..., extra flags
  label=$1; shift
  env "$@" timeout 900 python bench.py --config coco --steps 2 --warmup 2 --no-cpu-baseline $EXTRA > gpurun_out/coco_ab.json 2> gpurun_out/coco_ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/coco_ab.json')); e=d['e2e']
print('$label value', round(d['value'],1), round(d['ms_per_pass'],3), d['clocks']['sm_mhz'], '| e2e', round(e['value'],1), round(e['ms_per_pass'],3), e['clocks']['sm_mhz'])
PY
}
EXTRA="" run default X=1
EXTRA="--e2e-last" run e2e_last X=1
EXTRA="" run no_graph DSG_NO_GRAPH=1
EXTRA="--profile-stride 1000000" run no_eager_steps X=1
