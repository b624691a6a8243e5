mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json')); print('value', d['value'], d['ms_per_step'], d['clocks']); e=d['e2e']; print('e2e', e); print(d.get('strong_scaling'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 | head -c 400
