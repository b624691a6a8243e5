mkdir -p gpurun_out
# final state of round 2: the whole GPU suite, an ncu --set full capture of the MN-major weight-gradient GEMMs and of the
# attention backward, the sampling bench line
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gputests_r2_final.log; cat gpurun_out/gputests_r2_final.log
for spec in "wgrad_mn:gemm_kernel:150:16" "attention_bwd_final:window_attention_bwd_tc_kernel:0:2"; do
  IFS=: read name pat skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o gpurun_out/r2_train_full_$name -f python tools/train_one_step.py --steps 1 > gpurun_out/ncu_train_$name.log 2>&1
  tail -1 gpurun_out/ncu_train_$name.log
  echo "ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt python tools/train_one_step.py --steps 1   (B200, VG, batch 128, training iteration, eager, final state)" > gpurun_out/r2_train_ncu_full_$name.txt
  python tools/ncu_summary.py gpurun_out/r2_train_full_$name.ncu-rep --source 8 >> gpurun_out/r2_train_ncu_full_$name.txt 2>&1
  rm -f gpurun_out/r2_train_full_$name.ncu-rep
done
timeout 600 python bench.py > gpurun_out/bench_r2_v10.json 2> gpurun_out/bench_r2_v10.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_v10.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
