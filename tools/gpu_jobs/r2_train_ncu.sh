mkdir -p gpurun_out
# ncu --set full of the hot kernels of the training iteration (tools/train_one_step.py: VG, batch 128, eager launches)
for spec in "attention_bwd:window_attention_bwd_tc_kernel:2" "transpose:transpose_colsum_kernel:4" "ln_bwd:ln_bwd_kernel:2" "wgrad_gemm:gemm_kernel:24"; do
  IFS=: read name pat cnt <<< "$spec"
  skip=0; [ "$name" = wgrad_gemm ] && skip=150     # past the forward GEMMs: the dgrad / split-K wgrad launches of the last blocks
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o gpurun_out/r2_train_full_$name -f python tools/train_one_step.py --steps 1 > gpurun_out/ncu_train_$name.log 2>&1
  tail -1 gpurun_out/ncu_train_$name.log
  echo "ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt python tools/train_one_step.py --steps 1   (B200, VG, batch 128, training iteration, eager)" > gpurun_out/r2_train_ncu_full_$name.txt
  python tools/ncu_summary.py gpurun_out/r2_train_full_$name.ncu-rep --source 12 >> gpurun_out/r2_train_ncu_full_$name.txt 2>&1
  rm -f gpurun_out/r2_train_full_$name.ncu-rep
done
du -sh gpurun_out
