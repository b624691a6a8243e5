mkdir -p gpurun_out
# ncu --set full of launches of each hot kernel of the final state (tools/one_pass.py, VG batch 512, padding skipping on);
# the summaries are produced on the box, the reports themselves are only kept where they are small
for spec in "block_tail:block_tail_kernel:1" "block_head:block_head_kernel:1" "fused_mlp:fused_mlp_kernel:1" "gemm:gemm_kernel:14" "attention:window_attention_tc_kernel:9" "proj_ln:proj_ln_kernel:2" "patch_embed:patch_embed_kernel:4"; do
  IFS=: read name pat cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -c $cnt -o gpurun_out/r2_full_$name -f python tools/one_pass.py --passes 0 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  echo "ncu --set full --clock-control none --import-source on -k regex:$pat -c $cnt python tools/one_pass.py --passes 0   (B200, VG, batch 512, r2 final state, padding skipping on)" > gpurun_out/r2_ncu_full_$name.txt
  python tools/ncu_summary.py gpurun_out/r2_full_$name.ncu-rep --source 12 >> gpurun_out/r2_ncu_full_$name.txt 2>&1
  [ "$name" = gemm ] && rm -f gpurun_out/r2_full_gemm.ncu-rep
done
rm -f gpurun_out/r2_full_patch_embed.ncu-rep gpurun_out/r2_full_attention.ncu-rep
du -sh gpurun_out
